"""BASELINE.json configs[4]: 2000-phoneme utterance, chunked vocoder.  Reports time to first audio chunk and total RTF."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vosk_tts_b200 import config as C, synthetic, weights
from vosk_tts_b200.engine import Engine

def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    prec = int(os.environ.get("PRECISION", "1"))
    cfg = C.DEFAULT_CONFIG
    blob, man = weights.pack(weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234)), cfg)
    eng = Engine(cfg, blob, man, device=0, precision=prec)
    ids = np.random.RandomState(9).randint(0, cfg["n_vocab"], size=(1, T)).astype(np.int64)
    res = []
    for rep in range(3):
        t0 = time.perf_counter()
        first = None
        n = 0
        for c in eng.synthesize_stream(ids, 2, (0.8, 1.0, 0.8), chunk_frames=chunk, seed=3):
            if first is None:
                first = time.perf_counter() - t0
            n += c.size
        tot = time.perf_counter() - t0
        t1 = time.perf_counter()
        wav, yl = eng.infer(ids, [T], [2], (0.8, 1.0, 0.8), seed=3)
        mono = time.perf_counter() - t1
        res.append((first, tot, mono, n))
    first, tot, mono, n = min(res)
    print(json.dumps({"workload": "configs[4]: %d phonemes, %d-frame chunks, 24-frame halo" % (T, chunk), "precision_mode": prec,
                      "samples": n, "audio_s": n / 22050.0, "time_to_first_chunk_ms": first * 1e3, "streamed_total_ms": tot * 1e3,
                      "monolithic_ms": min(r[2] for r in res) * 1e3, "rtf_streamed": tot / (n / 22050.0),
                      "rtf_monolithic": min(r[2] for r in res) / (n / 22050.0)}))

if __name__ == "__main__":
    main()
