"""Secondary measurement: BASELINE.json configs[2] -- batch of 64 utterances of 64..256 phonemes in ONE call
(ragged, packed), Philox noise, precision mode from $PRECISION (default 1).  Prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vosk_tts_b200 import config as C, synthetic, weights
from vosk_tts_b200.engine import Engine

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    prec = int(os.environ.get("PRECISION", "1"))
    cfg = C.DEFAULT_CONFIG
    blob, man = weights.pack(weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234)), cfg)
    eng = Engine(cfg, blob, man, device=0, precision=prec)
    g = torch.Generator().manual_seed(1)
    lens = torch.randint(64, 257, (B,), generator=g).numpy().astype(np.int64)
    ids = torch.randint(0, cfg["n_vocab"], (B, int(lens.max())), generator=g).numpy().astype(np.int64)
    sid = torch.randint(0, 5, (B,), generator=g).numpy().astype(np.int64)
    dev = torch.device("cuda", 0)
    d_ids, d_sid = torch.as_tensor(ids, device=dev), torch.as_tensor(sid, device=dev)
    scales = np.array([0.8, 1.0, 0.8], np.float32)
    yl = eng.durations_dev(d_ids.data_ptr(), lens, d_sid.data_ptr(), B, ids.shape[1], scales, 0, seed=7)
    maxf = int(yl.max())
    d_wav = torch.zeros(B, maxf * eng.hop, device=dev)
    eng.synthesize_dev(d_wav.data_ptr(), maxf * eng.hop)
    est = torch.cuda.ExternalStream(eng.stream(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def step():
        return eng.infer_dev(d_ids.data_ptr(), lens, d_sid.data_ptr(), B, ids.shape[1], scales, d_wav.data_ptr(), maxf * eng.hop, seed=7)
    for _ in range(3):
        step()
    ms = []
    for _ in range(steps):
        flush.fill_(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(est); yl = step(); e1.record(est); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    eng.profile(True)
    for _ in range(3):
        step()
    prof = eng.profile_read(); eng.profile(False)
    stage = eng.stage_timings()
    n = int(yl.sum()) * eng.hop
    t = sum(ms) / len(ms)
    out = {"workload": "configs[2]: batch %d, 64-256 phonemes, one call" % B, "precision_mode": prec, "ms_per_step": t,
           "samples_per_step": n, "samples_per_s": n / (t / 1e3), "frames": int(yl.sum()), "phonemes": int(lens.sum()),
           "rtf": (t / 1e3) / (n / 22050.0), "stage_ms": stage,
           "conv_tc": {"ms_per_step": prof["tc_ms"] / 3, "tflops_algorithmic": prof["tc_flops"] / (prof["tc_ms"] / 1e3) / 1e12 if prof["tc_ms"] else 0,
                       "launches": prof["tc_launches"] / 3},
           "conv_ffma": {"ms_per_step": prof["conv_ms"] / 3, "tflops": prof["conv_flops"] / (prof["conv_ms"] / 1e3) / 1e12 if prof["conv_ms"] else 0,
                         "launches": prof["conv_launches"] / 3}}
    print(json.dumps(out))

if __name__ == "__main__":
    main()
