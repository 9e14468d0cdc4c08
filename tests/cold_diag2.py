"""TEST INFRASTRUCTURE (not collected): host-side breakdown of never-seen utterances through VitsSession.run (bench e2e_cold)."""
import os, sys, time, threading
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vosk_tts_b200 import config as C, synthetic, weights
from vosk_tts_b200.engine import Engine
from vosk_tts_b200.session import VitsSession
cfg = C.DEFAULT_CONFIG
blob, man = weights.pack(weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234)), cfg, precision=1)
eng = Engine(cfg, blob, man, device=0, precision=1)
sess = VitsSession.__new__(VitsSession)
sess.cfg, sess.engine, sess._lock, sess._seed, sess._calls = cfg, eng, threading.Lock(), 0, 0
sess.last_y_lengths = sess.last_wav_lengths = None
g = torch.Generator().manual_seed(4242)
scales = np.array([0.8, 1.0, 0.8], np.float32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def fresh():
    T = int(torch.randint(100, 129, (1,), generator=g))
    return {"input": torch.randint(0, 62, (1, T), generator=g).numpy().astype(np.int64), "input_lengths": np.array([T], np.int64),
            "scales": scales, "sid": np.array([int(torch.randint(0, 200, (1,), generator=g))], np.int64), "bert": None, "phone_duration_extra": None}
for _ in range(60):
    sess.run(None, fresh())
for do_flush in (False, True):
    rows = []
    for _ in range(30):
        f = fresh()
        if do_flush:
            flush.fill_(1)
        torch.cuda.synchronize()
        t = time.perf_counter(); sess.run(None, f); wall = (time.perf_counter() - t) * 1e3
        rows.append([wall] + [v / 1e3 for v in eng.host_timings()[:5]] + [eng.host_timings()[5], int(sess.last_y_lengths[0]), f["input"].shape[1]])
    rows.sort()
    med = rows[len(rows) // 2]
    print("flush=%d  median call: wall %.3f | C side: enqueue1 %.3f enqueue2 %.3f wait %.3f copy-out %.3f total %.3f | spec %d frames %d T %d"
          % tuple([do_flush] + med))
    print("   walls:", " ".join("%.2f" % r[0] for r in rows))
    print("   spec misses in sample:", sum(1 for r in rows if r[6] == 2.0), " non-speculative:", sum(1 for r in rows if r[6] == 0.0))
print("speculation", eng.speculation_stats(), "replays", eng.graph_replays())
