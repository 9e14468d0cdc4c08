"""TEST INFRASTRUCTURE (not collected): where does a never-seen utterance spend its time?  Per-call wall time of
VitsSession.run for (A) one utterance repeated with caller noise, (B) the same with engine-drawn noise, (C) distinct
utterances of one length, (D) distinct lengths -- each after its own warm-up."""
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
g = torch.Generator().manual_seed(7)
scales = np.array([0.8, 1.0, 0.8], np.float32)
def feeds(T):
    return {"input": torch.randint(0, 62, (1, T), generator=g).numpy().astype(np.int64), "input_lengths": np.array([T], np.int64),
            "scales": scales, "sid": np.array([2], np.int64), "bert": None, "phone_duration_extra": None}
def timed(fn, n, warm):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e3)
    ts.sort()
    return "median %.3f  min %.3f  p90 %.3f  max %.3f ms" % (ts[len(ts) // 2], ts[0], ts[int(len(ts) * 0.9)], ts[-1])
fA = feeds(128)
sess.run(None, fA)
Ty = int(sess.last_y_lengths[0])
noise = {"dp": torch.randn(1, 2, 128, generator=g).numpy(), "z": torch.randn(1, 192, 400, generator=g).numpy()}
sess.run(None, fA, noise={"dp": noise["dp"], "z": noise["z"]})
Ty = int(sess.last_y_lengths[0])
nz = {"dp": noise["dp"], "z": np.ascontiguousarray(noise["z"][:, :, :Ty])}
print("A  same utterance, caller noise     :", timed(lambda: sess.run(None, fA, noise=nz), 30, 5))
print("B  same utterance, engine noise     :", timed(lambda: sess.run(None, fA), 30, 5))
print("C  distinct utterances, T = 128     :", timed(lambda: sess.run(None, feeds(128)), 30, 30))
print("D  distinct utterances, T = 100..128:", timed(lambda: sess.run(None, feeds(int(torch.randint(100, 129, (1,), generator=g)))), 30, 60))
print("   speculation", eng.speculation_stats(), "replays", eng.graph_replays())
os.environ["X"] = "1"

# ---- GPU-side span (first to last kernel entry stamp, %globaltimer) next to the wall time of the same call
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def span(fn, label, do_flush):
    res = []
    for _ in range(8):
        if do_flush:
            flush.fill_(1)
        torch.cuda.synchronize()
        eng.timeline(1)
        t = time.perf_counter(); fn(); wall = (time.perf_counter() - t) * 1e3
        tl = eng.timeline(2); eng.timeline(0)
        ts = np.sort(tl[:, 1].astype(np.int64))
        gaps = np.diff(ts)
        res.append((wall, (ts[-1] - ts[0]) / 1e6, len(ts), gaps.max() / 1e3, int(np.argmax(gaps))))
    res.sort()
    w, s_, n, g_, gi = res[len(res) // 2]
    print("%-44s wall %.3f ms  gpu span %.3f ms  stamps %d  largest gap %.1f us at stamp %d" % (label, w, s_, n, g_, gi))
span(lambda: sess.run(None, fA, noise=nz), "A same utterance, caller noise", False)
span(lambda: sess.run(None, fA), "B same utterance, engine noise", False)
span(lambda: sess.run(None, feeds(128)), "C distinct utterances", False)
span(lambda: sess.run(None, fA, noise=nz), "A + L2 flush", True)
span(lambda: sess.run(None, fA), "B + L2 flush", True)
span(lambda: sess.run(None, feeds(128)), "C + L2 flush", True)

# ---- does switching between captured graphs cost anything?  Two FIXED utterances (caller noise => fixed durations)
def fixed(T, seed):
    gg = torch.Generator().manual_seed(seed)
    f = {"input": torch.randint(0, 62, (1, T), generator=gg).numpy().astype(np.int64), "input_lengths": np.array([T], np.int64),
         "scales": scales, "sid": np.array([2], np.int64), "bert": None, "phone_duration_extra": None}
    nd = {"dp": torch.randn(1, 2, T, generator=gg).numpy(), "z": torch.randn(1, 192, 600, generator=gg).numpy()}
    sess.run(None, f, noise=nd)
    ty = int(sess.last_y_lengths[0])
    return f, {"dp": nd["dp"], "z": np.ascontiguousarray(nd["z"][:, :, :ty])}, ty
u1, n1, t1 = fixed(128, 11)
u2, n2, t2 = fixed(128, 12)
u3, n3, t3 = fixed(100, 13)
print("frames of the three fixed utterances:", t1, t2, t3)
state = {"i": 0}
def alt(pairs):
    def f():
        u, n = pairs[state["i"] % len(pairs)]; state["i"] += 1
        sess.run(None, u, noise=n)
    return f
print("E1 one fixed utterance                     :", timed(alt([(u1, n1)]), 30, 6))
print("E2 two fixed utterances alternating (T=128):", timed(alt([(u1, n1), (u2, n2)]), 30, 6))
print("E3 two fixed utterances alternating (128/100):", timed(alt([(u1, n1), (u3, n3)]), 30, 6))
