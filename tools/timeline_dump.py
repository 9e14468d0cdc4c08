"""Raw in-graph timeline (source line of the stamp, %globaltimer ns) of one configs[1] utterance -> .npy (for offline A/B of two
builds: VTTS_LIB selects the library, the line numbers refer to that build's sources)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vosk_tts_b200 import config as C, synthetic, weights
from vosk_tts_b200.engine import Engine
import bench

cfg = C.DEFAULT_CONFIG
wl = bench.workload(cfg)
blob, man = weights.pack(weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234)), cfg)
eng = Engine(cfg, blob, man, device=0, precision=1)
run = lambda: eng.infer(wl["tok"], wl["lens"], wl["sid"], wl["scales"], wl["eps_dp"], lambda mf: wl["eps_z"][:, :, :mf])
for _ in range(4):
    run()
best = None
for rep in range(5):
    eng.timeline(1)
    run()
    tl = eng.timeline(2)
    eng.timeline(0)
    tl = tl[np.argsort(tl[:, 1])]
    span = int(tl[-1, 1]) - int(tl[0, 1])
    if best is None or span < best[0]:
        best = (span, tl)
np.save(sys.argv[1], best[1])
print(sys.argv[1], "stamps", len(best[1]), "span us", best[0] / 1e3)
