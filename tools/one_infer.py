"""Runs N inferences of the headline workload (for ncu / quick timing)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vosk_tts_b200 import config as C, synthetic, weights
from vosk_tts_b200.engine import Engine
import bench

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    cfg = C.DEFAULT_CONFIG
    wl = bench.workload(cfg)
    prec = int(os.environ.get('PRECISION', '1'))
    blob, man = weights.pack(weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234)), cfg, precision=prec)
    eng = Engine(cfg, blob, man, device=0, precision=prec)
    eng.set_graphs(False)
    for i in range(n):
        t0 = time.perf_counter()
        wav, yl = eng.infer(wl["tok"], wl["lens"], wl["sid"], wl["scales"], wl["eps_dp"], lambda mf: wl["eps_z"][:, :, :mf])
        print("infer %d: %.3f ms  %s launches=%d" % (i, (time.perf_counter() - t0) * 1e3, eng.stage_timings(), eng.kernel_launches()))

if __name__ == "__main__":
    main()
