"""In-graph kernel timeline of one utterance: entry-to-entry intervals of consecutive kernels (ns, %globaltimer)."""
import os, sys, re, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vosk_tts_b200 import config as C, synthetic, weights
from vosk_tts_b200.engine import Engine
import bench

def kernel_names():
    names = {}
    for fn in ("kernels.cuh", "conv_tc.cuh", "attn_tc.cuh", "wn_tc.cuh"):
        cur = None
        for i, line in enumerate(open(os.path.join(ROOT, "vosk_tts_b200", "csrc", fn)), 1):
            m = re.search(r"^(?:__global__.*?\s|)(\w+_kernel)\s*\(", line)
            if m: cur = m.group(1)
            if "PDL_LAUNCH();" in line and cur: names[(fn, i)] = cur
    return names

def main():
    cfg = C.DEFAULT_CONFIG
    wl = bench.workload(cfg)
    prec = int(os.environ.get("PRECISION", "1"))
    blob, man = weights.pack(weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234)), cfg)
    eng = Engine(cfg, blob, man, device=0, precision=prec)
    run = lambda: eng.infer(wl["tok"], wl["lens"], wl["sid"], wl["scales"], wl["eps_dp"], lambda mf: wl["eps_z"][:, :, :mf])
    for _ in range(4): run()          # graphs captured
    eng.timeline(1)
    run()
    tl = eng.timeline(2)
    eng.timeline(0)
    print("replays", eng.graph_replays(), "stamps", len(tl))
    names = {ln: nm for (fn, ln), nm in kernel_names().items()}
    tl = tl[np.argsort(tl[:, 1])]
    t0 = int(tl[0, 1])
    agg = collections.OrderedDict()
    for i in range(len(tl)):
        ln, t = int(tl[i, 0]), int(tl[i, 1])
        dt = (int(tl[i + 1, 1]) - t) / 1e3 if i + 1 < len(tl) else 0.0
        ln = ln - (1 << 64) if ln >= (1 << 63) else ln
        nm = names.get(ln, "  .stamp%d" % ln)
        if os.environ.get("VERBOSE"): print("%4d %9.2f us  +%7.2f  %s" % (i, (t - t0) / 1e3, dt, nm))
        if dt < 300: agg.setdefault(nm, []).append(dt)
    tot = sum(sum(v) for v in agg.values())
    print("total (first to last entry) %.1f us" % ((int(tl[-1, 1]) - t0) / 1e3))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-28s n=%3d sum %7.1f us  avg %6.2f  min %6.2f  max %6.2f" % (k, len(v), sum(v), sum(v) / len(v), min(v), max(v)))

if __name__ == "__main__":
    main()
