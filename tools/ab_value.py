"""A/B of two builds of libvtts.so on ONE box: device-resident configs[1] step time (CUDA events, L2 flush between steps),
alternating the libraries in separate processes.  usage: python tools/ab_value.py libA.so libB.so [rounds]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
from vosk_tts_b200 import config as C, synthetic, weights
from vosk_tts_b200.engine import Engine
cfg = C.DEFAULT_CONFIG
blob, man = weights.pack(weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234)), cfg)      # full pack: also what older builds expect
eng = Engine(cfg, blob, man, device=0, precision=1)
g = torch.Generator().manual_seed(0)
tok = torch.randint(0, cfg["n_vocab"], (1, 128), generator=g)
dev = torch.device("cuda", 0)
d_ids, d_sid = tok.to(dev), torch.tensor([2], device=dev)
lens = np.array([128], np.int64)
scales = np.array([0.8, 1.0, 0.8], np.float32)
yl = eng.durations_dev(d_ids.data_ptr(), lens, d_sid.data_ptr(), 1, 128, scales, 0, seed=3)
Ty = int(yl[0])
d_wav = torch.zeros(1, (Ty + 64) * eng.hop, device=dev)
est = torch.cuda.ExternalStream(eng.stream(), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    return eng.infer_dev(d_ids.data_ptr(), lens, d_sid.data_ptr(), 1, 128, scales, d_wav.data_ptr(), (Ty + 64) * eng.hop, seed=3)
for _ in range(8): step()
ms = []
for _ in range(40):
    flush.fill_(1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(est); step(); e1.record(est); e1.synchronize()
    ms.append(e0.elapsed_time(e1))
ms.sort()
print(json.dumps({"lib": os.environ.get("VTTS_LIB"), "frames": Ty, "median_ms": ms[len(ms)//2], "min_ms": ms[0], "mean_ms": sum(ms)/len(ms)}))
''' % ROOT


def main():
    libs = sys.argv[1:3]
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    for r in range(rounds):
        for lib in libs:
            env = dict(os.environ, VTTS_LIB=os.path.abspath(lib))
            out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
            print((out.stdout.strip().splitlines() or [out.stderr[-400:]])[-1])


if __name__ == "__main__":
    main()
