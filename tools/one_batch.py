"""Runs N calls of BASELINE configs[2] (64 utterances, 64-256 phonemes, one call, eager launches) -- for ncu captures of the
machine-filling launches (conv_tc_kernel<128>, attn_tc_kernel)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vosk_tts_b200 import config as C, synthetic, weights
from vosk_tts_b200.engine import Engine

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    prec = int(os.environ.get("PRECISION", "1"))
    cfg = C.DEFAULT_CONFIG
    blob, man = weights.pack(weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234)), cfg, precision=prec)
    eng = Engine(cfg, blob, man, device=0, precision=prec)
    eng.set_graphs(False)
    g = torch.Generator().manual_seed(1)
    lens = torch.randint(64, 257, (B,), generator=g).numpy().astype(np.int64)
    ids = torch.randint(0, cfg["n_vocab"], (B, int(lens.max())), generator=g).numpy().astype(np.int64)
    sid = torch.randint(0, 5, (B,), generator=g).numpy().astype(np.int64)
    for i in range(n):
        t0 = time.perf_counter()
        wav, yl = eng.infer(ids, lens, sid, (0.8, 1.0, 0.8), seed=7)
        print("call %d: %.2f ms, %d frames, launches=%d" % (i, (time.perf_counter() - t0) * 1e3, int(yl.sum()), eng.kernel_launches()))

if __name__ == "__main__":
    main()
