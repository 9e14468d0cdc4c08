"""Monotonic Alignment Search: CUDA kernel (device-resident scores, vtts_maximum_path_dev) vs the reference's compiled Cython
kernel (oracle/_ref, if present on this box) at a training-like shape.  Prints one JSON line."""
import glob
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from vosk_tts_b200 import monotonic_align as MA  # noqa: E402


def main():
    B, Ty, Tx = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (32, 1000, 200)))
    rng = np.random.RandomState(0)
    nc = (rng.randn(B, Ty, Tx) * 4).astype(np.float32)
    ty = rng.randint(Ty // 2, Ty + 1, size=B).astype(np.int32)
    tx = np.array([rng.randint(Tx // 3, min(Tx, t) + 1) for t in ty], np.int32)
    mask = np.zeros((B, Ty, Tx), np.float32)
    for b in range(B):
        mask[b, : ty[b], : tx[b]] = 1
    d_nc, d_mask = torch.as_tensor(nc, device="cuda"), torch.as_tensor(mask, device="cuda")
    for _ in range(3):
        attn = MA.maximum_path(d_nc, d_mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        attn = MA.maximum_path(d_nc, d_mask)
    e1.record()
    e1.synchronize()
    gpu_ms = e0.elapsed_time(e1) / 10
    out = {"shape": [B, Ty, Tx], "gpu_ms_per_call_incl_clone_and_mask_sums": gpu_ms,
           "algorithmic_bytes": int(sum(int(a) * int(b) for a, b in zip(ty, tx)) * 8 + B * Ty * Tx * 4)}
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_mas_core*.so"))
    if so:
        spec = importlib.util.spec_from_file_location("ref_mas_core", so[0])
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            v = d_nc.cpu().numpy().astype(np.float32)                  # what the reference wrapper does around the kernel
            p = np.zeros(v.shape, np.int32)
            ref.maximum_path_c(p, v, ty, tx)
            back = torch.from_numpy(p).to(device="cuda", dtype=d_nc.dtype)
            torch.cuda.synchronize()
            t.append(time.perf_counter() - t0)
        out["reference_cpu_ms_per_call_incl_copies"] = 1e3 * min(t)
        out["same_path"] = bool(torch.equal(back, attn))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
