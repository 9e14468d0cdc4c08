"""GPU Monotonic Alignment Search with the reference's call signature (training/vits2/monotonic_align/__init__.py:6-22):

    attn = maximum_path(neg_cent, mask)        # neg_cent, mask: [b, t_t, t_s]  ->  path [b, t_t, t_s], dtype of neg_cent

The reference copies neg_cent to the host, runs a Cython loop per batch item and copies the path back; here the scores stay
on the GPU when they are CUDA tensors (vtts_maximum_path_dev on the tensor's device and current stream); numpy / CPU inputs go
through the host entry point vtts_maximum_path.  There is no CPU fallback: without the CUDA library this raises.
"""
import ctypes as C

import numpy as np

from .engine import VttsError, load_library


def _lengths(mask):
    # t_t_max = mask.sum(1)[:, 0], t_s_max = mask.sum(2)[:, 0]   (__init__.py:17-18)
    return mask.sum(1)[:, 0], mask.sum(2)[:, 0]


def maximum_path_numpy(neg_cent, t_ys, t_xs, device=0):
    """neg_cent float32 [B, T_y, T_x], lengths int [B] -> int32 path [B, T_y, T_x]."""
    lib = load_library()
    v = np.ascontiguousarray(neg_cent, dtype=np.float32)
    B, Ty, Tx = v.shape
    ty = np.ascontiguousarray(t_ys, dtype=np.int32).reshape(B)
    tx = np.ascontiguousarray(t_xs, dtype=np.int32).reshape(B)
    path = np.zeros((B, Ty, Tx), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.vtts_maximum_path(p(v), p(ty), p(tx), B, Ty, Tx, p(path), int(device))
    if rc != 0:
        raise VttsError(rc, lib.vtts_last_error(None).decode())
    return path


def maximum_path(neg_cent, mask):
    """Drop-in for monotonic_align.maximum_path: torch tensors in, torch tensor (same device and dtype) out."""
    import torch
    t_t, t_s = _lengths(mask)
    if not neg_cent.is_cuda:
        path = maximum_path_numpy(neg_cent.detach().cpu().numpy(), t_t.cpu().numpy(), t_s.cpu().numpy())
        return torch.from_numpy(path).to(device=neg_cent.device, dtype=neg_cent.dtype)
    lib = load_library()
    v = neg_cent.detach().to(torch.float32).contiguous().clone()          # the kernel accumulates in place
    B, Ty, Tx = v.shape
    ty = t_t.to(device=v.device, dtype=torch.int32).contiguous()
    tx = t_s.to(device=v.device, dtype=torch.int32).contiguous()
    if bool((tx > ty).any()):
        raise VttsError(-1, "maximum_path: more tokens than frames in a batch item")
    path = torch.empty((B, Ty, Tx), dtype=torch.int32, device=v.device)
    with torch.cuda.device(v.device):
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.vtts_maximum_path_dev(C.c_void_p(v.data_ptr()), C.c_void_p(ty.data_ptr()), C.c_void_p(tx.data_ptr()), B, Ty, Tx,
                                       C.c_void_p(path.data_ptr()), C.c_void_p(stream))
    if rc != 0:
        raise VttsError(rc, lib.vtts_last_error(None).decode())
    return path.to(dtype=neg_cent.dtype)
