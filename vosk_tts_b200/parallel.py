"""Multi-GPU plumbing: one process per GPU, utterance-level sharding, ONE weight broadcast at init.

The path shards naturally: utterances share nothing but read-only weights, so
there is no collective on the per-utterance path.  `torch.distributed` is used only to (1) broadcast the
packed weight blob from rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests) and (2) barrier /
max-reduce timings in bench.py.
"""
import numpy as np
import torch
import torch.distributed as dist


def lpt_shards(lengths, world_size):
    """Longest-processing-time-first assignment of utterances to ranks (cost ~ phoneme count, a proxy
    for frames).  Returns a list of index lists, one per rank; deterministic for equal inputs."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards


def broadcast_packed(blob, manifest, src=0, device=None):
    """Broadcast (blob float32[n], manifest str) from `src` to every rank.

    On rank `src` pass the packed weights; other ranks pass (None, None).  Returns (tensor, manifest) where
    `tensor` lives on `device` (a CUDA device for NCCL, CPU for gloo).  One collective for the 127 MB blob,
    one small one for its size + manifest."""
    rank = dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device(device if device is not None else ("cuda" if backend == "nccl" else "cpu"))
    if rank == src:
        mbytes = manifest.encode()
        hdr = torch.tensor([int(blob.size), len(mbytes)], dtype=torch.int64, device=dev)
    else:
        hdr = torch.zeros(2, dtype=torch.int64, device=dev)
    dist.broadcast(hdr, src)
    n, mlen = int(hdr[0]), int(hdr[1])
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(blob, dtype=np.float32)).to(dev)
        m = torch.frombuffer(bytearray(mbytes), dtype=torch.uint8).to(dev)
    else:
        t = torch.empty(n, dtype=torch.float32, device=dev)
        m = torch.empty(mlen, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    dist.broadcast(m, src)
    return t, bytes(m.cpu().numpy().tobytes()).decode()
