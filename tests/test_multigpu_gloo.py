"""CPU, world_size=2 over gloo: the N>1 host path (weight broadcast from rank 0, utterance sharding)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vosk_tts_b200 import parallel
    blob = manifest = None
    if rank == 0:
        blob = np.arange(100003, dtype=np.float32) * 0.5
        manifest = "a 0 64\nb 64 99939\n"
    t, m = parallel.broadcast_packed(blob, manifest, src=0)
    lens = np.random.RandomState(7).randint(64, 257, size=16)
    shards = parallel.lpt_shards(lens, world)
    mine = torch.tensor([float(lens[shards[rank]].sum())])
    tot = mine.clone()
    dist.all_reduce(tot)
    q.put((rank, float(t.double().sum()), m, shards[rank], float(tot), float(lens.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    expect = float((np.arange(100003, dtype=np.float64) * 0.5).sum())
    for rank, s, m, shard, tot, total in res:
        assert abs(s - expect) < 1e-3 * expect
        assert m == "a 0 64\nb 64 99939\n"
        assert tot == total
    assert sorted(res[0][3] + res[1][3]) == list(range(16))
