"""Scratch (run under torchrun, one rank per GPU): aggregate device->host bandwidth of all ranks
copying 792 MB each (the three tables of a 1 000-draw batch) into (a) their own page-locked
buffer, (b) their gene block of host tables shared by the node (POSIX shm, page-locked)."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyspade_b200 import _lib, shard                      # noqa: E402
from pyspade_b200.engine import PinnedArray               # noqa: E402

rank, local_rank, world = shard.world()
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
lib = _lib.load()
S, G = 1000 * world, 33000
blocks = shard.gene_blocks(G, 1024, world)
g0, g1 = blocks[rank]
w = g1 - g0
src = torch.randn((S, max(w, 1)), dtype=torch.float64, device="cuda")
stream = torch.cuda.current_stream()
res = {}


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


own = [PinnedArray((S, max(w, 1)), np.float64) for _ in range(3)]
def to_own():
    for p in own:
        lib.spade_memcpy2d(ctypes.c_void_p(p.array.ctypes.data), 8 * max(w, 1), ctypes.c_void_p(src.data_ptr()),
                           8 * max(w, 1), 8 * w, S, 1, ctypes.c_void_p(stream.cuda_stream))
dt = timed(to_own)
res["own_pinned_GBps"] = 3 * S * G * 8 / dt / 1e9

for touch in (False, True):
    sh = shard.SharedHostTables(("up", "down", "cpm"), S, G, rank, world, device=local_rank,
                                touch_block=(g0, g1) if touch else None)
    def to_shared():
        for n in ("up", "down", "cpm"):
            lib.spade_memcpy2d(ctypes.c_void_p(sh.arrays[n].ctypes.data + 8 * g0), 8 * G, ctypes.c_void_p(src.data_ptr()),
                               8 * max(w, 1), 8 * w, S, 1, ctypes.c_void_p(stream.cuda_stream))
    dt = timed(to_shared)
    res[f"shared_shm_{'numa_touch' if touch else 'plain'}_GBps"] = 3 * S * G * 8 / dt / 1e9
    sh.close()
if rank == 0:
    res["world"] = world
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
