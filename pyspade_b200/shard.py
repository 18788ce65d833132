"""Sharding of cell sets over the GPUs of one node.

Every (gene, set) test is independent, so DErand shards by draw and DEobs by region: the
staged matrix is replicated on every GPU, each rank runs its own sets, and one gather of the
per-set result tables to rank 0 is the only communication (NCCL over NVLink for device
tensors, gloo for the CPU tests).  Sets are always generated / enumerated on every rank in
the same global order before they are split, so the tables do not depend on the number of
ranks: sums are integer, tails are a pure function of (k, M, n, N).
"""
import os

import numpy as np


def world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) without it."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def plan(sizes, world_size, balance=False):
    """Assign set i to a rank.  Default: round-robin (rank = i mod P), the DErand rule - all
    draws of a bin have the same size.  ``balance=True`` (DEobs, ragged regions): longest
    processing time first on the member counts, ties broken by set index, so the plan is a
    pure function of ``sizes``.  Returns a list of index arrays, one per rank."""
    S = len(sizes)
    if world_size <= 1:
        return [np.arange(S, dtype=np.int64)]
    if not balance:
        return [np.arange(r, S, world_size, dtype=np.int64) for r in range(world_size)]
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.lexsort((np.arange(S), -sizes))
    load = np.zeros(world_size, dtype=np.int64)
    owner = np.zeros(S, dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))            # first minimum: deterministic
        owner[i] = r
        load[r] += max(int(sizes[i]), 1)
    return [np.flatnonzero(owner == r).astype(np.int64) for r in range(world_size)]


def gather_rows(local, assignment, rank, world_size, n_cols, group=None):
    """Gather per-rank tables to rank 0 and put the rows back in global set order.

    ``local``: dict name -> torch tensor ``[len(assignment[rank]), n_cols]`` (CPU with gloo,
    CUDA with nccl).  Returns on rank 0 a dict name -> tensor ``[S, n_cols]`` on the same
    device; ``None`` on the other ranks.  One ``gather`` per table (NCCL has no native
    gather: torch issues grouped send/recv over NVLink)."""
    import torch
    import torch.distributed as dist

    if world_size == 1:
        return dict(local)
    S = int(sum(len(a) for a in assignment))
    rows_max = max(len(a) for a in assignment)
    out = {} if rank == 0 else None
    for name in sorted(local):
        t = local[name]
        pad = t
        if t.shape[0] != rows_max:                         # gather wants equal shapes
            pad = torch.zeros((rows_max, n_cols), dtype=t.dtype, device=t.device)
            pad[: t.shape[0]] = t
        bufs = None
        if rank == 0:
            bufs = [torch.empty((rows_max, n_cols), dtype=t.dtype, device=t.device)
                    for _ in range(world_size)]
        dist.gather(pad.contiguous(), bufs, dst=0, group=group)
        if rank == 0:
            full = torch.empty((S, n_cols), dtype=t.dtype, device=t.device)
            for r in range(world_size):
                rows = torch.as_tensor(assignment[r], device=t.device)
                full[rows] = bufs[r][: len(assignment[r])]
            out[name] = full
    return out


def run_sharded(engine, sets, tables=("up", "down", "fc", "cpm"), balance=False, want_k=False):
    """Run ``sets`` over all ranks of the torchrun job (or locally when there is one rank).

    Returns on rank 0 a dict of numpy arrays ``[S, G]`` in the order of ``sets``; ``None``
    elsewhere.  With one rank this is ``engine.run``: host tables through the pipelined
    device->host path.  With several ranks each rank keeps its tables on its GPU, one NCCL
    gather brings them to rank 0, rank 0 copies them to the host."""
    rank, local_rank, P = world()
    if P == 1:
        return engine.run(sets, tables=tables, want_k=want_k)
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assignment = plan([len(s) for s in sets], P, balance=balance)
    mine = [sets[i] for i in assignment[rank]]
    dev = torch.device("cuda", engine.device)
    members, offsets = engine.pack_sets(mine)
    engine._validate_members(members, offsets)
    out = {k: torch.empty((len(mine), engine.G), dtype=torch.float64, device=dev) for k in tables}
    kd = torch.empty((len(mine), engine.G), dtype=torch.int32, device=dev) if want_k else None
    dmem = torch.from_numpy(members.astype(np.int32)).to(dev)
    engine.run_device(dmem, offsets, out, k=kd)
    if want_k:
        out["k"] = kd
    full = gather_rows(out, assignment, rank, P, engine.G)
    torch.cuda.synchronize(dev)
    if rank != 0:
        return None
    return {k: v.cpu().numpy() for k, v in full.items()}
