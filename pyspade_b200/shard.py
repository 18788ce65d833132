"""Sharding of cell sets over the GPUs of one node.

Every (gene, set) test is independent, so DErand shards by draw and DEobs by region: the
staged matrix is replicated on every GPU, each rank runs its own sets, and one gather of the
per-set result tables to rank 0 is the only communication (NCCL over NVLink for device
tensors, gloo for the CPU tests).  Sets are always generated / enumerated on every rank in
the same global order before they are split, so the tables do not depend on the number of
ranks: sums are integer, tails are a pure function of (k, M, n, N).
"""
import ctypes
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


def expander_shares(total_sets, world_size, expand_cost=0.12):
    """Contiguous shares of ``total_sets`` for the raw-word gather, where rank 0 also expands
    the words of all other ranks (``spade_expand``, about ``expand_cost`` of the cost of
    testing a set): rank 0 takes fewer sets so that every rank finishes together,
        x (1 - e) + T e = (T - x) / (P - 1),   T = total_sets, e = expand_cost.
    Returns ``counts`` (list, rank 0 first); the ranks > 0 get equal counts."""
    P, T, e = world_size, total_sets, expand_cost
    if P <= 1:
        return [T]
    x = T * (1.0 / (P - 1) - e) / (1.0 - e + 1.0 / (P - 1))
    y = int(np.ceil((T - max(x, 0.0)) / (P - 1)))
    x = T - y * (P - 1)
    while x < 0:                                       # tiny T: give the remainder to the peers
        y -= 1
        x = T - y * (P - 1)
    return [x] + [y] * (P - 1)


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


class PeerTables:
    """Result tables that live in rank 0's HBM and are mapped into every rank of the node.

    Rank 0 allocates ``rows x cols`` elements per table (``spade_device_alloc``), exports
    CUDA IPC handles, the other ranks open them (``spade_ipc_open``, peer access over
    NVLink / NVSwitch).  A rank passes ``address(name, first_row)`` to
    ``Engine.run_device(..., row_stride=...)``: the test kernel then stores its results
    directly into rank 0's memory while it computes - the gather is fused into the kernel and
    needs no collective, only a barrier before rank 0 reads.  ``torch.distributed`` is used for
    the handle exchange and the barrier."""

    ITEM = {"up": 8, "down": 8, "fc": 8, "cpm": 8, "k": 4, "raw": 8}

    def __init__(self, names, rows, cols, rank, world_size, device):
        import torch.distributed as dist

        from . import _lib
        self._lib = _lib.load()
        self.names, self.rows, self.cols = list(names), int(rows), int(cols)
        self.rank, self.device = rank, device
        self.ptr = {}
        handles = [None]
        if rank == 0:
            try:
                hs = {}
                for name in self.names:
                    p = ctypes.c_void_p()
                    self._ok(self._lib.spade_device_alloc(device, ctypes.byref(p),
                                                          self.rows * self.cols * self.ITEM[name]))
                    self.ptr[name] = p.value
                    buf = ctypes.create_string_buffer(64)
                    self._ok(self._lib.spade_ipc_export(device, p, buf))
                    hs[name] = bytes(buf.raw)
                handles = [hs]
            except RuntimeError as exc:                # every rank must leave the broadcast together
                handles = [str(exc)]
        dist.broadcast_object_list(handles, src=0)
        if isinstance(handles[0], str):
            self.close()
            raise RuntimeError(handles[0])
        if rank != 0:
            for name in self.names:
                p = ctypes.c_void_p()
                self._ok(self._lib.spade_ipc_open(device, handles[0][name], ctypes.byref(p)))
                self.ptr[name] = p.value

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError("peer tables: " + self._lib.spade_last_error(None).decode())

    def address(self, name, first_row=0, first_col=0):
        return self.ptr[name] + (first_row * self.cols + first_col) * self.ITEM[name]

    def tensor(self, name):
        """Rank 0: the table as a torch tensor (no copy)."""
        import torch

        assert self.rank == 0
        dtype, typestr = {"k": (torch.int32, "<i4"), "raw": (torch.int64, "<i8")}.get(name, (torch.float64, "<f8"))

        class _View:
            pass

        v = _View()
        v.__cuda_array_interface__ = {"shape": (self.rows, self.cols), "typestr": typestr,
                                      "data": (self.ptr[name], False), "version": 2}
        t = torch.as_tensor(v, device=torch.device("cuda", self.device))
        assert t.dtype == dtype
        return t

    def close(self):
        for name, p in self.ptr.items():
            if self.rank == 0:
                self._lib.spade_device_free(self.device, ctypes.c_void_p(p))
            else:
                self._lib.spade_ipc_close(self.device, ctypes.c_void_p(p))
        self.ptr = {}


class _near_gpu:
    """Context manager: run the calling thread on the CPUs NVML reports as local to GPU
    ``device`` (so that first-touched pages come from that GPU's memory node), then restore the
    previous affinity.  A no-op when NVML or the affinity calls are unavailable."""

    def __init__(self, device):
        self.device, self.saved = device, None

    def __enter__(self):
        if self.device is None:
            return self
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[self.device]) if visible else self.device
            self.saved = os.sched_getaffinity(0)
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        except Exception:                                  # noqa: BLE001 - best effort only
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except OSError:
                pass
        return False


class SharedHostTables:
    """Result tables in host memory shared by the ranks of one node (POSIX shared memory,
    ``/dev/shm``), page-locked in every rank (``spade_host_register``).  Each rank hands
    strided views of its own rows to ``Engine.run(..., row_stride=...)``: every GPU then
    delivers its rows over its own PCIe link, in parallel, and rank 0 reads the complete
    tables from the same mapping after a barrier.  No GPU-to-GPU traffic at all - the right
    path when the tables are wanted on the host (file writers, the e2e benchmark).
    ``touch=(first_row, row_step, count)``: the rows this rank will write, zero-filled by it
    before the tables are pinned (NUMA first touch; with ``device`` the thread is moved to the
    CPUs next to that GPU while it does so)."""

    DTYPE = {"up": np.float64, "down": np.float64, "fc": np.float64, "cpm": np.float64, "k": np.int32}

    def __init__(self, names, rows, cols, rank, world_size, pin=True, directory="/dev/shm", touch=None,
                 device=None):
        import torch.distributed as dist

        from . import _lib
        self._lib = _lib.load()
        self.names, self.rows, self.cols, self.rank = list(names), int(rows), int(cols), rank
        need = sum(self.rows * self.cols * np.dtype(self.DTYPE[n]).itemsize for n in self.names)
        tag = [None]
        if rank == 0:
            st = os.statvfs(directory)
            if st.f_bavail * st.f_frsize < need + (64 << 20):
                tag = [""]                                   # not enough shared memory
            else:
                tag = [f"spade_b200_{os.getpid()}_{np.random.default_rng().integers(1 << 30)}"]
        dist.broadcast_object_list(tag, src=0)
        if not tag[0]:
            raise MemoryError(f"{directory} cannot hold {need} bytes of result tables")
        self.paths = {n: os.path.join(directory, f"{tag[0]}_{n}") for n in self.names}
        self.arrays, self._pinned = {}, []
        if rank == 0:
            for n in self.names:
                self.arrays[n] = np.memmap(self.paths[n], dtype=self.DTYPE[n], mode="w+",
                                           shape=(self.rows, self.cols))
        dist.barrier()
        if rank != 0:
            for n in self.names:
                self.arrays[n] = np.memmap(self.paths[n], dtype=self.DTYPE[n], mode="r+",
                                           shape=(self.rows, self.cols))
        if touch is not None:
            # first touch: a rank's own rows get their pages from the memory node the rank runs
            # on, instead of all tables landing on the node of whoever pins them first
            first_row, row_step, count = touch
            with _near_gpu(device):
                for n in self.names:
                    self.rows_view(n, first_row, row_step, count)[:] = 0
            dist.barrier()
        if pin:
            for n in self.names:
                a = self.arrays[n]
                if self._lib.spade_host_register(ctypes.c_void_p(a.ctypes.data), a.nbytes) == 0:
                    self._pinned.append(a.ctypes.data)
        dist.barrier()
        if rank == 0:                                        # mappings stay valid after unlink
            for pth in self.paths.values():
                os.unlink(pth)

    def rows_view(self, name, first_row, row_step=1, count=None):
        v = self.arrays[name][first_row::row_step]
        return v if count is None else v[:count]

    def close(self):
        for p in self._pinned:
            self._lib.spade_host_unregister(ctypes.c_void_p(p))
        self._pinned = []
        self.arrays = {}


_SHARED = {}


def _shared_tables(names, rows, cols, rank, world_size, touch=None):
    """One set of shared host tables per (names, cols), grown when more rows are needed."""
    key = (tuple(names), cols)
    cur = _SHARED.get(key)
    if cur is not None and cur.rows >= rows:
        return cur
    if cur is not None:
        cur.close()
    _SHARED[key] = SharedHostTables(names, rows, cols, rank, world_size, touch=touch, device=world()[1])
    return _SHARED[key]


def run_sharded(engine, sets, tables=("up", "down", "fc", "cpm"), balance=False, want_k=False,
                gather="host"):
    """Run ``sets`` over all ranks of the torchrun job (or locally when there is one rank).

    Returns on rank 0 a dict of numpy arrays ``[S, G]`` in the order of ``sets``; ``None``
    elsewhere.  With one rank this is ``engine.run``.  With several ranks:

    * ``gather="host"`` (default): every rank writes its rows into host tables shared by the
      node (:class:`SharedHostTables`), each GPU over its own PCIe link; falls back to
      ``"nccl"`` when ``/dev/shm`` is too small;
    * ``gather="nccl"``: every rank keeps its tables on its GPU, one NCCL gather per table
      brings them to rank 0 over NVLink, rank 0 copies them to the host."""
    rank, local_rank, P = world()
    if P == 1:
        return engine.run(sets, tables=tables, want_k=want_k)
    import torch
    import torch.distributed as dist

    gather = os.environ.get("SPADE_GATHER", gather)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assignment = plan([len(s) for s in sets], P, balance=balance)
    mine = [sets[i] for i in assignment[rank]]
    members, offsets = engine.pack_sets(mine)
    names = list(tables) + (["k"] if want_k else [])
    S, G = len(sets), engine.G

    if gather == "host":
        if balance:
            touch = (int(sum(len(a) for a in assignment[:rank])), 1, len(mine))
        else:
            touch = (rank, P, len(mine))
        try:
            shared = _shared_tables(names, S, G, rank, P, touch=touch)
        except MemoryError:
            gather = "nccl"
    if gather == "host":
        if balance:        # contiguous block per rank, rows put back in set order by rank 0
            first = int(sum(len(a) for a in assignment[:rank]))
            out = {n: shared.rows_view(n, first, 1, len(mine)) for n in names}
            stride = G
        else:              # round-robin: set i of rank r is row i*P + r
            out = {n: shared.rows_view(n, rank, P, len(mine)) for n in names}
            stride = P * G
        if len(mine):
            engine.run(members=members, offsets=offsets, out=out, tables=tables, want_k=want_k,
                       row_stride=stride)
        dist.barrier()
        res = None
        if rank == 0:
            if balance:
                order = np.concatenate(assignment)
                res = {}
                for n in names:
                    full = np.empty((S, G), dtype=shared.DTYPE[n])
                    full[order] = shared.arrays[n][:S]
                    res[n] = full
            else:
                res = {n: np.array(shared.arrays[n][:S]) for n in names}
        dist.barrier()                                   # tables may be overwritten after this
        return res

    engine._validate_members(members, offsets)
    dev = torch.device("cuda", engine.device)
    out = {k: torch.empty((len(mine), G), dtype=torch.float64, device=dev) for k in tables}
    kd = torch.empty((len(mine), G), dtype=torch.int32, device=dev) if want_k else None
    dmem = torch.from_numpy(members.astype(np.int32)).to(dev)
    engine.run_device(dmem, offsets, out, k=kd)
    if want_k:
        out["k"] = kd
    full = gather_rows(out, assignment, rank, P, G)
    torch.cuda.synchronize(dev)
    if rank != 0:
        return None
    return {k: v.cpu().numpy() for k, v in full.items()}
