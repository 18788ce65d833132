"""Sharding of cell sets over the GPUs of one node.

Every (gene, set) test is independent, so DErand shards by draw and DEobs by region: the
staged matrix is replicated on every GPU and each rank tests its own sets.  Sets are always
generated / enumerated on every rank in the same global order before they are split, so the
tables do not depend on the number of ranks: sums are integer, tails are a pure function of
(k, M, n, N).  What happens to the results:

* :class:`GeneExchange` / :func:`run_gene_sharded` - tables stay on the GPUs, re-sharded by
  GENE: the test kernel stores its raw 8-byte accumulator words straight into the memory of
  the GPU that owns the gene block (CUDA IPC over NVLink / NVSwitch, an all-to-all fused into
  the kernel), and every GPU finishes ALL sets x ITS genes (``spade_expand_block``).  No GPU
  is a funnel, and the layout is the one the per-gene consumers need (gamma fit over the
  draws, empirical p-values, column-compressed MAT tables).
* :func:`run_sharded` - tables wanted on the host of rank 0: every rank delivers its rows over
  its own PCIe link into host tables shared by the node (``gather="host"``), or one NCCL
  gather per table over NVLink (``gather="nccl"``, what north_star names; gloo in the CPU
  tests).
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


def gene_blocks(G, genes_per_part, world_size):
    """Gene range owned by each of ``world_size`` ranks in the gene-sharded exchange: consecutive
    runs of whole gene partitions of the staged layout, as even as the partition count allows.
    List of ``(gene_begin, gene_end)``; a rank may own nothing when there are fewer partitions
    than ranks."""
    nparts = (G + genes_per_part - 1) // genes_per_part
    cuts = [(q * nparts) // world_size for q in range(world_size + 1)]
    return [(min(G, cuts[q] * genes_per_part), min(G, cuts[q + 1] * genes_per_part))
            for q in range(world_size)]


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


class PeerBuffers:
    """``count`` device buffers of ``nbytes`` in EVERY rank's HBM, each mapped into every rank of
    the node: rank r allocates its own (``spade_device_alloc``), exports CUDA IPC handles, and
    opens everybody else's (``spade_ipc_open``, peer access over NVLink / NVSwitch).
    ``ptr[i][r]`` is the address, valid in this process, of buffer ``i`` of rank ``r``.
    ``torch.distributed`` carries the 64-byte handles."""

    def __init__(self, nbytes, rank, world_size, device, count=1):
        import torch.distributed as dist

        from . import _lib
        self._lib = _lib.load()
        self.rank, self.world, self.device, self.nbytes = rank, world_size, device, int(nbytes)
        self.ptr = [[None] * world_size for _ in range(count)]
        mine, err = [], None
        try:
            for i in range(count):
                p = ctypes.c_void_p()
                self._ok(self._lib.spade_device_alloc(device, ctypes.byref(p), self.nbytes))
                self.ptr[i][rank] = p.value
                buf = ctypes.create_string_buffer(64)
                self._ok(self._lib.spade_ipc_export(device, p, buf))
                mine.append(bytes(buf.raw))
        except RuntimeError as exc:                    # every rank must reach the exchange
            err = str(exc)
        everyone = [None] * world_size
        dist.all_gather_object(everyone, err if err else mine)
        bad = [e for e in everyone if isinstance(e, str)]
        try:
            if bad:
                raise RuntimeError(bad[0])
            for r in range(world_size):
                if r == rank:
                    continue
                for i in range(count):
                    p = ctypes.c_void_p()
                    self._ok(self._lib.spade_ipc_open(device, everyone[r][i], ctypes.byref(p)))
                    self.ptr[i][r] = p.value
        except RuntimeError:
            self.close()
            raise

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError("peer buffers: " + self._lib.spade_last_error(None).decode())

    def close(self):
        for row in self.ptr:
            for r, p in enumerate(row):
                if p is None:
                    continue
                if r == self.rank:
                    self._lib.spade_device_free(self.device, ctypes.c_void_p(p))
                else:
                    self._lib.spade_ipc_close(self.device, ctypes.c_void_p(p))
        self.ptr = []


class GeneExchange:
    """Multi-GPU form of the batched test with the tables re-sharded by gene (see the module
    docstring).  One object per rank and engine; ``max_rows`` = the largest number of sets (over
    ALL ranks) of one step.

    ``step(members, offsets, first_row, row_step, set_sizes)`` is asynchronous.  On the current
    stream: the routed raw launch of this step, then the expansion of the PREVIOUS step's gene block
    (the two compete for the same SMs, so they are not overlapped).  On a side stream: a 4-byte NCCL
    all-reduce per step ("every rank's words of this step have landed"), which the expansion of that
    step waits for one step later - by then it has normally passed, so the ranks do not run in
    lockstep and a rank may be up to one step behind the others.  Four word buffers are used in
    turn: the kernel of step i + 1 follows this rank's wait for the all-reduce of step i - 1, which
    completes only after every rank has launched its kernel of step i - 1, i.e. finished its
    expansion of step i - 3 - the last reader of the buffer step i + 1 writes into.  ``drain()``
    finishes the last step; then ``tables[name][:rows]`` holds ALL sets x the genes
    ``[gene_begin, gene_end)`` of this rank, rows in global set order."""

    BUFFERS = 4

    def __init__(self, engine, max_rows, tables=("up", "down", "cpm"), want_k=False, rank=None,
                 world_size=None, device=None):
        import torch
        import torch.distributed as dist

        r, lr, P = world()
        self.rank = r if rank is None else rank
        self.world = P if world_size is None else world_size
        self.device = lr if device is None else device
        self.engine, self.max_rows = engine, int(max_rows)
        if engine.wide_genes:
            raise RuntimeError("gene-sharded exchange needs the raw form, which this matrix cannot use "
                               "(wide genes, see spade_wide_genes); use run_sharded(gather='host')")
        self.blocks = gene_blocks(engine.G, engine.genes_per_part, self.world)
        self.gene_begin, self.gene_end = self.blocks[self.rank]
        self.width = max(1, max(b - a for a, b in self.blocks))
        self.words = PeerBuffers(self.max_rows * self.width * 8, self.rank, self.world, self.device,
                                 count=self.BUFFERS)
        dev = torch.device("cuda", self.device)
        cols = max(1, self.gene_end - self.gene_begin)
        self.cols = cols
        self.tables = {n: torch.empty((self.max_rows, cols), dtype=torch.float64, device=dev) for n in tables}
        self.k = torch.empty((self.max_rows, cols), dtype=torch.int32, device=dev) if want_k else None
        self.side = torch.cuda.Stream(device=dev, priority=-1)      # all-reduces (and the bench's table copies)
        self.flag = torch.zeros(1, device=dev)
        self.pending = None                            # (buffer, set sizes, rows, all-reduce event) of the last step
        self.steps = 0
        self._dist = dist

    def step(self, members, offsets, first_row, row_step, set_sizes):
        import torch

        rows = len(set_sizes)
        assert rows <= self.max_rows
        buf = self.steps % self.BUFFERS
        self.steps += 1
        main = torch.cuda.current_stream(self.device)
        blocks = [(self.words.ptr[buf][q], self.width, a, b) for q, (a, b) in enumerate(self.blocks)]
        if len(offsets) > 1:
            self.engine.run_raw_blocks(members, offsets, blocks, first_row, row_step, stream=main)
        stored = torch.cuda.Event()
        stored.record(main)
        self.side.wait_event(stored)
        with torch.cuda.stream(self.side):
            self.flag.zero_()
            self._dist.all_reduce(self.flag)           # passes once every rank's kernel of this step has finished
        landed = torch.cuda.Event()
        landed.record(self.side)
        self._finish_pending(main)                     # the previous step, behind this step's kernel
        self.pending = (buf, np.array(set_sizes, dtype=np.int64), rows, landed)
        self.rows = rows

    def _finish_pending(self, stream):
        if self.pending is None:
            return
        buf, set_sizes, rows, landed = self.pending
        self.pending = None
        stream.wait_event(landed)
        if self.gene_end > self.gene_begin and rows:
            self.engine.expand_block(self.words.ptr[buf][self.rank], set_sizes, self.gene_begin,
                                     self.gene_end - self.gene_begin, self.tables, k=self.k,
                                     raw_row_stride=self.width, row_stride=self.cols, stream=stream)

    def drain(self):
        """Finish the last step on the current stream."""
        import torch
        self._finish_pending(torch.cuda.current_stream(self.device))

    def close(self):
        import torch
        torch.cuda.synchronize(self.device)
        self.words.close()
        self.tables = {}


def run_gene_sharded(engine, sets, tables=("up", "down", "cpm"), balance=False, want_k=False, exchange=None):
    """Test ``sets`` over all ranks and leave the tables ON THE GPUS, sharded by gene.

    Returns ``(exchange, order)``: ``exchange.tables[name][:len(sets)]`` is a CUDA tensor
    ``[S, gene_end - gene_begin]`` of this rank's gene block; ``order[i]`` = index in ``sets`` of
    table row ``i`` (identity for round-robin sharding; with ``balance=True`` the rows are grouped
    by the rank that tested them).  Pass ``exchange`` to reuse the buffers of an earlier call."""
    import torch
    import torch.distributed as dist

    rank, local_rank, P = world()
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sizes = np.array([len(s) for s in sets], dtype=np.int64)
    assignment = plan(sizes, P, balance=balance)
    mine = [sets[i] for i in assignment[rank]]
    members, offsets = engine.pack_sets(mine)
    S = len(sets)
    if exchange is None:
        exchange = GeneExchange(engine, S, tables=tables, want_k=want_k)
    if balance:
        order = np.concatenate(assignment)
        first_row, row_step = int(sum(len(a) for a in assignment[:rank])), 1
    else:
        order = np.arange(S, dtype=np.int64)
        first_row, row_step = rank, P
    exchange.step(members.astype(np.int32), offsets, first_row, row_step, sizes[order])
    exchange.drain()
    failure = None
    try:
        engine.check()
    except Exception as exc:                               # noqa: BLE001
        failure = exc
    _raise_together(failure, rank)
    return exchange, order


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
    CPUs next to that GPU while it does so).  ``touch_block=(first_col, last_col)``: the same for
    a rank that will write those columns of every row."""

    DTYPE = {"up": np.float64, "down": np.float64, "fc": np.float64, "cpm": np.float64, "k": np.int32}

    def __init__(self, names, rows, cols, rank, world_size, pin=True, directory="/dev/shm", touch=None,
                 device=None, touch_block=None):
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
        if touch_block is not None:
            # the same for a rank that will write a block of COLUMNS of every row (gene block)
            with _near_gpu(device):
                for n in self.names:
                    self.arrays[n][:, touch_block[0]:touch_block[1]] = 0
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


def _raise_together(failure, rank):
    """Collective: every rank learns whether any rank failed, and all raise together (a bare
    barrier would leave the healthy ranks waiting forever for the one that raised)."""
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    bad = torch.tensor([0 if failure is None else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX)
    if failure is not None:
        raise failure
    if int(bad) != 0:
        raise RuntimeError(f"rank {rank}: another rank failed while running its share of the sets")


def _shared_tables(names, rows, cols, rank, world_size, touch=None):
    """One set of shared host tables per (names, cols), grown when more rows are needed."""
    key = (tuple(names), cols)
    cur = _SHARED.get(key)
    if cur is not None and cur.rows >= rows:
        return cur
    if cur is not None:
        _SHARED.pop(key)                 # a failed regrow must not leave a closed object behind
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
        failure = None
        try:
            if len(mine):
                engine.run(members=members, offsets=offsets, out=out, tables=tables, want_k=want_k,
                           row_stride=stride)
        except Exception as exc:                         # noqa: BLE001 - re-raised on every rank below
            failure = exc
        _raise_together(failure, rank)                   # also the barrier: all rows are in place
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
