#!/usr/bin/env python
"""Benchmark of the hypergeometric DE hot path (BASELINE.json metric: gene x draw
hypergeometric tests per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (``config.workload``): BASELINE.json ``configs[2]`` restricted to one size bin -
DErand, 1 000 random draws of 500 cells per GPU and step, 50 000 cells x 33 000 genes,
complement background, synthetic Poisson counts, draws from the
reference's RNG call sequence (np.random.seed(0); np.random.choice(arange(C), 500, False)).
One "step" = one pass of the hot path over that batch: 33 M (gene, draw) tests per GPU.

``value``   tests/s with member lists and result tables resident in HBM (CUDA events on the
            launch stream, barrier + synchronize on both sides, max over ranks); with N > 1
            every rank runs its own 1 000 draws (weak scaling) and the timed region includes
            the single NCCL gather of the up / down / cpm tables to rank 0.
``e2e``     the same metric through the public API with HOST buffers: member index lists in
            host memory, result tables delivered to pinned host memory (H2D + kernels +
            D2H inside the timed region; + the NCCL gather and rank 0's D2H when N > 1).
``roofline`` dominant kernel (k_de_fused: member gather + tails + table writes in one
            launch): algorithmic bytes of SURVEY section 8-d (rho * min(8 N, C/8) + 24 B per
            test) / its CUDA-event time, against the measured HBM copy bandwidth of
            MEASURED_PEAKS.json.
``cpu_baseline`` oracle per-gene port (oracle/de_port.py, kind "port": the Python reference
            cannot travel to the GPU box; the port is ~8x faster per core than the reference
            itself) on a bounded sample of the same workload, all host
            cores; the same sample is the parity check of the GPU tables (``parity``).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

G_GENES, C_CELLS, N_MEMBERS, S_DRAWS = 33000, 50000, 500, 1000
TABLES = ("up", "down", "cpm")               # DErand keeps these three (DErand.py:219-221)
B_OUT = 8 * len(TABLES)
SAMPLE_GENES, SAMPLE_SETS = 4096, 24      # ~10 s of CPU work on 16 cores; also the parity sample
METRIC = "gene x draw hypergeometric tests per second"
UNIT = "tests/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed regions."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.lines:
            if not any(a <= ts <= b + 0.1 for a, b in windows):
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except Exception:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(rho, sizes, C, G, b_out):
    gather = float(sum(rho * min(8.0 * n, C / 8.0) for n in sizes)) * G
    return gather, gather + float(len(sizes)) * G * b_out


def make_draws(total_sets, C, N):
    """The reference's draw sequence (DErand.py:27,197), generated once on the host so the
    stream does not depend on the number of ranks."""
    from pyspade_b200 import synth
    return synth.derand_draws(C, N, total_sets, seed=0)


def sample_matrix(indptr, indices, counts, G, C, n_genes):
    """Host float64 CPM rows of a density-stratified gene sample (``synth.sample_rows``)."""
    from pyspade_b200 import synth
    return synth.sample_rows(indptr, indices, counts, G, C, n_genes)


def cpu_port_leg(X, sets, nproc):
    """Time the oracle's per-gene port (reference structure: one pool per cell set)."""
    from oracle import de_port
    Gs = X.shape[0]
    res = {k: np.zeros((len(sets), Gs)) for k in ("up", "down", "fc", "cpm")}
    t0 = time.perf_counter()
    for s, members in enumerate(sets):
        N, up, down, fc, cpm = de_port.perform_DE_port(
            members, X, np.arange(Gs), nproc, np.zeros(Gs), np.zeros(Gs), np.ones(Gs), np.zeros(Gs))
        res["up"][s], res["down"][s], res["fc"][s], res["cpm"][s] = up, down, fc, cpm
    dt = time.perf_counter() - t0
    return dt, res


def rel_err(got, want):
    ok = (np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
          and np.array_equal(got == 0, want == 0))
    m = np.isfinite(want) & (want != 0)
    err = float(np.max(np.abs(got[m] - want[m]) / np.abs(want[m]))) if m.any() else 0.0
    return ok, err


# --------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores.  The Python
    reference tree does not exist on the GPU box, so this is the oracle per-gene port."""
    if rank != 0:
        return
    import torch
    nproc = os.cpu_count() or 1
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    from pyspade_b200 import synth
    if dev == "cuda":
        indptr, indices, counts = synth.counts_torch(G_GENES, C_CELLS, seed=0, device="cuda:0")
    else:
        m = synth.counts_numpy(SAMPLE_GENES, C_CELLS, seed=0)          # no GPU: sample rows only
        indptr, indices, counts = (torch.from_numpy(m.indptr.astype(np.int64)),
                                   torch.from_numpy(m.indices.astype(np.int32)),
                                   torch.from_numpy(m.data.astype(np.int32)))
    Gm = indptr.numel() - 1
    genes, X = sample_matrix(indptr, indices, counts, Gm, C_CELLS, min(SAMPLE_GENES, Gm))
    draws = make_draws(SAMPLE_SETS, C_CELLS, N_MEMBERS)
    per_step_sets = 2
    times = []
    for step in range(args.warmup + args.steps):
        sets = [draws[(step * per_step_sets + j) % len(draws)] for j in range(per_step_sets)]
        dt, _ = cpu_port_leg(X, sets, nproc)
        if step >= args.warmup:
            times.append(dt)
    tests = X.shape[0] * per_step_sets
    total = float(sum(times))
    value = tests * len(times) / total
    sample = (f"{X.shape[0]} density-stratified genes x {per_step_sets} draws of {N_MEMBERS} cells "
              f"per step at full C={C_CELLS}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"DErand bin N={N_MEMBERS}: {S_DRAWS} draws x {G_GENES} genes x "
                                   f"{C_CELLS} cells per GPU per step (CPU arm: bounded sample)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the per-rank tables reach rank 0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine, PinnedArray

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    G, C, N, S = G_GENES, C_CELLS, N_MEMBERS, S_DRAWS
    t0 = time.time()
    indptr, indices, counts = synth.counts_torch(G, C, seed=0, device=dev)      # replicated matrix
    torch.cuda.synchronize()
    nnz = int(indices.numel())
    rho = nnz / (G * C)
    eng = Engine.from_device_csr(G, C, indptr, indices, counts, device=local_rank)
    stage_s = time.time() - t0

    # draws: one global host stream, rank r takes draws r, r+P, ...
    n_batches = 2
    all_draws = make_draws(S * world * n_batches, C, N)
    batches = []
    for b in range(n_batches):
        mine = all_draws[b * S * world + rank: (b + 1) * S * world: world]
        members, offsets = synth.concat_sets(mine)
        batches.append((mine, members, offsets, torch.from_numpy(members).to(dev)))

    out_dev = {k: torch.empty((S, G), dtype=torch.float64, device=dev) for k in TABLES}
    gathered = None
    peer = None
    gather_mode = "none"
    if world > 1:
        # Preferred: two [world*S, G] buffers of raw accumulator words in rank 0's HBM mapped into
        # every rank (CUDA IPC).  A peer's test kernel stores its 8-byte words straight into rank
        # 0's memory over NVLink while it computes (3x fewer bytes than the three float64
        # tables); rank 0 writes its own rows directly and finishes the peers' rows with
        # k_expand; rank 0 takes a smaller share of the draws to make up for that work.
        # Fallback: one NCCL gather per table.
        from pyspade_b200 import shard
        ok = torch.ones(1, device=dev)
        if args.gather == "p2p":
            try:
                peer = [shard.PeerTables(["raw"], S * world, G, rank, world, local_rank) for _ in range(2)]
            except Exception as exc:                                   # noqa: BLE001
                print(f"[rank {rank}] peer tables unavailable: {exc}", file=sys.stderr)
                ok.zero_()
        else:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) > 0:
            gather_mode = ("raw accumulator words stored into rank 0 from inside the test kernel "
                           "(CUDA IPC over NVLink), finished on rank 0 by k_expand")
        else:
            if peer is not None:
                for pt in peer:
                    pt.close()
                peer = None
            gather_mode = "nccl gather of up/down/cpm tables to rank 0"
            if rank == 0:
                gathered = {k: [torch.empty((S, G), dtype=torch.float64, device=dev) for _ in range(world)]
                            for k in TABLES}
    step_flag = torch.zeros(1, device=dev)
    final = None
    batches_dev = batches
    if peer is not None:
        # rank 0 also expands everybody else's words: it takes a smaller share of the draws so
        # that all ranks finish together (shares are contiguous blocks of the global draw list:
        # rank r's sets are rows first[r] .. first[r] + share[r] - 1 of the tables on rank 0)
        share = shard.expander_shares(S * world, world)
        first = [int(sum(share[:r])) for r in range(world)]
        batches_dev = []
        for b in range(n_batches):
            mine = all_draws[b * S * world + first[rank]: b * S * world + first[rank] + share[rank]]
            members, offsets = synth.concat_sets(mine)
            batches_dev.append((mine, members, offsets, torch.from_numpy(members).to(dev)))
        if rank == 0:
            final = {k: torch.empty((S * world, G), dtype=torch.float64, device=dev) for k in TABLES}
            peer_sizes = np.full(S * world - share[0], N, dtype=np.int64)
    step_no = [0]

    side = torch.cuda.Stream(device=dev) if (peer is not None and rank == 0) else None
    expanded = [None]                       # event: the previous step's expansion has finished

    def step_peer(dmem, offsets):
        # Peers fill one word buffer while rank 0 expands the other.  Rank 0 expands step i on a
        # side stream while its main stream already computes step i+1 (the test kernel is bound
        # by the LSU data pipe, the expansion by memory); the expansion of step i has to finish
        # before rank 0 releases the peers into step i+2 (they reuse its word buffer) - i.e.
        # before rank 0 enters the all-reduce of step i+1.
        buf = peer[step_no[0] & 1]
        step_no[0] += 1
        if rank == 0:
            main = torch.cuda.current_stream(dev)
            eng.run_device(dmem, offsets, {k: final[k][:share[0]] for k in TABLES})     # rank 0's own rows
            if expanded[0] is not None:
                main.wait_event(expanded[0])
            dist.all_reduce(step_flag)      # passes once every peer's words have landed
            landed = torch.cuda.Event()
            landed.record(main)
            side.wait_event(landed)
            eng.expand(buf.address("raw", first_row=share[0]), peer_sizes,
                       {k: final[k].data_ptr() + 8 * share[0] * G for k in TABLES}, stream=side)
            expanded[0] = torch.cuda.Event()
            expanded[0].record(side)
        else:
            eng.run_raw(dmem, offsets, buf.address("raw", first_row=first[rank]))
            dist.all_reduce(step_flag)

    def drain_peer():
        if side is not None and expanded[0] is not None:
            torch.cuda.current_stream(dev).wait_event(expanded[0])

    def gather_tables():
        if world == 1:
            return
        for k in TABLES:
            dist.gather(out_dev[k], gathered[k] if rank == 0 else None, dst=0)

    def step_device(i):
        _, _, offsets, dmem = batches_dev[i % n_batches]
        if peer is not None:
            step_peer(dmem, offsets)
        else:
            eng.run_device(dmem, offsets, out_dev)
            gather_tables()

    # N > 1: host tables shared by the ranks of the node (/dev/shm, page-locked in every rank);
    # each GPU delivers its rows over its own PCIe link, rank 0 reads the same mapping
    shared = None
    e2e_mode = "Engine.run into pinned host tables"
    if world > 1:
        from pyspade_b200 import shard
        try:
            shared = shard.SharedHostTables(TABLES, S * world, G, rank, world, touch=(rank, world, S),
                                            device=local_rank)
            e2e_mode = ("Engine.run on every rank into host tables shared by the node "
                        "(POSIX shm, page-locked; rows interleaved by rank)")
        except Exception as exc:                                       # noqa: BLE001
            print(f"[rank {rank}] shared host tables unavailable: {exc}", file=sys.stderr)
            e2e_mode = "device tables gathered to rank 0, rank 0 copies them to pinned host memory"
    shared_out = ({k: shared.rows_view(k, rank, world, S) for k in TABLES} if shared is not None else None)

    # pinned host tables for the end-to-end leg (rank 0 receives every rank's tables)
    host_rows = S * world if (rank == 0 and shared is None) else S
    pinned = {k: PinnedArray((host_rows, G), np.float64) for k in TABLES}
    host_out = {k: pinned[k].array for k in TABLES}
    host_out_t = {k: torch.from_numpy(host_out[k]) for k in TABLES}

    def step_e2e(i):
        mine, members, offsets, _ = batches[i % n_batches]
        if world == 1:
            eng.run(members=members, offsets=offsets, out=host_out, validate=False, tables=TABLES)
            return
        if shared is not None:
            eng.run(members=members, offsets=offsets, out=shared_out, validate=False, tables=TABLES,
                    row_stride=world * G)
            dist.all_reduce(step_flag)              # every rank's rows are in the shared tables
            torch.cuda.synchronize()
            return
        dmem = torch.from_numpy(members).to(dev, non_blocking=True)          # H2D of the index lists
        if peer is not None:
            _, members, offsets, _ = batches_dev[i % n_batches]
            dmem = torch.from_numpy(members).to(dev, non_blocking=True)
            step_peer(dmem, offsets)
            drain_peer()
            if rank == 0:
                for k in TABLES:
                    host_out_t[k].copy_(final[k], non_blocking=True)
        else:
            eng.run_device(dmem, offsets, out_dev)
            gather_tables()
            if rank == 0:
                for k in TABLES:
                    for r in range(world):
                        host_out_t[k][r * S:(r + 1) * S].copy_(gathered[k][r], non_blocking=True)
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.timing(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        drain_peer()                        # N > 1: rank 0's last expansion belongs to the timed region
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        w1 = time.time()
        dev_ms = e0.elapsed_time(e1)
        kt = eng.timing(reset=True)
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), kt, (w0, w1)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, _, kt, win_a = timed(step_device, args.steps, args.warmup)
    _, e2e_ms, _, win_b = timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.stop([win_a, win_b]) if rank == 0 else None

    tests_per_step = S * G * world
    value = tests_per_step * args.steps / (dev_ms * 1e-3)
    e2e_value = tests_per_step * args.steps / (e2e_ms * 1e-3)

    # roofline of the dominant kernel (per rank; rank 0's kernel times)
    peak, peak_src = peaks()
    launches_per_step = 1                      # one fused launch per step (+ the tail-table fill)
    sets_rank0 = share[0] if peer is not None else S          # rank 0's kernel times are reported
    gather_bytes, path_bytes = algorithmic_bytes(rho, [N] * sets_rank0, C, G, B_OUT)
    test_ms = kt["test_ms"] / args.steps
    table_ms = kt["table_ms"] / args.steps
    achieved = path_bytes / (test_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("k_de_fused_dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "k_de_fused", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": path_bytes / launches_per_step,
                "launch_ms": test_ms / launches_per_step, "launches_per_step": launches_per_step,
                "bytes_per_test": path_bytes / (sets_rank0 * G), "gather_bytes_per_launch": gather_bytes,
                "sets_per_launch": sets_rank0, "other_kernels_ms_per_step": table_ms}

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"DErand bin N={N}: {S} draws x {G} genes x {C} cells per GPU per step "
                                       f"(BASELINE.json configs[2], complement background)",
                           "cells": C, "genes": G, "draws_per_gpu": S, "set_size": N, "nnz": nnz,
                           "density": rho, "sharding": f"draws round-robin over {world} rank(s), matrix replicated",
                           "gather": gather_mode,
                           "draws_per_rank": (share if peer is not None else [S] * world),
                           "l2": "inputs larger than L2: 1.0 GB cell-major matrix read + 0.8 GB of tables written per step",
                           "stage_seconds": stage_s},
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps, "path": e2e_mode,
                        "h2d_bytes_per_step": int(world * (4 * S * N + 8 * (S + 1))),
                        "d2h_bytes_per_step": int(world * S * G * B_OUT)},
                "gpu_launches": int(kt["launches"]),
                "roofline": roofline, "clocks": clocks}

    # CPU baseline + parity on a bounded sample (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu:
        mine, members, offsets, dmem = batches[0]
        kdev = torch.empty((S, G), dtype=torch.int32, device=dev)
        full = {k: torch.empty((S, G), dtype=torch.float64, device=dev) for k in ("up", "down", "fc", "cpm")}
        eng.run_device(dmem, offsets, full, k=kdev)
        torch.cuda.synchronize()
        genes, X = sample_matrix(indptr, indices, counts, G, C, SAMPLE_GENES)
        sets = mine[:SAMPLE_SETS]
        nproc = os.cpu_count() or 1
        dt, cpu = cpu_port_leg(X, sets, nproc)
        gsel = torch.from_numpy(genes).to(dev)
        parity = {}
        for name in ("up", "down", "fc", "cpm"):
            got = full[name][:SAMPLE_SETS][:, gsel].cpu().numpy()
            ok, err = rel_err(got, cpu[name])
            parity[name] = {"special_values_match": ok, "max_rel_err": err}
        from oracle import de_port
        ora = de_port.de_tables(X, sets)
        med, n_dev = eng.gene_cutoffs()
        parity["k_exact"] = bool(np.array_equal(kdev[:SAMPLE_SETS][:, gsel].cpu().numpy(), ora["k"]))
        parity["n_exact"] = bool(np.array_equal(n_dev[genes], ora["n"]))
        parity["median_exact"] = bool(np.array_equal(med[genes], ora["median"]))
        parity["tests_checked"] = int(X.shape[0] * len(sets))
        line["cpu_baseline"] = {"value": X.shape[0] * len(sets) / dt, "unit": UNIT, "cores": nproc,
                                "kind": "port",
                                "sample": f"{X.shape[0]} density-stratified genes x {len(sets)} draws of "
                                          f"{N} cells at full C={C} ({dt:.1f} s)"}
        line["parity"] = parity
    # N > 1: rank 0 recomputes a few sets of another rank and compares them bit for bit with the
    # rows that rank stored into rank 0's tables
    if world > 1 and peer is not None:
        step_device(0)
        drain_peer()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            other = world - 1
            theirs = all_draws[first[other]: first[other] + 4]
            m2, o2 = synth.concat_sets(theirs)
            chk = {k: torch.empty((4, G), dtype=torch.float64, device=dev) for k in TABLES}
            eng.run_device(torch.from_numpy(m2).to(dev), o2, chk)
            torch.cuda.synchronize()
            same = all(bool(torch.equal(chk[k].view(torch.int64),
                                        final[k][first[other]: first[other] + 4].view(torch.int64))) for k in TABLES)
            line["multi_gpu_check"] = {"rows_of_rank": other, "sets": 4, "bit_identical": same}
        dist.barrier()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        if peer is not None:
            for pt in peer:
                pt.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
