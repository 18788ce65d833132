#!/usr/bin/env python
"""Benchmark of the hypergeometric DE hot path (BASELINE.json metric: gene x draw
hypergeometric tests per second at 1/2/4/8 B200; DErand 1k-draw wall time).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--configs all|none]

Headline workload (``config.workload``): BASELINE.json ``configs[2]``, size bin N = 500 - DErand,
1 000 random draws of 500 cells per GPU and step, 50 000 cells x 33 000 genes, complement
background, synthetic Poisson counts, draws from the reference's RNG call
sequence (np.random.seed(0); np.random.choice(arange(C), 500, False)).  One "step" = one pass of
the hot path over that batch: 33 M (gene, draw) tests per GPU.

``value``   tests/s with member lists and result tables resident in HBM (CUDA events on the
            launch stream, barrier + synchronize on both sides, max over ranks).  N > 1: every
            rank tests its own 1 000 draws (weak scaling) and the timed region includes the
            exchange that re-shards the tables by gene (``shard.GeneExchange``: raw words stored
            into the owning GPU from inside the test kernel over NVLink, finished there) - at the
            end every GPU holds ALL draws x ITS genes, the layout the per-gene consumers need.
``e2e``     the same metric through the public API with HOST buffers: member index lists in host
            memory, tables delivered to page-locked host memory (H2D + kernels + exchange + D2H
            inside the timed region; N > 1: every rank delivers its gene block over its own PCIe
            link into host tables shared by the node).
``roofline`` the test itself - at N = 1 two launches per step, k_de_fused (member gather into
            shared-memory accumulators, one 8-byte accumulator word per test out) and k_expand (words
            -> tails + mean CPM tables): algorithmic bytes of SURVEY section 8-d (rho * min(8 N, C/8) +
            24 B per test) / their CUDA-event time (live, over the timed region), against the measured
            HBM copy bandwidth of MEASURED_PEAKS.json; ``kernels`` lists each launch with its own time,
            bytes and fraction, ``traffic`` the DRAM bytes ncu measured for the step.
``cpu_baseline`` the CPU implementation of the path on a bounded sample of the same workload, all
            host cores: the UNMODIFIED reference (``oracle/_ref``, kind "reference") when
            ``build()`` could stage it next to the oracle, else the oracle's per-gene port
            (kind "port"); the same sample is the parity check of the GPU tables (``parity``).
``configs`` every other BASELINE.json configuration, each with tests/s, the recomputed roofline
            fraction and a sampled-oracle parity verdict: cfg1 (reference-runnable), cfg2 DEobs 500
            ragged regions, cfg3 bins N = 100 / 2000, cfg4 200 000 cells (DErand N = 1 000 and DEobs
            6 000 regions), cfg5 set-size sweep 50..5 000 x 10 000 draws.
``derand_wall`` BASELINE's second metric: wall seconds of one DErand bin (staging, the reference's
            host RNG, tests, the three sparse MAT tables + cell-ID listing, gamma fit) through the
            product's own ``DErand.finish_draws``.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

G_GENES, C_CELLS, N_MEMBERS, S_DRAWS = 33000, 50000, 500, 1000
TABLES = ("up", "down", "cpm")               # DErand keeps these three (DErand.py:219-221)
ALL_TABLES = ("up", "down", "fc", "cpm")     # DEobs keeps all four (DEobs.py:208-221)
SAMPLE_GENES, SAMPLE_SETS = 4096, 24         # headline parity / cpu_baseline sample
CFG_SAMPLE_GENES, CFG_SAMPLE_SETS = 1024, 8  # per-config parity sample
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
    """per test rho * min(8 N, C/8) + B_out, summed over the sets."""
    gather = float(sum(rho * min(8.0 * n, C / 8.0) for n in sizes)) * G
    return gather, gather + float(len(sizes)) * G * b_out


def make_draws(total_sets, C, N):
    """The reference's draw sequence (DErand.py:27,197), generated once on the host so the
    stream does not depend on the number of ranks."""
    from pyspade_b200 import synth
    return synth.derand_draws(C, N, total_sets, seed=0)


def rel_err(got, want):
    ok = (np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
          and np.array_equal(got == 0, want == 0))
    m = np.isfinite(want) & (want != 0)
    err = float(np.max(np.abs(got[m] - want[m]) / np.abs(want[m]))) if m.any() else 0.0
    return ok, err


# --------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference when build() staged it under oracle/_ref, else the port
# --------------------------------------------------------------------------------------------
def reference_perform_de():
    """``(perform_DE, perform_DE_NC, kind)``: the reference's own functions (utils.py:296-341)
    imported from ``oracle/_ref`` (a verbatim, git-ignored copy ``build()`` makes where
    ``/root/reference`` exists; it travels with the snapshot), else the oracle's per-gene port."""
    ref_root = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isfile(os.path.join(ref_root, "pySpade", "utils.py")):
        try:
            os.environ["PYSPADE_REFERENCE_ROOT"] = ref_root
            from oracle import ref_loader
            ref_loader.REFERENCE_ROOT = ref_root
            utils, _, _ = ref_loader.load()
            return utils.perform_DE, utils.perform_DE_NC, "reference"
        except Exception as exc:                                       # noqa: BLE001
            print(f"[bench] oracle/_ref not importable ({exc}); using the port", file=sys.stderr)
    from oracle import de_port
    return de_port.perform_DE_port, de_port.perform_DE_NC_port, "port"


def cpu_leg(X, sets, nproc, perform_de):
    """Time the CPU implementation (reference structure: one process pool per cell set)."""
    Gs = X.shape[0]
    res = {k: np.zeros((len(sets), Gs)) for k in ALL_TABLES}
    Xc = X.tocoo()                                       # what cpm_normalization hands over (utils.py:155)
    t0 = time.perf_counter()
    for s, members in enumerate(sets):
        N, up, down, fc, cpm = perform_de(list(members), Xc, np.arange(Gs), nproc, np.zeros(Gs), np.zeros(Gs),
                                          np.ones(Gs), np.zeros(Gs))
        res["up"][s], res["down"][s], res["fc"][s], res["cpm"][s] = up, down, fc, cpm
    return time.perf_counter() - t0, res


def run_reference_arm(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores, bounded sample
    of the headline workload per step."""
    if rank != 0:
        return
    import torch
    nproc = os.cpu_count() or 1
    from pyspade_b200 import synth
    perform_de, _, kind = reference_perform_de()
    n_genes = 1024 if kind == "reference" else SAMPLE_GENES          # the reference is ~8x slower per test
    if torch.cuda.is_available():
        indptr, indices, counts = synth.counts_torch(G_GENES, C_CELLS, seed=0, device="cuda:0")
        genes, X = synth.sample_rows(indptr, indices, counts, G_GENES, C_CELLS, n_genes)
    else:
        import scipy.sparse as sp
        from oracle import de_port
        X = sp.csr_matrix(de_port.cpm_normalization_port(synth.counts_numpy(n_genes, C_CELLS, seed=0)))
    draws = make_draws(SAMPLE_SETS, C_CELLS, N_MEMBERS)
    per_step_sets = 2
    times = []
    for step in range(args.warmup + args.steps):
        sets = [draws[(step * per_step_sets + j) % len(draws)] for j in range(per_step_sets)]
        dt, _ = cpu_leg(X, sets, nproc, perform_de)
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
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# per-configuration measurement (one GPU)
# --------------------------------------------------------------------------------------------
class Staged:
    """A synthetic matrix on the device, its engine and a parity sample of its rows."""

    def __init__(self, G, C, dev, sample_genes):
        import torch
        from pyspade_b200 import synth
        from pyspade_b200.engine import Engine
        t0 = time.time()
        self.G, self.C, self.dev = G, C, dev
        self.csr = synth.counts_torch(G, C, seed=0, device=dev)
        torch.cuda.synchronize()
        self.nnz = int(self.csr[1].numel())
        self.rho = self.nnz / (G * C)
        self.synth_s = time.time() - t0
        t0 = time.time()
        self.eng = Engine.from_device_csr(G, C, *self.csr, device=dev.index)
        self.stage_s = time.time() - t0
        self.genes, self.X = synth.sample_rows(*self.csr, G, C, sample_genes)
        self.nc = np.random.default_rng(2).choice(C, C // 50, replace=False)     # 2 % of the cells

    def close(self):
        self.eng.close()
        self.csr = None


def measure_config(st, name, workload, sets, tables, reps, peak, background=None, parity_sets=CFG_SAMPLE_SETS,
                   l2_note="inputs larger than L2"):
    """Time ``reps`` passes of one batch with device-resident inputs and tables; parity of a
    sample of its rows against the oracle's all-genes form."""
    import torch
    from oracle import c_oracle, de_port
    from pyspade_b200 import synth
    eng, G, C, dev = st.eng, st.G, st.C, st.dev
    if (eng.n_nc > 0) != (background is not None):
        eng.set_background(background)
    S = len(sets)
    members, offsets = synth.concat_sets(sets)
    dm = torch.from_numpy(members).to(dev)
    out = {k: torch.empty((S, G), dtype=torch.float64, device=dev) for k in tables}
    eng.run_device(dm, offsets, out)                                   # warm-up (tail table, allocations)
    eng.run_device(dm, offsets, out)
    torch.cuda.synchronize()
    eng.timing(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.run_device(dm, offsets, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kt = eng.timing(reset=True)
    sizes = [len(s) for s in sets]
    b_out = 8 * len(tables)
    _, path_bytes = algorithmic_bytes(st.rho, sizes, C, G, b_out)
    test_ms = kt["test_ms"] / reps
    # parity: sampled genes x the first / largest sets of the batch
    pick = sorted(set(list(range(min(parity_sets - 1, S))) + [int(np.argmax(sizes))]))
    sub = [sets[i] for i in pick]
    m2, o2 = synth.concat_sets(sub)
    kd = torch.empty((len(sub), G), dtype=torch.int32, device=dev)
    chk = {k: torch.empty((len(sub), G), dtype=torch.float64, device=dev) for k in ALL_TABLES}
    eng.run_device(torch.from_numpy(m2).to(dev), o2, chk, k=kd)
    eng.check()
    gsel = torch.from_numpy(st.genes).to(dev)
    ora = de_port.de_tables(st.X, sub, background, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    parity = {"tests_checked": int(st.X.shape[0] * len(sub))}
    worst, special = 0.0, True
    for t in ALL_TABLES:
        ok, err = rel_err(chk[t][:, gsel].cpu().numpy(), ora[t])
        parity[t + "_max_rel_err"] = err
        worst, special = max(worst, err), special and ok
        if t in out:      # rows of the timed batch = rows of the small run, bit for bit
            special = special and bool(torch.equal(out[t][pick][:, gsel].view(torch.int64),
                                                   chk[t][:, gsel].view(torch.int64)))
    med, n_dev = eng.gene_cutoffs()
    parity["integers_exact"] = bool(np.array_equal(kd[:, gsel].cpu().numpy(), ora["k"])
                                    and np.array_equal(n_dev[st.genes], ora["n"])
                                    and np.array_equal(med[st.genes], ora["median"]))
    parity["special_values_match"] = special
    parity["ok"] = bool(parity["integers_exact"] and special and worst <= 1e-6)
    del out, chk
    return {"name": name, "workload": workload, "cells": C, "genes": G, "sets": S,
            "set_size": sizes[0] if len(set(sizes)) == 1 else f"{min(sizes)}..{max(sizes)} (mean {np.mean(sizes):.0f})",
            "background": "complement" if background is None else f"NC ({len(background)} cells)",
            "tables": list(tables), "value": S * G / (ms * 1e-3), "unit": UNIT, "ms_per_pass": ms,
            "kernel_ms": test_ms, "finish_ms": kt["finish_ms"] / reps, "other_kernels_ms": kt["table_ms"] / reps,
            "bytes_per_test": path_bytes / (S * G),
            "roofline_frac": path_bytes / ((test_ms + kt["table_ms"] / reps) * 1e-3) / 1e9 / peak if test_ms > 0 else None,
            "l2": l2_note, "parity": parity}


def run_configs(dev, peak, st50):
    """BASELINE.json configs other than the headline bin, one GPU."""
    import torch
    from pyspade_b200 import synth
    res = []
    # cfg1: the reference-runnable case (parity-test size; everything fits L2)
    st1 = Staged(5000, 2000, dev, 1024)
    sets = synth.deobs_sets(2000, 10, lo=30, hi=200, uniform_int=True)
    res.append(measure_config(st1, "cfg1", "DEobs 2 000 cells x 5 000 genes, 10 regions (BASELINE configs[0])",
                              sets, ALL_TABLES, 20, peak, l2_note="fits L2 (parity-sized case)"))
    st1.close()
    # cfg2: DEobs, 500 ragged regions
    sets = synth.deobs_sets(C_CELLS, 500)
    res.append(measure_config(st50, "cfg2", "DEobs 50 000 cells x 33 000 genes, 500 regions of 50..2 000 cells "
                              "(BASELINE configs[1])", sets, ALL_TABLES, 5, peak))
    res.append(measure_config(st50, "cfg2-nc", "the same with the negative-control background (2 % of the cells)",
                              sets, ALL_TABLES, 5, peak, background=st50.nc))
    # cfg3: the other size bins of the DErand background
    for N in (100, 2000):
        res.append(measure_config(st50, f"cfg3-N{N}", f"DErand bin N={N}: 1 000 draws x 33 000 genes x 50 000 cells "
                                  "(BASELINE configs[2])", make_draws(S_DRAWS, C_CELLS, N), TABLES, 5, peak))
    # cfg5: set-size sweep x 10 000 draws
    # (draws from a numpy Generator: the legacy stream of the reference costs ~5 s of host time per
    # 10 000 draws and is not what this configuration measures)
    rng = np.random.default_rng(3)
    for N in (50, 100, 200, 500, 1000, 2000, 5000):
        sweep = [rng.choice(C_CELLS, N, replace=False) for _ in range(10000)]
        res.append(measure_config(st50, f"cfg5-N{N}", f"sweep: 10 000 draws of {N} cells (BASELINE configs[4])",
                                  sweep, TABLES, 2, peak, parity_sets=4))
        torch.cuda.empty_cache()
    # cfg4: Gasperini scale
    st200 = Staged(G_GENES, 200000, dev, CFG_SAMPLE_GENES)
    res.append(measure_config(st200, "cfg4-DErand", "DErand 1 000 draws of 1 000 cells, 200 000 cells x 33 000 genes "
                              "(BASELINE configs[3])", make_draws(S_DRAWS, 200000, 1000), TABLES, 5, peak))
    res.append(measure_config(st200, "cfg4-DEobs", "DEobs 6 000 regions of 50..2 000 cells, 200 000 cells x 33 000 genes "
                              "(BASELINE configs[3])", synth.deobs_sets(200000, 6000), ALL_TABLES, 2, peak))
    for r in res[-2:]:
        r["stage_seconds"] = st200.stage_s
    st200.close()
    torch.cuda.empty_cache()
    return res


def derand_wall(dev, C, N, S, fit_sample=64):
    """BASELINE's second metric through the product's own pipeline (``DErand.finish_draws``):
    staging (K0, K1, layout) from device-resident counts, overlapped with the reference's host
    RNG draws in a second thread; then tests, files and the gamma fit.  The scipy gamma fit
    (the CLI default) is timed on a gene sample and extrapolated to the host cores."""
    import torch
    from pyspade_b200 import DErand, synth
    from pyspade_b200.engine import Engine
    indptr, indices, counts = synth.counts_torch(G_GENES, C, seed=0, device=dev)
    torch.cuda.synchronize()
    res = {"cells": C, "genes": G_GENES, "set_size": N, "draws": S}
    box = {}

    def draw():
        t = time.perf_counter()
        box["draws"] = synth.derand_draws(C, N, S)
        box["rng_s"] = time.perf_counter() - t

    t0 = time.perf_counter()
    th = threading.Thread(target=draw)
    th.start()
    eng = Engine.from_device_csr(G_GENES, C, indptr, indices, counts, device=dev.index)
    res["stage_s"] = time.perf_counter() - t0
    th.join()
    res["draws_host_rng_s"] = box["rng_s"]
    res["stage_and_rng_overlapped_s"] = time.perf_counter() - t0
    timings = {}
    with tempfile.TemporaryDirectory() as tmp:
        DErand.finish_draws(eng, box["draws"], N, tmp + "/", gamma_fit="device", timings=timings)
        res["file_bytes"] = int(sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp)))
    res.update({k: float(v) for k, v in timings.items()})
    res["gamma_fit"] = "device (spade_gamma_fit)"
    res["wall_s"] = time.perf_counter() - t0
    res["wall_without_gamma_fit_s"] = res["wall_s"] - res["gamma_fit_s"]
    # the CLI default tail: scipy unchanged, on a gene sample
    out = {k: torch.empty((S, G_GENES), dtype=torch.float64, device=dev) for k in ("up", "down")}
    members, offsets = synth.concat_sets(box["draws"])
    eng.run_device(torch.from_numpy(members).to(dev), offsets, out)
    pick = torch.linspace(0, G_GENES - 1, fit_sample, device=dev).round().long()
    up, down = out["up"][:, pick].cpu().numpy(), out["down"][:, pick].cpu().numpy()
    t1 = time.perf_counter()
    DErand.gamma_parameters(up, "up", processes=1)
    DErand.gamma_parameters(down, "down", processes=1)
    dt = time.perf_counter() - t1
    cores = os.cpu_count() or 1
    res["gamma_fit_scipy_ms_per_gene_and_direction"] = 1e3 * dt / (2 * fit_sample)
    res["gamma_fit_scipy_estimate_all_cores_s"] = dt / (2 * fit_sample) * 2 * G_GENES / cores
    res["host_cores"] = cores
    eng.close()
    return res


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--configs", default="all", choices=["all", "none"],
                    help="also measure the other BASELINE configurations and the DErand wall time")
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

    from pyspade_b200 import shard, synth
    from pyspade_b200.engine import PinnedArray

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    G, C, N, S = G_GENES, C_CELLS, N_MEMBERS, S_DRAWS
    peak, peak_src = peaks()
    st = Staged(G, C, dev, SAMPLE_GENES)                              # replicated matrix
    eng, nnz, rho = st.eng, st.nnz, st.rho

    # draws: one global host stream, rank r takes draws r, r+P, ... of every batch
    n_batches = 2
    all_draws = make_draws(S * world * n_batches, C, N)
    batches = []
    for b in range(n_batches):
        mine = all_draws[b * S * world + rank: (b + 1) * S * world: world]
        members, offsets = synth.concat_sets(mine)
        batches.append((mine, members, offsets, torch.from_numpy(members).to(dev)))
    sizes_all = np.full(S * world, N, dtype=np.int64)

    out_dev = {k: torch.empty((S, G), dtype=torch.float64, device=dev) for k in TABLES}
    exchange = shard.GeneExchange(eng, S * world, tables=TABLES) if world > 1 else None
    if exchange is not None:
        gather_mode = ("tables re-sharded by gene: raw accumulator words stored into the owning GPU from inside "
                       "the test kernel (CUDA IPC over NVLink), finished there by k_expand")
    else:
        gather_mode = "none (one GPU)"

    def step_device(i):
        _, _, offsets, dmem = batches[i % n_batches]
        if exchange is not None:
            exchange.step(dmem, offsets, rank, world, sizes_all)     # draw i of rank r = row i * P + r
        else:
            eng.run_device(dmem, offsets, out_dev)

    # host tables of the end-to-end leg.  N > 1: tables shared by the node (/dev/shm, page-locked
    # in every rank): each GPU delivers ITS GENE BLOCK of all draws over its own PCIe link
    shared, pinned = None, None
    if world > 1:
        shared = shard.SharedHostTables(TABLES, S * world, G, rank, world, device=local_rank,
                                        touch_block=(exchange.gene_begin, exchange.gene_end))
        e2e_mode = ("Engine members from host memory, gene-sharded exchange, every rank copies its gene block of "
                    "all draws into host tables shared by the node (POSIX shm, page-locked)")
    else:
        pinned = {k: PinnedArray((S, G), np.float64) for k in TABLES}
        host_out = {k: pinned[k].array for k in TABLES}
        e2e_mode = "Engine.run: host member lists -> page-locked host tables"
    lib = eng._lib

    def step_e2e(i):
        mine, members, offsets, _ = batches[i % n_batches]
        if world == 1:
            eng.run(members=members, offsets=offsets, out=host_out, validate=False, tables=TABLES)
            return
        import ctypes
        exchange.step(members, offsets, rank, world, sizes_all)       # host member lists: uploaded by the call
        exchange.drain()                                               # this step's gene block is finished
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        exchange.side.wait_event(ready)
        g0, g1 = exchange.gene_begin, exchange.gene_end
        if g1 > g0:
            for k in TABLES:
                dst = shared.arrays[k]
                rc = lib.spade_memcpy2d(ctypes.c_void_p(dst.ctypes.data + 8 * g0), 8 * G,
                                        ctypes.c_void_p(exchange.tables[k].data_ptr()), 8 * exchange.cols,
                                        8 * (g1 - g0), S * world, 1, ctypes.c_void_p(exchange.side.cuda_stream))
                assert rc == 0
        exchange.side.synchronize()
        dist.all_reduce(exchange.flag)                                 # every rank's block is in the shared tables
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps, warmup, ex=exchange):
        def drain():
            if ex is not None:
                ex.drain()
        for i in range(warmup):
            fn(i)
        drain()
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
        drain()                             # N > 1: the last expansion belongs to the timed region
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
    # what the link gives: the same tables as three plain device -> host copies, nothing else
    link_gbs = None
    if world == 1:
        import ctypes

        def plain_copies(i):
            for k in TABLES:
                lib.spade_memcpy2d(ctypes.c_void_p(host_out[k].ctypes.data), 8 * G, ctypes.c_void_p(out_dev[k].data_ptr()),
                                   8 * G, 8 * G, S, 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        copy_ms, _, _, _ = timed(plain_copies, 3, 1)
        link_gbs = len(TABLES) * S * G * 8 * 3 / (copy_ms * 1e-3) / 1e9
    clocks = sampler.stop([win_a, win_b]) if rank == 0 else None

    tests_per_step = S * G * world
    value = tests_per_step * args.steps / (dev_ms * 1e-3)
    e2e_value = tests_per_step * args.steps / (e2e_ms * 1e-3)

    # roofline of the dominant kernel (per rank; rank 0's kernel times).  N > 1: the kernel
    # stores 8-byte raw words instead of the three float64 tables
    b_out = 8 * len(TABLES) if world == 1 else 8
    gather_bytes, path_bytes = algorithmic_bytes(rho, [N] * S, C, G, b_out)
    test_ms = kt["test_ms"] / args.steps
    table_ms = kt["table_ms"] / args.steps
    finish_ms = kt["finish_ms"] / args.steps                       # N = 1: the k_expand launch of the two-phase form
    gather_ms = test_ms - finish_ms
    achieved = path_bytes / (test_ms * 1e-3) / 1e9
    tests = float(S) * G
    kernels = [{"name": "k_de_fused", "ms": gather_ms, "algorithmic_bytes": gather_bytes + 8.0 * tests,
                "role": "member gather into shared-memory accumulators, 8-byte accumulator word per test out"}]
    kernels[0]["frac"] = kernels[0]["algorithmic_bytes"] / (gather_ms * 1e-3) / 1e9 / peak
    if finish_ms > 0:
        kernels.append({"name": "k_expand", "ms": finish_ms, "algorithmic_bytes": (8.0 + b_out) * tests,
                        "role": "accumulator words -> (tails, mean CPM) tables, gene-major",
                        "frac": (8.0 + b_out) * tests / (finish_ms * 1e-3) / 1e9 / peak})
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("step_dram_bytes")        # both launches of the two-phase form
    roofline = {"bound": "hbm", "kernel": "k_de_fused" + (" + k_expand (two-phase form of the test)" if finish_ms > 0 else ""),
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic if world == 1 else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": path_bytes, "launch_ms": test_ms,
                "launches_per_step": len(kernels), "kernels": kernels,
                "bytes_per_test": path_bytes / (S * G), "gather_bytes_per_launch": gather_bytes,
                "sets_per_launch": S, "other_kernels_ms_per_step": table_ms,
                "note": "algorithmic bytes of SURVEY 8-d against the measured HBM copy bandwidth; the cell-major layout "
                        "(1.1 GB) is re-read by every set and served from L2 (measured DRAM traffic = `traffic`, about a "
                        "quarter of the algorithmic bytes), so the fraction of k_de_fused can exceed 1: its binding unit is "
                        "the SM's LSU data pipe (l1tex__throughput 88 % under ncu, profiles/r02_ncu_full_two_phase.json)"}

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
                           "gene_blocks": exchange.blocks if exchange is not None else [[0, G]],
                           "l2": "inputs larger than L2: 1.0 GB cell-major matrix read + 0.8 GB of tables written per step",
                           "stage_seconds": st.stage_s},
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps, "path": e2e_mode,
                        "h2d_bytes_per_step": int(world * (4 * S * N + 8 * (S + 1))),
                        "d2h_bytes_per_step": int(world * S * G * 8 * len(TABLES)),
                        "d2h_gbs": world * S * G * 8 * len(TABLES) / (e2e_ms / args.steps * 1e-3) / 1e9,
                        "d2h_gbs_plain_copy_of_the_tables": link_gbs},
                "gpu_launches": int(kt["launches"]),
                "roofline": roofline, "clocks": clocks}

    # ---------------- N = 1: CPU baseline + parity on a bounded sample, other configs, wall time
    if world == 1 and not args.no_cpu:
        mine, members, offsets, dmem = batches[0]
        kdev = torch.empty((S, G), dtype=torch.int32, device=dev)
        full = {k: torch.empty((S, G), dtype=torch.float64, device=dev) for k in ALL_TABLES}
        eng.run_device(dmem, offsets, full, k=kdev)
        eng.check()
        perform_de, _, kind = reference_perform_de()
        nproc = os.cpu_count() or 1
        genes, X = st.genes, st.X
        if kind == "reference":            # ~8x slower per test than the port: a smaller sample
            keep = np.arange(0, X.shape[0], 4)
            genes, X, sets = genes[keep], X[keep], mine[:12]
        else:
            sets = mine[:SAMPLE_SETS]
        dt, cpu = cpu_leg(X, sets, nproc, perform_de)
        gsel = torch.from_numpy(genes).to(dev)
        parity = {}
        for name in ALL_TABLES:
            got = full[name][:len(sets)][:, gsel].cpu().numpy()
            ok, err = rel_err(got, cpu[name])
            parity[name] = {"special_values_match": ok, "max_rel_err": err}
        from oracle import de_port
        ora = de_port.de_tables(X, sets)
        med, n_dev = eng.gene_cutoffs()
        parity["k_exact"] = bool(np.array_equal(kdev[:len(sets)][:, gsel].cpu().numpy(), ora["k"]))
        parity["n_exact"] = bool(np.array_equal(n_dev[genes], ora["n"]))
        parity["median_exact"] = bool(np.array_equal(med[genes], ora["median"]))
        parity["tests_checked"] = int(X.shape[0] * len(sets))
        parity["against"] = ("the unmodified reference's perform_DE (oracle/_ref)" if kind == "reference"
                             else "the oracle's per-gene port")
        line["cpu_baseline"] = {"value": X.shape[0] * len(sets) / dt, "unit": UNIT, "cores": nproc, "kind": kind,
                                "sample": f"{X.shape[0]} density-stratified genes x {len(sets)} draws of "
                                          f"{N} cells at full C={C} ({dt:.1f} s)"}
        line["parity"] = parity
        del full, kdev
    if world == 1 and args.configs == "all":
        del out_dev
        torch.cuda.empty_cache()
        line["configs"] = run_configs(dev, peak, st)
        st.close()
        torch.cuda.empty_cache()
        line["derand_wall"] = [derand_wall(dev, C_CELLS, N_MEMBERS, S_DRAWS), derand_wall(dev, 200000, 1000, S_DRAWS)]

    # ---------------- N > 1: bit-identity against a one-GPU recomputation, strong scaling, DEobs sharded
    if world > 1:
        step_device(0)
        exchange.drain()
        torch.cuda.synchronize()
        g0, g1 = exchange.gene_begin, exchange.gene_end
        theirs = all_draws[:2 * world]                                   # rows 0 .. 2P-1 of batch 0: two draws of every rank
        m2, o2 = synth.concat_sets(theirs)
        chk = {k: torch.empty((len(theirs), G), dtype=torch.float64, device=dev) for k in TABLES}
        eng.run_device(torch.from_numpy(m2).to(dev), o2, chk)
        torch.cuda.synchronize()
        same = all(bool(torch.equal(chk[k][:, g0:g1].contiguous().view(torch.int64),
                                    exchange.tables[k][:len(theirs), :g1 - g0].contiguous().view(torch.int64)))
                   for k in TABLES)
        flag = torch.tensor([1.0 if same else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            line["multi_gpu_check"] = {"rows": len(theirs), "gene_block_of_every_rank": True,
                                       "bit_identical": bool(float(flag) > 0)}
        if args.configs == "all":
            extra = []
            # strong scaling: BASELINE's "1 000 draws sharded over N GPUs"
            draws = all_draws[:S]
            mine = draws[rank::world]
            members, offsets = synth.concat_sets(mine)
            dm = torch.from_numpy(members).to(dev)
            sz = np.full(S, N, dtype=np.int64)

            def strong(i):
                exchange.step(dm, offsets, rank, world, sz)
            ms, _, kts, _ = timed(strong, args.steps, args.warmup)
            extra.append({"name": "strong-cfg3-N500", "workload": f"{S} draws of {N} cells in total, sharded over {world} GPUs",
                          "scaling": "strong", "value": S * G * args.steps / (ms * 1e-3), "unit": UNIT,
                          "ms_per_pass": ms / args.steps, "kernel_ms": kts["test_ms"] / args.steps})
            # DEobs: 500 ragged regions, balanced over the ranks by member count
            regions = synth.deobs_sets(C, 500)
            rsz = np.array([len(r) for r in regions], dtype=np.int64)
            assignment = shard.plan(rsz, world, balance=True)
            order = np.concatenate(assignment)
            members, offsets = synth.concat_sets([regions[i] for i in assignment[rank]])
            dm2 = torch.from_numpy(members).to(dev)
            first = int(sum(len(a) for a in assignment[:rank]))
            ex4 = shard.GeneExchange(eng, 500, tables=ALL_TABLES)

            def deobs(i):
                ex4.step(dm2, offsets, first, 1, rsz[order])
            ms, _, kts, _ = timed(deobs, args.steps, args.warmup, ex=ex4)
            extra.append({"name": "cfg2-sharded", "workload": f"DEobs 500 regions of 50..2 000 cells, balanced over {world} GPUs",
                          "scaling": "strong", "value": 500 * G * args.steps / (ms * 1e-3), "unit": UNIT,
                          "ms_per_pass": ms / args.steps, "kernel_ms": kts["test_ms"] / args.steps,
                          "member_cells_per_rank": [int(rsz[a].sum()) for a in assignment]})
            ex4.close()
            if rank == 0:
                line["configs"] = extra
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        exchange.close()
        shared.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
