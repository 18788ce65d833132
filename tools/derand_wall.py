"""DErand 1 000-draw wall time at full scale on one GPU (the second half of BASELINE.json's
metric): staging + the reference's RNG draws + the batched test into host tables + the
final files, and - reported separately, as SURVEY.md section 8-d asks - the unchanged
scipy gamma-fit tail timed on a gene sample and extrapolated.

    python tools/derand_wall.py [--C 50000] [--G 33000] [--N 500] [--S 1000] [--out DIR]

Inputs are the synthetic counts of section 8-d generated on the device (the reference's
DataFrame pickle of this size would be 6.6 GB; loading it is pandas time, not path time).
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyspade_b200 import DErand, synth                     # noqa: E402
from pyspade_b200 import io as spio                        # noqa: E402
from pyspade_b200.engine import Engine, PinnedArray        # noqa: E402

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=50000)
    ap.add_argument("--G", type=int, default=33000)
    ap.add_argument("--N", type=int, default=500)
    ap.add_argument("--S", type=int, default=1000)
    ap.add_argument("--fit-genes", type=int, default=96)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()

    dev = torch.device("cuda:0")
    indptr, indices, counts = synth.counts_torch(a.G, a.C, seed=0, device=dev)
    torch.cuda.synchronize()
    res = {"cells": a.C, "genes": a.G, "set_size": a.N, "draws": a.S, "nnz": int(indices.numel())}

    t0 = time.perf_counter()
    eng = Engine.from_device_csr(a.G, a.C, indptr, indices, counts, device=0)
    res["stage_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    np.random.seed(0)
    draws = [np.random.choice(np.arange(a.C), size=a.N, replace=False) for _ in range(a.S)]
    res["draws_host_rng_s"] = time.perf_counter() - t0

    tables = ("up", "down", "cpm")
    pinned = {k: PinnedArray((a.S, a.G), np.float64) for k in tables}
    out = {k: pinned[k].array for k in tables}
    eng.run(draws[:8], tables=tables)                          # warm-up (module load, allocations)
    t0 = time.perf_counter()
    got = eng.run(draws, out=out, tables=tables)
    res["tests_s"] = time.perf_counter() - t0
    res["tests_per_s"] = a.S * a.G / res["tests_s"]

    with tempfile.TemporaryDirectory() as tmp:
        od = a.out or tmp
        t0 = time.perf_counter()
        spio.write_cell_ids(os.path.join(od, f"{a.N}-rand_cell_ID.txt"), draws)
        spio.save_draw_tables([(os.path.join(od, f"{a.N}-up_log-pval"), got["up"]),
                               (os.path.join(od, f"{a.N}-down_log-pval"), got["down"]),
                               (os.path.join(od, f"{a.N}-cpm"), got["cpm"])])
        res["write_files_s"] = time.perf_counter() - t0
    res["wall_without_gamma_fit_s"] = (res["stage_s"] + res["draws_host_rng_s"] + res["tests_s"]
                                       + res["write_files_s"])

    # gamma-fit tail (DErand.py:250-292, scipy unchanged): serial cost per fit on an evenly spaced
    # gene sample, then the estimate for 2 x G fits spread over the host cores by the worker pool
    pick = np.linspace(0, a.G - 1, a.fit_genes).round().astype(int)
    t0 = time.perf_counter()
    DErand.gamma_parameters(got["up"][:, pick], "up", processes=1)
    DErand.gamma_parameters(got["down"][:, pick], "down", processes=1)
    dt = time.perf_counter() - t0
    res["gamma_fit_ms_per_gene_and_direction"] = 1e3 * dt / (2 * pick.size)
    res["gamma_fit_sample_genes"] = int(pick.size)
    res["gamma_fit_cores"] = os.cpu_count()
    res["gamma_fit_estimate_all_cores_s"] = dt / (2 * pick.size) * 2 * a.G / max(1, os.cpu_count())
    # the same tail on the GPU (spade_gamma_fit), all genes, both directions
    t0 = time.perf_counter()
    DErand.gamma_parameters_device(got["up"], "up")
    DErand.gamma_parameters_device(got["down"], "down")
    res["gamma_fit_device_s"] = time.perf_counter() - t0
    print(json.dumps(res))


if __name__ == "__main__":          # the gamma-fit pool uses the spawn start method
    main()
