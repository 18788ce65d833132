"""Scratch perf probe (not part of the bench contract): DErand-shaped batch on one GPU."""
import argparse
import time

import numpy as np
import torch

from pyspade_b200 import synth
from pyspade_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--G", type=int, default=33000)
ap.add_argument("--C", type=int, default=50000)
ap.add_argument("--N", type=int, default=500)
ap.add_argument("--S", type=int, default=200)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--sort", action="store_true")
ap.add_argument("--deobs", action="store_true", help="ragged region sizes 50..2000 instead of draws")
a = ap.parse_args()

dev = torch.device("cuda:0")
t0 = time.time()
indptr, indices, counts = synth.counts_torch(a.G, a.C, seed=0, device=dev)
torch.cuda.synchronize()
print(f"synth {time.time()-t0:.2f}s nnz={indices.numel()} density={indices.numel()/(a.G*a.C):.4f}")
t0 = time.time()
eng = Engine.from_device_csr(a.G, a.C, indptr, indices, counts, device=0)
print(f"stage {time.time()-t0:.2f}s genes_per_part={eng.genes_per_part}")
draws = synth.deobs_sets(a.C, a.S) if a.deobs else synth.derand_draws(a.C, a.N, a.S)
if a.sort:
    draws = [np.sort(d) for d in draws]
members, offsets = synth.concat_sets(draws)
dm = torch.from_numpy(members).to(dev)
out = {k: torch.empty((a.S, a.G), dtype=torch.float64, device=dev) for k in ("up", "down", "cpm", "fc")}
for rep in range(a.reps):
    eng.timing(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run_device(dm, offsets, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = eng.timing()
    print(f"rep{rep}: total {ms:.3f} ms  test {t['test_ms']:.3f} (finish {t['finish_ms']:.3f})  table {t['table_ms']:.2f}  "
          f"tests/s {a.S*a.G/ms*1e3:.3e}")
# host path
t0 = time.time()
res = eng.run(draws, validate=False)
print(f"host-output run {1e3*(time.time()-t0):.1f} ms")
print("nan frac", float(np.isnan(res['up']).mean()), "zero frac", float((res['up'] == 0).mean()))
