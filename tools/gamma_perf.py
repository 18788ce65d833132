"""Timing of spade_gamma_fit (row f1) on a DErand-shaped table: 1 000 draws x 33 000 genes of
-log p values produced by the engine itself (synthetic bench matrix, N = 500)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyspade_b200 import synth                                  # noqa: E402
from pyspade_b200.engine import Engine, gamma_fit               # noqa: E402


def main():
    G, C, N, S = 33000, 50000, 500, 1000
    dev = torch.device("cuda:0")
    indptr, indices, counts = synth.counts_torch(G, C, seed=0, device=dev)
    eng = Engine.from_device_csr(G, C, indptr, indices, counts)
    members, offsets = synth.concat_sets(synth.derand_draws(C, N, S))
    out = {k: torch.empty((S, G), dtype=torch.float64, device=dev) for k in ("up", "down", "cpm")}
    eng.run_device(torch.from_numpy(members).to(dev), offsets, out)
    torch.cuda.synchronize()
    for name in ("up", "down"):
        gamma_fit(out[name][:, :256].contiguous())                # warm-up
        t0 = time.perf_counter()
        A, B, C_, st = gamma_fit(out[name])
        dt = time.perf_counter() - t0
        print(f"{name}: {dt*1e3:.1f} ms for {G} genes x {S} draws; fitted {int((st == 0).sum())}, "
              f"skipped {int((st == 1).sum())}, constant {int((st == 2).sum())}, failed {int((st == 3).sum())}; "
              f"median shape {np.nanmedian(A[st == 0]):.3f}")


if __name__ == "__main__":
    main()
