"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck), one tool per
gpurun call:  compute-sanitizer --tool racecheck python tools/sanitize_case.py
Covers staging (K0, K1, layout, bank order, dense heads), the fused test kernel with the tail
table and with direct tails, the split path, the raw form + expand, host and device outputs."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyspade_b200 import synth                      # noqa: E402
from pyspade_b200.engine import Engine              # noqa: E402


def main():
    G, C = 1100, 900
    counts = synth.counts_numpy(G, C, seed=3)
    draws = synth.derand_draws(C, 60, 40)
    ragged = synth.deobs_sets(C, 5, lo=3, hi=400)
    nc = np.arange(0, C, 7)
    with Engine.from_counts(counts) as eng:
        a = eng.run(draws, want_k=True)                              # tail table, host outputs
        b = eng.run(ragged, want_k=True)                             # direct tails
        eng.set_background(nc)
        c = eng.run(draws, want_k=True)
        eng.set_option("member_block", 32)
        d = eng.run(ragged, want_k=True)                             # split path
        eng.set_option("member_block", 32768)
        eng.set_background(None)
        members, offsets = synth.concat_sets(draws)
        dm = torch.from_numpy(members).cuda()
        raw = torch.zeros((len(draws), G), dtype=torch.int64, device="cuda")
        eng.run_raw(dm, offsets, raw)
        out = {k: torch.empty((len(draws), G), dtype=torch.float64, device="cuda") for k in ("up", "down", "fc", "cpm")}
        kd = torch.empty((len(draws), G), dtype=torch.int32, device="cuda")
        eng.expand(raw, [len(s) for s in draws], out, k=kd)
        torch.cuda.synchronize()
        for name in out:
            assert np.array_equal(out[name].cpu().numpy(), a[name], equal_nan=True), name
        assert np.array_equal(kd.cpu().numpy(), a["k"])
        eng.set_option("genes_per_part", 64)
        e = eng.run(draws, want_k=True)
        for name in a:
            assert np.array_equal(a[name], e[name], equal_nan=True), name
    print("sanitize case ok", float(np.nanmean(a["up"])), float(np.nanmean(b["cpm"])), float(np.nanmean(c["fc"])),
          float(np.nanmean(d["down"][np.isfinite(d["down"])])))


if __name__ == "__main__":
    main()
