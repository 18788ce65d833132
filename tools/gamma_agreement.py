"""Row f1: how far the device gamma fit (``spade_gamma_fit``) is from scipy's on full-size DErand
tables, measured where it matters - on the significance score ``gamma.logsf(x, A, B, C)`` that
``pySpade global`` / ``local`` compute from the parameters (global.py:151-171).

    python tools/gamma_agreement.py [--C 50000] [--N 500] [--S 1000] [--genes 33000] > out.json

For every gene and both directions: scipy.stats.gamma.fit on the host (all cores) and the device
fit; then the score at the quantiles 0.5 / 0.9 / 0.99 / max of the gene's own background values
and at 2 x max (a clear hit), with both parameter sets.
"""
import argparse
import json
import multiprocessing
import os
import sys
import time

import numpy as np
import scipy.stats as stats

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def fit_chunk(neg):
    out = np.zeros((neg.shape[1], 3))
    for j in range(neg.shape[1]):
        col = neg[:, j]
        col = col[np.isfinite(col)]
        if col.shape[0] <= 50:
            continue
        if np.all(col == col[0]):
            out[j] = (np.nan, 0.0, 1.0)
            continue
        try:
            out[j] = stats.gamma.fit(col)
        except Exception:                                   # FitError
            out[j] = np.nan
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=50000)
    ap.add_argument("--N", type=int, default=500)
    ap.add_argument("--S", type=int, default=1000)
    ap.add_argument("--genes", type=int, default=33000)
    a = ap.parse_args()
    import torch
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine, gamma_fit
    dev = torch.device("cuda:0")
    G = 33000
    eng = Engine.from_device_csr(G, a.C, *synth.counts_torch(G, a.C, seed=0, device=dev))
    draws = synth.derand_draws(a.C, a.N, a.S)
    members, offsets = synth.concat_sets(draws)
    out = {k: torch.empty((a.S, G), dtype=torch.float64, device=dev) for k in ("up", "down")}
    eng.run_device(torch.from_numpy(members).to(dev), offsets, out)
    eng.check()
    pick = np.linspace(0, G - 1, a.genes).round().astype(int)
    res = {"cells": a.C, "set_size": a.N, "draws": a.S, "genes": int(pick.size), "cores": os.cpu_count()}
    for d in ("up", "down"):
        table = out[d][:, torch.from_numpy(pick).to(dev)].contiguous()
        t0 = time.perf_counter()
        A, B, C, st, evals = gamma_fit(table, with_evals=True)
        res[f"{d}_device_s"] = time.perf_counter() - t0
        host = table.cpu().numpy()
        neg = -host
        t0 = time.perf_counter()
        spans = [(g0, min(pick.size, g0 + 128)) for g0 in range(0, pick.size, 128)]
        with multiprocessing.get_context("fork").Pool(os.cpu_count()) as pool:
            parts = pool.map(fit_chunk, [neg[:, s:e] for s, e in spans])
        ref = np.vstack(parts)
        res[f"{d}_scipy_s"] = time.perf_counter() - t0
        fitted = (st == 0) & np.isfinite(ref).all(axis=1) & (ref[:, 2] > 0)
        res[f"{d}_fitted"] = int(fitted.sum())
        res[f"{d}_status_counts"] = {str(k): int((st == k).sum()) for k in range(4)}
        # same classification of the genes that are not fitted
        res[f"{d}_skipped_same"] = bool(np.array_equal((st == 1), (ref == 0).all(axis=1)))
        res[f"{d}_constant_same"] = bool(np.array_equal((st == 2), np.isnan(ref[:, 0]) & (ref[:, 1] == 0)))
        dev_p = np.stack([A, B, C], axis=1)
        rel = np.abs(dev_p[fitted] - ref[fitted]) / np.maximum(np.abs(ref[fitted]), 1e-12)
        for tol in (1e-6, 1e-4, 1e-3, 1e-2):
            res[f"{d}_params_within_{tol:g}"] = float(np.mean((rel <= tol).all(axis=1)))
        # the score both parameter sets give at the gene's own quantiles and beyond
        x = np.where(np.isfinite(neg), neg, np.nan)
        qs = {"q50": np.nanquantile(x, 0.5, axis=0), "q90": np.nanquantile(x, 0.9, axis=0),
              "q99": np.nanquantile(x, 0.99, axis=0), "max": np.nanmax(x, axis=0)}
        qs["2max"] = 2.0 * qs["max"]
        qs["4max"] = 4.0 * qs["max"]
        res[f"{d}_evals_hist"] = {str(b): int(((evals[fitted] >= b) & (evals[fitted] < b + 100)).sum())
                                  for b in range(0, 700, 100)}
        capped = evals[fitted] >= 600
        res[f"{d}_capped_frac"] = float(capped.mean())
        bad_params = ~(rel <= 1e-3).all(axis=1)
        res[f"{d}_params_disagree_frac"] = float(bad_params.mean())
        res[f"{d}_params_disagree_that_are_capped"] = float((bad_params & capped).sum() / max(1, bad_params.sum()))
        for thr in (300, 400, 500):
            flag = evals[fitted] >= thr
            res[f"{d}_evals_ge_{thr}"] = {"flagged_frac": float(flag.mean()),
                                          "covers_disagree": float((bad_params & flag).sum() / max(1, bad_params.sum()))}
        for name, q in qs.items():
            with np.errstate(all="ignore"):
                s_dev = stats.gamma.logsf(q[fitted], A[fitted], B[fitted], C[fitted])
                s_ref = stats.gamma.logsf(q[fitted], ref[fitted, 0], ref[fitted, 1], ref[fitted, 2])
            diff = np.abs(s_dev - s_ref)
            ok = np.isfinite(diff)
            reld = diff[ok] / np.maximum(np.abs(s_ref[ok]), 1e-3)
            bad_score = np.zeros(fitted.sum(), dtype=bool)
            bad_score[ok] = reld > 1e-3
            res[f"{d}_score_{name}_disagree_that_are_capped"] = float((bad_score & capped).sum() / max(1, bad_score.sum()))
            res[f"{d}_score_{name}_uncapped_frac_rel_le_1e-3"] = float(np.mean(~bad_score[~capped])) if (~capped).any() else None
            res[f"{d}_score_{name}"] = {
                "abs_median": float(np.median(diff[ok])), "abs_p99": float(np.quantile(diff[ok], 0.99)),
                "abs_max": float(diff[ok].max()),
                "frac_abs_le_1e-3": float(np.mean(diff[ok] <= 1e-3)), "frac_rel_le_1e-3": float(np.mean(reld <= 1e-3)),
                "frac_rel_le_1e-2": float(np.mean(reld <= 1e-2)), "frac_rel_le_1e-1": float(np.mean(reld <= 1e-1)),
                "both_finite": float(np.mean(ok))}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
