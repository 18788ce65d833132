"""Golden vectors for row f4 (significance scoring, ``pySpade global`` / ``local``).

Run in the dev container (needs scipy, the dependency that holds the arithmetic, and
``tests/golden/cli_case.npz``, which the UNMODIFIED reference wrote):

    python tests/golden/make_score_golden.py

``global_analysis`` / ``local_analysis`` cannot be executed as a whole here (they need the
genome annotation, HDF5 inputs through PyTables and the hit-calling flow around them), so the
lines that do the scoring are executed as they stand in the reference on the reference's own
DEobs / DErand outputs:
    pval[np.argwhere(pval == 0)] = 1; pval[np.isinf(pval)] = 0; pval[np.isnan(pval)] = 0
                                                                  (global.py:118-123)
    scipy.stats.gamma.logsf(-pval[i], A[i], B[i], C[i])           (global.py:153-159, local.py:153-156)
    rand = sp_sparse.vstack([loadmat(...)['matrix']])
    np.sum(np.asarray(rand.tocsr()[:, idx].todense()) < pval[idx], axis=0) / iter_num
                                                                  (global.py:177-181)
with idx = every gene.  Plus a grid of scipy.stats.gamma.logsf values over all regimes.
"""
import os

import numpy as np
import scipy.sparse as sp_sparse
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_lines(pval, rand_dense, rand_stored, A, B, C):
    pval = np.array(pval, dtype=np.float64, copy=True)
    pval[np.argwhere(pval == 0)] = 1
    pval[np.isinf(pval)] = 0
    pval[np.isnan(pval)] = 0
    idx = list(range(pval.shape[0]))
    with np.errstate(all="ignore"):
        score = np.array([scipy.stats.gamma.logsf(-pval[i], A[i], B[i], C[i]) for i in idx])
    # the table as the reference holds it after loadmat: a sparse matrix of the stored entries
    rows, cols = np.nonzero(rand_stored)
    rand_matrix = sp_sparse.vstack([sp_sparse.csc_matrix((rand_dense[rows, cols], (rows, cols)),
                                                         shape=rand_dense.shape)])
    iter_num, gene_num = rand_matrix.shape
    with np.errstate(invalid="ignore"):
        emp = np.sum(np.asarray(rand_matrix.tocsr()[:, idx].todense()) < pval[idx], axis=0) / iter_num
    return pval, score, np.asarray(emp).ravel()


def main():
    z = np.load(os.path.join(HERE, "cli_case.npz"))
    out = {}
    for bg in ("complement", "NonTarget"):
        regions = sorted({k.split("mat:")[1].rsplit("-up_log-pval", 1)[0] for k in z.files
                          if k.startswith(f"DEobs/{bg}/mat:") and k.endswith("-up_log-pval")})
        out[f"{bg}/regions"] = np.array(regions)
        for mode in ("equal", "sgrna"):
            for d, D in (("up", "Up"), ("down", "Down")):
                rand = z[f"DErand/{bg}/{mode}/matsp:20-{d}_log-pval:dense"]
                stored = z[f"DErand/{bg}/{mode}/matsp:20-{d}_log-pval:stored"]
                A, B, C = (z[f"DErand/{bg}/{mode}/npy:{D}_dist_gamma-20-{x}.npy"] for x in "ABC")
                ps, scores, emps = [], [], []
                for reg in regions:
                    p, s, e = reference_lines(z[f"DEobs/{bg}/mat:{reg}-{d}_log-pval"][0], rand, stored, A, B, C)
                    ps.append(p); scores.append(s); emps.append(e)
                out[f"{bg}/{mode}/{d}/p"] = np.stack(ps)
                out[f"{bg}/{mode}/{d}/score"] = np.stack(scores)
                out[f"{bg}/{mode}/{d}/emp"] = np.stack(emps)
    # gamma.logsf grid: shapes 0.03 .. 3000, y = (x - loc) / scale from deep inside the body to the
    # far tail, both sides of loc, invalid parameters
    rng = np.random.default_rng(0)
    n = 6000
    a = np.exp(rng.uniform(np.log(0.03), np.log(3000.0), n))
    scale = np.exp(rng.uniform(np.log(1e-3), np.log(1e3), n))
    loc = rng.normal(0, 3, n)
    rel = np.exp(rng.uniform(np.log(1e-9), np.log(40.0), n))            # y / max(a, 1): body and tail
    y = rel * np.maximum(a, 1.0)
    x = loc + y * scale
    x[::50] = loc[::50] - 1.0                                           # below the support
    x[1::97] = loc[1::97]                                               # at its lower end
    a[2::211] = 0.0; a[3::223] = -1.0; a[5::227] = np.nan               # invalid shapes
    scale[7::229] = 0.0; scale[11::233] = -2.0
    x[13::239] = np.inf; x[17::241] = np.nan
    # the parameter patterns DErand writes: (0, 0, 0) skipped gene, (nan, 0, 1) constant gene
    a[19::251], loc[19::251], scale[19::251] = 0.0, 0.0, 0.0
    a[23::257], loc[23::257], scale[23::257] = np.nan, 0.0, 1.0
    with np.errstate(all="ignore"):
        out["grid/logsf"] = scipy.stats.gamma.logsf(x, a, loc, scale)
    out["grid/x"], out["grid/a"], out["grid/loc"], out["grid/scale"] = x, a, loc, scale
    np.savez_compressed(os.path.join(HERE, "score_case.npz"), **out)
    print("wrote score_case.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
