"""Restatement of the pySpade hot path (``/root/reference/pySpade/utils.py``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Two forms of the same algorithm:

* per-gene form (``gene_test`` / ``gene_test_nc`` fanned out by ``perform_DE_port`` /
  ``perform_DE_NC_port``): follows the reference step by step and keeps its cost
  structure (one Python call per gene and cell-set, scipy per test, a process
  pool over genes).  This is what ``bench.py`` times as the CPU baseline
  (``cpu_baseline.kind == "port"``) because the reference tree cannot travel to
  the GPU box.
* all-genes form (``de_tables``): integer tuple ``(k, M, n, N)`` for every gene and
  set with sparse algebra, p-values evaluated once per unique tuple.  Used to
  check the CUDA path at sizes where the per-gene form would take days.

Reference lines followed:
  utils.py:150-155  cpm_normalization      value = fl(fl(x*1e6) * fl(1/colsum))
  utils.py:194-215  hypergeo_test_sparse   median over all C cells (zeros included),
                                           low set {x <= median}, k = |low ∩ set|,
                                           fc = (mean_set+.01)/(mean_rest+.01),
                                           nan for both tails if any of k,M,n,N is 0,
                                           return order (down, up, fc, cpm)
  utils.py:248-266  hypergeo_test_NC_sparse median over NC cells only, low set over
                                           ALL cells, fc denominator = NC mean
  utils.py:296-341  perform_DE / perform_DE_NC  argument order (.., down, up, fc, cpm),
                                           return order (N, up, down, fc, cpm)
  scipy/sparse/_base.py:1561-1577 mean     sum(x_i * fl(1/count))
"""
import itertools
import math
import multiprocessing

import numpy as np
import scipy.sparse as sp

from . import hypergeom_port

try:  # live scipy is the un-vendored dependency itself; preferred when present
    from scipy.stats import hypergeom as _scipy_hypergeom
except Exception:  # pragma: no cover
    _scipy_hypergeom = None


# --------------------------------------------------------------------------- #
# a1: CPM normalisation (utils.py:150-155)
# --------------------------------------------------------------------------- #
def cpm_normalization_port(counts) -> sp.csr_matrix:
    """genes x cells integer counts -> CSR float64 CPM, reference operation order."""
    counts = sp.csr_matrix(counts)
    col_tot = np.asarray(counts.sum(axis=0)).ravel()           # int64 per cell
    with np.errstate(divide="ignore"):
        inv = 1.0 / col_tot                                        # fl(1/colsum)
    out = counts.astype(np.float64).tocsr(copy=True)
    out.data = (out.data * 1e6) * inv[out.indices]                 # fl(fl(x*1e6)*inv)
    return out


# --------------------------------------------------------------------------- #
# a3..a7: one gene, one cell set
# --------------------------------------------------------------------------- #
def _tails(k, M, n, N, use_scipy=True):
    if not all([k, M, n, N]):
        return math.nan, math.nan
    if use_scipy and _scipy_hypergeom is not None:
        return (float(_scipy_hypergeom.logsf(k, M, n, N)),
                float(_scipy_hypergeom.logcdf(k, M, n, N)))
    return hypergeom_port.logsf(k, M, n, N), hypergeom_port.logcdf(k, M, n, N)


def _mean_like_scipy_sparse(values, count):
    # scipy/sparse/_base.py:1574  (x * (1.0/denom)).sum()
    return float(np.sum(np.asarray(values, dtype=np.float64) * (1.0 / count)))


def gene_test(row, member_idx, use_scipy=True):
    """One gene (dense float64 vector over all cells) against one cell set.

    Complement background, ``utils.py:194-215``.  Returns ``(down, up, fc, cpm)``.
    """
    row = np.asarray(row, dtype=np.float64).ravel()
    members = np.asarray(member_idx, dtype=np.int64)
    C = row.shape[0]
    cutoff = np.median(row)
    low = row <= cutoff
    in_set = np.zeros(C, dtype=bool)
    in_set[members] = True
    k = int(np.count_nonzero(low & in_set))          # |low ∩ set| (set semantics)
    M, n, N = C, int(np.count_nonzero(low)), int(members.shape[0])
    cpm = _mean_like_scipy_sparse(row[members], N)
    rest = _mean_like_scipy_sparse(row[~in_set], C - int(np.count_nonzero(in_set)))
    fc = (cpm + 0.01) / (rest + 0.01)
    down, up = _tails(k, M, n, N, use_scipy)
    return down, up, fc, cpm


def gene_test_nc(row, member_idx, nc_idx, use_scipy=True):
    """Negative-control background, ``utils.py:248-266``."""
    row = np.asarray(row, dtype=np.float64).ravel()
    members = np.asarray(member_idx, dtype=np.int64)
    nc = np.asarray(nc_idx, dtype=np.int64)
    C = row.shape[0]
    cutoff = np.median(row[nc])
    low = row <= cutoff
    in_set = np.zeros(C, dtype=bool)
    in_set[members] = True
    k = int(np.count_nonzero(low & in_set))
    M, n, N = C, int(np.count_nonzero(low)), int(members.shape[0])
    cpm = _mean_like_scipy_sparse(row[members], N)
    bg = _mean_like_scipy_sparse(row[nc], nc.shape[0])
    fc = (cpm + 0.01) / (bg + 0.01)
    down, up = _tails(k, M, n, N, use_scipy)
    return down, up, fc, cpm


def _gene_task(args):
    data, cols, C, members, nc = args
    row = np.zeros(C, dtype=np.float64)
    row[cols] = data
    if nc is None:
        return gene_test(row, members)
    return gene_test_nc(row, members, nc)


def _fan_out(member_idx, nc_idx, X, idx, num_processes, down, up, fc, cpm):
    X = sp.csr_matrix(X)
    C = X.shape[1]
    members = np.asarray(member_idx, dtype=np.int64)
    nc = None if nc_idx is None else np.asarray(nc_idx, dtype=np.int64)
    tasks = ((X.data[X.indptr[g]:X.indptr[g + 1]], X.indices[X.indptr[g]:X.indptr[g + 1]],
              C, members, nc) for g in range(X.shape[0]))
    if num_processes and num_processes > 1:
        with multiprocessing.Pool(processes=num_processes) as pool:
            results = pool.map(_gene_task, tasks, chunksize=64)
    else:
        results = [_gene_task(t) for t in tasks]
    for i, (d, u, f, c) in zip(idx, results):
        down[i], up[i], fc[i], cpm[i] = d, u, f, c
    return len(member_idx), up, down, fc, cpm


def perform_DE_port(sgrna_idx, input_array, idx, num_processes,
                    pval_list_down, pval_list_up, fc_list, cpm_list):
    """Same contract as ``utils.perform_DE`` (utils.py:296-317)."""
    return _fan_out(sgrna_idx, None, input_array, idx, num_processes,
                    pval_list_down, pval_list_up, fc_list, cpm_list)


def perform_DE_NC_port(sgrna_idx, NC_idx, input_array, idx, num_processes,
                       pval_list_down, pval_list_up, fc_list, cpm_list):
    """Same contract as ``utils.perform_DE_NC`` (utils.py:319-341)."""
    return _fan_out(sgrna_idx, NC_idx, input_array, idx, num_processes,
                    pval_list_down, pval_list_up, fc_list, cpm_list)


# --------------------------------------------------------------------------- #
# all-genes form
# --------------------------------------------------------------------------- #
def _order_stat(neg_sorted, zeros, pos_sorted, i):
    """i-th smallest (0-based) of a row given as sorted negatives, a zero count,
    and sorted positives."""
    if i < neg_sorted.shape[0]:
        return float(neg_sorted[i])
    i -= neg_sorted.shape[0]
    if i < zeros:
        return 0.0
    return float(pos_sorted[i - zeros])


def gene_cutoffs(X, nc_idx=None):
    """Per-gene ``(median, n)``: median over all cells (or over ``nc_idx`` cells),
    ``n = #{cells : x <= median}`` over ALL cells (utils.py:199-200 / :253-254)."""
    X = sp.csr_matrix(X)
    G, C = X.shape
    med = np.zeros(G, dtype=np.float64)
    n = np.zeros(G, dtype=np.int64)
    if nc_idx is not None:
        nc = np.asarray(nc_idx, dtype=np.int64)
        in_nc = np.zeros(C, dtype=bool)
        in_nc[nc] = True
        L = nc.shape[0]
    else:
        in_nc, L = None, C
    lo_i, hi_i = (L - 1) // 2, L // 2
    for g in range(G):
        s, e = X.indptr[g], X.indptr[g + 1]
        vals = X.data[s:e]
        if in_nc is not None:
            sub = vals[in_nc[X.indices[s:e]]]
        else:
            sub = vals
        sub = sub[sub != 0.0]
        neg = np.sort(sub[sub < 0])
        pos = np.sort(sub[sub > 0])
        zeros = L - neg.shape[0] - pos.shape[0]
        a = _order_stat(neg, zeros, pos, lo_i)
        b = _order_stat(neg, zeros, pos, hi_i)
        m = a if lo_i == hi_i else (a + b) / 2.0          # np.median: fl(a+b)/2
        med[g] = m
        stored_low = int(np.count_nonzero(vals <= m))
        implicit = C - (e - s)
        n[g] = stored_low + (implicit if 0.0 <= m else 0)
    return med, n


def member_matrix(sets, C):
    """cells x sets 0/1 CSC indicator; duplicates inside a set are rejected."""
    rows, cols = [], []
    for s, members in enumerate(sets):
        members = np.asarray(members, dtype=np.int64)
        if np.unique(members).shape[0] != members.shape[0]:
            raise ValueError("duplicate cell index inside a set")
        rows.append(members)
        cols.append(np.full(members.shape[0], s, dtype=np.int64))
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    return sp.csc_matrix((np.ones(rows.shape[0], dtype=np.float64), (rows, cols)),
                         shape=(C, len(sets)))


def unique_tails(k, n, N, M, tail_fn=None):
    """Evaluate ``(down, up)`` once per distinct ``(k, n, N)`` and scatter back.

    ``k``: [S, G], ``n``: [G], ``N``: [S].  ``tail_fn(k, M, n, N) -> (logcdf, logsf)``
    arrays; default is live scipy (falls back to the python port)."""
    S, G = k.shape
    kk = k.astype(np.int64).ravel()
    nn = np.broadcast_to(n.astype(np.int64)[None, :], (S, G)).ravel()
    NN = np.broadcast_to(N.astype(np.int64)[:, None], (S, G)).ravel()
    key = np.stack([kk, nn, NN], axis=1)
    uniq, inv = np.unique(key, axis=0, return_inverse=True)
    inv = inv.ravel()
    uk, un, uN = uniq[:, 0], uniq[:, 1], uniq[:, 2]
    if tail_fn is not None:
        ucdf, usf = tail_fn(uk, M, un, uN)
    elif _scipy_hypergeom is not None:
        ucdf = _scipy_hypergeom.logcdf(uk, M, un, uN)
        usf = _scipy_hypergeom.logsf(uk, M, un, uN)
    else:
        ucdf = np.array([hypergeom_port.logcdf(a, M, b, c) for a, b, c in uniq])
        usf = np.array([hypergeom_port.logsf(a, M, b, c) for a, b, c in uniq])
    ucdf = np.asarray(ucdf, dtype=np.float64).copy()
    usf = np.asarray(usf, dtype=np.float64).copy()
    degenerate = (uk == 0) | (un == 0) | (uN == 0) | (M == 0)      # all([k,M,n,N]) guard
    ucdf[degenerate] = np.nan
    usf[degenerate] = np.nan
    return usf[inv].reshape(S, G), ucdf[inv].reshape(S, G)


def de_tables(X, sets, nc_idx=None, tail_fn=None):
    """All genes x all sets at once.  Returns a dict with integer arrays
    ``k`` [S,G], ``n`` [G], ``N`` [S], ``M`` and float64 ``down, up, fc, cpm`` [S,G]."""
    X = sp.csr_matrix(X).astype(np.float64)
    G, C = X.shape
    med, n = gene_cutoffs(X, nc_idx)
    ind = member_matrix(sets, C)                               # C x S
    N = np.asarray(ind.sum(axis=0)).ravel().astype(np.int64)
    # low-set membership of stored entries; implicit zeros are low iff 0 <= median
    row_of = np.repeat(np.arange(G), np.diff(X.indptr))
    stored_low = sp.csr_matrix(((X.data <= med[row_of]).astype(np.float64), X.indices, X.indptr),
                               shape=X.shape)
    stored_any = sp.csr_matrix((np.ones_like(X.data), X.indices, X.indptr), shape=X.shape)
    low_hits = np.asarray((stored_low @ ind).todense()).T          # S x G
    any_hits = np.asarray((stored_any @ ind).todense()).T
    zero_low = (0.0 <= med)[None, :]
    k = np.rint(low_hits + np.where(zero_low, N[:, None] - any_hits, 0.0)).astype(np.int64)
    sums = np.asarray((X @ ind).todense()).T                       # S x G member sums
    with np.errstate(divide="ignore", invalid="ignore"):
        cpm = sums * (1.0 / N[:, None])
        if nc_idx is None:
            rowsum = np.asarray(X.sum(axis=1)).ravel()
            rest = (rowsum[None, :] - sums) * (1.0 / (C - N[:, None]))
            # every stored value inside the set -> the rest is exactly zero
            nnz_g = np.diff(X.indptr)
            rest = np.where(any_hits == nnz_g[None, :], 0.0, rest)
            # rowsum - sums cancels when the set holds nearly all of a row's mass (rows with a few
            # huge cells): those entries are summed directly over the complement, as the
            # reference does (utils.py:206-207)
            abs_rowsum = np.asarray(abs(X).sum(axis=1)).ravel()
            risky = np.abs(rowsum[None, :] - sums) < 1e-5 * abs_rowsum[None, :]
            risky &= any_hits != nnz_g[None, :]
            for s_i, g_i in zip(*np.nonzero(risky)):
                row = np.asarray(X[g_i].todense()).ravel()
                outside = np.ones(C, dtype=bool)
                outside[np.asarray(sets[s_i], dtype=np.int64)] = False
                rest[s_i, g_i] = _mean_like_scipy_sparse(row[outside], C - int(N[s_i]))
        else:
            nc = np.asarray(nc_idx, dtype=np.int64)
            bg = np.asarray(X[:, nc].sum(axis=1)).ravel() * (1.0 / nc.shape[0])
            rest = np.broadcast_to(bg[None, :], sums.shape)
        fc = (cpm + 0.01) / (rest + 0.01)
    down, up = unique_tails(k, n, N, C, tail_fn)
    return dict(k=k, n=n, N=N, M=C, median=med, down=down, up=up, fc=fc, cpm=cpm)
