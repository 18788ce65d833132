"""GPU tier: seeded random matrices and cell sets against the oracle (general float64 matrices
as perform_DE accepts them: negative values, explicitly stored zeros, ties, empty rows and
columns, unsorted column indices), plus argument errors of the C ABI."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import LOGP_RTOL, VALUE_RTOL, assert_tables_close, max_rel_err
from oracle import c_oracle, de_port

pytestmark = pytest.mark.gpu


def _random_case(rng):
    G = int(rng.integers(1, 140))
    C = int(rng.integers(2, 400))
    kind = rng.choice(["counts", "signed", "ties", "sparse"])
    dens = float(rng.choice([0.02, 0.15, 0.5, 0.95]))
    mask = rng.random((G, C)) < dens
    if kind == "counts":
        dense = rng.poisson(3.0, (G, C)).astype(np.float64) * mask
    elif kind == "signed":
        dense = rng.normal(0, 2, (G, C)) * mask
    elif kind == "ties":
        dense = rng.integers(-2, 3, (G, C)).astype(np.float64) * mask
    else:
        dense = rng.gamma(0.3, 50.0, (G, C)) * mask
    if G > 2:
        dense[rng.integers(G)] = 0.0                     # an all-zero gene
        dense[rng.integers(G)] = -np.abs(dense[rng.integers(G)]) - 1.0     # an all-negative gene
    X = sp.csr_matrix(dense)
    if X.nnz and rng.random() < 0.5:                     # explicitly stored zeros
        X.data[rng.integers(0, X.nnz, max(1, X.nnz // 20))] = 0.0
    if rng.random() < 0.5:                               # column indices in arbitrary order
        for g in range(G):
            lo, hi = X.indptr[g], X.indptr[g + 1]
            perm = rng.permutation(hi - lo)
            X.indices[lo:hi] = X.indices[lo:hi][perm]
            X.data[lo:hi] = X.data[lo:hi][perm]
    n_sets = int(rng.integers(1, 7))
    sizes = [int(rng.integers(0, C + 1)) for _ in range(n_sets)]
    if rng.random() < 0.4:
        sizes = [sizes[0]] * 40                          # uniform sizes: the tail-table path
    sets = [rng.choice(C, n, replace=False) for n in sizes]
    nc = rng.choice(C, int(rng.integers(1, C)), replace=False) if rng.random() < 0.5 else None
    return X, sets, nc


@pytest.mark.parametrize("seed", range(24))
def test_random_matrix_vs_oracle(seed):
    from pyspade_b200.engine import Engine
    rng = np.random.default_rng(1000 + seed)
    X, sets, nc = _random_case(rng)
    Xo = X.copy()
    Xo.sort_indices()
    live = [i for i, s in enumerate(sets) if len(s) > 0 and len(s) < X.shape[1]]
    with Engine.from_cpm(X) as eng:
        eng.set_background(nc)
        if seed % 3 == 0:
            eng.set_option("genes_per_part", 32)
        got = eng.run(sets, want_k=True)
        med, n = eng.gene_cutoffs()
    ora = de_port.de_tables(Xo, [sets[i] for i in live], nc, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    assert np.array_equal(med, ora["median"]) and np.array_equal(n, ora["n"])
    if live:
        sub = {k: v[live] for k, v in got.items()}
        assert np.array_equal(sub["k"], ora["k"])
        assert max_rel_err(sub["up"], ora["up"]) <= LOGP_RTOL
        assert max_rel_err(sub["down"], ora["down"]) <= LOGP_RTOL
        # means of signed values can cancel: compare on the scale of the row's magnitude
        scale = np.maximum(np.asarray(abs(Xo).max(axis=1).todense()).ravel(), 1e-300)[None, :]
        assert np.max(np.abs(sub["cpm"] - ora["cpm"]) / scale) <= VALUE_RTOL
        if (Xo.data >= 0).all():
            assert_tables_close(sub, ora, names=("fc", "cpm"))
    for i, s in enumerate(sets):                          # degenerate sizes
        if len(s) == 0:
            assert np.isnan(got["up"][i]).all() and np.isnan(got["cpm"][i]).all()


def test_c_abi_argument_errors():
    from pyspade_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    indptr = np.array([0, 1, 2], dtype=np.int64)
    idx = np.array([0, 9], dtype=np.int32)                # cell 9 does not exist (C = 3)
    val = np.array([1.0, 2.0])
    p = lambda a: ctypes.c_void_p(a.ctypes.data)          # noqa: E731
    assert lib.spade_create(0, 2, 3, p(indptr), p(idx), p(val), _lib.SPADE_HOST, ctypes.byref(h)) == _lib.SPADE_ERR_ARG
    assert b"out of range" in lib.spade_last_error(None)
    assert lib.spade_create(0, 2, 3, None, p(idx), p(val), _lib.SPADE_HOST, ctypes.byref(h)) == _lib.SPADE_ERR_ARG
    assert lib.spade_create(99, 2, 3, p(indptr), p(idx), p(val), _lib.SPADE_HOST, ctypes.byref(h)) == _lib.SPADE_ERR_ARG
    assert lib.spade_create(0, -1, 3, p(indptr), p(idx), p(val), _lib.SPADE_HOST, ctypes.byref(h)) == _lib.SPADE_ERR_ARG
    bad_ptr = np.array([1, 1, 2], dtype=np.int64)         # indptr must start at 0
    assert lib.spade_create(0, 2, 3, p(bad_ptr), p(idx), p(val), _lib.SPADE_HOST, ctypes.byref(h)) == _lib.SPADE_ERR_ARG
    idx[1] = 2
    assert lib.spade_create(0, 2, 3, p(indptr), p(idx), p(val), _lib.SPADE_HOST, ctypes.byref(h)) == _lib.SPADE_OK
    try:
        out = np.empty((1, 2))
        mem = np.array([0, 1], dtype=np.int32)
        off = np.array([0, 2], dtype=np.int64)
        run = lambda o, S, loc=_lib.SPADE_HOST: lib.spade_run(h, p(mem), _lib.SPADE_HOST, p(o), S, p(out), None,   # noqa: E731
                                                              None, None, None, loc, None)
        assert run(off, 1) == _lib.SPADE_OK
        assert run(np.array([1, 2], dtype=np.int64), 1) == _lib.SPADE_ERR_ARG          # offsets[0] != 0
        assert run(np.array([0, 2, 1], dtype=np.int64), 2) == _lib.SPADE_ERR_ARG       # decreasing
        assert run(off, -1) == _lib.SPADE_ERR_ARG
        assert run(off, 1, loc=7) == _lib.SPADE_ERR_ARG
        assert lib.spade_set_background(h, p(np.array([0, 0], dtype=np.int32)), 2) == _lib.SPADE_ERR_ARG   # duplicate
        assert lib.spade_set_background(h, p(np.array([5], dtype=np.int32)), 1) == _lib.SPADE_ERR_ARG
        assert lib.spade_set_option(h, 99, 1) == _lib.SPADE_ERR_ARG
        assert lib.spade_set_option(h, _lib.SPADE_OPT_MEMBER_BLOCK, 0) == _lib.SPADE_ERR_ARG
        assert lib.spade_run_strided(h, p(mem), _lib.SPADE_HOST, p(off), 1, p(out), None, None, None, None,
                                     _lib.SPADE_HOST, 1, None) == _lib.SPADE_ERR_ARG           # stride < G
        assert run(off, 1) == _lib.SPADE_OK                                            # still usable
    finally:
        lib.spade_destroy(h)
    assert lib.spade_run(None, None, 0, None, 0, None, None, None, None, None, 0, None) == _lib.SPADE_ERR_ARG
