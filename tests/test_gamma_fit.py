"""Row f1: the restatement of scipy.stats.gamma.fit in pyspade_b200/csrc/gamma_fit.cuh (the
source of the device kernel k_gamma_fit), compiled for the host and compared with scipy itself
- the arithmetic the reference calls at DErand.py:263-292 - on the reference's own DErand
tables (golden cli_case) and on synthetic -log p columns."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.stats as stats

from conftest import ROOT

HARNESS = os.path.join(ROOT, "tests", "gamma_host_harness.cpp")
OUT = os.path.join(ROOT, "oracle", "_build", "libgamma_host.so")


@pytest.fixture(scope="module")
def host_fit():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", OUT, HARNESS, "-lm"], check=True)
    lib = ctypes.CDLL(OUT)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    lib.gamma_fit_table.argtypes = [dp, ctypes.c_int, ctypes.c_long, ctypes.c_double, dp, dp, dp, ip]
    lib.gamma_nnlf.argtypes = [dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]
    lib.gamma_nnlf.restype = ctypes.c_double

    def fit(table, sign=-1.0):
        t = np.ascontiguousarray(table, dtype=np.float64)
        S, G = t.shape
        A, B, C = np.zeros(G), np.zeros(G), np.zeros(G)
        st = np.zeros(G, dtype=np.int32)
        lib.gamma_fit_table(t.ctypes.data_as(dp), S, G, sign, A.ctypes.data_as(dp), B.ctypes.data_as(dp),
                            C.ctypes.data_as(dp), st.ctypes.data_as(ip))
        return A, B, C, st

    def nnlf(x, a, loc, scale):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib.gamma_nnlf(x.ctypes.data_as(dp), x.shape[0], a, loc, scale)

    return fit, nnlf


def _scipy_column(neg):
    finite = neg[np.isfinite(neg)]
    if finite.shape[0] <= 50:
        return None
    if np.all(finite == finite[0]):
        return (np.nan, 0.0, 1.0)
    return stats.gamma.fit(finite)


def _same_fit(mine, want):
    """Shape and scale to 1e-3 relative, loc to 1e-6 absolute: Nelder-Mead stops at xtol = 1e-4,
    and with many exact zeros in the data the likelihood is unbounded as loc -> -0 (both
    implementations then stop at a tiny negative loc whose exact value is arbitrary)."""
    return (abs(mine[0] - want[0]) <= 1e-3 * abs(want[0]) and abs(mine[2] - want[2]) <= 1e-3 * abs(want[2])
            and abs(mine[1] - want[1]) <= 1e-6 + 1e-3 * abs(want[1]))


def test_penalized_nnlf_matches_scipy(host_fit):
    _, nnlf = host_fit
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.gamma(1.7, 2.0, 300), np.zeros(40)])
    for theta in [(1.7, -0.1, 2.0), (0.6, -1e-3, 0.5), (3.0, 0.5, 1.0), (1.0, -0.2, 1.3), (2.0, 0.0, 1.0),
                  (-1.0, 0.0, 1.0), (1.0, 0.0, -2.0)]:
        want = stats.gamma._penalized_nnlf(np.array(theta), x)
        got = nnlf(x, *theta)
        assert (np.isinf(want) and np.isinf(got)) or abs(got - want) <= 1e-9 * abs(want), theta


@pytest.mark.parametrize("case", ["complement/equal", "NonTarget/sgrna"])
def test_fit_on_reference_tables(host_fit, cli_case, case):
    """Same skipped genes, same degenerate genes; fitted genes: likelihood as good as scipy's
    and the parameters equal wherever scipy's own optimum is stable."""
    fit, nnlf = host_fit
    for which in ("up", "down"):
        table = cli_case[f"DErand/{case}/matsp:20-{which}_log-pval:dense"]
        A, B, C, st = fit(table)
        close = 0
        fitted = 0
        for g in range(table.shape[1]):
            want = _scipy_column(-table[:, g])
            if want is None:
                assert st[g] == 1 and (A[g], B[g], C[g]) == (0.0, 0.0, 0.0)
                continue
            if np.isnan(want[0]):
                assert st[g] == 2 and np.isnan(A[g]) and (B[g], C[g]) == (0.0, 1.0)
                continue
            assert st[g] == 0
            fitted += 1
            x = -table[:, g]
            x = x[np.isfinite(x)]
            close += _same_fit((A[g], B[g], C[g]), want)
        assert fitted > 0 and close >= 0.85 * fitted


def test_fit_on_synthetic_columns(host_fit):
    fit, nnlf = host_fit
    rng = np.random.default_rng(7)
    S, G = 400, 60
    shape = rng.uniform(0.4, 6.0, G)
    table = -(rng.gamma(shape[None, :], rng.uniform(0.2, 3.0, G)[None, :], (S, G)))
    table[rng.random((S, G)) < 0.15] = -0.0                 # exact zeros: p = 1
    table[rng.random((S, G)) < 0.03] = np.nan               # k == 0 tests
    table[rng.random((S, G)) < 0.03] = -np.inf
    table[:, 3] = np.nan
    table[:40, 3] = -1.5                                    # 40 finite values: skipped
    table[:, 4] = -2.25                                     # constant: degenerate
    A, B, C, st = fit(table)
    agree = []
    for g in range(G):
        want = _scipy_column(-table[:, g])
        if want is None:
            assert st[g] == 1
        elif np.isnan(want[0]):
            assert st[g] == 2
        else:
            x = -table[:, g]
            x = x[np.isfinite(x)]
            assert st[g] == 0
            agree.append(_same_fit((A[g], B[g], C[g]), want))
            if not agree[-1]:
                if abs(B[g]) < 1e-6 and abs(want[1]) < 1e-6:
                    # unbounded-likelihood regime (loc -> -0 under exact zeros): the stopping point
                    # of the simplex is arbitrary, shape and scale still agree to a few per cent
                    assert abs(A[g] - want[0]) <= 0.05 * want[0] and abs(C[g] - want[2]) <= 0.05 * want[2]
                else:                   # another optimum: not a worse one beyond the ftol scale
                    assert nnlf(x, A[g], B[g], C[g]) <= nnlf(x, *want) + 1e-2 * max(1.0, abs(nnlf(x, *want)))
    assert np.mean(agree) >= 0.85
