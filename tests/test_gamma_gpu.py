"""GPU tier, row f1: spade_gamma_fit (k_gamma_fit) against the host build of the same source and
against scipy.stats.gamma.fit on the reference's DErand tables and on synthetic columns."""
import numpy as np
import pytest

from test_gamma_fit import _same_fit, _scipy_column, host_fit   # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def _tables(cli_case):
    rng = np.random.default_rng(11)
    S, G = 300, 150
    shape = rng.uniform(0.4, 6.0, G)
    syn = -(rng.gamma(shape[None, :], rng.uniform(0.2, 3.0, G)[None, :], (S, G)))
    syn[rng.random((S, G)) < 0.15] = -0.0
    syn[rng.random((S, G)) < 0.03] = np.nan
    syn[rng.random((S, G)) < 0.03] = -np.inf
    syn[:, 3] = np.nan
    syn[:40, 3] = -1.5
    syn[:, 4] = -2.25
    return [cli_case["DErand/complement/equal/matsp:20-up_log-pval:dense"],
            cli_case["DErand/NonTarget/sgrna/matsp:20-down_log-pval:dense"], syn]


def test_device_fit_vs_host_build_and_scipy(host_fit, cli_case):   # noqa: F811
    from pyspade_b200.engine import gamma_fit
    fit, _ = host_fit
    for table in _tables(cli_case):
        A, B, C, st = gamma_fit(table)
        hA, hB, hC, hst = fit(table)
        assert np.array_equal(st, hst)
        fitted = st == 0
        # same source, device libm instead of the host's: the simplex paths coincide almost always
        same = [_same_fit((A[g], B[g], C[g]), (hA[g], hB[g], hC[g])) for g in np.flatnonzero(fitted)]
        assert np.mean(same) >= 0.8
        for g in np.flatnonzero(~fitted):
            assert np.array_equal([A[g], B[g], C[g]], [hA[g], hB[g], hC[g]], equal_nan=True)
        agree = []
        for g in range(table.shape[1]):
            want = _scipy_column(-table[:, g])
            if want is None:
                assert st[g] == 1 and (A[g], B[g], C[g]) == (0.0, 0.0, 0.0)
            elif np.isnan(want[0]):
                assert st[g] == 2 and np.isnan(A[g]) and (B[g], C[g]) == (0.0, 1.0)
            else:
                agree.append(_same_fit((A[g], B[g], C[g]), want))
        assert np.mean(agree) >= 0.8


def test_device_table_input_and_cli_option(tmp_path, cli_case):
    import torch
    from pyspade_b200 import DErand
    from pyspade_b200.engine import gamma_fit
    from test_cli_gpu import _read_outputs, _write_inputs
    table = cli_case["DErand/complement/equal/matsp:20-up_log-pval:dense"]
    a = gamma_fit(table)
    b = gamma_fit(torch.from_numpy(np.ascontiguousarray(table)).cuda())
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    tpk, spk, dtx = _write_inputs(str(tmp_path), cli_case)
    od = str(tmp_path / "rand") + "/"
    import os
    os.makedirs(od)
    np.random.seed(0)
    DErand.DE_random_cells(tpk, spk, dtx, 20, "equal", od, iteration=56, background_cells="complement",
                           gamma_fit="device")
    got = _read_outputs(od)
    want = cli_case["DErand/complement/equal/npy:Up_dist_gamma-20-A.npy"]
    mine = got["npy:Up_dist_gamma-20-A.npy"]
    assert np.array_equal(np.isnan(mine), np.isnan(want)) and np.array_equal(mine == 0, want == 0)
    m = ~np.isnan(want) & (want != 0)
    assert np.median(np.abs(mine[m] - want[m]) / np.abs(want[m])) <= 1e-3
