"""CPU tier: the oracle restatement of the significance scoring (row f4; pySpade/global.py:107-187,
local.py:104-177) against vectors produced by the reference's own lines under scipy 1.18.1
(``tests/golden/score_case.npz``, ``tests/golden/make_score_golden.py``)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import score_port

# gamma.logsf: 1e-6 relative; next to 0 scipy's value is log of a double-rounded sf, so an absolute
# 1e-15 is allowed there
LOGSF_RTOL, LOGSF_ATOL = 1e-6, 1e-15


@pytest.fixture(scope="module")
def score_case():
    return np.load(os.path.join(GOLDEN, "score_case.npz"))


def logsf_close(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.isneginf(got), np.isneginf(want))
    m = np.isfinite(want)
    err = np.abs(got[m] - want[m])
    ok = err <= LOGSF_RTOL * np.abs(want[m]) + LOGSF_ATOL
    assert ok.all(), (float(err.max()), int((~ok).sum()))


def test_gamma_logsf_port_vs_scipy_grid(score_case):
    z = score_case
    got = [score_port.gamma_logsf_port(*t) for t in zip(z["grid/x"], z["grid/a"], z["grid/loc"], z["grid/scale"])]
    logsf_close(got, z["grid/logsf"])


@pytest.mark.parametrize("bg", ["complement", "NonTarget"])
@pytest.mark.parametrize("mode", ["equal", "sgrna"])
@pytest.mark.parametrize("d,D", [("up", "Up"), ("down", "Down")])
def test_score_port_vs_reference_lines(score_case, cli_case, bg, mode, d, D):
    z, c = score_case, cli_case
    obs = np.stack([c[f"DEobs/{bg}/mat:{r}-{d}_log-pval"][0] for r in z[f"{bg}/regions"]])
    rand = c[f"DErand/{bg}/{mode}/matsp:20-{d}_log-pval:dense"]
    A, B, C = (c[f"DErand/{bg}/{mode}/npy:{D}_dist_gamma-20-{x}.npy"] for x in "ABC")
    p, score, emp = score_port.score_tables(obs, rand, A, B, C)
    assert np.array_equal(p, z[f"{bg}/{mode}/{d}/p"])
    logsf_close(score, z[f"{bg}/{mode}/{d}/score"])
    assert np.array_equal(emp, z[f"{bg}/{mode}/{d}/emp"])
