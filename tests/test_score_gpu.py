"""GPU tier: ``spade_score`` (row f4: gamma significance score and empirical p-value of
pySpade/global.py:107-187, local.py:104-177) against the reference's own lines run on the
reference's DEobs / DErand outputs (``tests/golden/score_case.npz``), against the oracle
restatement, and against live scipy on a bench-shaped table."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import score_port
from test_score import logsf_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def score_case():
    return np.load(os.path.join(GOLDEN, "score_case.npz"))


def test_gamma_logsf_kernel_vs_scipy_grid(score_case):
    """One 'region' whose raw p-value is -x: the kernel's preprocessing leaves negative finite
    values alone, so score = gamma.logsf(x, a, loc, scale)."""
    from pyspade_b200.score import significance
    z = score_case
    x = z["grid/x"]
    ok = np.isfinite(x) & (x > 0)                  # p = -x must survive the preprocessing unchanged
    score, _ = significance(-x[ok][None, :], None, z["grid/a"][ok], z["grid/loc"][ok], z["grid/scale"][ok])
    logsf_close(score[0], z["grid/logsf"][ok])


@pytest.mark.parametrize("bg", ["complement", "NonTarget"])
@pytest.mark.parametrize("mode", ["equal", "sgrna"])
@pytest.mark.parametrize("d,D", [("up", "Up"), ("down", "Down")])
def test_score_vs_reference_lines(score_case, cli_case, bg, mode, d, D):
    from pyspade_b200.score import preprocess_pvals, significance
    z, c = score_case, cli_case
    obs = np.stack([c[f"DEobs/{bg}/mat:{r}-{d}_log-pval"][0] for r in z[f"{bg}/regions"]])
    rand = c[f"DErand/{bg}/{mode}/matsp:20-{d}_log-pval:dense"]
    A, B, C = (c[f"DErand/{bg}/{mode}/npy:{D}_dist_gamma-20-{x}.npy"] for x in "ABC")
    score, emp = significance(obs, rand, A, B, C)
    assert np.array_equal(preprocess_pvals(obs), z[f"{bg}/{mode}/{d}/p"])
    logsf_close(score, z[f"{bg}/{mode}/{d}/score"])
    assert np.array_equal(emp, z[f"{bg}/{mode}/{d}/emp"])          # counts / S: exact


def test_score_large_table_vs_scipy_and_port():
    """1 000 draws x 3 000 genes with nan / -inf / 0 entries, 40 regions, device inputs."""
    import scipy.stats
    import torch
    from pyspade_b200.score import significance
    rng = np.random.default_rng(3)
    S, G, R = 1000, 3000, 40
    rand = -rng.gamma(rng.uniform(0.5, 4, G)[None, :], rng.uniform(0.2, 3, G)[None, :], size=(S, G))
    rand[rng.random((S, G)) < 0.02] = np.nan
    rand[rng.random((S, G)) < 0.02] = -np.inf
    rand[rng.random((S, G)) < 0.05] = 0.0
    rand[:, 5] = np.nan                                            # a gene never tested
    obs = -rng.gamma(2.0, 2.0, size=(R, G)) * rng.choice([1, 1, 1, 10], size=(R, G))
    obs[rng.random((R, G)) < 0.03] = np.nan
    obs[rng.random((R, G)) < 0.03] = -np.inf
    obs[rng.random((R, G)) < 0.03] = 0.0
    obs[0, :50] = rand[17, :50]                                    # ties: strict '<'
    A, B, C = rng.uniform(0.3, 30, G), rng.normal(0, 0.2, G), rng.uniform(0.1, 5, G)
    A[:3], B[:3], C[:3] = 0.0, 0.0, 0.0                            # genes DErand skipped
    A[3], B[3], C[3] = np.nan, 0.0, 1.0                            # constant gene
    score, emp = significance(torch.from_numpy(obs).cuda(), torch.from_numpy(rand).cuda(), A, B, C)
    p = score_port.preprocess(obs)
    with np.errstate(all="ignore"):
        want = scipy.stats.gamma.logsf(-p, A[None, :], B[None, :], C[None, :])
        want_emp = np.stack([np.sum(rand < p[r][None, :], axis=0) / S for r in range(R)])
    logsf_close(score, want)
    assert np.array_equal(emp, want_emp)
    _, port, port_emp = score_port.score_tables(obs[:2, :300], rand[:, :300], A[:300], B[:300], C[:300])
    logsf_close(score[:2, :300], port)
    assert np.array_equal(emp[:2, :300], port_emp)
