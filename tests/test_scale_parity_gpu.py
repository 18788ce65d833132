"""GPU tier: sampled-oracle parity at the sizes of BASELINE.json configs[1..3] - 200 000 and
50 000 cells x 33 000 genes - for both backgrounds (``hypergeo_test_sparse`` utils.py:194-215,
``hypergeo_test_NC_sparse`` utils.py:248-266) and both batch shapes (DErand: uniform set size,
tail table; DEobs: ragged regions, direct tails).  The oracle gets a density-stratified sample
of 4 096 genes at full C (the whole matrix would take it hours); integers exact, floats within
1e-6 relative (``helpers``)."""
import numpy as np
import pytest

from helpers import assert_tables_close
from oracle import c_oracle, de_port

pytestmark = pytest.mark.gpu

G = 33000
SAMPLE_GENES = 4096


def _staged(C):
    import torch
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    indptr, indices, counts = synth.counts_torch(G, C, seed=0, device="cuda")
    genes, X = synth.sample_rows(indptr, indices, counts, G, C, SAMPLE_GENES)
    eng = Engine.from_device_csr(G, C, indptr, indices, counts)
    del indptr, indices, counts
    torch.cuda.empty_cache()
    return eng, genes, X


@pytest.fixture(scope="module")
def staged_200k():
    eng, genes, X = _staged(200000)
    yield eng, genes, X
    eng.close()


@pytest.fixture(scope="module")
def staged_50k():
    eng, genes, X = _staged(50000)
    yield eng, genes, X
    eng.close()


def _compare(eng, genes, X, sets, nc):
    import torch
    from pyspade_b200 import synth
    members, offsets = synth.concat_sets(sets)
    S = len(sets)
    out = {k: torch.empty((S, G), dtype=torch.float64, device="cuda") for k in ("up", "down", "fc", "cpm")}
    kd = torch.empty((S, G), dtype=torch.int32, device="cuda")
    eng.run_device(torch.from_numpy(members).cuda(), offsets, out, k=kd)
    eng.check()
    gsel = torch.from_numpy(genes).cuda()
    got = {k: v[:, gsel].cpu().numpy() for k, v in out.items()}
    ora = de_port.de_tables(X, sets, nc, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    med, n = eng.gene_cutoffs()
    assert np.array_equal(med[genes], ora["median"])
    assert np.array_equal(n[genes], ora["n"])
    assert np.array_equal(kd[:, gsel].cpu().numpy(), ora["k"])
    assert_tables_close(got, ora)


def _cases(C, N):
    from pyspade_b200 import synth
    derand = synth.derand_draws(C, N, 8)                                   # DErand.py:195-197
    deobs = synth.deobs_sets(C, 7, lo=50, hi=2000) + [np.random.default_rng(9).choice(C, 5000, replace=False)]
    nc = np.random.default_rng(2).choice(C, C // 50, replace=False)        # 2 % of the cells
    return derand, deobs, nc


@pytest.mark.parametrize("background", ["complement", "nc"])
@pytest.mark.parametrize("shape", ["derand", "deobs"])
def test_parity_200k_cells(staged_200k, background, shape):
    eng, genes, X = staged_200k
    derand, deobs, nc = _cases(200000, 1000)
    bg = nc if background == "nc" else None
    if (eng.n_nc > 0) != (bg is not None):
        eng.set_background(bg)
    assert eng.wide_genes == 0
    _compare(eng, genes, X, derand if shape == "derand" else deobs, bg)


@pytest.mark.parametrize("shape", ["derand", "deobs"])
def test_parity_nc_background_50k_cells(staged_50k, shape):
    eng, genes, X = staged_50k
    derand, deobs, nc = _cases(50000, 500)
    if eng.n_nc == 0:
        eng.set_background(nc)
    _compare(eng, genes, X, derand if shape == "derand" else deobs, nc)


def test_parity_200k_full_batch_tail_table_window(staged_200k):
    """1 000 draws at 200 000 cells (cfg4's DErand batch): the sampled genes of all draws against
    the oracle's integers, and the tables of the first draws bit-identical to the 8-draw run
    (tail-table entries are the direct evaluation itself)."""
    import torch
    from pyspade_b200 import synth
    eng, genes, X = staged_200k
    if eng.n_nc:
        eng.set_background(None)
    draws = synth.derand_draws(200000, 1000, 1000)
    members, offsets = synth.concat_sets(draws)
    out = {k: torch.empty((1000, G), dtype=torch.float64, device="cuda") for k in ("up", "down", "cpm")}
    kd = torch.empty((1000, G), dtype=torch.int32, device="cuda")
    eng.run_device(torch.from_numpy(members).cuda(), offsets, out, k=kd)
    eng.check()
    m8, o8 = synth.concat_sets(draws[:8])
    out8 = {k: torch.empty((8, G), dtype=torch.float64, device="cuda") for k in ("up", "down", "cpm")}
    eng.set_option("tail_table", 0)
    eng.run_device(torch.from_numpy(m8).cuda(), o8, out8)
    eng.set_option("tail_table", 1)
    eng.check()
    for name in out8:
        assert bool(torch.equal(out8[name].view(torch.int64), out[name][:8].view(torch.int64))), name
    ind = de_port.member_matrix(draws, 200000)
    med, n = de_port.gene_cutoffs(X)
    import scipy.sparse as sp
    stored_any = sp.csr_matrix((np.ones_like(X.data), X.indices, X.indptr), shape=X.shape)
    row_of = np.repeat(np.arange(X.shape[0]), np.diff(X.indptr))
    stored_low = sp.csr_matrix(((X.data <= med[row_of]).astype(np.float64), X.indices, X.indptr), shape=X.shape)
    k = np.rint(np.asarray((stored_low @ ind).todense()).T
                + np.where((0.0 <= med)[None, :], 1000 - np.asarray((stored_any @ ind).todense()).T, 0.0))
    gsel = torch.from_numpy(genes).cuda()
    assert np.array_equal(kd[:, gsel].cpu().numpy(), k.astype(np.int64))
