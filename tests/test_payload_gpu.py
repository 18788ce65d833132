"""GPU tier: the wire format of the DErand tables built on the GPU (``spade_csc_count`` /
``spade_csc_fill``, pySpade/DErand.py:230-248) and ``Cpm_mean`` (DErand.py:252)."""
import os

import numpy as np
import pytest
from scipy import io as sio

pytestmark = pytest.mark.gpu


def _table(S, G, seed):
    rng = np.random.default_rng(seed)
    t = rng.normal(size=(S, G))
    t[rng.random((S, G)) < 0.45] = 0.0
    t[rng.random((S, G)) < 0.05] = -0.0
    t[rng.random((S, G)) < 0.05] = np.nan
    t[rng.random((S, G)) < 0.05] = -np.inf
    t[:, G // 2] = 0.0                                     # an empty column
    return t


@pytest.mark.parametrize("S,G", [(56, 30), (1000, 777), (1, 5), (33, 1)])
def test_device_payload_equals_host_payload(S, G):
    import torch
    from pyspade_b200 import io as spio
    t = _table(S, G, S + G)
    jc, ir, pr = spio.dense_payload(t)
    dj, di, dp = spio.device_payload(torch.from_numpy(t).cuda())
    assert np.array_equal(jc, dj) and np.array_equal(ir, di)
    assert np.array_equal(pr.view(np.int64), dp.view(np.int64))           # values bit for bit, nan included
    # a block of columns of a wider table (what a rank holds after the gene-sharded exchange)
    wide = torch.from_numpy(np.hstack([t, t[:, ::-1]])).cuda()
    bj, bi, bp = spio.device_payload(wide[:, :G])
    assert np.array_equal(jc, bj) and np.array_equal(ir, bi) and np.array_equal(pr.view(np.int64), bp.view(np.int64))


def test_saved_table_loads_like_scipys(tmp_path):
    import scipy.sparse as sp
    import torch
    from pyspade_b200 import io as spio
    t = _table(200, 91, 4)
    spio.save_draw_table(str(tmp_path / "dev"), torch.from_numpy(t).cuda())
    sio.savemat(str(tmp_path / "ref.mat"), {"matrix": sp.csr_matrix(t)})
    a, b = sio.loadmat(str(tmp_path / "dev"))["matrix"].tocsc(), sio.loadmat(str(tmp_path / "ref.mat"))["matrix"].tocsc()
    assert a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    assert np.array_equal(a.data.view(np.int64), b.data.view(np.int64))
    # all zeros: scipy keeps one explicit slot
    spio.save_draw_table(str(tmp_path / "z"), torch.zeros((7, 3), dtype=torch.float64, device="cuda"))
    z = sio.loadmat(str(tmp_path / "z"))["matrix"]
    assert z.shape == (7, 3) and np.count_nonzero(z.toarray()) == 0


def test_column_mean():
    import torch
    from pyspade_b200.engine import column_mean
    rng = np.random.default_rng(1)
    t = rng.gamma(2.0, 50.0, size=(1000, 515)) * (rng.random((1000, 515)) < 0.3)
    got = column_mean(torch.from_numpy(t).cuda())
    want = np.mean(t, axis=0)
    assert np.array_equal(got == 0, want == 0)
    assert np.max(np.abs(got - want) / np.maximum(want, 1e-300)) <= 1e-12
