"""GPU tier: the accuracy contract of fc / cpm for ANY float64 matrix (``perform_DE``'s
``input_array``, pySpade/utils.py:296).  Member sums are carried in a per-gene fixed point;
genes whose values span too many binary orders of magnitude for it ("wide" genes) are summed
in float64 by the side path (``csrc/wide.cuh``), like the reference does (scipy sparse mean,
utils.py:206-208, 258-259).  Tolerance everywhere: 1e-6 relative (north_star), written in
``helpers.VALUE_RTOL``; integers exact."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import assert_tables_close, max_rel_err
from oracle import c_oracle, de_port

pytestmark = pytest.mark.gpu


def _adversarial(C, rng):
    """Rows built to break a fixed point with one scale per gene."""
    rows = []
    a = np.zeros(C); a[rng.choice(C, C // 3, replace=False)] = 1.2345; a[7] = 1e9      # one huge outlier
    rows.append(a)
    b = np.zeros(C); b[: C // 2] = 2e4; b[C // 2: C // 2 + 200] = 7.0 / 3.0            # CPM-range + tiny tail
    rows.append(b)
    c = np.exp(rng.uniform(np.log(1e-6), np.log(1e6), C)) * (rng.random(C) < 0.6)      # log-uniform 1e-6..1e6
    rows.append(c)
    d = np.zeros(C); d[rng.choice(C, 50, replace=False)] = 3e-300; d[11] = 1e200        # 500 binary orders
    rows.append(d)
    # a ladder of dynamic ranges 2^0 .. 2^60 around the wide / not-wide decision
    for e in range(0, 61, 4):
        r = np.zeros(C)
        idx = rng.choice(C, C // 4, replace=False)
        r[idx] = rng.uniform(1.0, 2.0, idx.shape[0])
        r[idx[:3]] *= 2.0 ** e
        rows.append(r)
    return np.stack(rows)


@pytest.mark.parametrize("mode", ["complement", "nc"])
def test_wide_dynamic_range_genes(mode):
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    rng = np.random.default_rng(77)
    C = 3000
    normal = de_port.cpm_normalization_port(synth.counts_numpy(200, C, seed=5)).toarray()
    dense = np.vstack([normal[:100], _adversarial(C, rng), normal[100:]])
    X = sp.csr_matrix(dense)
    n_adv = dense.shape[0] - 200
    sets = [rng.choice(C, n, replace=False) for n in (40, 40, 500, 1500)]
    sets.append(np.flatnonzero(dense[100] == 1.2345)[:60])           # members all tiny, outlier outside
    sets.append(np.r_[7, np.flatnonzero(dense[100] == 1.2345)[:59]])  # ... and with the outlier inside
    sets.append(np.arange(C // 2, C // 2 + 150))                      # the 7/3 tail of the second gene
    nc = rng.choice(C, 300, replace=False) if mode == "nc" else None
    ora = de_port.de_tables(X, sets, nc, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    with Engine.from_cpm(X) as eng:
        eng.set_background(nc)
        wide = eng.wide_genes
        assert 4 <= wide <= n_adv          # the four hand-built genes at least; no ordinary CPM gene
        host = eng.run(sets, want_k=True)
        import torch
        members, offsets = synth.concat_sets(sets)
        dev = {k: torch.empty((len(sets), X.shape[0]), dtype=torch.float64, device="cuda")
               for k in ("up", "down", "fc", "cpm")}
        eng.run_device(torch.from_numpy(members).cuda(), offsets, dev)
        eng.check()
        eng.set_option("member_block", 64)                            # split path
        split = eng.run(sets, want_k=True)
        with pytest.raises(Exception, match="float64 side path"):
            eng.run_raw(torch.from_numpy(members).cuda(), offsets,
                        torch.zeros((len(sets), X.shape[0]), dtype=torch.int64, device="cuda"))
    assert np.array_equal(host["k"], ora["k"])
    assert_tables_close(host, ora)
    # the side path sums in float64: far inside the contract
    adv = slice(100, 100 + n_adv)
    for name in ("fc", "cpm"):
        assert max_rel_err(host[name][:, adv], ora[name][:, adv]) <= 1e-6
        assert np.array_equal(dev[name].cpu().numpy(), host[name], equal_nan=True)
        assert np.array_equal(split[name], host[name], equal_nan=True)


def test_count_data_has_no_wide_genes():
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    with Engine.from_counts(synth.counts_numpy(3000, 4000, seed=1)) as eng:
        assert eng.wide_genes == 0
        eng.set_background(np.arange(0, 4000, 7))
        assert eng.wide_genes == 0


def test_device_output_errors_surface_in_check():
    """run_device returns before its kernels have run: a bad member list is reported by
    ``spade_check`` (or by the next call once the kernels have finished)."""
    import torch
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C = 300, 900
    with Engine.from_counts(synth.counts_numpy(G, C, seed=2)) as eng:
        out = {k: torch.empty((2, G), dtype=torch.float64, device="cuda") for k in ("up", "cpm")}
        offsets = np.array([0, 3, 6], dtype=np.int64)
        eng.run_device(torch.tensor([1, 2, 3, 4, C + 5, 6], dtype=torch.int32, device="cuda"), offsets, out)
        with pytest.raises(IndexError):
            eng.check()
        eng.check()                                                    # reported once
        eng.run_device(torch.tensor([1, 2, 2, 4, 5, 6], dtype=torch.int32, device="cuda"), offsets, out)
        torch.cuda.synchronize()
        with pytest.raises(ValueError):                                # surfaces at the next call
            eng.run_device(torch.tensor([1, 2, 3, 4, 5, 6], dtype=torch.int32, device="cuda"), offsets, out)
        eng.run_device(torch.tensor([1, 2, 3, 4, 5, 6], dtype=torch.int32, device="cuda"), offsets, out)
        eng.check()
