"""GPU tier at BASELINE.json's full size (50 000 cells x 33 000 genes, the bench matrix):
size-independent properties instead of a CPU oracle that would need days.

* a partition of ALL cells into disjoint sets: the per-set counts k add up to n_g exactly and
  the per-set CPM sums add up to the row sums (a checksum of checksums over every stored value);
* logcdf and logsf are complementary: exp(up) + exp(down) = 1;
* fold change re-derived from the same run's mean CPM and the row sums;
* sharding invariance: four shards of a batch give the bytes of the whole batch;
* host tables (pipelined strips) equal device tables bit for bit.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G, C = 33000, 50000


@pytest.fixture(scope="module")
def staged():
    import torch
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    indptr, indices, counts = synth.counts_torch(G, C, seed=0, device="cuda")
    eng = Engine.from_device_csr(G, C, indptr, indices, counts)
    colsum = torch.zeros(C, dtype=torch.int64, device="cuda").index_add_(0, indices.long(), counts.long())
    val = (counts.double() * 1e6) * (1.0 / colsum.double())[indices.long()]
    rows = torch.repeat_interleave(torch.arange(G, device="cuda"), indptr[1:] - indptr[:-1])
    rowsum = torch.zeros(G, dtype=torch.float64, device="cuda").index_add_(0, rows, val)
    del val, rows
    yield eng, rowsum
    eng.close()


def _run(eng, sets, want=("up", "down", "fc", "cpm")):
    import torch
    from pyspade_b200 import synth
    members, offsets = synth.concat_sets(sets)
    out = {k: torch.empty((len(sets), G), dtype=torch.float64, device="cuda") for k in want}
    k = torch.empty((len(sets), G), dtype=torch.int32, device="cuda")
    eng.run_device(torch.from_numpy(members).cuda(), offsets, out, k=k)
    torch.cuda.synchronize()
    return out, k


def test_partition_of_all_cells_checksums(staged):
    import torch
    eng, rowsum = staged
    N = 500
    perm = np.random.default_rng(5).permutation(C)
    sets = [perm[i * N:(i + 1) * N] for i in range(C // N)]
    out, k = _run(eng, sets)
    med, n = eng.gene_cutoffs()
    assert (med >= 0).all()
    # every cell is in exactly one set: the low-side counts add up to n_g, exactly
    assert np.array_equal(k.sum(dim=0, dtype=torch.int64).cpu().numpy(), n.astype(np.int64))
    # ... and the member sums add up to the row sum of the CPM matrix
    total = (out["cpm"] * N).sum(dim=0)
    rel = ((total - rowsum).abs() / rowsum.clamp_min(1e-300))[rowsum > 0]
    assert float(rel.max()) <= 1e-9
    assert bool((total[rowsum == 0] == 0).all())
    # complementary tails where both are finite; nan exactly where k == 0
    up, down = out["up"], out["down"]
    assert bool(torch.equal(torch.isnan(up), k == 0)) and bool(torch.equal(torch.isnan(down), k == 0))
    fin = torch.isfinite(up) & torch.isfinite(down)
    assert float((torch.exp(up[fin]) + torch.exp(down[fin]) - 1.0).abs().max()) <= 1e-9
    edge = ~torch.isnan(up) & ~fin                      # support edge: (0, -inf)
    assert bool((up[edge] == 0).all()) and bool(torch.isneginf(down[edge]).all())
    # fold change from the same run's mean CPM and the row sums (utils.py:206-207)
    rest = (rowsum[None, :] - out["cpm"] * N).clamp_min(0) / (C - N)
    fc = (out["cpm"] + 0.01) / (rest + 0.01)
    assert float(((out["fc"] - fc).abs() / fc).max()) <= 1e-9


def test_sharding_invariance_full_size(staged):
    import torch
    from pyspade_b200 import synth
    eng, _ = staged
    draws = synth.derand_draws(C, 500, 256)
    whole, kw = _run(eng, draws, want=("up", "down", "cpm"))
    for r in range(4):
        part, kp = _run(eng, draws[r::4], want=("up", "down", "cpm"))
        for name in part:
            assert bool(torch.equal(part[name].view(torch.int64), whole[name][r::4].view(torch.int64))), name
        assert bool(torch.equal(kp, kw[r::4]))


def test_host_tables_equal_device_tables_full_size(staged):
    from pyspade_b200 import synth
    eng, _ = staged
    draws = synth.derand_draws(C, 500, 64)
    dev, kd = _run(eng, draws)
    host = eng.run(draws, want_k=True)
    for name in dev:
        assert np.array_equal(dev[name].cpu().numpy(), host[name], equal_nan=True), name
    assert np.array_equal(kd.cpu().numpy(), host["k"])
