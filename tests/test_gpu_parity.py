"""GPU tier: the CUDA path (through the C ABI) against the oracle and the golden
fixtures.  Integers (k, n) bit-exact; log p-values, fold changes and mean CPM within
1e-6 relative (north_star); nan / -inf / 0.0 positions identical."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import LOGP_RTOL, assert_tables_close, max_rel_err, parse_float, sets_from_packed
from oracle import c_oracle, de_port

pytestmark = pytest.mark.gpu


def _engine(*a, **k):
    from pyspade_b200.engine import Engine
    return Engine


@pytest.mark.parametrize("tag", ["odd", "even"])
def test_cpm_on_device_bit_identical(small_case, tag):
    from pyspade_b200.engine import Engine
    z = small_case
    counts = sp.csr_matrix(z[f"{tag}_counts"])
    counts.sort_indices()
    with Engine.from_counts(counts) as eng:
        assert np.array_equal(eng.cpm_values(), z[f"{tag}_cpm_data"])     # utils.py:150-155, bit-exact


@pytest.mark.parametrize("tag", ["odd", "even"])
@pytest.mark.parametrize("mode", ["complement", "nc"])
@pytest.mark.parametrize("source", ["counts", "cpm"])
def test_small_case_vs_reference(small_case, tag, mode, source):
    from pyspade_b200.engine import Engine
    z = small_case
    counts = sp.csr_matrix(z[f"{tag}_counts"])
    if source == "counts":
        eng = Engine.from_counts(counts)
    else:
        eng = Engine.from_cpm(de_port.cpm_normalization_port(counts))
    sets = sets_from_packed(z[f"{tag}_set_members"], z[f"{tag}_set_offsets"])
    nc = z[f"{tag}_nc_idx"] if mode == "nc" else None
    with eng:
        eng.set_background(nc)
        got = eng.run(sets, want_k=True)
        med, n = eng.gene_cutoffs()
    X = de_port.cpm_normalization_port(counts)
    ora = de_port.de_tables(X, sets, nc)
    assert np.array_equal(med, ora["median"])          # float64 cutoffs, bit-exact
    assert np.array_equal(n, ora["n"])                 # integers, bit-exact
    assert np.array_equal(got["k"], ora["k"])
    want = {name: z[f"{tag}_{mode}_{name}"] for name in ("up", "down", "fc", "cpm")}
    assert_tables_close(got, want)                     # against the reference's own output
    assert_tables_close(got, ora)


def test_kat_genes(kat):
    from pyspade_b200.engine import Engine
    C = kat["C"]
    A = np.zeros(C); A[:600] = np.arange(600) + 1.0
    B = np.zeros(C); B[::7] = 3.0; B[::35] = 9.0
    Z = np.zeros(C)
    X = sp.csr_matrix(np.stack([A, B, Z]))
    row = {"A": 0, "B": 1, "Z": 2}
    sets = [np.array(c["members"]) for c in kat["cases"]]
    nc = np.arange(*kat["nc"])
    with Engine.from_cpm(X) as eng:
        for mode in ("complement", "nc"):
            eng.set_background(nc if mode == "nc" else None)
            got = eng.run(sets)
            for s, case in enumerate(kat["cases"]):
                g = row[case["gene"]]
                for name in ("up", "down", "fc", "cpm"):
                    want = parse_float(case[mode][name])
                    err = max_rel_err([got[name][s, g]], [want])
                    assert err <= 1e-6, (case["gene"], mode, name, got[name][s, g], want)


def test_hypergeom_kernel_vs_scipy_grid(hypergeom_grid, kat):
    from pyspade_b200.engine import hypergeom_logcdf_logsf
    g = hypergeom_grid
    lc, ls = hypergeom_logcdf_logsf(g["k"], g["M"], g["n"], g["N"])
    assert max_rel_err(lc, g["logcdf"]) <= LOGP_RTOL
    assert max_rel_err(ls, g["logsf"]) <= LOGP_RTOL
    for t in kat["hypergeom"]:
        lc, ls = hypergeom_logcdf_logsf([t["k"]], [t["M"]], [t["n"]], [t["N"]])
        assert max_rel_err(lc, [parse_float(t["logcdf"])]) <= LOGP_RTOL
        assert max_rel_err(ls, [parse_float(t["logsf"])]) <= LOGP_RTOL


def test_hypergeom_kernel_dense_sweep():
    """every k of a few (M, n, N): support edges, both branches, far tails."""
    from pyspade_b200.engine import hypergeom_logcdf_logsf
    for (M, n, N) in [(60, 25, 17), (1000, 990, 50), (2000, 1100, 300), (50000, 46000, 500),
                      (50000, 25001, 2000), (200000, 180000, 1000), (200000, 100003, 5000)]:
        k = np.arange(-1, min(n, N) + 2)
        lc, ls = hypergeom_logcdf_logsf(k, M, n, N)
        wc, ws = c_oracle.hypergeom_logcdf_logsf(k, M, np.full_like(k, n), np.full_like(k, N))
        assert max_rel_err(lc, wc) <= LOGP_RTOL, (M, n, N)
        assert max_rel_err(ls, ws) <= LOGP_RTOL, (M, n, N)


@pytest.mark.parametrize("mode", ["complement", "nc"])
def test_mid_size_vs_oracle(mode):
    """3000 cells x 900 genes, ragged sets incl. a single cell and C-1 cells."""
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C = 900, 3000
    counts = synth.counts_numpy(G, C, seed=4)
    rng = np.random.default_rng(8)
    sizes = [1, 2, 33, 150, 700, 1500, 2999, 64, 64]
    sets = [rng.choice(C, n, replace=False) for n in sizes]
    nc = rng.choice(C, 120, replace=False) if mode == "nc" else None
    with Engine.from_counts(counts) as eng:
        eng.set_background(nc)
        got = eng.run(sets, want_k=True)
        med, n = eng.gene_cutoffs()
        X = sp.csr_matrix((eng.cpm_values(), counts.indices, counts.indptr), shape=counts.shape)
    Xo = de_port.cpm_normalization_port(counts)
    assert np.array_equal(X.data, Xo.data)
    ora = de_port.de_tables(Xo, sets, nc, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    assert np.array_equal(med, ora["median"])
    assert np.array_equal(n, ora["n"])
    assert np.array_equal(got["k"], ora["k"])
    assert_tables_close(got, ora)


def test_background_switch_is_idempotent():
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    counts = synth.counts_numpy(300, 800, seed=2)
    sets = synth.deobs_sets(800, 5, lo=20, hi=200)
    nc = np.arange(0, 800, 9)
    with Engine.from_counts(counts) as eng:
        a = eng.run(sets, want_k=True)
        eng.set_background(nc)
        b = eng.run(sets, want_k=True)
        eng.set_background(None)
        c = eng.run(sets, want_k=True)
        eng.set_background(nc)
        d = eng.run(sets, want_k=True)
    for name in a:
        assert np.array_equal(a[name], c[name], equal_nan=True)
        assert np.array_equal(b[name], d[name], equal_nan=True)
    assert not np.array_equal(a["fc"], b["fc"], equal_nan=True)


def test_empty_and_degenerate_inputs():
    from pyspade_b200.engine import Engine, SpadeError
    X = sp.csr_matrix(np.array([[0, 0, 0, 0, 0], [1.5, 0, 2.5, 0, 0], [1, 1, 1, 1, 1]], dtype=np.float64))
    with Engine.from_cpm(X) as eng:
        out = eng.run([])                                   # no sets
        assert out["up"].shape == (0, 3)
        out = eng.run([np.array([], dtype=np.int64), np.array([0, 2])], want_k=True)
        # empty set: N = 0 -> nan tails (all([k,M,n,N]) guard); means are 0/0 = nan like the reference
        assert np.isnan(out["up"][0]).all() and np.isnan(out["down"][0]).all()
        ora = de_port.de_tables(X, [np.array([0, 2])])
        assert np.array_equal(out["k"][1], ora["k"][0])
        assert_tables_close({k: v[1:] for k, v in out.items() if k != "k"}, ora)
        with pytest.raises(IndexError):
            eng.run([np.array([0, 5])])
        with pytest.raises(ValueError):
            eng.run([np.array([1, 1])])
    # a matrix with no stored entries at all
    with Engine.from_cpm(sp.csr_matrix((4, 7), dtype=np.float64)) as eng:
        out = eng.run([np.array([1, 2, 3])], want_k=True)
        assert (out["k"] == 3).all() and (out["up"] == 0).all() and np.isneginf(out["down"]).all()
        assert (out["fc"] == 1.0).all() and (out["cpm"] == 0.0).all()


def test_negative_values_general_matrix():
    """The boundary takes any float64 matrix (perform_DE's input_array), not only CPM."""
    from pyspade_b200.engine import Engine
    rng = np.random.default_rng(12)
    dense = rng.normal(0, 1, (40, 90)) * (rng.random((40, 90)) < 0.7)
    dense[3] = -np.abs(dense[3]) - 0.1         # all negative: median < 0, zeros never low
    dense[4, :] = 0.0
    dense[4, :10] = -2.0
    X = sp.csr_matrix(dense)
    sets = [rng.choice(90, n, replace=False) for n in (5, 20, 45)]
    with Engine.from_cpm(X) as eng:
        got = eng.run(sets, want_k=True)
        med, n = eng.gene_cutoffs()
    ora = de_port.de_tables(X, sets)
    assert np.array_equal(med, ora["median"]) and np.array_equal(n, ora["n"])
    assert np.array_equal(got["k"], ora["k"])
    for name in ("up", "down"):
        assert max_rel_err(got[name], ora[name]) <= LOGP_RTOL


def test_sharding_is_bit_identical():
    """Sets split over several calls (what ranks do) give byte-identical tables."""
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    counts = synth.counts_numpy(500, 1200, seed=6)
    sets = synth.derand_draws(1200, 60, 16)
    with Engine.from_counts(counts) as eng:
        whole = eng.run(sets, want_k=True)
        parts = [eng.run(sets[r::4], want_k=True) for r in range(4)]
    for name in whole:
        for r in range(4):
            assert np.array_equal(whole[name][r::4], parts[r][name], equal_nan=True)


def _tables_equal(a, b):
    for name in a:
        assert np.array_equal(a[name], b[name], equal_nan=True), name


def test_tail_table_and_direct_paths_are_bit_identical():
    """Uniform-N batch (DErand shape): per-run tail table vs direct evaluation; includes
    extreme (non-random) sets whose k falls outside the table window."""
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C, N = 700, 4000, 150
    counts = synth.counts_numpy(G, C, seed=9)
    sets = synth.derand_draws(C, N, 40)
    depth = np.asarray(counts.sum(axis=0)).ravel()
    order = np.argsort(depth)
    sets += [order[:N], order[-N:], order[C // 2:C // 2 + N]]     # shallowest / deepest cells
    X = de_port.cpm_normalization_port(counts)
    ora = de_port.de_tables(X, sets, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    with Engine.from_counts(counts) as eng:
        res = {}
        for mode in (0, 1, 2):
            eng.set_option("tail_table", mode)
            res[mode] = eng.run(sets, want_k=True)
        assert eng.timing()["table_ms"] > 0.0
    _tables_equal(res[0], res[1])
    _tables_equal(res[0], res[2])
    assert np.array_equal(res[1]["k"], ora["k"])
    assert_tables_close(res[1], ora)


@pytest.mark.parametrize("mode", ["complement", "nc"])
def test_split_path_small_member_block(mode):
    """Sets larger than the member block go through global accumulators (split path)."""
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C = 400, 1500
    counts = synth.counts_numpy(G, C, seed=13)
    rng = np.random.default_rng(3)
    sets = [rng.choice(C, n, replace=False) for n in (5, 64, 65, 300, 1499, 0)]
    nc = rng.choice(C, 90, replace=False) if mode == "nc" else None
    with Engine.from_counts(counts) as eng:
        eng.set_background(nc)
        whole = eng.run(sets, want_k=True)
        eng.set_option("member_block", 64)
        split = eng.run(sets, want_k=True)
    _tables_equal(whole, split)
    X = de_port.cpm_normalization_port(counts)
    ora = de_port.de_tables(X, sets[:-1], nc, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    assert np.array_equal(split["k"][:-1], ora["k"])
    assert_tables_close({k: v[:-1] for k, v in split.items() if k != "k"}, ora)


def test_set_larger_than_default_member_block():
    """N > 32768 member cells: the count field of a warp accumulator would overflow, the
    split path must take over on its own."""
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C = 64, 40000
    counts = synth.counts_numpy(G, C, seed=5)
    counts = sp.csr_matrix(counts.toarray() + (np.arange(G)[:, None] % 4 == 0))   # some dense genes
    rng = np.random.default_rng(1)
    sets = [rng.choice(C, n, replace=False) for n in (36000, 100, 39999)]
    with Engine.from_counts(counts) as eng:
        got = eng.run(sets, want_k=True)
    X = de_port.cpm_normalization_port(counts)
    ora = de_port.de_tables(X, sets, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    assert np.array_equal(got["k"], ora["k"])
    assert_tables_close(got, ora)


@pytest.mark.parametrize("gp", [32, 96, 1024, 4096])
def test_partition_width_does_not_change_results(gp):
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C = 1300, 2500
    counts = synth.counts_numpy(G, C, seed=21)
    sets = synth.deobs_sets(C, 12, lo=10, hi=900)
    nc = np.arange(3, C, 17)
    with Engine.from_counts(counts) as eng:
        eng.set_background(nc)
        base = eng.run(sets, want_k=True)
        eng.set_option("genes_per_part", gp)
        assert eng.genes_per_part == min(gp, 1312)
        eng.set_background(nc)
        other = eng.run(sets, want_k=True)
    _tables_equal(base, other)


def test_device_outputs_match_host_outputs():
    """spade_run with device members / device tables (bench path) vs host tables."""
    import torch
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C = 2100, 3000
    counts = synth.counts_numpy(G, C, seed=31)
    sets = synth.derand_draws(C, 120, 70)
    members, offsets = synth.concat_sets(sets)
    with Engine.from_counts(counts) as eng:
        host = eng.run(sets, want_k=True)
        out = {k: torch.empty((len(sets), G), dtype=torch.float64, device="cuda") for k in ("up", "down", "fc", "cpm")}
        kd = torch.empty((len(sets), G), dtype=torch.int32, device="cuda")
        eng.run_device(torch.from_numpy(members).cuda(), offsets, out, k=kd)
        torch.cuda.synchronize()
    for name in out:
        assert np.array_equal(out[name].cpu().numpy(), host[name], equal_nan=True), name
    assert np.array_equal(kd.cpu().numpy(), host["k"])


def test_non_finite_values_are_refused():
    from pyspade_b200.engine import Engine, SpadeError
    X = sp.csr_matrix(np.array([[1.0, np.inf, 0.0], [0.0, 2.0, 3.0]]))
    with pytest.raises(SpadeError):
        Engine.from_cpm(X)


def test_layout_switches_do_not_change_results(monkeypatch):
    """Dense blocks (register accumulation of each partition's densest genes) and the
    bank-aware segment order are pure layout choices: tables are bit-identical without them."""
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C = 2300, 1800
    counts = synth.counts_numpy(G, C, seed=17)
    sets = synth.deobs_sets(C, 9, lo=5, hi=1200) + synth.derand_draws(C, 77, 5)
    nc = np.arange(1, C, 13)
    res = {}
    for dense, order in ((1, 1), (0, 1), (1, 0), (0, 0), (1, "two"), (0, "two"), (1, "hot"), (0, "hot")):
        monkeypatch.setenv("SPADE_DENSE_BLOCKS", str(dense))
        monkeypatch.setenv("SPADE_BANK_ORDER", "1" if order in ("two", "hot") else str(order))
        # a second accumulator for every gene ("two") or for the 64 hottest genes of a partition
        monkeypatch.setenv("SPADE_TWO_CHOICE", {"two": "1", "hot": "64"}.get(order, "0"))
        with Engine.from_counts(counts) as eng:
            a = eng.run(sets, want_k=True)
            eng.set_background(nc)
            b = eng.run(sets, want_k=True)
            if order in ("two", "hot"):                     # re-staging and other partition widths
                eng.set_background(None)
                _tables_equal(a, eng.run(sets, want_k=True))
                eng.set_option("genes_per_part", 96)
                _tables_equal(a, eng.run(sets, want_k=True))
        res[(dense, order)] = (a, b)
    monkeypatch.setenv("SPADE_TWO_CHOICE", "0")
    with Engine.from_counts(counts) as eng:
        eng.set_option("sort_members", 0)               # caller's member order
        res["unsorted"] = (eng.run(sets, want_k=True), None)
        eng.set_background(nc)
        res["unsorted"] = (res["unsorted"][0], eng.run(sets, want_k=True))
    for key in res:
        _tables_equal(res[(1, 1)][0], res[key][0])
        _tables_equal(res[(1, 1)][1], res[key][1])
    X = de_port.cpm_normalization_port(counts)
    ora = de_port.de_tables(X, sets, nc, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    assert np.array_equal(res[(1, 1)][1]["k"], ora["k"])
    assert_tables_close(res[(1, 1)][1], ora)


def test_kernel_forms_do_not_change_results(monkeypatch):
    """Device tables of a uniform batch come from two launches (k_de_fused leaves accumulator
    words, k_expand finishes them); SPADE_TWO_PHASE=0 keeps the finalize inside the test kernel
    and SPADE_TMA=1 stages the segments through the TMA unit (k_de_tma): all bit-identical to
    the host-output run, which always finishes its tests inside the kernel."""
    import torch
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C = 2300, 1800
    counts = synth.counts_numpy(G, C, seed=29)
    nc = np.arange(3, C, 11)
    batches = (synth.derand_draws(C, 77, 40), synth.deobs_sets(C, 9, lo=5, hi=1200))
    with Engine.from_counts(counts) as eng:
        want = [eng.run(sets, want_k=True) for sets in batches]
        eng.set_background(nc)
        want_nc = [eng.run(sets, want_k=True) for sets in batches]
    for two_phase, tma, two_choice, coop in ((1, 0, 0, 1), (0, 0, 0, 1), (1, 1, 0, 0), (0, 1, 0, 0), (1, 2, 1, 0),
                                             (0, 2, 1, 0), (1, 0, 112, 2), (0, 0, 112, 4), (1, 0, 112, 4)):
        monkeypatch.setenv("SPADE_TWO_PHASE", str(two_phase))
        monkeypatch.setenv("SPADE_TMA", str(tma))              # 1 = cp.async.bulk ring, 2 = cp.async ring
        monkeypatch.setenv("SPADE_TWO_CHOICE", str(two_choice))
        monkeypatch.setenv("SPADE_COOP", str(coop))            # warps per (partition, set) task, 0 = automatic
        with Engine.from_counts(counts) as eng:
            for bg, ref in ((None, want), (nc, want_nc)):
                eng.set_background(bg)
                for sets, w in zip(batches, ref):
                    members, offsets = synth.concat_sets(sets)
                    dm = torch.from_numpy(members).cuda()
                    out = {t: torch.empty((len(sets), G), dtype=torch.float64, device="cuda")
                           for t in ("up", "down", "fc", "cpm")}
                    kd = torch.empty((len(sets), G), dtype=torch.int32, device="cuda")
                    eng.run_device(dm, offsets, out, k=kd)
                    eng.check()
                    got = {t: v.cpu().numpy() for t, v in out.items()}
                    got["k"] = kd.cpu().numpy()
                    _tables_equal({t: w[t] for t in got}, got)


@pytest.mark.parametrize("mode", ["complement", "nc"])
def test_raw_form_and_expand_equal_direct_run(mode):
    """spade_run_raw + spade_expand (what a peer GPU sends / what rank 0 finishes) give the
    same bits as spade_run, uniform-N (tail table) and ragged (direct tails) batches."""
    import torch
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C = 1500, 2600
    counts = synth.counts_numpy(G, C, seed=23)
    nc = np.arange(2, C, 19) if mode == "nc" else None
    for sets in (synth.derand_draws(C, 90, 40), synth.deobs_sets(C, 7, lo=3, hi=700)):
        members, offsets = synth.concat_sets(sets)
        with Engine.from_counts(counts) as eng:
            eng.set_background(nc)
            want = eng.run(sets, want_k=True)
            dm = torch.from_numpy(members).cuda()
            raw = torch.zeros((len(sets), G + 5), dtype=torch.int64, device="cuda")       # padded rows
            eng.run_raw(dm, offsets, raw, row_stride=G + 5)
            out = {k: torch.full((len(sets), 2 * G), -7.0, dtype=torch.float64, device="cuda")
                   for k in ("up", "down", "fc", "cpm")}
            kd = torch.zeros((len(sets), 2 * G), dtype=torch.int32, device="cuda")
            # tables interleaved with another "rank": this rank owns columns G..2G of each row
            eng.expand(raw, [len(s) for s in sets], {k: v.data_ptr() + 8 * G for k, v in out.items()},
                       k=kd.data_ptr() + 4 * G, raw_row_stride=G + 5, row_stride=2 * G)
            torch.cuda.synchronize()
        for name in out:
            got = out[name].cpu().numpy()
            assert np.array_equal(got[:, G:], want[name], equal_nan=True), name
            assert (got[:, :G] == -7.0).all()
        assert np.array_equal(kd.cpu().numpy()[:, G:], want["k"])


def test_fully_dense_matrix_long_segments():
    """Every gene expressed in every cell: segments of ~1000 entries (the long-segment loop of
    the test kernel), medians > 0 everywhere, ties, G not a multiple of the partition width."""
    from pyspade_b200.engine import Engine
    rng = np.random.default_rng(33)
    G, C = 2091, 400
    dense = rng.integers(1, 6, size=(G, C)).astype(np.float64)      # many ties
    dense[5] = 3.0                                                  # constant gene: n = C
    X = sp.csr_matrix(dense)
    sets = [rng.choice(C, n, replace=False) for n in (1, 17, 200, 399)]
    nc = rng.choice(C, 57, replace=False)
    with Engine.from_cpm(X) as eng:
        for bg in (None, nc):
            eng.set_background(bg)
            got = eng.run(sets, want_k=True)
            med, n = eng.gene_cutoffs()
            ora = de_port.de_tables(X, sets, bg, tail_fn=c_oracle.hypergeom_logcdf_logsf)
            assert np.array_equal(med, ora["median"]) and np.array_equal(n, ora["n"])
            assert np.array_equal(got["k"], ora["k"])
            assert_tables_close(got, ora)


def test_strided_host_tables_interleave_like_two_ranks():
    """Engine.run(row_stride=...): two 'ranks' (two calls) interleave their rows in one host
    table, the layout shard.SharedHostTables gives every rank; plus the split path."""
    from pyspade_b200 import synth
    from pyspade_b200.engine import Engine
    G, C = 1100, 1500
    counts = synth.counts_numpy(G, C, seed=41)
    sets = synth.derand_draws(C, 70, 9)
    with Engine.from_counts(counts) as eng:
        whole = eng.run(sets, want_k=True)
        for block in (None, 32):
            if block:
                eng.set_option("member_block", block)           # sets of 70 cells take the split path
            full = {n: np.full((9, G), -3.0) for n in ("up", "down", "fc", "cpm")}
            full["k"] = np.full((9, G), -3, dtype=np.int32)
            for r in range(2):
                mine = sets[r::2]
                out = {n: a[r::2] for n, a in full.items()}
                eng.run(mine, out=out, want_k=True, row_stride=2 * G)
            for name in whole:
                assert np.array_equal(full[name], whole[name], equal_nan=True), (name, block)
