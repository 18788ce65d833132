"""CPU tier: pin the oracle against the golden fixtures produced by the unmodified
reference (tests/golden/make_golden.py) and against live scipy."""
import math
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import assert_tables_close, max_rel_err, parse_float, sets_from_packed
from oracle import c_oracle, de_port, hypergeom_port, ref_loader


def _kat_genes(C):
    A = np.zeros(C); A[:600] = np.arange(600) + 1.0
    B = np.zeros(C); B[::7] = 3.0; B[::35] = 9.0
    Z = np.zeros(C)
    return {"A": A, "B": B, "Z": Z}


def test_kat_per_gene_port(kat):
    genes = _kat_genes(kat["C"])
    nc = np.arange(*kat["nc"])
    for case in kat["cases"]:
        row = genes[case["gene"]]
        for mode in ("complement", "nc"):
            want = {k: parse_float(v) for k, v in case[mode].items()}
            if mode == "complement":
                d, u, fc, cpm = de_port.gene_test(row, case["members"])
            else:
                d, u, fc, cpm = de_port.gene_test_nc(row, case["members"], nc)
            got = dict(down=d, up=u, fc=fc, cpm=cpm)
            for name in ("down", "up", "fc", "cpm"):
                assert max_rel_err([got[name]], [want[name]]) <= 1e-12, (case["gene"], mode, name)


def test_kat_survey_values(kat):
    """The fixture itself must reproduce the numbers recorded in SURVEY.md section 8-c."""
    first = kat["cases"][0]["complement"]
    assert first["down"] == pytest.approx(-0.10045204690821853, rel=1e-12)
    assert first["up"] == pytest.approx(-2.3478804271089664, rel=1e-12)
    assert first["fc"] == pytest.approx(0.5726556764523566, rel=1e-12)
    assert first["cpm"] == pytest.approx(105.5, rel=1e-12)
    nan_case = kat["cases"][3]["complement"]
    assert nan_case["down"] == "nan" and nan_case["up"] == "nan"
    assert kat["rng_seed0_choice500_60"] == [90, 254, 283, 445, 461, 15]


def test_kat_hypergeom_ports(kat):
    for t in kat["hypergeom"]:
        k, M, n, N = t["k"], t["M"], t["n"], t["N"]
        want_c, want_s = parse_float(t["logcdf"]), parse_float(t["logsf"])
        assert max_rel_err([hypergeom_port.logcdf(k, M, n, N)], [want_c]) <= 1e-8
        assert max_rel_err([hypergeom_port.logsf(k, M, n, N)], [want_s]) <= 1e-8
        lc, ls = c_oracle.hypergeom_logcdf_logsf(np.array([k]), M, np.array([n]), np.array([N]))
        assert max_rel_err(lc, [want_c]) <= 1e-8
        assert max_rel_err(ls, [want_s]) <= 1e-8


def test_grid_c_oracle(hypergeom_grid):
    g = hypergeom_grid
    for M in np.unique(g["M"]):
        sel = g["M"] == M
        lc, ls = c_oracle.hypergeom_logcdf_logsf(g["k"][sel], int(M), g["n"][sel], g["N"][sel])
        assert max_rel_err(lc, g["logcdf"][sel]) <= 1e-8
        assert max_rel_err(ls, g["logsf"][sel]) <= 1e-8


def test_grid_python_port_sample(hypergeom_grid):
    g = hypergeom_grid
    rng = np.random.default_rng(0)
    for i in rng.choice(g["k"].shape[0], 120, replace=False):
        k, M, n, N = (int(g[x][i]) for x in ("k", "M", "n", "N"))
        assert max_rel_err([hypergeom_port.logcdf(k, M, n, N)], [g["logcdf"][i]]) <= 1e-8
        assert max_rel_err([hypergeom_port.logsf(k, M, n, N)], [g["logsf"][i]]) <= 1e-8


def test_live_scipy_matches_fixture(hypergeom_grid):
    """scipy on this box is the same dependency the fixture was generated with."""
    from scipy.stats import hypergeom
    g = hypergeom_grid
    sel = slice(0, 300)
    assert max_rel_err(hypergeom.logcdf(g["k"][sel], g["M"][sel], g["n"][sel], g["N"][sel]),
                       g["logcdf"][sel]) <= 1e-12


@pytest.mark.parametrize("tag", ["odd", "even"])
def test_cpm_port_bit_identical(small_case, tag):
    z = small_case
    X = de_port.cpm_normalization_port(sp.csr_matrix(z[f"{tag}_counts"]))
    X.sort_indices()
    assert np.array_equal(X.indptr, z[f"{tag}_cpm_indptr"])
    assert np.array_equal(X.indices, z[f"{tag}_cpm_indices"])
    assert np.array_equal(X.data, z[f"{tag}_cpm_data"])        # bit-identical


@pytest.mark.parametrize("tag", ["odd", "even"])
@pytest.mark.parametrize("mode", ["complement", "nc"])
def test_all_genes_port_vs_reference(small_case, tag, mode):
    z = small_case
    X = de_port.cpm_normalization_port(sp.csr_matrix(z[f"{tag}_counts"]))
    sets = sets_from_packed(z[f"{tag}_set_members"], z[f"{tag}_set_offsets"])
    nc = z[f"{tag}_nc_idx"] if mode == "nc" else None
    want = {name: z[f"{tag}_{mode}_{name}"] for name in ("up", "down", "fc", "cpm")}
    got = de_port.de_tables(X, sets, nc)
    assert_tables_close(got, want)
    got_c = de_port.de_tables(X, sets, nc, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    assert_tables_close(got_c, want)


@pytest.mark.parametrize("mode", ["complement", "nc"])
def test_per_gene_port_vs_reference(small_case, mode):
    z = small_case
    tag = "odd"
    X = de_port.cpm_normalization_port(sp.csr_matrix(z[f"{tag}_counts"]))
    G = X.shape[0]
    sets = sets_from_packed(z[f"{tag}_set_members"], z[f"{tag}_set_offsets"])
    for s in (0, 3, 7):
        args = (np.arange(G), 1, np.zeros(G), np.zeros(G), np.ones(G), np.zeros(G))
        if mode == "nc":
            N, up, down, fc, cpm = de_port.perform_DE_NC_port(sets[s], z[f"{tag}_nc_idx"], X, *args)
        else:
            N, up, down, fc, cpm = de_port.perform_DE_port(sets[s], X, *args)
        assert N == len(sets[s])
        want = {name: z[f"{tag}_{mode}_{name}"][s] for name in ("up", "down", "fc", "cpm")}
        assert_tables_close(dict(up=up, down=down, fc=fc, cpm=cpm), want)


def test_per_gene_port_pool_matches_serial(small_case):
    z = small_case
    X = de_port.cpm_normalization_port(sp.csr_matrix(z["even_counts"]))
    G = X.shape[0]
    members = sets_from_packed(z["even_set_members"], z["even_set_offsets"])[4]
    a = de_port.perform_DE_port(members, X, np.arange(G), 1, np.zeros(G), np.zeros(G), np.ones(G), np.zeros(G))
    b = de_port.perform_DE_port(members, X, np.arange(G), 2, np.zeros(G), np.zeros(G), np.ones(G), np.zeros(G))
    for x, y in zip(a[1:], b[1:]):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the dev container")
def test_live_reference_agrees_with_port():
    utils, _, _ = ref_loader.load()
    rng = np.random.default_rng(9)
    C, G = 180, 12
    counts = rng.poisson(np.exp(rng.normal(-1, 2, (G, 1))) * np.ones((1, C))).astype(np.int32)
    counts[:, 3] += 1
    X = de_port.cpm_normalization_port(sp.csr_matrix(counts))
    members = list(int(v) for v in rng.choice(C, 25, replace=False))
    nc = list(int(v) for v in rng.choice(C, 30, replace=False))
    for g in range(G):
        row = X[g]
        want = utils.hypergeo_test_sparse(row, members, g)
        got = de_port.gene_test(row.toarray().ravel(), members)
        want_nc = utils.hypergeo_test_NC_sparse(row, members, nc, g)
        got_nc = de_port.gene_test_nc(row.toarray().ravel(), members, nc)
        for a, b in zip(got + got_nc, want + want_nc):
            assert max_rel_err([a], [b]) <= 1e-12
