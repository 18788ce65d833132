"""Generate the golden fixtures by running the UNMODIFIED reference.

Run in the dev container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so the
pins are outputs of the reference's functions executed under this image's
numpy 2.3.5 / scipy 1.18.1 / pandas 3.0.2.  Nothing is copied from the reference;
it is imported from where it lies (``oracle/ref_loader.py``).

Files written next to this script:
  kat.json            hand-built genes A/B/Z at C=1000 through
                      utils.hypergeo_test_sparse / hypergeo_test_NC_sparse, pure
                      hypergeom tuples through scipy, the legacy RNG prefix
  small_case.npz      counts (48 genes x 241 and x 240 cells), reference CPM values,
                      reference perform_DE / perform_DE_NC tables for 8 cell sets
  hypergeom_grid.npz  (k, M, n, N) tuples with scipy logcdf / logsf
  cli_case.npz        tiny DEobs / DErand end-to-end inputs and every output file's
                      content as produced by DE_observe_cells / DE_random_cells
"""
import glob
import json
import os
import sys
import tempfile

import numpy as np
import pandas as pd
import scipy.sparse as sp
from scipy import io as sio
from scipy.stats import hypergeom

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

utils, DEobs, DErand = ref_loader.load()
NPROC = 4


def f(x):
    """json-safe float"""
    x = float(x)
    if np.isnan(x):
        return "nan"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    return x


# --------------------------------------------------------------------------- kat.json
def make_kat():
    C = 1000
    A = np.zeros(C)
    A[:600] = np.arange(600) + 1.0
    B = np.zeros(C)
    B[::7] = 3.0
    B[::35] = 9.0
    Z = np.zeros(C)
    genes = {"A": A, "B": B, "Z": Z}
    nc = list(range(700, 1000))
    cases = []
    for gname, members in [("A", list(range(80, 130))), ("A", list(range(590, 640))),
                           ("A", list(range(0, 1000, 10))), ("A", list(range(100, 150))),
                           ("B", list(range(0, 1000, 5))), ("Z", list(range(50))),
                           ("B", list(range(3, 1000, 7))), ("A", list(range(0, 600))),
                           ("A", [999]), ("A", [0])]:
        row = sp.csr_matrix(genes[gname][None, :])
        down, up, fc, cpm = utils.hypergeo_test_sparse(row, members, 0)
        d2, u2, fc2, cpm2 = utils.hypergeo_test_NC_sparse(row, members, nc, 0)
        cases.append(dict(gene=gname, members=members,
                          complement=dict(down=f(down), up=f(up), fc=f(fc), cpm=f(cpm)),
                          nc=dict(down=f(d2), up=f(u2), fc=f(fc2), cpm=f(cpm2))))
    tuples = [(880, 200000, 180000, 1000), (4300, 200000, 180000, 5000), (40, 1000, 990, 50),
              (50, 1000, 990, 50), (39, 1000, 990, 50), (1, 1000, 500, 50), (49, 1000, 500, 50),
              (25, 1000, 500, 50), (24, 1000, 500, 50), (500, 50000, 46000, 500),
              (455, 50000, 46000, 500), (1, 50000, 25000, 2), (2000, 200000, 100000, 5000),
              (2500, 200000, 100000, 5000), (3000, 200000, 100000, 5000)]
    hg = [dict(k=k, M=M, n=n, N=N, logcdf=f(hypergeom.logcdf(k, M, n, N)),
               logsf=f(hypergeom.logsf(k, M, n, N))) for (k, M, n, N) in tuples]
    np.random.seed(0)
    rng_prefix = [int(v) for v in np.random.choice(np.arange(500), 60, replace=False)[:6]]
    np.random.seed(0)
    rng_second = [int(v) for v in np.random.choice(np.arange(50000), 500, replace=False)[:4]]
    out = dict(C=C, nc=[700, 1000], genes=dict(A="x[j]=j+1 for j<600 else 0",
                                              B="x[::7]=3 then x[::35]=9", Z="all zero"),
               cases=cases, hypergeom=hg, rng_seed0_choice500_60=rng_prefix,
               rng_seed0_choice50000_500=rng_second,
               versions=dict(numpy=np.__version__, scipy=__import__("scipy").__version__,
                             pandas=pd.__version__))
    with open(os.path.join(HERE, "kat.json"), "w") as fh:
        json.dump(out, fh, indent=1)


# --------------------------------------------------------------------------- small_case.npz
def small_counts(G, C, seed):
    rng = np.random.default_rng(seed)
    mu = np.exp(rng.normal(-1.0, 2.2, size=G))
    s = np.exp(rng.normal(0.0, 0.35, size=C))
    x = rng.poisson(mu[:, None] * s[None, :]).astype(np.int32)
    x[0] = 0                                    # all-zero gene
    x[1] = 0
    x[1, : C // 2] = 2                          # exactly floor(C/2) expressing cells, equal counts
    x[2] = 1                                    # every cell expresses (ties everywhere)
    x[3] = 0
    x[3, : C // 2 + 1] = np.arange(C // 2 + 1) % 3 + 1   # just over half expressing
    x[4] = 0
    x[4, 5] = 7                                 # single expressing cell
    x[:, 7] += 1                                # no empty cell columns
    x[0] = 0
    return x


def run_reference_tables(counts, sets, nc_idx):
    G, C = counts.shape
    df = pd.DataFrame(counts, index=[f"g{i}" for i in range(G)],
                      columns=[f"c{i}" for i in range(C)]).astype(pd.SparseDtype("int32", 0))
    cpm = utils.cpm_normalization(df)                           # COO float64
    cpm_csr = sp.csr_matrix(cpm)
    cpm_csr.sort_indices()
    idx = np.arange(G)
    tabs = {m: np.zeros((4, len(sets), G)) for m in ("complement", "nc")}
    for s, members in enumerate(sets):
        members = list(int(v) for v in members)
        N, up, down, fc, cp = utils.perform_DE(members, cpm, idx, NPROC, np.zeros(G), np.zeros(G),
                                               np.ones(G), np.zeros(G))
        assert N == len(members)
        tabs["complement"][:, s] = np.stack([up, down, fc, cp])
        N, up, down, fc, cp = utils.perform_DE_NC(members, list(int(v) for v in nc_idx), cpm, idx,
                                                  NPROC, np.zeros(G), np.zeros(G), np.ones(G),
                                                  np.zeros(G))
        tabs["nc"][:, s] = np.stack([up, down, fc, cp])
    return cpm_csr, tabs


def make_small():
    out = {}
    for tag, C in (("odd", 241), ("even", 240)):
        G = 48
        counts = small_counts(G, C, seed=11 + C)
        rng = np.random.default_rng(5)
        sizes = [1, 2, 7, 30, 60, 120, 200, C - 1]
        sets = [np.sort(rng.choice(C, n, replace=False)) if i % 2 == 0 else rng.choice(C, n, replace=False)
                for i, n in enumerate(sizes)]
        nc_idx = rng.choice(C, 41, replace=False)
        cpm_csr, tabs = run_reference_tables(counts, sets, nc_idx)
        out[f"{tag}_counts"] = counts
        out[f"{tag}_cpm_indptr"] = cpm_csr.indptr.astype(np.int64)
        out[f"{tag}_cpm_indices"] = cpm_csr.indices.astype(np.int32)
        out[f"{tag}_cpm_data"] = cpm_csr.data
        out[f"{tag}_set_members"] = np.concatenate(sets).astype(np.int64)
        out[f"{tag}_set_offsets"] = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        out[f"{tag}_nc_idx"] = nc_idx.astype(np.int64)
        for mode in ("complement", "nc"):
            for j, name in enumerate(("up", "down", "fc", "cpm")):
                out[f"{tag}_{mode}_{name}"] = tabs[mode][j]
    np.savez_compressed(os.path.join(HERE, "small_case.npz"), **out)


# --------------------------------------------------------------------------- hypergeom_grid.npz
def make_grid():
    rng = np.random.default_rng(3)
    ks, Ms, ns, Ns = [], [], [], []
    for M in (30, 1000, 2000, 50000, 200000):
        cnt = 400
        n = rng.integers(1, M + 1, cnt)
        N = rng.integers(1, min(M, 5000) + 1, cnt)
        mean = n * N / M
        sd = np.sqrt(N * (n / M) * (1 - n / M)) + 1
        spread = rng.choice([0.3, 2.0, 8.0, 30.0], cnt)
        k = np.rint(mean + rng.normal(0, 1, cnt) * spread * sd).astype(np.int64)
        lo = np.maximum(0, N - (M - n)) - 1
        hi = np.minimum(n, N) + 1
        k = np.clip(k, np.maximum(lo, 0), hi)
        # sprinkle exact support edges
        k[:20] = np.maximum(0, N - (M - n))[:20]
        k[20:40] = np.minimum(n, N)[20:40]
        k[40:60] = np.maximum(np.minimum(n, N)[40:60] - 1, 0)
        ks.append(k); Ms.append(np.full(cnt, M)); ns.append(n); Ns.append(N)
    k = np.concatenate(ks); M = np.concatenate(Ms); n = np.concatenate(ns); N = np.concatenate(Ns)
    np.savez_compressed(os.path.join(HERE, "hypergeom_grid.npz"), k=k, M=M, n=n, N=N,
                        logcdf=hypergeom.logcdf(k, M, n, N), logsf=hypergeom.logsf(k, M, n, N))


# --------------------------------------------------------------------------- cli_case.npz
def cli_inputs():
    G, C, R = 30, 120, 12
    rng = np.random.default_rng(21)
    counts = small_counts(G, C, seed=77)
    genes = [f"GENE{i}" for i in range(G)]
    cells = [f"CELL{i:04d}-1" for i in range(C)]
    sg_names = [f"sg{i:02d}" for i in range(R)]
    sg = (rng.random((R, C)) < 0.12).astype(np.int64) * rng.integers(1, 9, (R, C))
    sg[10] = 0                                    # sgRNA present in the matrix but in no cell
    dict_lines = ["chr1:1000-1500\tsg00;sg01;sg02",
                  "chr2:2000-2500\tsg03;sgMISSING",
                  "chr3:3000-3500\tsgGONE1;sgGONE2",
                  "chr4:4000-4500\tsg10",
                  "chr5:5000-5500\tsg04;sg05;sg06;sg07",
                  "NonTarget\tsg08;sg09;sg11"]
    return counts, genes, cells, sg, sg_names, dict_lines


def read_outputs(out_dir):
    res = {}
    for path in sorted(glob.glob(os.path.join(out_dir, "*"))):
        name = os.path.basename(path)
        if name.endswith(".npy"):
            res["npy:" + name] = np.load(path)
        elif name.endswith(".txt"):
            with open(path) as fh:
                res["txt:" + name] = np.array(fh.read())
        else:
            m = sio.loadmat(path)["matrix"]
            if sp.issparse(m):
                m = m.tocsr()
                res["matsp:" + name + ":dense"] = np.asarray(m.todense())
                mask = np.zeros(m.shape, dtype=bool)
                coo = m.tocoo()
                mask[coo.row, coo.col] = True
                res["matsp:" + name + ":stored"] = mask
            else:
                res["mat:" + name] = np.asarray(m)
    return res


def make_cli():
    counts, genes, cells, sg, sg_names, dict_lines = cli_inputs()
    out = dict(counts=counts, sg=sg, genes=np.array(genes), cells=np.array(cells),
               sg_names=np.array(sg_names), dict_text=np.array("\n".join(dict_lines) + "\n"))
    with tempfile.TemporaryDirectory() as tmp:
        tpk = os.path.join(tmp, "trans.pkl")
        spk = os.path.join(tmp, "sgrna.pkl")
        dtx = os.path.join(tmp, "dict.txt")
        pd.DataFrame(counts, index=genes, columns=cells).to_pickle(tpk)
        pd.DataFrame(sg, index=sg_names, columns=cells).to_pickle(spk)
        with open(dtx, "w") as fh:
            fh.write("\n".join(dict_lines) + "\n")
        for bg in ("complement", "NonTarget"):
            od = os.path.join(tmp, f"obs_{bg}")
            os.makedirs(od)
            DEobs.DE_observe_cells(tpk, spk, dtx, od, num_processing=1, norm="cpm", background_cells=bg)
            for key, val in read_outputs(od).items():
                out[f"DEobs/{bg}/{key}"] = val
            for method in ("equal", "sgrna"):
                od = os.path.join(tmp, f"rand_{bg}_{method}") + "/"
                os.makedirs(od)
                np.random.seed(0)              # a fresh CLI process starts here (DErand.py:27)
                DErand.DE_random_cells(tpk, spk, dtx, 20, method, od, iteration=56,
                                       num_processing=1, norm="cpm", background_cells=bg)
                for key, val in read_outputs(od).items():
                    out[f"DErand/{bg}/{method}/{key}"] = val
    np.savez_compressed(os.path.join(HERE, "cli_case.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["kat", "small", "grid", "cli"]
    if "kat" in which:
        make_kat()
    if "small" in which:
        make_small()
    if "grid" in which:
        make_grid()
    if "cli" in which:
        make_cli()
    print("golden fixtures written to", HERE)


# --------------------------------------------------------------------------- cfg1_case.npz
def make_cfg1():
    """BASELINE.json configs[0]: DEobs on synthetic 2 000 cells x 5 000 genes, 10 regions of
    30..200 cells - the reference's own CPU-runnable case.  Inputs come from the seeded
    generators of pyspade_b200/synth.py (not stored); stored: the reference's perform_DE tables
    (complement background) and perform_DE_NC tables (NC = every 50th cell) for all genes."""
    from pyspade_b200 import synth
    G, C, R = 5000, 2000, 10
    counts = synth.counts_numpy(G, C, seed=0)
    sets = synth.deobs_sets(C, R, lo=30, hi=200, uniform_int=True)
    nc_idx = np.arange(0, C, 50)
    df = pd.DataFrame.sparse.from_spmatrix(counts.astype(np.int32))
    cpm = utils.cpm_normalization(df)
    idx = np.arange(G)
    out = {"nc_idx": nc_idx.astype(np.int64), "sizes": np.array([len(s) for s in sets])}
    tabs = {m: np.zeros((4, R, G)) for m in ("complement", "nc")}
    for s, members in enumerate(sets):
        members = [int(v) for v in members]
        N, up, down, fc, cp = utils.perform_DE(members, cpm, idx, os.cpu_count(), np.zeros(G), np.zeros(G),
                                               np.ones(G), np.zeros(G))
        tabs["complement"][:, s] = np.stack([up, down, fc, cp])
        N, up, down, fc, cp = utils.perform_DE_NC(members, [int(v) for v in nc_idx], cpm, idx, os.cpu_count(),
                                                  np.zeros(G), np.zeros(G), np.ones(G), np.zeros(G))
        tabs["nc"][:, s] = np.stack([up, down, fc, cp])
        print("region", s, "done", flush=True)
    for mode in tabs:
        for j, name in enumerate(("up", "down", "fc", "cpm")):
            out[f"{mode}_{name}"] = tabs[mode][j]
    np.savez_compressed(os.path.join(HERE, "cfg1_case.npz"), **out)


if __name__ == "__main__" and "cfg1" in sys.argv[1:]:
    make_cfg1()
    print("cfg1_case.npz written")
