"""CPU tier: host logic of the reference-facing layer (no compute calls, no GPU):
command-line flags, file formats against the reference's own outputs (golden cli_case),
the RNG call sequence of DErand, the sharding plan and the world_size-2 gather over gloo."""
import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp
from scipy import io as sio

from pyspade_b200 import io as spio
from pyspade_b200 import parsers, shard


# --------------------------------------------------------------------------- flags
def test_deobs_flags_match_reference_parser():
    a = parsers.parse_args(["DEobs", "-t", "t.pkl", "-s", "s.pkl", "-d", "d.txt", "-o", "out"])
    assert (a.command, a.transcriptome_df, a.input_sgrna, a.dict, a.output_dir) == \
        ("DEobs", "t.pkl", "s.pkl", "d.txt", "out")
    assert (a.threads, a.norm_method, a.bg) == (1, "cpm", "complement")      # parsers.py:199-219
    a = parsers.parse_args(["DEobs", "--transcriptome_df", "t", "--sgrna", "s", "-dict", "d",
                            "--threads", "8", "--norm", "cpm", "--bg", "NonTarget", "--output_dir", "o"])
    assert (a.threads, a.bg, a.dict) == (8, "NonTarget", "d")


def test_derand_flags_match_reference_parser():
    a = parsers.parse_args(["DErand", "-t", "t", "-s", "s", "-d", "d", "-m", "500", "-a", "equal",
                            "-o", "out/"])
    assert (a.num, a.iteration, a.randomization_method, a.bg, a.norm_method, a.threads) == \
        (500, 1000, "equal", "complement", "cpm", 1)                         # parsers.py:262-304
    a = parsers.parse_args(["DErand", "-t", "t", "-s", "s", "-d", "d", "--num", "20", "--iteration", "56",
                            "--randomization_method", "sgrna", "-b", "NonTarget", "-o", "out/"])
    assert (a.num, a.iteration, a.randomization_method, a.bg) == (20, 56, "sgrna", "NonTarget")
    with pytest.raises(SystemExit):
        parsers.parse_args(["DErand", "-t", "t", "-s", "s", "-d", "d", "-o", "out/"])   # -m, -a required


def test_version_flag(capsys):
    with pytest.raises(SystemExit):
        parsers.parse_args(["--version"])
    assert "version" in capsys.readouterr().out


# --------------------------------------------------------------------------- formats
def test_deobs_vector_roundtrip_like_reference(tmp_path, cli_case):
    want = cli_case["DEobs/complement/mat:chr1:1000-1500-up_log-pval"]
    path = str(tmp_path / "chr1:1000-1500-up_log-pval")          # no .mat extension, ':' and '-' in names
    spio.save_vector(path, want[0])
    assert os.path.isfile(path) and not os.path.isfile(path + ".mat")
    got = sio.loadmat(path)["matrix"]
    assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True)


def test_derand_table_roundtrip_like_reference(tmp_path, cli_case):
    dense = cli_case["DErand/complement/equal/matsp:20-up_log-pval:dense"]
    stored = cli_case["DErand/complement/equal/matsp:20-up_log-pval:stored"]
    path = str(tmp_path / "20-up_log-pval")
    spio.save_draw_table(path, dense)
    m = sio.loadmat(path)["matrix"]
    assert sp.issparse(m)
    coo = m.tocoo()
    mask = np.zeros(m.shape, dtype=bool)
    mask[coo.row, coo.col] = True
    assert np.array_equal(mask, stored)                         # exact zeros dropped, nan/-inf kept
    assert np.array_equal(np.asarray(m.todense()), dense, equal_nan=True)


@pytest.mark.parametrize("method", ["equal", "sgrna"])
def test_draw_sequence_and_cell_id_text(tmp_path, cli_case, method):
    """DErand.py:27,195-199,224-228: seed 0, one np.random.choice per draw, text listing."""
    sg = cli_case["sg"]
    c = sg.shape[1]
    state = np.random.get_state()
    try:
        np.random.seed(0)
        if method == "sgrna":
            per_cell = (sg > 0).sum(axis=0)
            w = np.array(list(per_cell / per_cell.sum()))
            w = w / np.sum(w)
            draws = [np.random.choice(np.arange(c), size=20, replace=False, p=w) for _ in range(56)]
        else:
            draws = [np.random.choice(np.arange(c), size=20, replace=False) for _ in range(56)]
    finally:
        np.random.set_state(state)
    path = str(tmp_path / "ids.txt")
    spio.write_cell_ids(path, draws)
    with open(path) as fh:
        assert fh.read() == str(cli_case[f"DErand/complement/{method}/txt:20-rand_cell_ID.txt"])


def test_frame_to_csr_dense_and_sparse_frames(cli_case):
    counts = cli_case["counts"]
    df = pd.DataFrame(counts, index=list(cli_case["genes"]), columns=list(cli_case["cells"]))
    a = spio.frame_to_csr(df)
    b = spio.frame_to_csr(df.astype(pd.SparseDtype("int32", 0)))
    assert a.dtype == np.int32 and np.array_equal(a.toarray(), counts) and np.array_equal(b.toarray(), counts)
    with pytest.raises(ValueError):
        spio.read_frame("matrix.csv")


def test_membership_lookup_matches_reference_semantics(cli_case):
    from pyspade_b200.utils import find_all_sgrna_cells, region_cell_sets
    sg = cli_case["sg"]
    df = pd.DataFrame(sg, index=list(cli_case["sg_names"]), columns=list(cli_case["cells"])) > 0
    for names in (["sg00", "sg01", "sg02"], ["sg03"], ["sg10"]):
        rows = [list(cli_case["sg_names"]).index(s) for s in names]
        want = sorted(np.flatnonzero((sg[rows] > 0).sum(axis=0) > 0).tolist())
        assert find_all_sgrna_cells(df, names) == want
    with pytest.raises(KeyError):
        find_all_sgrna_cells(df, ["sgMISSING"])
    d = {"r1": ["sg00", "sg01", "sg02"], "r2": ["sg03", "sgMISSING"], "r3": ["sgGONE1", "sgGONE2"],
         "r4": ["sg10"], "NT": ["sg08"]}
    regions = region_cell_sets(df, d, skip=("NT",))
    assert [k for k, _ in regions] == ["r1", "r2"]              # r3: no sgRNA left, r4: no cell
    assert d["r2"] == ["sg03"]                                  # dict rewritten like DEobs.py:181-182
    # region sizes recorded in the reference's file names
    assert len(regions[0][1]) == 39 and len(regions[1][1]) == 15


@pytest.mark.parametrize("case", ["complement/equal", "NonTarget/sgrna"])
def test_gamma_tail_reproduces_reference_parameters(cli_case, case):
    """DErand.py:250-292 on the reference's own tables gives the reference's own A/B/C."""
    from pyspade_b200 import DErand
    for which, names in (("up", "Up"), ("down", "Down")):
        table = cli_case[f"DErand/{case}/matsp:20-{which}_log-pval:dense"]
        A, B, C = DErand.gamma_parameters(table, which, processes=1)
        for got, letter in ((A, "A"), (B, "B"), (C, "C")):
            want = cli_case[f"DErand/{case}/npy:{names}_dist_gamma-20-{letter}.npy"]
            assert np.array_equal(got, want, equal_nan=True), (which, letter)
    cpm = cli_case[f"DErand/{case}/matsp:20-cpm:dense"]
    assert np.array_equal(np.mean(cpm, axis=0), cli_case[f"DErand/{case}/npy:Cpm_mean-20.npy"])


def test_gamma_tail_worker_pool_equals_serial(cli_case):
    from pyspade_b200 import DErand
    table = np.tile(cli_case["DErand/complement/equal/matsp:20-up_log-pval:dense"], (1, 3))   # 90 genes
    pooled = DErand.gamma_parameters(table, "up", processes=2, min_pool_work=0)
    serial = DErand.gamma_parameters(table, "up", processes=1)
    for x, y in zip(pooled, serial):
        assert x.shape == (90,) and np.array_equal(x, y, equal_nan=True)


# --------------------------------------------------------------------------- sharding
def test_plan_round_robin_and_balanced():
    p = shard.plan([500] * 10, 4)
    assert [list(a) for a in p] == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    assert [list(a) for a in shard.plan([3, 4], 1)] == [[0, 1]]
    sizes = [2000, 50, 60, 1900, 70, 1000, 900, 80]
    p = shard.plan(sizes, 2, balance=True)
    assert sorted(np.concatenate(p).tolist()) == list(range(8))
    loads = [sum(sizes[i] for i in a) for a in p]
    assert abs(loads[0] - loads[1]) <= 100
    assert [list(a) for a in shard.plan(sizes, 2, balance=True)] == [list(a) for a in p]   # deterministic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gather_worker(rank, world, port, sizes, balance, n_cols, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assignment = shard.plan(sizes, world, balance=balance)
        mine = assignment[rank]
        # row s of the global table is filled with s + column/1000: position encodes identity
        local = {"up": torch.tensor(mine[:, None] + np.arange(n_cols)[None, :] / 1000.0, dtype=torch.float64),
                 "k": torch.tensor(mine[:, None] * 7 + np.arange(n_cols)[None, :], dtype=torch.int32)}
        full = shard.gather_rows(local, assignment, rank, world, n_cols)
        if rank == 0:
            q.put({k: v.numpy() for k, v in full.items()})
        else:
            assert full is None
            q.put(None)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balance", [False, True])
def test_gather_rows_world_size_2_gloo(balance):
    """The N>1 exchange step on CPU: ragged per-rank row counts, rows back in global order."""
    import torch.multiprocessing as mp
    sizes = [30, 500, 20, 700, 10, 40, 900]
    n_cols, world = 17, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, sizes, balance, n_cols, q))
             for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = next(g for g in got if g is not None)
    S = len(sizes)
    assert np.array_equal(full["up"], np.arange(S)[:, None] + np.arange(n_cols)[None, :] / 1000.0)
    assert np.array_equal(full["k"], np.arange(S)[:, None] * 7 + np.arange(n_cols)[None, :])


def _shm_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S, G = 7, 11                                              # 4 rows for rank 0, 3 for rank 1
        t = shard.SharedHostTables(("up", "k"), S, G, rank, world, pin=False)
        mine = np.arange(rank, S, world)
        t.rows_view("up", rank, world, len(mine))[:] = mine[:, None] + np.arange(G)[None, :] / 100.0
        t.rows_view("k", rank, world, len(mine))[:] = mine[:, None] * 3 + np.arange(G)[None, :]
        assert t.rows_view("up", rank, world, len(mine)).strides == (8 * world * G, 8)
        dist.barrier()
        if rank == 0:
            q.put({"up": np.array(t.arrays["up"]), "k": np.array(t.arrays["k"])})
        else:
            q.put(None)
        dist.barrier()
        t.close()
    finally:
        dist.destroy_process_group()


def test_shared_host_tables_world_size_2_gloo():
    """Rows written by two processes into the node-shared host tables, read back by rank 0."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = next(g for g in got if g is not None)
    assert np.array_equal(full["up"], np.arange(7)[:, None] + np.arange(11)[None, :] / 100.0)
    assert np.array_equal(full["k"], np.arange(7)[:, None] * 3 + np.arange(11)[None, :])


def test_mat_writers_load_like_scipy_written_files(tmp_path):
    """The direct MAT-v5 writers against scipy.io.savemat: loadmat gives identical arrays,
    identical sparse structure (stored nan / -inf, dropped +-0.0), also for empty tables."""
    rng = np.random.default_rng(4)
    vec = rng.normal(size=301)
    vec[[0, 5]] = [np.nan, -np.inf]
    vec[7] = 0.0
    spio.save_vector(str(tmp_path / "a-b:c-up_log-pval"), vec)
    sio.savemat(str(tmp_path / "ref_vec"), {"matrix": vec})
    a, b = sio.loadmat(str(tmp_path / "a-b:c-up_log-pval"))["matrix"], sio.loadmat(str(tmp_path / "ref_vec"))["matrix"]
    assert a.shape == b.shape == (1, 301) and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)
    for shape in ((56, 30), (3, 1), (1, 7), (40, 33)):
        t = rng.normal(size=shape)
        t[rng.random(shape) < 0.4] = 0.0
        t[rng.random(shape) < 0.05] = -0.0
        t[rng.random(shape) < 0.05] = np.nan
        t[rng.random(shape) < 0.05] = -np.inf
        for tab, tag in ((t, "x"), (np.zeros(shape), "z")):
            spio.save_draw_table(str(tmp_path / f"mine_{tag}"), tab)
            sio.savemat(str(tmp_path / f"ref_{tag}"), {"matrix": sp.csr_matrix(tab)})
            m, r = sio.loadmat(str(tmp_path / f"mine_{tag}"))["matrix"], sio.loadmat(str(tmp_path / f"ref_{tag}"))["matrix"]
            assert sp.issparse(m) and m.shape == r.shape and m.dtype == r.dtype and m.format == r.format
            m, r = m.tocsc(), r.tocsc()
            assert np.array_equal(m.indptr, r.indptr) and np.array_equal(m.indices, r.indices)
            assert np.array_equal(m.data, r.data, equal_nan=True)
    spio.save_draw_tables([(str(tmp_path / "p1"), t), (str(tmp_path / "p2"), -t)])
    assert np.array_equal(np.asarray(sio.loadmat(str(tmp_path / "p2"))["matrix"].todense()), -t + 0.0, equal_nan=True)


def test_gene_blocks_tile_all_genes_on_partition_boundaries():
    from pyspade_b200 import shard
    for G, Gp in ((33000, 1024), (1000, 1024), (5000, 96), (64, 32), (0, 1024)):
        for P in (1, 2, 3, 4, 8, 16):
            blocks = shard.gene_blocks(G, Gp, P)
            assert len(blocks) == P and blocks[0][0] == 0 and blocks[-1][1] == G
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and a <= b
            assert all(a % Gp == 0 for a, b in blocks if b > a)
            widths = [b - a for a, b in blocks]
            assert max(widths) - min(widths) <= Gp                 # as even as whole partitions allow

def _slices_worker(rank, world, port, path, q):
    """One 'rank' of the gene-sharded DErand finish on the CPU: it owns a block of gene columns of
    the table, builds their payload, learns the others' entry counts, and fills its slice of the
    one shared file (what ``DErand.finish_draws`` does with GPU-built payloads)."""
    import torch.distributed as dist
    from pyspade_b200 import io as spio
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)                       # the same table on every rank
        S, G = 40, 23
        t = rng.normal(size=(S, G))
        t[rng.random((S, G)) < 0.5] = 0.0
        t[rng.random((S, G)) < 0.05] = np.nan
        t[rng.random((S, G)) < 0.05] = -np.inf
        blocks = shard.gene_blocks(G, 8, world)              # partitions of 8 genes
        g0, g1 = blocks[rank]
        jc, ir, pr = spio.dense_payload(t[:, g0:g1])
        counts = [None] * world
        dist.all_gather_object(counts, (int(jc[-1]), jc))
        layout = spio.SparseMatLayout(S, G, sum(c[0] for c in counts))
        if rank == 0:
            full, base = [np.zeros(1, dtype=np.int64)], 0
            for n, j in counts:
                full.append(j[1:].astype(np.int64) + base)
                base += n
            fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
            layout.write_frame(fd, np.concatenate(full))
            os.close(fd)
        dist.barrier()
        if jc[-1] > 0:
            fd = os.open(path, os.O_RDWR)
            layout.write_entries(fd, sum(c[0] for c in counts[:rank]), ir, pr)
            os.close(fd)
        dist.barrier()
        # one rank failing makes every rank raise
        try:
            shard._raise_together(ValueError("boom") if rank == 1 else None, rank)
            q.put("no error")
        except ValueError:
            q.put("ValueError" if rank == 1 else "wrong")
        except RuntimeError:
            q.put("RuntimeError" if rank != 1 else "wrong")
        if rank == 0:
            q.put(t)
    finally:
        dist.destroy_process_group()


def test_file_slices_and_joint_failure_world_size_2_gloo(tmp_path):
    import scipy.sparse as sp
    import torch.multiprocessing as mp
    from scipy import io as sio
    world, path = 2, str(tmp_path / "20-up_log-pval")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slices_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world + 1)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(g for g in got if isinstance(g, str)) == ["RuntimeError", "ValueError"]
    t = next(g for g in got if not isinstance(g, str))
    a = sio.loadmat(path)["matrix"].tocsc()
    b = sp.csc_matrix(sp.csr_matrix(t))
    assert a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    assert np.array_equal(a.data.view(np.int64), b.data.view(np.int64))
