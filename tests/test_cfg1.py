"""BASELINE.json configs[0] - DEobs on synthetic 2 000 cells x 5 000 genes, 10 regions of
30..200 cells, the reference's own CPU-runnable case - against the tables the UNMODIFIED
reference produced for it (tests/golden/cfg1_case.npz, written by
``python tests/golden/make_golden.py cfg1``; inputs are regenerated from the seeded
generators, only the reference's outputs are stored)."""
import os

import numpy as np
import pytest

from helpers import assert_tables_close
from oracle import c_oracle, de_port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg1_case.npz")
G, C, R = 5000, 2000, 10


@pytest.fixture(scope="module")
def cfg1():
    from pyspade_b200 import synth
    z = np.load(GOLDEN)
    counts = synth.counts_numpy(G, C, seed=0)
    sets = synth.deobs_sets(C, R, lo=30, hi=200, uniform_int=True)
    assert [len(s) for s in sets] == list(z["sizes"])
    want = {m: {n: z[f"{m}_{n}"] for n in ("up", "down", "fc", "cpm")} for m in ("complement", "nc")}
    return counts, sets, z["nc_idx"], want


@pytest.mark.parametrize("mode", ["complement", "nc"])
def test_oracle_reproduces_reference_at_cfg1(cfg1, mode):
    counts, sets, nc, want = cfg1
    X = de_port.cpm_normalization_port(counts)
    ora = de_port.de_tables(X, sets, nc if mode == "nc" else None, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    assert_tables_close(ora, want[mode])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["complement", "nc"])
def test_engine_reproduces_reference_at_cfg1(cfg1, mode):
    from pyspade_b200.engine import Engine
    counts, sets, nc, want = cfg1
    with Engine.from_counts(counts) as eng:
        eng.set_background(nc if mode == "nc" else None)
        got = eng.run(sets, want_k=True)
    assert_tables_close(got, want[mode])
    X = de_port.cpm_normalization_port(counts)
    ora = de_port.de_tables(X, sets, nc if mode == "nc" else None, tail_fn=c_oracle.hypergeom_logcdf_logsf)
    assert np.array_equal(got["k"], ora["k"])
