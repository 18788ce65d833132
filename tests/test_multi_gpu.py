"""GPU tier, needs >= 2 GPUs (skipped otherwise): DEobs / DErand launched under torchrun with
2 ranks write the same files as the reference (and therefore as the 1-rank run)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from test_cli_gpu import _compare, _read_outputs, _write_inputs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(nproc, args, port, gather="genes"):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), SPADE_GATHER=gather)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "pyspade_b200"] + args
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("bg,gather", [("complement", "genes"), ("NonTarget", "genes"), ("complement", "host"),
                                       ("NonTarget", "host"), ("complement", "nccl")])
def test_two_rank_cli_matches_reference_files(tmp_path, cli_case, bg, gather):
    tpk, spk, dtx = _write_inputs(str(tmp_path), cli_case)
    od = str(tmp_path / "obs")
    os.makedirs(od)
    _torchrun(2, ["DEobs", "-t", tpk, "-s", spk, "-d", dtx, "-o", od, "-b", bg], 29611, gather)
    _compare(_read_outputs(od), cli_case, f"DEobs/{bg}/")
    od = str(tmp_path / "rand") + "/"
    os.makedirs(od)
    _torchrun(2, ["DErand", "-t", tpk, "-s", spk, "-d", dtx, "-o", od, "-b", bg, "-m", "20", "-i", "56",
                  "-a", "equal"], 29612, gather)
    _compare(_read_outputs(od), cli_case, f"DErand/{bg}/equal/")


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_gene_sharded_exchange_bit_identical():
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(_ngpu(), 4)}",
           "--master-addr", "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "tests", "_exchange_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
