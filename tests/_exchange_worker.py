"""Worker of tests/test_multi_gpu.py::test_gene_sharded_exchange (run under torchrun, one rank
per GPU): the gene-sharded exchange (``shard.run_gene_sharded``: raw words routed to the owner
of each gene block from inside the test kernel, finished there) must give, on every rank, the
bytes of the single-GPU tables for its gene block - DErand-shaped (round-robin, uniform N,
tail table) and DEobs-shaped (balanced, ragged) batches, twice over the same buffers."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pyspade_b200 import shard, synth                      # noqa: E402
from pyspade_b200.engine import Engine                     # noqa: E402


def main():
    rank, local_rank, world = shard.world()
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    G, C = 2500, 3000
    counts = synth.counts_numpy(G, C, seed=11)
    eng = Engine.from_counts(counts, device=local_rank)
    eng.set_option("genes_per_part", 512)                  # 5 gene partitions over the ranks
    nc = np.arange(1, C, 23)
    cases = [("derand", synth.derand_draws(C, 80, 37), False, None),
             ("deobs", synth.deobs_sets(C, 11, lo=5, hi=900), True, None),
             ("derand-nc", synth.derand_draws(C, 50, 64), False, nc)]
    for name, sets, balance, bg in cases:
        eng.set_background(bg)
        want = eng.run(sets, want_k=True, tables=("up", "down", "fc", "cpm"))
        ex = None
        for rep in range(2):
            ex, order = shard.run_gene_sharded(eng, sets, tables=("up", "down", "fc", "cpm"), balance=balance,
                                               want_k=True, exchange=ex)
            torch.cuda.synchronize()
            g0, g1 = ex.gene_begin, ex.gene_end
            for t in ("up", "down", "fc", "cpm"):
                got = ex.tables[t][:len(sets), :g1 - g0].cpu().numpy()
                ref = want[t][order][:, g0:g1]
                assert np.array_equal(got.view(np.int64), ref.view(np.int64)), (name, t, rank, rep)
            assert np.array_equal(ex.k[:len(sets), :g1 - g0].cpu().numpy(), want["k"][order][:, g0:g1]), (name, rank)
        ex.close()
    # a bad member list on ONE rank makes every rank raise together
    # (round-robin: set 0 is tested by rank 0, whose copy of it holds a cell index out of range)
    bad = [np.array([1, 2, C + 7] if rank == 0 else [1, 2, 3]), np.array([4, 5, 6])]
    try:
        shard.run_gene_sharded(eng, bad, tables=("up",))
        raise SystemExit("expected an error")
    except (IndexError, RuntimeError):
        pass
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
