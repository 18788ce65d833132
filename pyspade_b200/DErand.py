"""``pySpade DErand``: DE of random cell draws, the empirical background (mirror of
``pySpade/DErand.py:30-304``).  The draws come from the same legacy numpy RNG calls in the
same order, so the selected cells are identical to the reference's; all draws are then
tested in one batch on the GPU (sharded by draw when launched under torchrun), and the
final files are written once instead of after every draw (``DErand.py:224-248`` rewrites
them each iteration; the last write is what stays on disk).
"""
import multiprocessing
import os
import sys

import numpy as np
import scipy.stats as stats

from . import io as spio
from . import shard
from .engine import Engine
from .utils import find_all_sgrna_cells, get_logger, get_num_processes, read_sgrna_dict

np.random.seed(0)                       # DErand.py:27
logger = get_logger(logger_name=__name__)


def _gamma_fit_column(neg_logp):
    """One gene of ``DErand.py:263-276``: ``None`` = leave the zeros, else (a, b, c)."""
    finite = neg_logp[np.isfinite(neg_logp)]
    if finite.shape[0] <= 50:
        return None
    if np.all(finite == finite[0]):
        return (np.nan, 0.0, 1.0)
    return stats.gamma.fit(finite)


def _gamma_fit_chunk(neg_table):
    return [_gamma_fit_column(neg_table[:, j]) for j in range(neg_table.shape[1])]


def gamma_parameters(table, what, processes=None, min_pool_work=200000):
    """Per-gene ``scipy.stats.gamma.fit`` of ``-log p`` over the draws (the unchanged
    arithmetic of ``DErand.py:250-292``), genes farmed out to worker processes."""
    S, G = table.shape
    A, B, C = np.zeros(G), np.zeros(G), np.zeros(G)
    neg = -table
    processes = processes or get_num_processes()
    chunk = 256 if G >= 2048 else 32
    spans = [(g0, min(G, g0 + chunk)) for g0 in range(0, G, chunk)]
    if processes > 1 and G * S >= min_pool_work:
        with multiprocessing.get_context("spawn").Pool(processes) as pool:
            parts = pool.map(_gamma_fit_chunk, [neg[:, a:b] for a, b in spans])
    else:
        parts = [_gamma_fit_chunk(neg[:, a:b]) for a, b in spans]
    for (a, b), part in zip(spans, parts):
        for j, fit in enumerate(part):
            if (a + j) % 10000 == 0:
                logger.info('Finished ' + str(a + j) + ' genes.')
            if fit is not None:
                A[a + j], B[a + j], C[a + j] = fit
    return A, B, C


def gamma_parameters_device(table, what, device=0):
    """The same tail on the GPU (``spade_gamma_fit``: scipy's start values, penalised likelihood
    and Nelder-Mead restated in ``csrc/gamma_fit.cuh``, one gene per warp) - a fraction of a
    second instead of minutes for 2 x 33 000 fits.  Opt-in (``--gamma_fit device``): the
    3-parameter likelihood is ill-conditioned for genes with few distinct values, and there the
    simplex of this implementation and scipy's end at different points (see DESIGN.md for the
    measured agreement); where scipy would raise ``FitError`` nan is written."""
    from .engine import gamma_fit
    A, B, C, status = gamma_fit(table, sign=-1.0, device=device)
    logger.info(f'{what}: {int(np.count_nonzero(status == 0))} genes fitted on the GPU, '
                f'{int(np.count_nonzero(status == 1))} skipped, {int(np.count_nonzero(status == 2))} constant.')
    bad = int(np.count_nonzero(status == 3))
    if bad:
        logger.critical(f'{bad} genes: the {what} gamma fit did not converge to valid parameters (nan written).')
    return A, B, C


def DE_random_cells(sub_df_file,
                    sgrna_df,
                    sgrnas_file,
                    num_cell,
                    RAND_M,
                    output_dir,
                    iteration=1000,
                    num_processing=1,
                    norm='cpm',
                    background_cells='complement',
                    gamma_fit=None):
    rank, local_rank, world = shard.world()
    logger.info(f'B200 engine on {world} GPU rank(s); --threads is ignored (DErand.py:41).')

    if (norm != 'cpm') and (norm != 'metacell'):
        logger.critical("Incorrect normalization method. Has to be either 'cpm' or 'metacell'")
        sys.exit(0)
    if norm == 'metacell':
        logger.critical("'metacell' normalization is not available (broken in pySpade 0.1.7). Job cancels.")
        sys.exit(0)
    if (RAND_M != 'equal') and (RAND_M != 'sgrna'):
        logger.critical("Incorrect randomization method. Has to be either 'equal' or 'sgrna'")
        sys.exit(0)

    # The sgRNA table comes first: it fixes the population, the weights and the background
    # cells, i.e. everything the draws depend on.  The draws - the reference's sequential legacy
    # RNG stream, the largest host cost of a bin - then run in a second thread while this thread
    # reads and stages the transcriptome (numpy releases the GIL inside the shuffles).
    sgrna_df_adj_bool = (spio.read_frame(sgrna_df) > 0)
    [g, c] = sgrna_df_adj_bool.shape
    if RAND_M == 'sgrna':
        per_cell = np.asarray(sgrna_df_adj_bool.sum(axis=0).values, dtype=np.int64)
        total_choose_num = per_cell.sum()
        sgrna_weight_list_un = np.array(list(per_cell / total_choose_num))
        sgrna_weight_list = sgrna_weight_list_un / np.sum(sgrna_weight_list_un)
    logger.info('Finished processing sgRNA df.')

    population = np.arange(c)
    weights = sgrna_weight_list if RAND_M == 'sgrna' else None
    NC_idx = None
    if background_cells != 'complement':
        sgrna_dict = read_sgrna_dict(sgrnas_file)
        if background_cells not in sgrna_dict:
            logger.critical('Background region is not in sgRNA dictionary text file. Please fix the '
                            'dictionary file or use default parameter. Job cancels.')
            raise KeyError(background_cells)
        missing_NC_list = list(set(sgrna_dict[background_cells]) - set(sgrna_df_adj_bool.index))
        remain_NC_list = list(set(sgrna_dict[background_cells]) - (set(missing_NC_list)))
        logger.info(str(len(remain_NC_list)) + ' sgRNAs remain to find background cells.')
        sgrna_dict.update({background_cells: remain_NC_list})
        NC_idx = find_all_sgrna_cells(sgrna_df_adj_bool, sgrna_dict[background_cells])
        if len(NC_idx) == 0:
            logger.critical('No cell as background. Please consider other background set. Job cancels.')
            sys.exit(0)
        logger.info(str(len(NC_idx)) + ' cells as background for differential expression analysis.')
        population = np.setxor1d(np.arange(c), NC_idx)              # DErand.py:129-133
        if weights is not None:
            weights = sgrna_weight_list[list(population)]
            weights /= weights.sum()
    sgrna_columns = sgrna_df_adj_bool.columns
    del sgrna_df_adj_bool

    # The draws: the reference's RNG calls, in the reference's order, before any sharding.
    logger.info('Start randomization Calculation.')
    draws, failure = [], []

    def draw_all():
        try:
            for i in np.arange(iteration):
                if i % 25 == 0:
                    logger.info('Finished ' + str(i) + ' iterations.')
                if weights is not None:
                    draws.append(np.random.choice(population, size=num_cell, replace=False, p=weights))
                else:
                    draws.append(np.random.choice(population, size=num_cell, replace=False))
        except BaseException as exc:                                 # noqa: BLE001 - re-raised by the main thread
            failure.append(exc)

    import threading
    drawer = threading.Thread(target=draw_all, name="derand-draws")
    drawer.start()
    try:
        logger.info('Reading transcriptome file.')
        trans_df = spio.read_frame(sub_df_file)
        columns = trans_df.columns
        assert len(columns) == c and np.sum(columns == sgrna_columns) == c
        counts = spio.frame_to_csr(trans_df)
        del trans_df
        engine = Engine.from_counts(counts, device=local_rank)
        del counts
        logger.info('Finished transcriptome normalization.')
        if NC_idx is not None:
            engine.set_background(NC_idx)
    finally:
        drawer.join()
    if failure:
        raise failure[0]

    gamma_fit = gamma_fit or os.environ.get("SPADE_GAMMA_FIT", "scipy")
    if gamma_fit not in ("scipy", "device"):
        logger.critical("Incorrect gamma fit method. Has to be either 'scipy' or 'device'")
        sys.exit(0)
    finish_draws(engine, draws, num_cell, output_dir, gamma_fit)
    engine.close()
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    if rank == 0:
        logger.info('Job is done.')


def _fit_block(up, down, gamma_fit, device):
    """Gamma parameters of a block of gene columns: ``up`` / ``down`` CUDA tensors ``[S, g]``.
    Returns ``[A, B, C, D, E, F]`` (up then down), each ``[g]``."""
    if up.shape[1] == 0:
        return [np.zeros(0)] * 6
    if gamma_fit == "device":
        return (list(gamma_parameters_device(up, 'up', device))
                + list(gamma_parameters_device(down, 'down', device)))
    return (list(gamma_parameters(up.cpu().numpy(), 'up'))
            + list(gamma_parameters(down.cpu().numpy(), 'down')))


def finish_draws(engine, draws, num_cell, output_dir, gamma_fit="scipy", timings=None):
    """Everything after the draws (``DErand.py:190-302``): the tests of all draws in one batch,
    the three sparse tables + the cell-ID listing, the gamma parameters and ``Cpm_mean``.

    The dense tables stay on the GPU(s).  One rank: device tables ``[S, G]``.  Several ranks:
    draws sharded round-robin, tables re-sharded by gene through the fused exchange
    (``shard.run_gene_sharded``), so that every rank holds all draws of its genes - what the
    column-compressed MAT tables, the per-gene gamma fit and ``Cpm_mean`` need.  Each rank builds
    the payload of its gene columns on its GPU (``spade_csc_*``), and writes it into its slice
    of the shared files; rank 0 writes the headers.  Only the stored entries cross PCIe (and the
    dense ``up`` / ``down`` blocks when the unchanged scipy fit on the host is chosen).
    ``timings``: optional dict that receives the wall seconds of the three phases."""
    import time

    import torch

    t_start = time.perf_counter()

    rank, local_rank, world = shard.world()
    S, G = len(draws), engine.G
    names = ("up", "down", "cpm")
    files = {"up": '%s/%s-up_log-pval' % (output_dir, num_cell),
             "down": '%s/%s-down_log-pval' % (output_dir, num_cell),
             "cpm": '%s/%s-cpm' % (output_dir, num_cell)}
    use_exchange = world > 1 and not engine.wide_genes and os.environ.get("SPADE_GATHER", "genes") == "genes"

    if world > 1 and not use_exchange:
        # tables gathered on the host of rank 0 (matrices with wide genes, or SPADE_GATHER=host|nccl)
        res = shard.run_sharded(engine, draws, tables=names)
        if rank != 0:
            return
        block = {n: torch.from_numpy(res[n]).cuda(local_rank) for n in names}
        g0, g1, world_files = 0, G, 1
    elif world == 1:
        dev = torch.device("cuda", local_rank)
        members, offsets = engine.pack_sets(draws)
        block = {n: torch.empty((S, G), dtype=torch.float64, device=dev) for n in names}
        engine.run_device(torch.from_numpy(members.astype(np.int32)).to(dev), offsets, block)
        engine.check()
        g0, g1, world_files = 0, G, 1
    else:
        exchange, _ = shard.run_gene_sharded(engine, draws, tables=names)
        g0, g1 = exchange.gene_begin, exchange.gene_end
        block = {n: exchange.tables[n][:S, :g1 - g0] for n in names}
        world_files = world

    torch.cuda.synchronize(local_rank)
    t_tests = time.perf_counter()

    # --- the three sparse tables: payload of this rank's gene columns, written into its slice.
    # One thread per table (payload on the GPU, device->host copy, file write) so that the copy
    # of one table overlaps the file write of another; the cell-ID listing goes alongside.
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=4)
    jobs = []
    if rank == 0:
        jobs.append(pool.submit(spio.write_cell_ids, output_dir + '/' + str(num_cell) + '-rand_cell_ID.txt', draws))
    if world_files == 1:
        def one_table(n):
            jc, ir, pr = spio.device_payload(block[n], local_rank)
            spio.save_sparse_payload(files[n], S, G, jc, ir, pr)
        jobs += [pool.submit(one_table, n) for n in names]
        for j in jobs:
            j.result()
    else:
        import torch.distributed as dist
        payload = {n: spio.device_payload(block[n], local_rank) for n in names}
        counts = [None] * world
        dist.all_gather_object(counts, {n: (int(payload[n][0][-1]), payload[n][0]) for n in names})
        layouts = {}
        for n in names:
            layouts[n] = spio.SparseMatLayout(S, G, int(sum(c[n][0] for c in counts)))
            if rank == 0:
                jc, base = [np.zeros(1, dtype=np.int64)], 0
                for r in range(world):
                    jc.append(counts[r][n][1][1:].astype(np.int64) + base)
                    base += counts[r][n][0]
                fd = os.open(files[n], os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
                layouts[n].write_frame(fd, np.concatenate(jc))
                os.close(fd)
        dist.barrier()                                       # the files exist and have their size

        def one_slice(n):
            if payload[n][0][-1] > 0:
                fd = os.open(files[n], os.O_RDWR)
                layouts[n].write_entries(fd, int(sum(counts[r][n][0] for r in range(rank))), payload[n][1], payload[n][2])
                os.close(fd)
        jobs += [pool.submit(one_slice, n) for n in names]
        for j in jobs:
            j.result()
        del payload
    pool.shutdown()
    t_files = time.perf_counter()

    # --- gamma parameters and Cpm_mean of this rank's genes
    if rank == 0:
        logger.info('Start modeling randomization distribution.')
    from .engine import column_mean
    cpm_mean = column_mean(block["cpm"], local_rank)
    if rank == 0:
        logger.info('Process up-regulation.')
        logger.info('Process down-regulation.')
    parts = _fit_block(block["up"], block["down"], gamma_fit, local_rank) + [cpm_mean]
    if world_files > 1:
        import torch.distributed as dist
        everyone = [None] * world if rank == 0 else None
        dist.gather_object(parts, everyone, dst=0)
        if rank != 0:
            return
        parts = [np.concatenate([everyone[r][i] for r in range(world)]) for i in range(7)]
    A, B, C, D, E, F, cpm_mean = parts
    if timings is not None:
        timings.update(tests_s=t_tests - t_start, files_s=t_files - t_tests,
                       gamma_fit_s=time.perf_counter() - t_files)

    # string concatenation as in DErand.py:295-302: output_dir must end with '/'
    np.save(output_dir + 'Up_dist_gamma-%s-A' % (num_cell), A)
    np.save(output_dir + 'Up_dist_gamma-%s-B' % (num_cell), B)
    np.save(output_dir + 'Up_dist_gamma-%s-C' % (num_cell), C)
    np.save(output_dir + 'Down_dist_gamma-%s-A' % (num_cell), D)
    np.save(output_dir + 'Down_dist_gamma-%s-B' % (num_cell), E)
    np.save(output_dir + 'Down_dist_gamma-%s-C' % (num_cell), F)
    np.save(output_dir + 'Cpm_mean-%s' % (num_cell), cpm_mean)


if __name__ == '__main__':
    pass
