"""``pySpade DEobs``: genome-wide DE of every perturbation region (mirror of
``pySpade/DEobs.py:28-223``).  Same arguments, same log messages for the per-region
bookkeeping, same output files; the per-gene tests of all regions run as batches on the GPU
(``Engine.run`` / ``shard.run_sharded``) instead of one process pool per region.
"""
import os
import sys

import numpy as np

from . import io as spio
from . import shard
from .engine import Engine
from .utils import find_all_sgrna_cells, get_logger, read_sgrna_dict, region_cell_sets

np.random.seed(0)                       # DEobs.py:25
logger = get_logger(logger_name=__name__)

REGION_BATCH = 512                      # regions per GPU batch (bounds the host tables)


def DE_observe_cells(sub_df_file,
                     sgrna_df,
                     sgrnas_file,
                     output_dir,
                     num_processing=1,
                     norm='cpm',
                     background_cells='complement'):
    rank, local_rank, world = shard.world()
    logger.info(f'B200 engine on {world} GPU rank(s); --threads is ignored (DEobs.py:36).')

    if (norm != 'cpm') and (norm != 'metacell'):
        logger.critical("Incorrect normalization method. Has to be either 'cpm' or 'metacell'")
        sys.exit(0)
    if norm == 'metacell':
        # utils.py:158-170 has no return statement: the reference crashes on it
        logger.critical("'metacell' normalization is not available (broken in pySpade 0.1.7). Job cancels.")
        sys.exit(0)

    logger.info('Reading transcriptome file')
    trans_df = spio.read_frame(sub_df_file)
    columns = trans_df.columns
    counts = spio.frame_to_csr(trans_df)
    del trans_df
    g, c = counts.shape

    logger.info('Start transcriptome normalization.')
    engine = Engine.from_counts(counts, device=local_rank)         # K0 + K1 on the GPU
    del counts
    logger.info('Finished transcriptome normalization.')

    logger.info('Loading sgRNA df.')
    sgrna_df_adj_bool = (spio.read_frame(sgrna_df) > 0)
    if sgrna_df_adj_bool.shape[1] != c or np.sum(columns == sgrna_df_adj_bool.columns) != c:
        logger.critical('The cell ID sequences of transcriptome df and sgrna df are different.')
        sgrna_df_adj_bool = sgrna_df_adj_bool.reindex(columns=columns, fill_value=False)

    sgrna_dict = read_sgrna_dict(sgrnas_file)

    skip = ()
    if background_cells != 'complement':
        if background_cells not in sgrna_dict:
            logger.critical('Background region is not in sgRNA dictionary text file. Please fix the '
                            'dictionary file or use default parameter. Job cancels.')
            raise KeyError(background_cells)           # what DEobs.py:91 ends in
        missing_NC_list = list(set(sgrna_dict[background_cells]) - set(sgrna_df_adj_bool.index))
        remain_NC_list = list(set(sgrna_dict[background_cells]) - (set(missing_NC_list)))
        logger.info(str(len(remain_NC_list)) + ' sgRNAs remain to find background cells.')
        sgrna_dict.update({background_cells: remain_NC_list})
        NC_idx = find_all_sgrna_cells(sgrna_df_adj_bool, sgrna_dict[background_cells])
        if len(NC_idx) == 0:
            logger.critical('No cell as background. Job cancels.')
            sys.exit(0)
        logger.info(str(len(NC_idx)) + ' cells as background for differential expression analysis.')
        engine.set_background(NC_idx)
        skip = (background_cells,)

    regions = region_cell_sets(sgrna_df_adj_bool, sgrna_dict, skip=skip, logger=logger)
    pinned = None
    if world == 1 and regions:
        # one set of page-locked host tables, reused by every batch (fast device->host copies)
        from .engine import PinnedArray
        rows = min(REGION_BATCH, len(regions))
        pinned = {n: PinnedArray((rows, g), np.float64) for n in ("up", "down", "fc", "cpm")}
    for b0 in range(0, len(regions), REGION_BATCH):
        batch = regions[b0:b0 + REGION_BATCH]
        sets = [cells for _, cells in batch]
        if pinned is not None:
            res = engine.run(sets, out={n: p.array[:len(sets)] for n, p in pinned.items()})
        else:
            res = shard.run_sharded(engine, sets, balance=True)
        if rank != 0:
            continue
        for j, (k, cells) in enumerate(batch):
            num_sgrna_cell = len(cells)
            spio.save_vector('%s/%s-up_log-pval' % (output_dir, k), res["up"][j])
            spio.save_vector('%s/%s-down_log-pval' % (output_dir, k), res["down"][j])
            spio.save_vector('%s/%s-%s-foldchange' % (output_dir, k, num_sgrna_cell), res["fc"][j])
            spio.save_vector('%s/%s-cpm' % (output_dir, k), res["cpm"][j])
    engine.close()
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    logger.info('Job is done.')


if __name__ == '__main__':
    pass
