"""``python -m pyspade_b200 DEobs|DErand ...`` - the two hot-path sub-commands of
``pySpade`` (dispatch mirrors ``pySpade/__main__.py:63-93``).  Under
``python -m torch.distributed.run --nproc-per-node N -m pyspade_b200 ...`` the cell sets
are sharded over N GPUs and rank 0 writes the files.
"""
import importlib
import sys

from . import __version__
from .parsers import parse_args
from .utils import get_logger


def main(argv=None):
    args = parse_args(argv)
    logger = get_logger(logger_name=__name__)
    logger.info(f'pySpade (B200 engine) version: {__version__}')
    logger.info(f'Python version: {sys.version_info.major}.{sys.version_info.minor}')

    if args.command == 'DEobs':
        logger.info('Running DEobs command ...')
        m = importlib.import_module(name=f'{__package__}.{args.command}')
        m.DE_observe_cells(
            sub_df_file=args.transcriptome_df,
            sgrna_df=args.input_sgrna,
            sgrnas_file=args.dict,
            num_processing=args.threads,
            norm=args.norm_method,
            background_cells=args.bg,
            output_dir=args.output_dir)
        logger.info('Done.')
    elif args.command == 'DErand':
        logger.info('Running DErand command ...')
        m = importlib.import_module(name=f'{__package__}.{args.command}')
        m.DE_random_cells(
            sub_df_file=args.transcriptome_df,
            sgrna_df=args.input_sgrna,
            sgrnas_file=args.dict,
            num_cell=args.num,
            iteration=args.iteration,
            num_processing=args.threads,
            norm=args.norm_method,
            RAND_M=args.randomization_method,
            background_cells=args.bg,
            output_dir=args.output_dir,
            gamma_fit=args.gamma_fit)
        logger.info('Done.')
    else:
        logger.critical('Only the DEobs and DErand sub-commands are provided by this engine.')
        sys.exit(1)


if __name__ == '__main__':
    main()
