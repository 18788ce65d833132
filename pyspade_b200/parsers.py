"""Command-line flags of the two sub-commands on the hot path, identical to the
reference's (``pySpade/parsers.py:167-227`` DEobs, ``:230-312`` DErand): same option
strings, destinations, types, defaults and required-ness.  The other seven sub-commands of
pySpade are outside this engine's scope.
"""
import argparse
import sys

from . import __version__


def parse_args(args=None):
    argv = sys.argv[1:] if args is None else list(args)
    parser = argparse.ArgumentParser(
        prog="pySpade",
        formatter_class=argparse.RawDescriptionHelpFormatter,
        description="pySpade (B200 engine for DEobs / DErand)\nVersion: " + __version__)
    subparsers = parser.add_subparsers(title="functions", dest="command", metavar="")
    add_DEobs_subparser(subparsers)
    add_DErand_subparser(subparsers)

    if len(argv) > 0:
        if argv[0] in ("--version", "-v"):
            print(" version: %s" % __version__)
            sys.exit()
    else:
        parser.parse_args(["-h"])
        sys.exit()
    return parser.parse_args(argv)


def _common(parser, threads_help):
    parser.add_argument("-t", "--transcriptome_df", dest="transcriptome_df", required=True, type=str,
                        help="processed transcriptome matrix (genes x cells DataFrame, .pkl or .h5)")
    parser.add_argument("-s", "--sgrna", dest="input_sgrna", required=True, type=str,
                        help="processed sgRNA matrix (sgRNAs x cells DataFrame, .pkl or .h5)")
    parser.add_argument("-d", "-dict", dest="dict", required=True, type=str,
                        help="text file: region <TAB> sgRNA;sgRNA;... per line")


def _tail(parser):
    parser.add_argument("-r", "--threads", dest="threads", required=False, type=int, default=1,
                        help="accepted for compatibility and ignored (the reference ignores it too)")
    parser.add_argument("-n", "--norm", dest="norm_method", required=False, type=str, default="cpm",
                        help='normalization method: "cpm" ("metacell" is refused: broken upstream)')


def add_DEobs_subparser(subparsers):
    parser = subparsers.add_parser(
        "DEobs",
        help="differential expression of the observed cells of every perturbation region",
        description="Genome-wide hypergeometric test of every gene for every region of the sgRNA "
                    "dictionary. Writes, per region: up / down log p-values, fold change, mean CPM.")
    _common(parser, True)
    _tail(parser)
    parser.add_argument("-b", "--bg", dest="bg", required=False, type=str, default="complement",
                        help="background cells: 'complement' (all other cells) or a key of the "
                             "sgRNA dictionary (negative-control cells)")
    parser.add_argument("-o", "--output_dir", dest="output_dir", required=True,
                        help="output directory")


def add_DErand_subparser(subparsers):
    parser = subparsers.add_parser(
        "DErand",
        help="differential expression of random cell draws (empirical background)",
        description="The same test on randomly drawn cell sets of a given size; draws are uniform "
                    "('equal') or weighted by the number of sgRNAs per cell ('sgrna').")
    _common(parser, True)
    parser.add_argument("-m", "--num", dest="num", required=True, type=int,
                        help="number of cells per random draw")
    parser.add_argument("-i", "--iteration", dest="iteration", required=False, type=int, default=1000,
                        help="number of draws (default 1000)")
    _tail(parser)
    parser.add_argument("-a", "--randomization_method", dest="randomization_method", required=True,
                        type=str, help='"equal" or "sgrna"')
    parser.add_argument("-b", "--bg", dest="bg", required=False, type=str, default="complement",
                        help="background cells: 'complement' or a key of the sgRNA dictionary")
    parser.add_argument("--gamma_fit", dest="gamma_fit", required=False, type=str, default=None,
                        choices=["scipy", "device"],
                        help="not in the reference: 'scipy' (default) = the reference's per-gene "
                             "scipy.stats.gamma.fit on the host cores (exact), 'device' = the restated procedure on the GPU "
                             "(0.2 s instead of a minute; parameters within 1e-3 of scipy's for 97-99 %% of the "
                             "genes - the other fits stop at a different simplex point, see DESIGN.md)")
    parser.add_argument("-o", "--output_dir", dest="output_dir", required=True,
                        help="output directory (must end with '/': the .npy names are concatenated)")
