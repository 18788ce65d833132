"""ctypes binding of ``libspade_b200.so`` (C ABI declared in ``include/spade.h``).

There is no CPU fallback: if the shared library is missing the import of the
engine fails with an explicit message telling how to build it.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPADE_LIB") or os.path.join(_PKG, "libspade_b200.so")   # SPADE_LIB: tuning variants

SPADE_OK, SPADE_ERR_ARG, SPADE_ERR_CUDA, SPADE_ERR_STATE = 0, 1, 2, 3
SPADE_HOST, SPADE_DEVICE = 0, 1
SPADE_ABI_VERSION = 5
SPADE_OPT_TAIL_TABLE, SPADE_OPT_MEMBER_BLOCK, SPADE_OPT_GENES_PER_PART, SPADE_OPT_SORT_MEMBERS = 1, 2, 3, 4

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_vpp = ctypes.POINTER(ctypes.c_void_p)


class GeneBlock(ctypes.Structure):
    """``spade_gene_block`` of ``include/spade.h``."""
    _fields_ = [("base", ctypes.c_void_p), ("ld", ctypes.c_int64), ("gene_begin", ctypes.c_int64),
                ("gene_end", ctypes.c_int64)]

# name -> (restype, argtypes); every symbol include/spade.h declares
SIGNATURES = {
    "spade_abi_version": (ctypes.c_int, []),
    "spade_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "spade_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_vpp]),
    "spade_create_from_counts": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_int, c_vpp]),
    "spade_destroy": (None, [ctypes.c_void_p]),
    "spade_set_background": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    "spade_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64]),
    "spade_dims": (ctypes.c_int, [ctypes.c_void_p, c_i64p, c_i64p, c_i64p, c_i32p]),
    "spade_gene_cutoffs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "spade_cpm_values": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "spade_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "spade_run_strided": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                         ctypes.c_void_p]),
    "spade_run_raw": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "spade_expand": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "spade_run_raw_blocks": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                            ctypes.c_int64, ctypes.c_void_p]),
    "spade_expand_block": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int64, ctypes.c_void_p]),
    "spade_csc_count": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_int64, ctypes.c_void_p]),
    "spade_csc_fill": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_int]),
    "spade_column_mean": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_int64, ctypes.c_void_p]),
    "spade_memcpy2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]),
    "spade_device_alloc": (ctypes.c_int, [ctypes.c_int, c_vpp, ctypes.c_int64]),
    "spade_device_free": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "spade_ipc_export": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p]),
    "spade_ipc_open": (ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, c_vpp]),
    "spade_ipc_close": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "spade_check": (ctypes.c_int, [ctypes.c_void_p]),
    "spade_wide_genes": (ctypes.c_int, [ctypes.c_void_p, c_i64p]),
    "spade_timing": (ctypes.c_int, [ctypes.c_void_p, c_f64p, c_f64p, c_i64p, ctypes.c_int]),
    "spade_timing_finish": (ctypes.c_int, [ctypes.c_void_p, c_f64p]),
    "spade_hypergeom": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "spade_gamma_fit": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "spade_score": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_int64]),
    "spade_host_alloc": (ctypes.c_int, [c_vpp, ctypes.c_int64]),
    "spade_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "spade_host_register": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "spade_host_unregister": (ctypes.c_int, [ctypes.c_void_p]),
}

_lib = None


class SpadeLibraryMissing(ImportError):
    pass


def load():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SpadeLibraryMissing(
            f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU "
            "fallback. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `python pyspade_b200/_build.py`) with nvcc on PATH.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.spade_abi_version()
    if got != SPADE_ABI_VERSION:
        raise ImportError(f"libspade_b200 ABI {got} != expected {SPADE_ABI_VERSION}; rebuild")
    _lib = lib
    return lib
