"""Host-side mirror of the hot-path helpers of ``pySpade/utils.py`` (reference v0.1.7).

Same names, argument order and return order as the reference; the arithmetic runs on the
GPU through ``libspade_b200`` (no CPU fallback - without the CUDA extension or a device
every function here raises).

=====================================  ==========================================
here                                   reference
=====================================  ==========================================
``get_logger``                         ``utils.py:22-40``
``read_sgrna_dict``                    ``utils.py:113-120``
``find_all_sgrna_cells``               ``utils.py:141-148`` (vectorised, sorted)
``region_cell_sets``                   ``DEobs.py:165-187`` bookkeeping, all regions at once
``cpm_normalization``                  ``utils.py:150-155``
``perform_DE`` / ``perform_DE_NC``     ``utils.py:296-317`` / ``:319-341``
``get_num_processes``                  ``utils.py:190-192``
=====================================  ==========================================
"""
import logging
import multiprocessing
import weakref

import numpy as np
import scipy.sparse as sp

from .engine import Engine

_FORMATTER = "%(asctime)s - %(name)s - %(levelname)s - %(message)s"


def get_logger(logger_name, log_file=False):
    """Same handler layout and message format as ``utils.py:22-40``."""
    logger = logging.getLogger(logger_name)
    logger.setLevel(level=logging.DEBUG)
    if not any(getattr(h, "_spade", False) for h in logger.handlers):
        console = logging.StreamHandler()
        console.setLevel(logging.INFO)
        console.setFormatter(logging.Formatter(_FORMATTER))
        console._spade = True
        logger.addHandler(console)
        if log_file:
            fh = logging.FileHandler(filename="pySpade.log")
            fh.setLevel(logging.DEBUG)
            fh.setFormatter(logging.Formatter(_FORMATTER))
            fh._spade = True
            logger.addHandler(fh)
    return logger


def get_num_processes():
    return multiprocessing.cpu_count()


def read_sgrna_dict(SGRNA_FILE):
    """``region<TAB>sg1;sg2;...`` per line (``utils.py:113-120``); a malformed line raises
    ``ValueError`` exactly like the reference's tuple unpacking."""
    sgrna_dict = {}
    with open(SGRNA_FILE) as fh:
        for line in fh:
            region_id, sgrna_string = line.strip().split("\t")
            sgrna_dict.update({region_id: sgrna_string.split(";")})
    return sgrna_dict


def _bool_matrix(sgRNA_df):
    """sgRNAs x cells boolean CSR of a (possibly pandas-sparse) DataFrame."""
    try:
        coo = sgRNA_df.sparse.to_coo()
        return sp.csr_matrix(coo != 0)
    except (AttributeError, TypeError):
        return sp.csr_matrix(np.asarray(sgRNA_df) != 0)


def find_all_sgrna_cells(sgRNA_df, sgRNA_dict):
    """Integer positions of the cells that carry any sgRNA of the list
    (``utils.py:141-148``).  The reference resolves every barcode with
    ``np.argwhere(columns == bc)`` (seconds per region at 50k cells); here it is one boolean
    reduction.  The reference returns the positions in ``set`` order; only the set matters
    downstream, so they come back sorted.  Unknown sgRNA names raise ``KeyError`` like
    ``.loc``."""
    rows = sgRNA_df.index.get_indexer(list(sgRNA_dict))
    if (rows < 0).any():
        missing = [s for s, r in zip(sgRNA_dict, rows) if r < 0]
        raise KeyError(f"{missing} not in index")
    mat = _bool_matrix(sgRNA_df)
    hit = np.asarray(mat[rows].sum(axis=0)).ravel() > 0
    if sgRNA_df.columns.is_unique:
        return [int(i) for i in np.flatnonzero(hit)]
    # duplicated barcodes: the reference's argwhere(...).item() raises on them
    raise ValueError("can only convert an array of size 1 to a Python scalar")


def region_cell_sets(sgrna_df_bool, sgrna_dict, skip=(), logger=None):
    """The per-region bookkeeping of ``DEobs.py:165-187`` for every region at once:
    drop sgRNAs absent from the matrix, skip regions left with no sgRNA or no cell.
    Returns ``[(region, cell_index_list), ...]`` in dictionary order and rewrites
    ``sgrna_dict`` entries the way the reference does."""
    mat = _bool_matrix(sgrna_df_bool)
    index = sgrna_df_bool.index
    present = set(index)
    out = []
    for k in sgrna_dict:
        if k in skip:
            continue
        if logger:
            logger.info(f"Start DE analysis of region: {k}")
        missing = list(set(sgrna_dict[k]) - present)
        if len(missing) == len(sgrna_dict[k]):
            if logger:
                logger.info("Missing all the sgRNA in this region.")
            continue
        if len(missing) > 0:
            if logger:
                left = len(sgrna_dict[k]) - len(missing)
                logger.info("Missing " + str(len(missing)) + " sgrna, " + str(left) + " left for analysis.")
                for i in missing:
                    logger.info("Missing sgrna: " + str(i))
            sgrna_dict.update({k: list(set(sgrna_dict[k]) - set(missing))})
        rows = index.get_indexer(list(sgrna_dict[k]))
        hit = np.asarray(mat[rows].sum(axis=0)).ravel() > 0
        cells = [int(i) for i in np.flatnonzero(hit)]
        if len(cells) == 0:
            if logger:
                logger.info("No cell in this region.")
            continue
        out.append((k, cells))
    return out


# ------------------------------------------------------------------------------------------
# staged matrices: one engine per input matrix object, kept while the matrix is alive
# ------------------------------------------------------------------------------------------
_ENGINES = {}          # id(matrix) -> (weakref to the matrix, engine, fingerprint)


def _fingerprint(matrix):
    """Cheap guard against in-place edits of a cached matrix (the reference re-reads the
    matrix on every call): shape, nnz and a strided sample of the stored values."""
    data = getattr(matrix, "data", None)
    if not isinstance(data, np.ndarray):
        return None
    step = max(1, data.shape[0] // 4096)
    sample = data[::step]
    return (tuple(matrix.shape), int(data.shape[0]), float(np.nansum(sample)), int(np.isnan(sample).sum()))


def _drop_engine(key):
    hit = _ENGINES.pop(key, None)
    if hit is not None:
        hit[1].close()


def _remember(matrix, engine):
    """Keep ``engine`` for as long as ``matrix`` lives.  Objects that cannot be weakly
    referenced are not cached at all (``id()`` values are recycled)."""
    try:
        ref = weakref.ref(matrix)
        key = id(matrix)
        weakref.finalize(matrix, _drop_engine, key)
    except TypeError:
        return engine
    _ENGINES[key] = (ref, engine, _fingerprint(matrix))
    return engine


def engine_for(input_array, device=0):
    """The engine staged for this matrix object (staged on first use).  The cache entry is
    used only when it was made for this very object and the stored values still look the same;
    the matrix must not be edited in place between calls."""
    key = id(input_array)
    hit = _ENGINES.get(key)
    if hit is not None:
        if hit[0]() is input_array and hit[2] == _fingerprint(input_array):
            return hit[1]
        _drop_engine(key)
    return _remember(input_array, Engine.from_cpm(input_array, device=device))


def _release_if_uncached(input_array, eng):
    hit = _ENGINES.get(id(input_array))
    if hit is None or hit[1] is not eng:
        eng.close()


def cpm_normalization(trans_df, device=0):
    """``utils.py:150-155``: genes x cells counts -> sparse float64 CPM (COO, like the
    reference).  The normalisation runs on the GPU (K0) and is bit-identical to
    ``X.multiply(1e6).multiply(1 / colsum)``; the returned matrix remembers its staged
    engine, so ``perform_DE`` on it does not upload anything again."""
    try:
        counts = sp.csr_matrix(trans_df.sparse.to_coo())
    except (AttributeError, TypeError):
        counts = sp.csr_matrix(trans_df)
    counts.sum_duplicates()
    counts.sort_indices()
    eng = Engine.from_counts(counts, device=device)
    cpm = sp.csr_matrix((eng.cpm_values(), counts.indices, counts.indptr), shape=counts.shape).tocoo()
    _remember(cpm, eng)
    return cpm


def _scatter(idx, res, pval_list_down, pval_list_up, fc_list, cpm_list):
    # the reference zips the matrix rows with idx: row j's result lands at position idx[j]
    idx = np.asarray(idx, dtype=np.int64)
    rows = np.arange(idx.shape[0])
    pval_list_up[idx] = res["up"][0, rows]
    pval_list_down[idx] = res["down"][0, rows]
    fc_list[idx] = res["fc"][0, rows]
    cpm_list[idx] = res["cpm"][0, rows]


def _as_out(a):
    return a if isinstance(a, np.ndarray) else np.asarray(a, dtype=np.float64)


def perform_DE(sgrna_idx, input_array, idx, num_processes, pval_list_down, pval_list_up, fc_list,
               cpm_list):
    """Drop-in for ``utils.py:296-317`` (complement background): every gene of
    ``input_array`` against the cell set ``sgrna_idx``.  ``num_processes`` is accepted and
    ignored.  Returns ``(len(sgrna_idx), up, down, fc, cpm)`` - up before down, as the
    reference does - after filling the caller's arrays at positions ``idx``."""
    eng = engine_for(input_array)
    if eng.n_nc:
        eng.set_background(None)
    res = eng.run([np.asarray(sgrna_idx, dtype=np.int64)])
    pval_list_down, pval_list_up, fc_list, cpm_list = map(_as_out, (pval_list_down, pval_list_up,
                                                                    fc_list, cpm_list))
    _release_if_uncached(input_array, eng)
    _scatter(idx, res, pval_list_down, pval_list_up, fc_list, cpm_list)
    return len(sgrna_idx), pval_list_up, pval_list_down, fc_list, cpm_list


def perform_DE_NC(sgrna_idx, NC_idx, input_array, idx, num_processes, pval_list_down, pval_list_up,
                  fc_list, cpm_list):
    """Drop-in for ``utils.py:319-341`` (negative-control background)."""
    eng = engine_for(input_array)
    nc = np.asarray(NC_idx, dtype=np.int64)
    if eng.background_key != nc.tobytes():
        eng.set_background(nc)
    res = eng.run([np.asarray(sgrna_idx, dtype=np.int64)])
    pval_list_down, pval_list_up, fc_list, cpm_list = map(_as_out, (pval_list_down, pval_list_up,
                                                                    fc_list, cpm_list))
    _release_if_uncached(input_array, eng)
    _scatter(idx, res, pval_list_down, pval_list_up, fc_list, cpm_list)
    return len(sgrna_idx), pval_list_up, pval_list_down, fc_list, cpm_list
