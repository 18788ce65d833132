"""Input / output formats of DEobs and DErand.

Inputs (``DEobs.py:47-50,69-72``): genes x cells and sgRNAs x cells DataFrames as ``.pkl``
(``pd.read_pickle``) or ``.h5`` key ``'df'`` (``pd.read_hdf``, needs PyTables).
Outputs: MATLAB v5 files written by ``scipy.io.savemat`` WITHOUT a ``.mat`` extension,
variable name ``'matrix'`` (``DEobs.py:208-221``, ``DErand.py:230-248``); ``.npy`` vectors;
the ``rand_cell_ID.txt`` listing.  Readers downstream: ``utils.load_data``
(``utils.py:343-361``), ``local.py:81-110``, ``global.py:84-111``.
"""
import struct

import numpy as np
import pandas as pd
import scipy.sparse as sp
from scipy import io as sio

ROW_CHUNK = 2048


def read_frame(path):
    """The reference reads ``.pkl`` and ``.h5`` and leaves anything else undefined (a
    ``NameError`` later); here an unknown extension is an explicit ``ValueError``."""
    if path.endswith("pkl"):
        return pd.read_pickle(path)
    if path.endswith("h5"):
        return pd.read_hdf(path, "df")
    raise ValueError(f"{path}: expected a .pkl or .h5 DataFrame file")


def frame_to_csr(df, dtype=np.int32):
    """genes x cells DataFrame -> CSR without materialising pandas' per-column sparse
    blocks (what ``astype(SparseDtype)`` + ``csr_matrix(df)`` do in ``DEobs.py:52,60``)."""
    try:
        m = sp.csr_matrix(df.sparse.to_coo(), dtype=dtype)
    except (AttributeError, TypeError):
        blocks = []
        for r0 in range(0, df.shape[0], ROW_CHUNK):
            blocks.append(sp.csr_matrix(df.iloc[r0:r0 + ROW_CHUNK].to_numpy().astype(dtype, copy=False)))
        m = sp.vstack(blocks, format="csr") if blocks else sp.csr_matrix(df.shape, dtype=dtype)
    m.sum_duplicates()
    m.sort_indices()
    return m


# ------------------------------------------------------------------------------------------
# MAT-v5 writers.  The files are what ``scipy.io.savemat(path, {'matrix': x})`` produces
# (Level-5 MAT, little endian, one miMATRIX element named 'matrix'), written directly with
# numpy: ``savemat`` spends seconds per DErand table in Python-level conversions, and a
# 6 000-region DEobs run writes 24 000 small files.  ``tests/test_host.py`` checks that
# ``scipy.io.loadmat`` returns the same arrays as for scipy-written files.
# ------------------------------------------------------------------------------------------
_MI_INT8, _MI_INT32, _MI_UINT32, _MI_DOUBLE, _MI_MATRIX = 1, 5, 6, 9, 14
_MX_DOUBLE, _MX_SPARSE = 6, 5


def _mat_header():
    text = b"MATLAB 5.0 MAT-file Platform: posix, Created by spade-b200"
    return text.ljust(116, b" ") + b"\x00" * 8 + struct.pack("<H2s", 0x0100, b"IM")


def _element(mdtype, payload):
    """Tag + data padded to 8 bytes (always the long form, as scipy writes arrays)."""
    n = len(payload)
    return struct.pack("<II", mdtype, n) + payload + b"\x00" * ((-n) % 8)


def _matrix_prefix(mclass, nzmax, dims, name=b"matrix"):
    return (_element(_MI_UINT32, struct.pack("<II", mclass, nzmax))
            + _element(_MI_INT32, struct.pack("<ii", *dims))
            + _element(_MI_INT8, name))


def _array_tag(mdtype, nbytes):
    return struct.pack("<II", mdtype, nbytes)


def save_vector(path, vec):
    """One DEobs table row: 1-D float64 under ``'matrix'`` (loads back as ``(1, G)``),
    ``DEobs.py:208-221``."""
    data = np.ascontiguousarray(vec, dtype="<f8").ravel()
    body = _matrix_prefix(_MX_DOUBLE, 0, (1, data.shape[0])) + _array_tag(_MI_DOUBLE, data.nbytes)
    pad = (-data.nbytes) % 8
    total = len(body) + data.nbytes + pad
    if total >= 1 << 32:
        raise ValueError("array too large for a MAT-v5 element")
    with open(path, "wb") as fh:
        fh.write(_mat_header())
        fh.write(struct.pack("<II", _MI_MATRIX, total))
        fh.write(body)
        fh.write(data.tobytes())
        fh.write(b"\x00" * pad)


def save_draw_table(path, table):
    """One DErand table: the sparse matrix ``csr_matrix(dense)`` (exact zeros dropped, nan /
    -inf kept) under ``'matrix'`` - what ``DErand.py:230-248`` leaves on disk after the last
    draw.  MAT sparse storage is column-compressed: row indices, column pointers, values."""
    table = np.asarray(table, dtype=np.float64)
    S, G = table.shape
    tt = np.ascontiguousarray(table.T)                      # [G, S]: one column of the table per row
    mask = tt != 0                                          # nan and -inf are stored, +-0.0 are not
    jc = np.zeros(G + 1, dtype="<i4")
    np.cumsum(np.count_nonzero(mask, axis=1), out=jc[1:])
    ir = np.nonzero(mask)[1].astype("<i4")                  # row index of every stored value, column-major
    pr = tt[mask].astype("<f8", copy=False)
    nnz = int(pr.shape[0])
    if nnz == 0:
        # scipy keeps one explicit slot for an empty sparse matrix (nzmax = 1)
        ir, pr, nzmax = np.zeros(1, "<i4"), np.zeros(1, "<f8"), 1
    else:
        nzmax = nnz
    parts = [_matrix_prefix(_MX_SPARSE, nzmax, (S, G))]
    for mdtype, arr in ((_MI_INT32, ir), (_MI_INT32, jc), (_MI_DOUBLE, pr)):
        parts.append((_array_tag(mdtype, arr.nbytes), arr, b"\x00" * ((-arr.nbytes) % 8)))
    total = len(parts[0]) + sum(len(t) + a.nbytes + len(pd_) for t, a, pd_ in parts[1:])
    if total >= 1 << 32:
        raise ValueError("table too large for a MAT-v5 element (same limit as scipy.io.savemat)")
    with open(path, "wb") as fh:
        fh.write(_mat_header())
        fh.write(struct.pack("<II", _MI_MATRIX, total))
        fh.write(parts[0])
        for tag, arr, padding in parts[1:]:
            fh.write(tag)
            arr.tofile(fh)
            fh.write(padding)


def save_draw_tables(items):
    """``[(path, table), ...]`` written concurrently (the numpy passes release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, len(items))) as pool:
        list(pool.map(lambda pt: save_draw_table(*pt), items))


def write_cell_ids(path, draws):
    """``<i>\\t<repr of the index list>`` per draw (``DErand.py:224-228``): the reference
    joins a one-element list holding the list, so the text is Python's ``str(list(...))``
    of numpy integers."""
    with open(path, "w") as fh:
        for i, d in enumerate(draws):
            fh.write(str(i) + "\t" + ",".join(str(j) for j in [list(d)]) + "\n")
