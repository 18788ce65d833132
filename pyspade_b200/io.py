"""Input / output formats of DEobs and DErand.

Inputs (``DEobs.py:47-50,69-72``): genes x cells and sgRNAs x cells DataFrames as ``.pkl``
(``pd.read_pickle``) or ``.h5`` key ``'df'`` (``pd.read_hdf``, needs PyTables).
Outputs: MATLAB v5 files written by ``scipy.io.savemat`` WITHOUT a ``.mat`` extension,
variable name ``'matrix'`` (``DEobs.py:208-221``, ``DErand.py:230-248``); ``.npy`` vectors;
the ``rand_cell_ID.txt`` listing.  Readers downstream: ``utils.load_data``
(``utils.py:343-361``), ``local.py:81-110``, ``global.py:84-111``.
"""
import os
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


class SparseMatLayout:
    """Byte layout of a MAT-v5 file holding one sparse ``[S, G]`` float64 matrix named
    ``'matrix'`` with ``nnz`` stored entries (column-compressed: ``ir`` row indices, ``jc``
    column pointers, ``pr`` values) - what ``scipy.io.savemat(path, {'matrix': csr_matrix(x)})``
    writes.  Knowing ``nnz`` fixes the offset of every array, so several writers (one per GPU
    rank, each holding a block of gene columns) can fill one file with ``os.pwrite``."""

    def __init__(self, S, G, nnz):
        self.S, self.G, self.nnz = int(S), int(G), int(nnz)
        self.nzmax = max(self.nnz, 1)            # scipy keeps one explicit slot for an empty matrix
        self.prefix = _matrix_prefix(_MX_SPARSE, self.nzmax, (self.S, self.G))
        ir_bytes, jc_bytes, pr_bytes = 4 * self.nzmax, 4 * (self.G + 1), 8 * self.nzmax
        pad = lambda n: (-n) % 8                 # noqa: E731
        at = 128 + 8 + len(self.prefix)
        self.ir_tag, self.ir_off = at, at + 8
        at = self.ir_off + ir_bytes + pad(ir_bytes)
        self.jc_tag, self.jc_off = at, at + 8
        at = self.jc_off + jc_bytes + pad(jc_bytes)
        self.pr_tag, self.pr_off = at, at + 8
        self.size = self.pr_off + pr_bytes + pad(pr_bytes)
        self.element_bytes = self.size - 128 - 8
        if self.element_bytes >= 1 << 32:
            raise ValueError("table too large for a MAT-v5 element (same limit as scipy.io.savemat)")

    def write_frame(self, fd, jc):
        """Size the file and write everything except ``ir`` / ``pr``: headers, tags, ``jc``."""
        os.ftruncate(fd, self.size)              # padding and the empty-matrix slot read as zeros
        os.pwrite(fd, _mat_header() + struct.pack("<II", _MI_MATRIX, self.element_bytes) + self.prefix, 0)
        os.pwrite(fd, _array_tag(_MI_INT32, 4 * self.nzmax), self.ir_tag)
        os.pwrite(fd, _array_tag(_MI_INT32, 4 * (self.G + 1)), self.jc_tag)
        os.pwrite(fd, np.ascontiguousarray(jc, dtype="<i4").tobytes(), self.jc_off)
        os.pwrite(fd, _array_tag(_MI_DOUBLE, 8 * self.nzmax), self.pr_tag)

    def write_entries(self, fd, first, ir, pr):
        """Stored entries ``first .. first + len(ir)`` of the file (column-major order)."""
        _pwrite_all(fd, memoryview(np.ascontiguousarray(ir, dtype="<i4")).cast("B"), self.ir_off + 4 * first)
        _pwrite_all(fd, memoryview(np.ascontiguousarray(pr, dtype="<f8")).cast("B"), self.pr_off + 8 * first)


def _pwrite_all(fd, view, offset):
    done = 0
    while done < len(view):
        done += os.pwrite(fd, view[done:done + (1 << 30)], offset + done)


def save_sparse_payload(path, S, G, jc, ir, pr):
    """One DErand table from its column-compressed payload (``spade_csc_count`` /
    ``spade_csc_fill`` on the GPU, or ``dense_payload`` on the host)."""
    layout = SparseMatLayout(S, G, int(jc[-1]))
    fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        layout.write_frame(fd, jc)
        if layout.nnz:
            layout.write_entries(fd, 0, ir, pr)
    finally:
        os.close(fd)


def dense_payload(table):
    """``(jc, ir, pr)`` of a dense host table: exact zeros dropped, nan / -inf kept."""
    table = np.asarray(table, dtype=np.float64)
    S, G = table.shape
    tt = np.ascontiguousarray(table.T)                      # [G, S]: one column of the table per row
    mask = tt != 0                                          # nan and -inf are stored, +-0.0 are not
    jc = np.zeros(G + 1, dtype="<i4")
    np.cumsum(np.count_nonzero(mask, axis=1), out=jc[1:])
    ir = np.nonzero(mask)[1].astype("<i4")                  # row index of every stored value, column-major
    pr = tt[mask].astype("<f8", copy=False)
    return jc, ir, pr


def device_payload(table, device=0):
    """``(jc, ir, pr)`` numpy arrays of a CUDA float64 tensor ``[S, G]`` (any row stride), built
    on the GPU: only the stored entries cross PCIe."""
    import ctypes

    from . import _lib
    lib = _lib.load()
    assert table.is_cuda and table.dim() == 2 and table.stride(1) == 1
    S, G = table.shape
    ld = table.stride(0) if S > 1 else max(G, 1)
    tp = ctypes.c_void_p(table.data_ptr())
    jc = np.zeros(G + 1, dtype=np.int32)

    def ok(rc):
        if rc != _lib.SPADE_OK:
            raise RuntimeError("device payload: " + lib.spade_last_error(None).decode())

    ok(lib.spade_csc_count(device, tp, S, G, ld, ctypes.c_void_p(jc.ctypes.data)))
    nnz = int(jc[-1])
    ir, pr = np.empty(nnz, dtype=np.int32), np.empty(nnz, dtype=np.float64)
    ok(lib.spade_csc_fill(device, tp, S, G, ld, ctypes.c_void_p(jc.ctypes.data), ctypes.c_void_p(ir.ctypes.data),
                          ctypes.c_void_p(pr.ctypes.data), _lib.SPADE_HOST))
    return jc, ir, pr


def save_draw_table(path, table):
    """One DErand table: the sparse matrix ``csr_matrix(dense)`` (exact zeros dropped, nan /
    -inf kept) under ``'matrix'`` - what ``DErand.py:230-248`` leaves on disk after the last
    draw.  ``table``: host array, or a CUDA tensor (payload built on the GPU)."""
    if isinstance(table, np.ndarray):
        S, G = table.shape
        jc, ir, pr = dense_payload(table)
    else:
        S, G = table.shape
        jc, ir, pr = device_payload(table, table.device.index)
    save_sparse_payload(path, S, G, jc, ir, pr)


def save_draw_tables(items):
    """``[(path, table), ...]`` written concurrently (the numpy passes release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, len(items))) as pool:
        list(pool.map(lambda pt: save_draw_table(*pt), items))


def save_payloads(items):
    """``[(path, S, G, jc, ir, pr), ...]`` written concurrently."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, len(items))) as pool:
        list(pool.map(lambda it: save_sparse_payload(*it), items))


def write_cell_ids(path, draws):
    """``<i>\\t<repr of the index list>`` per draw (``DErand.py:224-228``): the reference
    joins a one-element list holding the list, so the text is Python's ``str(list(...))``
    of numpy integers."""
    with open(path, "w") as fh:
        for i, d in enumerate(draws):
            fh.write(str(i) + "\t" + ",".join(str(j) for j in [list(d)]) + "\n")
