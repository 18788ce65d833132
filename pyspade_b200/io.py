"""Input / output formats of DEobs and DErand.

Inputs (``DEobs.py:47-50,69-72``): genes x cells and sgRNAs x cells DataFrames as ``.pkl``
(``pd.read_pickle``) or ``.h5`` key ``'df'`` (``pd.read_hdf``, needs PyTables).
Outputs: MATLAB v5 files written by ``scipy.io.savemat`` WITHOUT a ``.mat`` extension,
variable name ``'matrix'`` (``DEobs.py:208-221``, ``DErand.py:230-248``); ``.npy`` vectors;
the ``rand_cell_ID.txt`` listing.  Readers downstream: ``utils.load_data``
(``utils.py:343-361``), ``local.py:81-110``, ``global.py:84-111``.
"""
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


def save_vector(path, vec):
    """One DEobs table row: 1-D float64 under ``'matrix'`` (loads back as ``(1, G)``)."""
    sio.savemat(path, {"matrix": np.ascontiguousarray(vec, dtype=np.float64)})


def save_draw_table(path, table):
    """One DErand table: ``csr_matrix(dense)`` (exact zeros dropped, nan / -inf kept) under
    ``'matrix'`` - what ``DErand.py:230-248`` leaves on disk after the last draw."""
    sio.savemat(path, {"matrix": sp.csr_matrix(table)})


def write_cell_ids(path, draws):
    """``<i>\\t<repr of the index list>`` per draw (``DErand.py:224-228``): the reference
    joins a one-element list holding the list, so the text is Python's ``str(list(...))``
    of numpy integers."""
    with open(path, "w") as fh:
        for i, d in enumerate(draws):
            fh.write(str(i) + "\t" + ",".join(str(j) for j in [list(d)]) + "\n")
