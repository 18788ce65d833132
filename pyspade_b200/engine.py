"""Python face of one staged matrix on one GPU (``spade_engine`` of ``include/spade.h``).

``Engine.run`` is the batched form of the reference's ``perform_DE`` /
``perform_DE_NC`` (``/root/reference/pySpade/utils.py:296-341``): every gene
against S cell sets in one call.
"""
import ctypes

import numpy as np
import scipy.sparse as sp

from . import _lib

TABLES = ("up", "down", "fc", "cpm")


class SpadeError(RuntimeError):
    pass


def _ptr(arr):
    if arr is None:
        return None
    return ctypes.c_void_p(arr.ctypes.data)


class PinnedArray:
    """numpy view of page-locked host memory from ``spade_host_alloc``."""

    def __init__(self, shape, dtype):
        lib = _lib.load()
        self._lib = lib
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = ctypes.c_void_p()
        rc = lib.spade_host_alloc(ctypes.byref(p), nbytes)
        if rc != _lib.SPADE_OK:
            raise SpadeError(lib.spade_last_error(None).decode())
        self._p = p
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p is not None:
            self.array = None
            self._lib.spade_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One genes x cells matrix staged in the HBM of one GPU."""

    def __init__(self, handle, device):
        self._lib = _lib.load()
        self._h = handle
        self.device = device
        G, C, nnz = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        f = ctypes.c_int32()
        self._check(self._lib.spade_dims(self._h, ctypes.byref(G), ctypes.byref(C),
                                         ctypes.byref(nnz), ctypes.byref(f)))
        self.G, self.C, self.nnz, self.genes_per_part = G.value, C.value, nnz.value, f.value
        self.n_nc = 0
        self.background_key = None
        # out-of-range and duplicated member cells are caught on the device (range check in the
        # test kernel, adjacent-equal check after the member sort) unless the sort is disabled
        self.device_validation = True

    # ------------------------------------------------------------------ construction
    @staticmethod
    def _csr_parts(matrix, value_dtype):
        m = sp.csr_matrix(matrix)
        if not m.has_canonical_format:
            m = m.copy()
            m.sum_duplicates()
        indptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(m.indices, dtype=np.int32)
        data = np.ascontiguousarray(m.data, dtype=value_dtype)
        return m.shape, indptr, indices, data

    @classmethod
    def _create(cls, fn_name, device, G, C, indptr_p, indices_p, data_p, location):
        lib = _lib.load()
        h = ctypes.c_void_p()
        rc = getattr(lib, fn_name)(int(device), int(G), int(C), indptr_p, indices_p, data_p,
                                   int(location), ctypes.byref(h))
        if rc != _lib.SPADE_OK:
            raise SpadeError(f"{fn_name} failed ({rc}): {lib.spade_last_error(None).decode()}")
        return cls(h, device)

    @classmethod
    def from_cpm(cls, matrix, device=0):
        """Stage already-normalised float64 values (the ``input_array`` of
        ``perform_DE``, utils.py:296)."""
        (G, C), indptr, indices, data = cls._csr_parts(matrix, np.float64)
        return cls._create("spade_create", device, G, C, _ptr(indptr), _ptr(indices), _ptr(data),
                           _lib.SPADE_HOST)

    @classmethod
    def from_counts(cls, counts, device=0):
        """Stage raw integer counts; CPM normalisation (utils.py:150-155) runs on the GPU."""
        (G, C), indptr, indices, data = cls._csr_parts(counts, np.int32)
        return cls._create("spade_create_from_counts", device, G, C, _ptr(indptr), _ptr(indices),
                           _ptr(data), _lib.SPADE_HOST)

    @classmethod
    def from_device_csr(cls, G, C, indptr, indices, data, device=0):
        """CSR already resident on the GPU as torch tensors (int64 / int32 / float64 or
        int32 counts)."""
        import torch

        assert indptr.dtype == torch.int64 and indices.dtype == torch.int32
        fn = "spade_create_from_counts" if data.dtype == torch.int32 else "spade_create"
        if fn == "spade_create":
            assert data.dtype == torch.float64
        torch.cuda.synchronize(device)
        return cls._create(fn, device, G, C, ctypes.c_void_p(indptr.data_ptr()),
                           ctypes.c_void_p(indices.data_ptr()), ctypes.c_void_p(data.data_ptr()),
                           _lib.SPADE_DEVICE)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != _lib.SPADE_OK:
            msg = self._lib.spade_last_error(self._h).decode()
            if "duplicate cell index inside a cell set" in msg:
                raise ValueError(msg)                  # what the host-side validation raises
            if "member cell index out of range" in msg:
                raise IndexError(msg)
            raise SpadeError(f"libspade_b200 error {rc}: {msg}")

    def close(self):
        if self._h is not None:
            self._lib.spade_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ knobs
    def set_option(self, name, value):
        """``tail_table`` (0 direct / 1 auto / 2 always-when-uniform), ``member_block``
        (cells per warp before draining; larger sets take the split path),
        ``genes_per_part`` (re-stages the layout).  See ``spade_set_option``."""
        opt = {"tail_table": _lib.SPADE_OPT_TAIL_TABLE, "member_block": _lib.SPADE_OPT_MEMBER_BLOCK,
               "genes_per_part": _lib.SPADE_OPT_GENES_PER_PART,
               "sort_members": _lib.SPADE_OPT_SORT_MEMBERS}[name]
        self._check(self._lib.spade_set_option(self._h, opt, int(value)))
        if name == "sort_members":
            self.device_validation = bool(value)
        if name == "genes_per_part":
            f = ctypes.c_int32()
            self._check(self._lib.spade_dims(self._h, None, None, None, ctypes.byref(f)))
            self.genes_per_part = f.value

    # ------------------------------------------------------------------ background
    def set_background(self, nc_idx=None):
        """``None`` / empty: complement background (utils.py:194-215); otherwise the
        negative-control cells of ``hypergeo_test_NC_sparse`` (utils.py:248-266)."""
        if nc_idx is None or len(nc_idx) == 0:
            self._check(self._lib.spade_set_background(self._h, None, 0))
            self.n_nc = 0
            self.background_key = None
            return
        key = np.ascontiguousarray(np.asarray(nc_idx), dtype=np.int64)
        nc = key.astype(np.int32)
        self._check(self._lib.spade_set_background(self._h, _ptr(nc), nc.shape[0]))
        self.n_nc = int(nc.shape[0])
        self.background_key = key.tobytes()

    # ------------------------------------------------------------------ parity probes
    def gene_cutoffs(self):
        med = np.empty(self.G, dtype=np.float64)
        n = np.empty(self.G, dtype=np.int32)
        self._check(self._lib.spade_gene_cutoffs(self._h, _ptr(med), _ptr(n)))
        return med, n

    def cpm_values(self):
        vals = np.empty(self.nnz, dtype=np.float64)
        self._check(self._lib.spade_cpm_values(self._h, _ptr(vals)))
        return vals

    # ------------------------------------------------------------------ the batched test
    @staticmethod
    def pack_sets(sets):
        offsets = np.zeros(len(sets) + 1, dtype=np.int64)
        if len(sets):
            offsets[1:] = np.cumsum([len(s) for s in sets])
        if offsets[-1] > 0:
            members = np.concatenate([np.asarray(s, dtype=np.int64).ravel() for s in sets])
        else:
            members = np.zeros(0, dtype=np.int64)
        return members, offsets

    def _validate_members(self, members, offsets):
        if members.shape[0] == 0:
            return
        if members.min() < 0 or members.max() >= self.C:
            raise IndexError("cell index out of range in a cell set")
        # N counts duplicates while k de-duplicates in the reference (utils.py:203,211);
        # draws and set()-derived region lists never hold duplicates, so they are refused.
        set_id = np.repeat(np.arange(offsets.shape[0] - 1, dtype=np.int64), np.diff(offsets))
        key = set_id * np.int64(self.C) + members
        if np.unique(key).shape[0] != key.shape[0]:
            raise ValueError("duplicate cell index inside a cell set")

    def run(self, sets=None, members=None, offsets=None, want_k=False, out=None, validate=True,
            tables=TABLES, row_stride=None):
        """Test every gene against each cell set.

        ``sets``: list of index arrays, or pass packed ``members`` + ``offsets``.
        Returns a dict of float64 arrays ``[S, G]`` (``up, down, fc, cpm``; plus int32 ``k``
        when ``want_k``).  ``out`` may hold preallocated arrays (e.g. pinned) to fill.  With
        ``row_stride`` (elements) the arrays in ``out`` are strided views ``[S, G]`` of a larger
        host table (row i of this call = row i of the view), e.g. a table shared by the ranks of
        a node (:class:`pyspade_b200.shard.SharedHostTables`)."""
        if sets is not None:
            members, offsets = self.pack_sets(sets)
        members = np.asarray(members)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        S = offsets.shape[0] - 1
        if validate and not self.device_validation:
            self._validate_members(members.astype(np.int64, copy=False), offsets)
        elif validate and members.shape[0] and (members.min() < -(1 << 31) or members.max() >= (1 << 31)):
            raise IndexError("cell index out of range in a cell set")
        members = np.ascontiguousarray(members, dtype=np.int32)
        res = {}
        for name in tables:
            if out is not None and name in out:
                arr = out[name]
                assert arr.dtype == np.float64 and arr.shape == (S, self.G)
                if row_stride is None:
                    assert arr.flags.c_contiguous
                else:
                    assert S <= 1 or arr.strides == (8 * row_stride, 8)
            else:
                arr = np.empty((S, self.G), dtype=np.float64)
            res[name] = arr
        if want_k:
            res["k"] = (out["k"] if out is not None and "k" in out
                        else np.empty((S, self.G), dtype=np.int32))
        args = [self._h, _ptr(members), _lib.SPADE_HOST, _ptr(offsets), S,
                _ptr(res.get("up")), _ptr(res.get("down")), _ptr(res.get("fc")), _ptr(res.get("cpm")),
                _ptr(res.get("k")), _lib.SPADE_HOST]
        if row_stride is None:
            self._check(self._lib.spade_run(*args, None))
        else:
            if want_k and S > 1:
                assert res["k"].strides == (4 * row_stride, 4)
            self._check(self._lib.spade_run_strided(*args, int(row_stride), None))
        return res

    def run_device(self, members, offsets_host, out, k=None, stream=None, row_stride=None):
        """Device-resident form: ``members`` int32 torch tensor on this GPU,
        ``offsets_host`` int64 numpy, ``out`` dict of float64 torch tensors ``[S, G]`` - or raw
        device addresses (``int``), e.g. into a table that lives on another GPU and was opened
        with :class:`pyspade_b200.shard.PeerTables`.  ``row_stride`` (elements between the rows
        of consecutive sets, default G) lets several ranks interleave their rows in one table.
        Ordered on ``stream`` (a ``torch.cuda.Stream``; default: torch's current stream);
        returns without synchronising."""
        import torch

        offsets = np.ascontiguousarray(offsets_host, dtype=np.int64)
        S = offsets.shape[0] - 1
        if stream is None:
            stream = torch.cuda.current_stream(self.device)

        def dp(t, dtype=None):
            if t is None:
                return None
            if isinstance(t, int):
                return ctypes.c_void_p(t)
            assert t.is_cuda and t.is_contiguous()
            if dtype is not None:
                assert t.dtype == dtype and tuple(t.shape) == (S, self.G)
            return ctypes.c_void_p(t.data_ptr())

        assert members.dtype == torch.int32
        args = [self._h, dp(members), _lib.SPADE_DEVICE, _ptr(offsets), S,
                dp(out.get("up"), torch.float64), dp(out.get("down"), torch.float64),
                dp(out.get("fc"), torch.float64), dp(out.get("cpm"), torch.float64), dp(k, torch.int32)]
        if row_stride is None:
            self._check(self._lib.spade_run(*args, _lib.SPADE_DEVICE, ctypes.c_void_p(stream.cuda_stream)))
        else:
            self._check(self._lib.spade_run_strided(*args, _lib.SPADE_DEVICE, int(row_stride),
                                                    ctypes.c_void_p(stream.cuda_stream)))

    def run_raw(self, members, offsets_host, raw, row_stride=None, stream=None):
        """Raw form (``spade_run_raw``): accumulator words, 8 bytes per test, into ``raw`` - a
        uint64/int64 torch tensor ``[S, G]`` or a device address (e.g. on a peer GPU)."""
        import torch

        offsets = np.ascontiguousarray(offsets_host, dtype=np.int64)
        S = offsets.shape[0] - 1
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        rp = ctypes.c_void_p(raw if isinstance(raw, int) else raw.data_ptr())
        assert members.dtype == torch.int32
        self._check(self._lib.spade_run_raw(self._h, ctypes.c_void_p(members.data_ptr()), _lib.SPADE_DEVICE,
                                            _ptr(offsets), S, rp, int(row_stride or self.G),
                                            ctypes.c_void_p(stream.cuda_stream)))

    def expand(self, raw, set_sizes, out, k=None, raw_row_stride=None, row_stride=None, stream=None):
        """``spade_expand``: finish the tests of ``len(set_sizes)`` rows of raw words into the
        tables of ``out`` (torch tensors or device addresses)."""
        import torch

        sizes = np.ascontiguousarray(set_sizes, dtype=np.int64)
        if stream is None:
            stream = torch.cuda.current_stream(self.device)

        def dp(t):
            if t is None:
                return None
            return ctypes.c_void_p(t if isinstance(t, int) else t.data_ptr())

        self._check(self._lib.spade_expand(self._h, dp(raw), sizes.shape[0], int(raw_row_stride or self.G),
                                           _ptr(sizes), dp(out.get("up")), dp(out.get("down")),
                                           dp(out.get("fc")), dp(out.get("cpm")), dp(k),
                                           int(row_stride or self.G), ctypes.c_void_p(stream.cuda_stream)))

    def run_raw_blocks(self, members, offsets_host, blocks, first_row=0, row_step=1, stream=None):
        """``spade_run_raw_blocks``: raw words routed by gene block.  ``members``: int32 torch
        tensor on this GPU, or a numpy array (uploaded by the call).  ``blocks``: list of
        ``(device address, ld, gene_begin, gene_end)``, one per owner, tiling ``[0, G)``."""
        import torch

        offsets = np.ascontiguousarray(offsets_host, dtype=np.int64)
        S = offsets.shape[0] - 1
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        arr = (_lib.GeneBlock * len(blocks))()
        for i, (base, ld, g0, g1) in enumerate(blocks):
            arr[i].base, arr[i].ld, arr[i].gene_begin, arr[i].gene_end = int(base), int(ld), int(g0), int(g1)
        if isinstance(members, np.ndarray):
            members = np.ascontiguousarray(members, dtype=np.int32)
            mp, loc = _ptr(members), _lib.SPADE_HOST
        else:
            assert members.dtype == torch.int32
            mp, loc = ctypes.c_void_p(members.data_ptr()), _lib.SPADE_DEVICE
        self._check(self._lib.spade_run_raw_blocks(self._h, mp, loc, _ptr(offsets), S, arr, len(blocks),
                                                   int(first_row), int(row_step),
                                                   ctypes.c_void_p(stream.cuda_stream)))

    def expand_block(self, raw, set_sizes, gene_begin, gene_count, out, k=None, raw_row_stride=None,
                     row_stride=None, stream=None):
        """``spade_expand_block``: finish ``len(set_sizes)`` rows x ``gene_count`` words (column j
        = gene ``gene_begin + j``) into the tables of ``out`` (torch tensors or device addresses)."""
        import torch

        sizes = np.ascontiguousarray(set_sizes, dtype=np.int64)
        if stream is None:
            stream = torch.cuda.current_stream(self.device)

        def dp(t):
            if t is None:
                return None
            return ctypes.c_void_p(t if isinstance(t, int) else t.data_ptr())

        self._check(self._lib.spade_expand_block(
            self._h, dp(raw), sizes.shape[0], int(raw_row_stride or gene_count), int(gene_begin), int(gene_count),
            _ptr(sizes), dp(out.get("up")), dp(out.get("down")), dp(out.get("fc")), dp(out.get("cpm")), dp(k),
            int(row_stride or gene_count), ctypes.c_void_p(stream.cuda_stream)))

    def check(self):
        """``spade_check``: wait for the device and raise what the runs issued so far flagged
        (``IndexError`` for a member cell out of range, ``ValueError`` for a duplicated cell) -
        ``run_device`` / ``run_raw`` return before their kernels have run."""
        self._check(self._lib.spade_check(self._h))

    @property
    def wide_genes(self):
        """Number of genes whose means take the float64 side path (``spade_wide_genes``)."""
        n = ctypes.c_int64()
        self._check(self._lib.spade_wide_genes(self._h, ctypes.byref(n)))
        return n.value

    def timing(self, reset=False):
        """Kernel times since the last reset (``spade_timing``): ``test_ms`` = the test kernels,
        of which ``finish_ms`` in the finishing launch of two-phase runs; ``table_ms`` = the rest."""
        a, f, n, x = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
        self._check(self._lib.spade_timing_finish(self._h, ctypes.byref(x)))
        self._check(self._lib.spade_timing(self._h, ctypes.byref(a), ctypes.byref(f),
                                           ctypes.byref(n), 1 if reset else 0))
        return dict(test_ms=a.value, table_ms=f.value, finish_ms=x.value, launches=n.value)


def hypergeom_logcdf_logsf(k, M, n, N, device=0):
    """K3 in isolation on the GPU (``spade_hypergeom``): scipy semantics, no nan guard."""
    lib = _lib.load()
    shape = np.broadcast_shapes(np.shape(k), np.shape(M), np.shape(n), np.shape(N))
    arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.int64), shape)).ravel()
            for a in (k, M, n, N)]
    count = arrs[0].shape[0]
    lc = np.empty(count, dtype=np.float64)
    ls = np.empty(count, dtype=np.float64)
    rc = lib.spade_hypergeom(int(device), _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _ptr(arrs[3]),
                             count, _ptr(lc), _ptr(ls))
    if rc != _lib.SPADE_OK:
        raise SpadeError(f"spade_hypergeom failed ({rc}): {lib.spade_last_error(None).decode()}")
    return lc.reshape(shape), ls.reshape(shape)


def gamma_fit(table, sign=-1.0, device=0, with_evals=False):
    """``spade_gamma_fit``: per-gene ``scipy.stats.gamma.fit`` of ``sign * table[:, g]`` (finite
    values only) on the GPU.  ``table``: float64 ``[S, G]`` numpy array (or CUDA torch tensor).
    Returns ``(A, B, C, status)``, plus ``evals`` (likelihood evaluations per gene) on request."""
    lib = _lib.load()
    if isinstance(table, np.ndarray):
        t = np.ascontiguousarray(table, dtype=np.float64)
        S, G = t.shape
        ptr, loc = _ptr(t), _lib.SPADE_HOST
    else:
        assert table.is_cuda and table.dim() == 2 and table.stride(1) == 1
        S, G = table.shape
        ptr, loc = ctypes.c_void_p(table.data_ptr()), _lib.SPADE_DEVICE
    ld = G if isinstance(table, np.ndarray) or S <= 1 else table.stride(0)
    A, B, C = np.zeros(G), np.zeros(G), np.zeros(G)
    status = np.zeros(G, dtype=np.int32)
    evals = np.zeros(G, dtype=np.int32)
    rc = lib.spade_gamma_fit(int(device), ptr, loc, S, G, ld, float(sign), _ptr(A), _ptr(B), _ptr(C), _ptr(status),
                             _ptr(evals))
    if rc != _lib.SPADE_OK:
        raise SpadeError(f"spade_gamma_fit failed ({rc}): {lib.spade_last_error(None).decode()}")
    return (A, B, C, status, evals) if with_evals else (A, B, C, status)


def column_mean(table, device=0):
    """``spade_column_mean``: mean over the rows of every column of a CUDA float64 tensor
    ``[S, G]`` (``Cpm_mean`` of ``DErand.py:252``).  Returns numpy ``[G]``."""
    lib = _lib.load()
    assert table.is_cuda and table.dim() == 2 and table.stride(1) == 1
    S, G = table.shape
    out = np.zeros(G, dtype=np.float64)
    if G == 0:
        return out
    ld = table.stride(0) if S > 1 else max(G, 1)
    rc = lib.spade_column_mean(int(device), ctypes.c_void_p(table.data_ptr()), S, G, ld, _ptr(out))
    if rc != _lib.SPADE_OK:
        raise SpadeError(f"spade_column_mean failed ({rc}): {lib.spade_last_error(None).decode()}")
    return out
