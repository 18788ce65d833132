"""Significance scoring of observed regions against the DErand background - the arithmetic
``pySpade global`` / ``pySpade local`` run on the tables this package produces
(``/root/reference/pySpade/global.py:107-187``, ``local.py:104-177``), batched over regions and
genes on the GPU (``spade_score``).  Annotation joins, hit filtering by genomic window and the
CSV layout of those commands are host bookkeeping and stay out of scope.

============================  =====================================================
here                          reference
============================  =====================================================
``preprocess_pvals``          ``global.py:118-123`` / ``local.py:106-114``
``candidates``                ``global.py:111,124-131``
``significance``              ``gamma.logsf`` per hit ``global.py:151-171`` ("Significance_score")
                              and ``sum(rand[:, g] < p) / iter`` ``global.py:177-187``
                              ("pval-empirical")
============================  =====================================================
"""
import ctypes

import numpy as np

from . import _lib
from .engine import SpadeError, _ptr


def preprocess_pvals(p):
    """0 -> 1, +-inf -> 0, nan -> 0, in the reference's order (a copy)."""
    p = np.array(p, dtype=np.float64, copy=True)
    p[p == 0] = 1
    p[np.isinf(p)] = 0
    p[np.isnan(p)] = 0
    return p


def candidates(p_up, p_down, cpm, cpm_mean, fc_cutoff=0.0, pval_cutoff=0.0):
    """Masks of the genes ``global`` goes on to score: ``(up, down, fc_by_rand_dist_cpm)`` with
    ``fc = (cpm + 0.01) / (cpm_mean + 0.01)`` (``global.py:111``).  ``p_*`` preprocessed."""
    fc_cpm = (np.asarray(cpm) + 0.01) / (np.asarray(cpm_mean) + 0.01)
    up = (fc_cpm > 1 + fc_cutoff) & (np.asarray(p_up) < pval_cutoff)
    down = (fc_cpm < 1 - fc_cutoff) & (np.asarray(p_down) < pval_cutoff)
    return up, down, fc_cpm


def _loc(x):
    if x is None:
        return None, _lib.SPADE_HOST
    if isinstance(x, np.ndarray):
        return _ptr(x), _lib.SPADE_HOST
    assert x.is_cuda and x.is_contiguous()
    return ctypes.c_void_p(x.data_ptr()), _lib.SPADE_DEVICE


def significance(obs, rand=None, A=None, B=None, C=None, device=0):
    """``obs``: raw (not preprocessed) log p-values ``[R, G]`` of R observed regions (a DEobs
    table direction); ``rand``: the DErand table ``[S, G]`` of the same direction; ``A, B, C``:
    its gamma parameters ``[G]``.  numpy arrays or CUDA torch tensors (obs / rand).
    Returns ``(score, emp)`` as numpy ``[R, G]`` (``None`` for the one whose inputs are missing):
    ``score = gamma.logsf(-p, A, B, C)``, ``emp = #{rand[:, g] < p} / S``."""
    lib = _lib.load()
    if isinstance(obs, np.ndarray):
        obs = np.ascontiguousarray(np.atleast_2d(obs), dtype=np.float64)
    R, G = obs.shape
    if isinstance(rand, np.ndarray):
        rand = np.ascontiguousarray(rand, dtype=np.float64)
    S = 0 if rand is None else rand.shape[0]
    if rand is not None:
        assert rand.shape[1] == G
    want_score = A is not None
    abc = [np.ascontiguousarray(v, dtype=np.float64) for v in (A, B, C)] if want_score else [None] * 3
    score = np.empty((R, G)) if want_score else None
    emp = np.empty((R, G)) if rand is not None else None
    op, oloc = _loc(obs)
    rp, rloc = _loc(rand)
    rc = lib.spade_score(int(device), rp, rloc, S, G, G, op, oloc, R, G, _ptr(abc[0]), _ptr(abc[1]), _ptr(abc[2]),
                         _ptr(score), _ptr(emp), _lib.SPADE_HOST, G)
    if rc != _lib.SPADE_OK:
        raise SpadeError(f"spade_score failed ({rc}): {lib.spade_last_error(None).decode()}")
    return score, emp
