"""Restatement of ``scipy.stats.hypergeom.logcdf`` / ``logsf`` (scipy 1.18.1).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The p-value arithmetic of the hot path is not in ``/root/reference``: the
reference calls ``stats.hypergeom.logcdf(k, M, n, N)`` and ``.logsf`` at
``pySpade/utils.py:212-213`` (NC variant ``:263-264``).  scipy is an un-vendored,
un-pinned dependency (``setup.py`` has no ``install_requires``; README lists
"Scipy (<=1.6.2)"); the operative version is the one in this image, 1.18.1.
Its published algorithm, restated here:

* ``scipy/stats/_distn_infrastructure.py:3625-3663`` (``rv_discrete.logcdf``) and
  ``:3703-3741`` (``logsf``): support ``a = max(N-(M-n), 0)``, ``b = min(n, N)``;
  ``logcdf`` is ``-inf`` for ``k < a``, ``0`` for ``k >= b``; ``logsf`` is ``0`` for
  ``k < a``, ``-inf`` for ``k >= b``; invalid shape arguments give ``nan``.
* ``scipy/stats/_discrete_distns.py:708-731`` (``hypergeom_gen._logsf/_logcdf``):
  the tail on the far side of the mean is summed directly with ``logsumexp`` of
  ``_logpmf``; the near side is ``log1p(-exp(other tail))``.  Branch tests:
  ``logcdf`` takes the complement when ``(k+.5)(M+.5) > (n-.5)(N-.5)``, ``logsf``
  when ``(k+.5)(M+.5) < (n-.5)(N-.5)``.
* ``:669-675`` (``_logpmf``): ``ln C(n,k) + ln C(M-n,N-k) - ln C(M,N)`` written
  with ``betaln``; here with ``lgamma`` (same quantity, rounding noise ~1e-10
  absolute at M = 200 000, far inside the 1e-6 relative contract).

Pure Python / numpy, meant for small cases and as the readable statement; the
bulk checker is the C form in ``hypergeom_oracle.c``.
"""
import math

import numpy as np


def _lnchoose(a: int, b: int) -> float:
    if b < 0 or b > a:
        return -math.inf
    return math.lgamma(a + 1) - math.lgamma(b + 1) - math.lgamma(a - b + 1)


def logpmf(k: int, M: int, n: int, N: int) -> float:
    """``hypergeom_gen._logpmf`` (_discrete_distns.py:669-675)."""
    return _lnchoose(n, k) + _lnchoose(M - n, N - k) - _lnchoose(M, N)


def _logsumexp(vals) -> float:
    vals = np.asarray(vals, dtype=np.float64)
    if vals.size == 0:
        return -math.inf
    m = float(np.max(vals))
    if not math.isfinite(m):
        return m
    return m + math.log(float(np.sum(np.exp(vals - m))))


def _argcheck(M, n, N) -> bool:
    return (M > 0) and (n >= 0) and (N >= 0) and (n <= M) and (N <= M)


def _above_mean(k, M, n, N) -> bool:
    return (k + 0.5) * (M + 0.5) > (n - 0.5) * (N - 0.5)


def _below_mean(k, M, n, N) -> bool:
    return (k + 0.5) * (M + 0.5) < (n - 0.5) * (N - 0.5)


def logcdf(k: int, M: int, n: int, N: int) -> float:
    """ln P(X <= k), X ~ Hypergeom(M total, n marked, N drawn)."""
    k, M, n, N = int(k), int(M), int(n), int(N)
    if not _argcheck(M, n, N):
        return math.nan
    a, b = max(N - (M - n), 0), min(n, N)
    if k >= b:
        return 0.0
    if k < a:
        return -math.inf
    if _above_mean(k, M, n, N):
        return math.log1p(-math.exp(logsf(k, M, n, N)))
    return _logsumexp([logpmf(j, M, n, N) for j in range(0, k + 1)])


def logsf(k: int, M: int, n: int, N: int) -> float:
    """ln P(X > k)."""
    k, M, n, N = int(k), int(M), int(n), int(N)
    if not _argcheck(M, n, N):
        return math.nan
    a, b = max(N - (M - n), 0), min(n, N)
    if k < a:
        return 0.0
    if k >= b:
        return -math.inf
    if _below_mean(k, M, n, N):
        return math.log1p(-math.exp(logcdf(k, M, n, N)))
    return _logsumexp([logpmf(j, M, n, N) for j in range(k + 1, N + 1)])


def pvalues_like_reference(k: int, M: int, n: int, N: int):
    """``(down, up)`` exactly as ``utils.py:212-213`` would produce them.

    The reference returns ``nan`` for both tails as soon as any of
    ``k, M, n, N`` is zero (``all([k, M, n, N])`` guard).
    """
    if not all([k, M, n, N]):
        return math.nan, math.nan
    return logsf(k, M, n, N), logcdf(k, M, n, N)
