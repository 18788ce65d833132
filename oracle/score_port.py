"""Restatement of the significance scoring that ``pySpade global`` / ``pySpade local`` apply to
the DEobs tables against the DErand background (row f4 of the scope table).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference lines followed (``/root/reference/pySpade``):
  global.py:118-123, local.py:106-114   p: 0 -> 1, +-inf -> 0, nan -> 0 (in this order)
  global.py:111                          fc_cpm = (cpm + 0.01) / (cpm_mean + 0.01)
  local.py:150                           fc_rand = cpm / cpm_mean
  global.py:127-131                      candidates: fc_cpm > 1 & p_up < 0 ; fc_cpm < 1 & p_down < 0
  global.py:151-171, local.py:153-165    score = scipy.stats.gamma.logsf(-p, A, B, C)
  global.py:177-187, local.py:167-177    emp = sum(rand[:, g] < p[g]) / iter_num, rand = the saved
                                         sparse table made dense (unsaved entries are 0.0)

The arithmetic of ``gamma.logsf`` lives in scipy (un-vendored, unpinned; 1.18.1 in this image):
``rv_continuous.logsf`` (scipy/stats/_distn_infrastructure.py) - nan for invalid shape / scale,
0 at or below the lower end of the support, -inf at +inf, else the generic ``_logsf``:
``log(sf(y))`` above the median and ``log1p(-cdf(y))`` below, with ``_sf = gammaincc`` and
``_cdf = gammainc`` (cephes igamc / igam).  ``incomplete_gamma_pair`` restates the regularised
incomplete gamma functions with the textbook pair - power series of P below a + 1, continued
fraction of Q above - independent of scipy.special; ``tests/test_score.py`` pins it on
scipy's values (``tests/golden/score_case.npz``, written by ``tests/golden/make_score_golden.py``).
"""
import math

import numpy as np


def preprocess(p):
    p = np.array(p, dtype=np.float64, copy=True)
    p[p == 0] = 1
    p[np.isinf(p)] = 0
    p[np.isnan(p)] = 0
    return p


MAXLOG = 7.09782712893383996843e2       # cephes: exp() underflow guard of the prefactor


def incomplete_gamma_pair(a, x):
    """``(P(a, x), Q(a, x))`` for scalars a > 0, x >= 0: the side computed directly keeps its
    relative accuracy (P from the series below a + 1, Q from the continued fraction above)."""
    if x == 0.0:
        return 0.0, 1.0
    if math.isinf(x):
        return 1.0, 0.0
    lfac = a * math.log(x) - x - math.lgamma(a)
    if x < a + 1.0:
        term = 1.0 / a
        total = term
        ap = a
        for _ in range(100000):
            ap += 1.0
            term *= x / ap
            total += term
            if term < total * 1e-17:
                break
        p = total * math.exp(lfac)
        return p, 1.0 - p
    if lfac < -MAXLOG:                      # cephes igam_fac: the prefactor underflows -> Q = 0
        return 1.0, 0.0
    tiny = 1e-300
    b = x + 1.0 - a
    c = 1.0 / tiny
    d = 1.0 / b
    h = d
    for i in range(1, 100000):
        an = -i * (i - a)
        b += 2.0
        d = an * d + b
        if abs(d) < tiny:
            d = tiny
        c = b + an / c
        if abs(c) < tiny:
            c = tiny
        d = 1.0 / d
        delta = d * c
        h *= delta
        if abs(delta - 1.0) < 1e-16:
            break
    q = math.exp(lfac) * h
    return 1.0 - q, q


def gamma_logsf_port(x, a, loc, scale):
    """``scipy.stats.gamma.logsf(x, a, loc, scale)`` for scalars.  scipy 1.18's generic
    ``rv_continuous._logsf``: ``log(sf)`` above the median, ``log1p(-cdf)`` at or below it
    (above the median <=> cdf > 1/2)."""
    if not (a > 0.0) or not (scale > 0.0) or math.isinf(a) or not math.isfinite(loc) or math.isinf(scale):
        return math.nan
    if math.isnan(x):
        return math.nan
    y = (x - loc) / scale
    if y <= 0.0:
        return 0.0
    if math.isinf(y):
        return -math.inf
    p, q = incomplete_gamma_pair(a, y)
    if p > 0.5:
        return math.log(q) if q > 0.0 else -math.inf
    return math.log1p(-p)


def score_tables(obs, rand, A, B, C):
    """``obs`` [R, G] raw log p of observed regions, ``rand`` [S, G] the DErand table of the same
    direction, A/B/C [G].  Returns ``(p, score, emp)``, each [R, G]."""
    p = preprocess(obs)
    R, G = p.shape
    S = rand.shape[0]
    score = np.empty((R, G))
    for r in range(R):
        for g in range(G):
            score[r, g] = gamma_logsf_port(-p[r, g], A[g], B[g], C[g])
    with np.errstate(invalid="ignore"):
        emp = np.stack([np.sum(rand < p[r][None, :], axis=0) / S for r in range(R)])
    return p, score, emp


def candidates(p_up, p_down, cpm, cpm_mean):
    """global.py:111,124-131: boolean masks (up, down) of the genes that go on to scoring."""
    fc_cpm = (cpm + 0.01) / (cpm_mean + 0.01)
    return (fc_cpm > 1) & (p_up < 0), (fc_cpm < 1) & (p_down < 0), fc_cpm
