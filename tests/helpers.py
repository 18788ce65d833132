"""Comparison helpers shared by the CPU and GPU parity tests."""
import numpy as np

LOGP_RTOL = 1e-6     # north_star: log p-values within 1e-6 relative of scipy
VALUE_RTOL = 1e-6    # north_star: fold changes (and mean CPM) within 1e-6 relative


def parse_float(x):
    if isinstance(x, str):
        return {"nan": np.nan, "inf": np.inf, "-inf": -np.inf}[x]
    return float(x)


def max_rel_err(got, want):
    """nan / +-inf / exact-zero positions must coincide; returns the max relative error
    over the remaining finite entries."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "nan positions differ"
    assert np.array_equal(np.isposinf(got), np.isposinf(want)), "+inf positions differ"
    assert np.array_equal(np.isneginf(got), np.isneginf(want)), "-inf positions differ"
    assert np.array_equal(got == 0.0, want == 0.0), "exact-zero positions differ"
    m = np.isfinite(want) & (want != 0.0)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(got[m] - want[m]) / np.abs(want[m])))


def assert_tables_close(got, want, names=("up", "down", "fc", "cpm")):
    for name in names:
        tol = LOGP_RTOL if name in ("up", "down") else VALUE_RTOL
        err = max_rel_err(got[name], want[name])
        assert err <= tol, f"{name}: max relative error {err:.3e} > {tol:g}"


def sets_from_packed(members, offsets):
    return [members[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]
