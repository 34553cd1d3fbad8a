"""Comparison helpers shared by the tests."""
import numpy as np


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-for-bit equality, except that any NaN equals any NaN (payloads are not part of the contract)."""
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    ia = a.view(np.uint32 if a.dtype == np.float32 else np.uint64)
    ib = b.view(np.uint32 if b.dtype == np.float32 else np.uint64)
    return bool(np.array_equal(ia[~na], ib[~nb]))


def describe_mismatch(a, b):
    na = np.isnan(a) | np.isnan(b)
    d = np.where(na, 0, np.abs(a.astype(np.float64) - b.astype(np.float64)))
    i = np.unravel_index(np.argmax(d), d.shape)
    return f"max|d|={d.max():.3e} at {i}: {a[i]!r} vs {b[i]!r}; nan mismatch={np.sum(np.isnan(a) != np.isnan(b))}"


def max_rel(a, b):
    """max|a-b| / max|b|  -- the per-tensor tolerance form of SURVEY.md section 8d."""
    fin = np.isfinite(a) & np.isfinite(b)
    scale = np.max(np.abs(b[fin])) if fin.any() else 1.0
    scale = scale if scale > 0 else 1.0
    return float(np.max(np.abs(a[fin].astype(np.float64) - b[fin].astype(np.float64))) / scale) if fin.any() else 0.0
