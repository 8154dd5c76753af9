"""Shared test helpers: product-model -> oracle-model conversion and the parity metrics."""
import numpy as np

import chromatography_b200 as cb
from oracle import oracle as O

# north_star tolerances
PROFILE_RTOL = 1e-9   # max relative error on concentration profiles
MOMENT_RTOL = 1e-7    # retention time and peak area
TINY = 1e-290         # below this both sides are in / next to the subnormal range: compare absolutely


def to_oracle(model):
    return O.from_model(model)


def max_rel_err(a, b) -> float:
    """max |a-b|/|b| over entries with |b| > TINY; entries at/below TINY must agree to TINY absolutely."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    big = np.abs(b) > TINY
    small_ok = np.all(np.abs(a[~big] - b[~big]) <= TINY)
    assert small_ok, "mismatch in the subnormal tail"
    if not big.any():
        return 0.0
    return float(np.max(np.abs(a[big] - b[big]) / np.abs(b[big])))


def assert_bit_equal(a, b, what=""):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = a.view(np.uint64) == b.view(np.uint64)
    if not same.all():
        idx = np.argwhere(~same)
        i = tuple(idx[0])
        raise AssertionError(f"{what}: {idx.shape[0]} of {a.size} values differ bitwise; first at {i}: "
                             f"{a[i]!r} vs {b[i]!r}")
