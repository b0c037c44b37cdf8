"""Shared helpers for the GPU parity tests (host-side conversions only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cbind, field  # noqa: E402


def fr_to_mont_array(vals, mod=field.R_MOD):
    """ints -> (n,4) uint64 Montgomery limbs (the ABI layout)."""
    return cbind.ints_to_u64x4([field.to_mont(v % mod, mod) for v in vals])


def mont_array_to_fr(arr, mod=field.R_MOD):
    return [field.from_mont_fast(v, mod) for v in cbind.u64x4_to_ints(arr)]


def random_fr_mont(rng, count):
    """Uniform-ish Fr elements directly in Montgomery layout (any value < r is a valid Montgomery residue)."""
    a = rng.integers(0, 1 << 63, size=(count, 4), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(count, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)      # < 2^252 < r
    return np.ascontiguousarray(a)


_SRS_CACHE = {}


def toy_srs(k, tau=0x5EED5EED5EED5EED5EED5EED):
    """(g, g_lagrange) as (n,8) uint64 arrays from the C oracle; tau is public: tests only."""
    key = (k, tau)
    if key not in _SRS_CACHE:
        _SRS_CACHE[key] = cbind.srs(k, tau)
    return _SRS_CACHE[key]
