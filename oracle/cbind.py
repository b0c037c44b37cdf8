"""ctypes loader for oracle/c (the C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libzkfhe_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "c", "zkfhe_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        # -march=native is resolved on the machine that runs the build; the GPU box may differ,
        # so fall back to a portable build there if the prebuilt file fails to load.
        subprocess.run(["make", "-C", os.path.join(HERE, "c"), "-B", "-s"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        vp, u32, u64, ci = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
        L.orc_field_mul.argtypes = [ci, vp, vp, vp]
        L.orc_field_inv.argtypes = [ci, vp, vp]
        L.orc_to_mont.argtypes = [ci, vp, vp]
        L.orc_from_mont.argtypes = [ci, vp, vp]
        L.orc_to_mont_array.argtypes = [ci, vp, vp, u64]
        L.orc_ntt.argtypes = [vp, u32, u32, ci, ci, ci]
        L.orc_msm.argtypes = [vp, vp, u32, u32, vp, ci]
        L.orc_srs.argtypes = [u32, vp, vp, vp, ci]
        L.orc_poly_mul.argtypes = [vp, vp, u32, vp]
        L.orc_poly_reduce.argtypes = [vp, u32, u64, vp]
        L.orc_divide_by_cyclo.argtypes = [vp, u32, vp, u32, u64, vp, vp]
        L.orc_divide_by_cyclo.restype = ci
        L.orc_num_threads.restype = ci
        L.orc_from_mont_array.argtypes = [ci, vp, vp, u64]
        L.orc_bfv_witness.argtypes = [vp, u32, u64, u64, u64, u32, vp, vp, u64, vp, u64, vp, u64, vp, u64, vp]
        L.orc_bfv_witness.restype = ci
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data


def ints_to_u64x4(vals):
    """list of ints (< 2^256) -> (n,4) uint64 little-endian limbs."""
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def u64x4_to_ints(arr):
    b = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def field_mul(which, a, b):
    A, B = ints_to_u64x4([a]), ints_to_u64x4([b])
    R = np.zeros((1, 4), np.uint64)
    lib().orc_field_mul(which, _p(R), _p(A), _p(B))
    return u64x4_to_ints(R)[0]


def to_mont(which, arr):
    out = np.empty_like(arr)
    L = lib()
    for i in range(arr.shape[0]):
        L.orc_to_mont(which, out[i].ctypes.data, arr[i].ctypes.data)
    return out


def ntt(data_mont, log_n, batch, inverse=False, coset=False, threads=0):
    """data_mont: (batch*n, 4) uint64 Montgomery, transformed in place."""
    assert data_mont.dtype == np.uint64 and data_mont.flags.c_contiguous
    lib().orc_ntt(_p(data_mont), log_n, batch, int(inverse), int(coset), threads)
    return data_mont


def msm(scalars_mont, bases_mont, n, batch, threads=0):
    out = np.zeros((batch, 8), np.uint64)
    lib().orc_msm(_p(scalars_mont), _p(bases_mont), n, batch, _p(out), threads)
    return out


def srs(k, tau, want_g=True, want_gl=True, threads=0):
    n = 1 << k
    t = ints_to_u64x4([tau])
    g = np.zeros((n, 8), np.uint64) if want_g else None
    gl = np.zeros((n, 8), np.uint64) if want_gl else None
    lib().orc_srs(k, _p(t), _p(g) if want_g else None, _p(gl) if want_gl else None, threads)
    return g, gl


def poly_mul(a, b):
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    n = len(a)
    out = np.zeros((2 * n - 1, 2), np.uint64)
    lib().orc_poly_mul(_p(a), _p(b), n, _p(out))
    return [int(lo) | (int(hi) << 64) for lo, hi in out]


def poly_reduce(vals, q):
    arr = np.array([[v & (2**64 - 1), v >> 64] for v in vals], np.uint64)
    out = np.zeros(len(vals), np.uint64)
    lib().orc_poly_reduce(_p(arr), len(vals), q, _p(out))
    return [int(x) for x in out]


def divide_by_cyclo(dividend, cyclo, q):
    d = np.ascontiguousarray(dividend, np.uint64)
    c = np.ascontiguousarray(cyclo, np.uint64)
    deg = len(c) - 1
    quo = np.zeros(deg + 1, np.uint64)
    rem = np.zeros(2 * deg + 1, np.uint64)
    rc = lib().orc_divide_by_cyclo(_p(d), len(d), _p(c), len(c), q, _p(quo), _p(rem))
    if rc != 0:
        from .poly import OracleError
        raise OracleError("divide_by_cyclo: reference panics on this input")
    return [int(x) for x in quo], [int(x) for x in rem]


def bfv_witness(inp, N, Q, T, B, gamma, lookup_bits=8, caps=None):
    """Stage (1) on the CPU in C (orc_bfv_witness): `inp` is a bfv.in dict (decimal strings or ints).  Returns
    (adv0, adv1, adv2, lookups) as (cells, 4) uint64 Montgomery arrays -- the flat advice of the phase-0 gate,
    phase-1 gate and phase-1 RLC contexts and the cells_to_lookup values, in creation order."""
    from .bfv import INPUT_KEYS
    from .poly import OracleError
    arrs = [np.array([int(x) for x in inp[k]], dtype=np.uint64) for k in INPUT_KEYS]
    ptrs = (ctypes.c_void_p * 9)(*[a.ctypes.data for a in arrs])
    if caps is None:       # generous bounds from the cell-cost model: <= 1300 cells per input coefficient
        caps = (32 * N, 1300 * N, 40 * N, 320 * N)
    bufs = [np.zeros((c, 4), np.uint64) for c in caps]
    g = ints_to_u64x4([gamma * (1 << 256) % int("30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001", 16)])
    counts = np.zeros(4, np.uint64)
    rc = lib().orc_bfv_witness(ctypes.addressof(ptrs), N, Q, T, B, lookup_bits, _p(g), _p(bufs[0]), caps[0], _p(bufs[1]), caps[1],
                               _p(bufs[2]), caps[2], _p(bufs[3]), caps[3], _p(counts))
    if rc == -1:
        raise OracleError("stage (1): the reference asserts / panics on this input")
    if rc != 0:
        raise RuntimeError(f"orc_bfv_witness: buffer too small or inconsistent witness (rc={rc})")
    return tuple(b[:int(c)] for b, c in zip(bufs, counts))


def from_mont_array(arr, which=0):
    out = np.empty_like(arr)
    lib().orc_from_mont_array(which, _p(arr), _p(out), arr.shape[0])
    return out
