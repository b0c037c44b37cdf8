"""ctypes binding of libzkfhe_b200.so (the C ABI in include/zkfhe_b200.h).

This is the Python face of the drop-in boundary: plain pointers and sizes, no
torch types.  There is no CPU fallback -- if the library is missing it is
built with nvcc (zk-fhe_b200/build.py); if that fails, or no CUDA device is
present when a context is created, an exception is raised.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "zkfhe_b200.h")
LIB_PATH = os.path.join(HERE, "lib", "libzkfhe_b200.so")

OK = 0
ERR_CUDA, ERR_ARG, ERR_STATE, ERR_ASSERT, ERR_OVERFLOW, ERR_UNSATISFIED = -1, -2, -3, -4, -5, -6
ERROR_NAMES = {ERR_CUDA: "ZKFHE_ERR_CUDA", ERR_ARG: "ZKFHE_ERR_ARG", ERR_STATE: "ZKFHE_ERR_STATE",
               ERR_ASSERT: "ZKFHE_ERR_ASSERT", ERR_OVERFLOW: "ZKFHE_ERR_OVERFLOW",
               ERR_UNSATISFIED: "ZKFHE_ERR_UNSATISFIED"}


class ZkfheError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


def declared_symbols():
    """Every function name declared in include/zkfhe_b200.h."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkfhe_[a-z0-9_]+)\s*\(", text)))


_lib = None

_c = ctypes
_u8p = _c.c_void_p          # raw addresses (host or device) are passed as integers
_SIGNATURES = {
    "zkfhe_init": (_c.c_int, [_c.c_int, _c.POINTER(_c.c_void_p)]),
    "zkfhe_destroy": (None, [_c.c_void_p]),
    "zkfhe_last_error": (_c.c_char_p, [_c.c_void_p]),
    "zkfhe_version": (_c.c_char_p, []),
    "zkfhe_set_stream": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "zkfhe_sync": (_c.c_int, [_c.c_void_p]),
    "zkfhe_set_blocking_sync": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "zkfhe_launch_count": (_c.c_uint64, [_c.c_void_p]),
    "zkfhe_selftest": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_uint64, _c.POINTER(_c.c_uint32)]),
    "zkfhe_pairing_check": (_c.c_int, [_u8p, _u8p, _c.c_uint32, _c.POINTER(_c.c_int)]),
    "zkfhe_pairing": (_c.c_int, [_u8p, _u8p, _c.c_int, _u8p]),
    "zkfhe_srs_g2": (_c.c_int, [_u8p, _u8p]),
    "zkfhe_reference_test_tau": (_c.c_int, [_u8p, _u8p]),
    "zkfhe_point_compress": (_c.c_int, [_u8p, _u8p]),
    "zkfhe_point_decompress": (_c.c_int, [_u8p, _u8p]),
    "zkfhe_vk_export": (_c.c_int, [_c.c_void_p, _u8p, _c.c_size_t, _c.POINTER(_c.c_size_t)]),
    "zkfhe_verify": (_c.c_int, [_c.c_void_p, _u8p, _c.c_size_t, _u8p, _c.c_uint32, _u8p, _c.c_size_t, _u8p, _c.c_int,
                                _c.POINTER(_c.c_int)]),
    "zkfhe_poseidon_permute": (_c.c_int, [_u8p, _c.c_int]),
    "zkfhe_comm_unique_id": (_c.c_int, [_u8p]),
    "zkfhe_comm_init": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _u8p]),
    "zkfhe_comm_destroy": (_c.c_int, [_c.c_void_p]),
    "zkfhe_comm_info": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _c.POINTER(_c.c_int)]),
    "zkfhe_set_virtual_ranks": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "zkfhe_shard_range": (_c.c_int, [_c.c_uint32, _c.c_uint32, _c.c_uint32, _c.POINTER(_c.c_uint32), _c.POINTER(_c.c_uint32)]),
    "zkfhe_host_microbench": (_c.c_int, [_c.c_int, _c.c_uint32, _c.POINTER(_c.c_double), _c.c_char_p, _c.c_size_t]),
    "zkfhe_transcript_replay": (_c.c_int, [_c.c_int, _u8p, _c.c_size_t, _u8p, _c.c_size_t, _c.POINTER(_c.c_size_t)]),
    "zkfhe_pk_export": (_c.c_int, [_c.c_void_p, _u8p, _c.c_size_t, _c.POINTER(_c.c_size_t)]),
    "zkfhe_pk_import": (_c.c_int, [_c.c_void_p, _u8p, _c.c_size_t, _c.POINTER(_c.c_void_p)]),
    "zkfhe_microbench": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_uint32, _c.POINTER(_c.c_float), _c.POINTER(_c.c_uint64)]),
    "zkfhe_ntt_fr": (_c.c_int, [_c.c_void_p, _u8p, _c.c_uint32, _c.c_uint32, _c.c_int, _c.c_int]),
    "zkfhe_ntt_fr_dev": (_c.c_int, [_c.c_void_p, _u8p, _c.c_uint32, _c.c_uint32, _c.c_int, _c.c_int]),
    "zkfhe_coeff_to_extended_dev": (_c.c_int, [_c.c_void_p, _u8p, _c.c_uint32, _u8p, _c.c_uint32, _c.c_uint32]),
    "zkfhe_load_srs": (_c.c_int, [_c.c_void_p, _c.c_uint32, _u8p, _u8p]),
    "zkfhe_srs_setup": (_c.c_int, [_c.c_void_p, _c.c_uint32, _u8p, _u8p, _u8p]),
    "zkfhe_fr_convert_dev": (_c.c_int, [_c.c_void_p, _u8p, _c.c_uint64, _c.c_int]),
    "zkfhe_msm_g1": (_c.c_int, [_c.c_void_p, _u8p, _c.c_uint32, _c.c_int, _u8p]),
    "zkfhe_msm_g1_dev": (_c.c_int, [_c.c_void_p, _u8p, _c.c_uint32, _c.c_int, _u8p]),
    "zkfhe_msm_g1_dev_ex": (_c.c_int, [_c.c_void_p, _u8p, _c.c_uint32, _c.c_int, _c.c_int, _u8p]),
    "zkfhe_last_kernel_ms": (_c.c_float, [_c.c_void_p]),
    "zkfhe_timing_reset": (_c.c_int, [_c.c_void_p]),
    "zkfhe_timing_get": (_c.c_int, [_c.c_void_p, _c.c_int, _c.POINTER(_c.c_float), _c.POINTER(_c.c_uint32),
                                    _c.POINTER(_c.c_uint64)]),
    # stage (1a): Poly
    "zkfhe_poly_from_u64": (_c.c_int, [_c.c_void_p, _u8p, _c.c_uint32, _c.c_uint64, _c.POINTER(_c.c_void_p)]),
    "zkfhe_poly_from_decimal": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_size_t, _c.c_uint32, _c.c_uint64, _c.POINTER(_c.c_void_p)]),
    "zkfhe_poly_from_u256": (_c.c_int, [_c.c_void_p, _u8p, _c.c_uint32, _c.c_uint64, _c.POINTER(_c.c_void_p)]),
    "zkfhe_poly_mul": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p)]),
    "zkfhe_poly_reduce_by_modulus": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_uint64, _c.POINTER(_c.c_void_p)]),
    "zkfhe_poly_divide_by_cyclo": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_uint64,
                                              _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p)]),
    "zkfhe_poly_len": (_c.c_uint32, [_c.c_void_p]),
    "zkfhe_poly_max_bits": (_c.c_uint64, [_c.c_void_p]),
    "zkfhe_poly_download": (_c.c_int, [_c.c_void_p, _c.c_void_p, _u8p]),
    "zkfhe_poly_free": (None, [_c.c_void_p]),
    "zkfhe_status": (_c.c_int, [_c.c_void_p]),
    # stage (1b): PolyChip witness
    "zkfhe_witness_new": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.POINTER(_c.c_void_p)]),
    "zkfhe_witness_free": (None, [_c.c_void_p]),
    "zkfhe_witness_reset": (_c.c_int, [_c.c_void_p]),
    "zkfhe_chip_from_poly": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_void_p]),
    "zkfhe_chip_load_constant": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_uint64, _c.c_void_p]),
    "zkfhe_chip_to_public": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "zkfhe_chip_set_challenge": (_c.c_int, [_c.c_void_p, _u8p]),
    "zkfhe_chip_constrain_mul": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_uint32, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "zkfhe_chip_add": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "zkfhe_chip_scalar_mul": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_void_p, _c.c_uint64, _c.c_void_p]),
    "zkfhe_chip_reduce_by_cyclo": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_uint32, _c.c_void_p, _c.c_void_p,
                                              _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_uint64, _c.c_void_p]),
    "zkfhe_chip_reduce_by_modulo": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_uint64, _c.c_void_p]),
    "zkfhe_chip_constrain_equality": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_void_p]),
    "zkfhe_chip_constrain_coefficients_in_range": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_uint64, _c.c_uint64]),
    "zkfhe_chip_constrain_from_distribution_chi_key": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_uint64]),
    "zkfhe_chip_constrain_coefficients_in_modulus_field": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_uint64]),
    "zkfhe_chip_safe_trim_leading_zeroes": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_uint32, _c.c_void_p]),
    "zkfhe_witness_set_recording": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "zkfhe_witness_download_structure": (_c.c_int, [_c.c_void_p, _c.c_uint32, _u8p, _u8p]),
    "zkfhe_witness_download_lookup_sources": (_c.c_int, [_c.c_void_p, _u8p]),
    "zkfhe_witness_public_cells": (_c.c_int, [_c.c_void_p, _u8p]),
    "zkfhe_witness_mock": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_uint64), _c.POINTER(_c.c_uint64)]),
    "zkfhe_keygen": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_uint32, _c.POINTER(_c.c_void_p)]),
    "zkfhe_pk_free": (None, [_c.c_void_p]),
    "zkfhe_pk_pinning_json": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_size_t, _c.POINTER(_c.c_size_t)]),
    "zkfhe_pk_info": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_uint32)]),
    "zkfhe_pk_download_fixed": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_uint32, _u8p]),
    "zkfhe_pk_fixed_commitments": (_c.c_int, [_c.c_void_p, _u8p]),
    "zkfhe_prove_begin": (_c.c_int, [_c.c_void_p, _c.c_void_p, _u8p, _c.c_int, _c.POINTER(_c.c_void_p)]),
    "zkfhe_share_srs": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "zkfhe_prove_phase0": (_c.c_int, [_c.c_void_p, _c.c_void_p, _u8p]),
    "zkfhe_prove_finish": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_size_t)]),
    "zkfhe_prove_reset": (_c.c_int, [_c.c_void_p, _u8p]),
    "zkfhe_prover_round_ms": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_double)]),
    "zkfhe_prover_free": (None, [_c.c_void_p]),
    "zkfhe_proof_free": (None, [_c.c_void_p]),
    "zkfhe_witness_counts": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_uint64), _c.POINTER(_c.c_uint64), _c.POINTER(_c.c_uint64)]),
    "zkfhe_witness_download": (_c.c_int, [_c.c_void_p, _c.c_uint32, _u8p]),
    "zkfhe_witness_device_ptr": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.POINTER(_c.c_void_p)]),
}


class AssignedPoly(_c.Structure):
    """zkfhe_assigned_poly (= PolyChip { assigned_coefficients, max_num_bits, degree })."""
    _fields_ = [("ctx_id", _c.c_uint32), ("stride", _c.c_uint32), ("base", _c.c_uint64), ("len", _c.c_uint32),
                ("reserved", _c.c_uint32), ("max_num_bits", _c.c_uint64)]


class Cell(_c.Structure):
    """zkfhe_cell (= AssignedValue)."""
    _fields_ = [("ctx_id", _c.c_uint32), ("reserved", _c.c_uint32), ("offset", _c.c_uint64)]


FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def fr_mont_bytes(x):
    """Canonical int -> the 32-byte ABI layout of an Fr element (Montgomery, R = 2^256)."""
    return bytearray((((int(x) % FR_MODULUS) << 256) % FR_MODULUS).to_bytes(32, "little"))


def reference_test_tau():
    """The trapdoor of the reference's own fallback SRS (halo2-scaffold `gen_srs` without a params file:
    `ParamsKZG::setup(k, ChaCha20Rng::from_seed([0; 32]))`), as a canonical int.  INSECURE: tests and benchmarks."""
    out = bytearray(32)
    rc = load_library().zkfhe_reference_test_tau(_addr(out), None)
    if rc != 0:
        raise ZkfheError(rc, "zkfhe_reference_test_tau failed")
    return int.from_bytes(out, "little") * pow(1 << 256, -1, FR_MODULUS) % FR_MODULUS


def load_library(build_if_missing=True):
    """dlopen the in-tree library (building it first if needed) and type its symbols."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise FileNotFoundError(f"{LIB_PATH} not built; run python zk-fhe_b200/build.py")
        import importlib.util
        spec = importlib.util.spec_from_file_location("_zkfhe_build", os.path.join(HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _addr(buf):
    """Address of a host buffer (bytes-like / numpy) or pass through an int device pointer."""
    if isinstance(buf, int):
        return buf
    if buf is None:
        return None
    if hasattr(buf, "ctypes"):                      # numpy
        return buf.ctypes.data
    if isinstance(buf, (bytearray, memoryview)):
        return ctypes.addressof((ctypes.c_char * len(buf)).from_buffer(buf))
    if isinstance(buf, bytes):
        return ctypes.cast(ctypes.c_char_p(buf), ctypes.c_void_p).value
    raise TypeError(type(buf))


def comm_unique_id():
    """128 bytes identifying a new NCCL communicator (ncclGetUniqueId); create on one rank, broadcast to the others."""
    out = bytearray(128)
    rc = load_library().zkfhe_comm_unique_id(_addr(out))
    if rc != OK:
        raise ZkfheError(rc, "zkfhe_comm_unique_id failed (is libnccl.so.2 loadable?)")
    return bytes(out)


def shard_range(count, n_ranks, rank):
    """[lo, hi): the block of `count` items shard `rank` of `n_ranks` owns."""
    lo, hi = ctypes.c_uint32(), ctypes.c_uint32()
    rc = load_library().zkfhe_shard_range(count, n_ranks, rank, ctypes.byref(lo), ctypes.byref(hi))
    if rc != OK:
        raise ZkfheError(rc, "zkfhe_shard_range: bad arguments")
    return lo.value, hi.value


class Context:
    """One GPU, one stream.  Mirrors the zkfhe_ctx lifetime."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.zkfhe_init(device, ctypes.byref(h))
        if rc != OK:
            raise ZkfheError(rc, f"zkfhe_init(device={device}) failed: no usable CUDA device (no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.zkfhe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise ZkfheError(rc, self.lib.zkfhe_last_error(self.h).decode())

    # -- plumbing -----------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._check(self.lib.zkfhe_set_stream(self.h, cuda_stream))

    def sync(self):
        self._check(self.lib.zkfhe_sync(self.h))

    def set_blocking_sync(self, on):
        """Host waits inside the library sleep on a blocking event (True, default) or spin (False)."""
        self._check(self.lib.zkfhe_set_blocking_sync(self.h, int(bool(on))))

    def status(self):
        """Synchronise and raise if a data-dependent reference assert fired on the device."""
        self._check(self.lib.zkfhe_status(self.h))

    def launch_count(self):
        return int(self.lib.zkfhe_launch_count(self.h))

    def last_kernel_ms(self):
        return float(self.lib.zkfhe_last_kernel_ms(self.h))

    def timing_reset(self):
        self._check(self.lib.zkfhe_timing_reset(self.h))

    def timing(self, category):
        """(ms, spans, units) accumulated since timing_reset for a kernel category
        (0 MSM accumulate, 1 NTT passes, 2 MSM sort + reduce)."""
        ms, sp, un = ctypes.c_float(), ctypes.c_uint32(), ctypes.c_uint64()
        self._check(self.lib.zkfhe_timing_get(self.h, category, ctypes.byref(ms), ctypes.byref(sp), ctypes.byref(un)))
        return float(ms.value), int(sp.value), int(un.value)

    def selftest(self, n_cases=1 << 16, seed=1):
        bad = ctypes.c_uint32(0)
        self._check(self.lib.zkfhe_selftest(self.h, n_cases, seed, ctypes.byref(bad)))
        return bad.value

    def microbench(self, kind, iters):
        """(ms, ops) of an arithmetic micro-benchmark: kind 0 Montgomery-product throughput on a full
        GPU; 1..5 one-warp dependent chains (XYZZ add, mixed add, product, inversion, Fermat inversion)."""
        ms, ops = ctypes.c_float(), ctypes.c_uint64()
        self._check(self.lib.zkfhe_microbench(self.h, kind, iters, ctypes.byref(ms), ctypes.byref(ops)))
        return float(ms.value), int(ops.value)

    # -- one proof over several GPUs ---------------------------------------------
    def comm_init(self, rank, n_ranks, unique_id):
        """Collective: bind an NCCL communicator to this context (unique_id: 128 bytes from comm_unique_id on one rank)."""
        assert len(unique_id) == 128
        self._check(self.lib.zkfhe_comm_init(self.h, rank, n_ranks, _addr(bytes(unique_id))))

    def comm_destroy(self):
        self._check(self.lib.zkfhe_comm_destroy(self.h))

    def comm_info(self):
        """(rank, n_ranks, virtual_ranks)"""
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self.lib.zkfhe_comm_info(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def set_virtual_ranks(self, n_ranks):
        """Testing on one GPU: compute the shards of an `n_ranks`-way sharded proof one after the other here."""
        self._check(self.lib.zkfhe_set_virtual_ranks(self.h, n_ranks))

    # -- stage (3) ----------------------------------------------------------
    def ntt_fr(self, data, log_n, batch, inverse=False, coset=False):
        """In place on a host buffer (bytearray / numpy uint8) of batch*2^log_n*32 bytes."""
        self._check(self.lib.zkfhe_ntt_fr(self.h, _addr(data), log_n, batch, int(inverse), int(coset)))

    def ntt_fr_dev(self, d_ptr, log_n, batch, inverse=False, coset=False):
        self._check(self.lib.zkfhe_ntt_fr_dev(self.h, d_ptr, log_n, batch, int(inverse), int(coset)))

    def coeff_to_extended_dev(self, d_coeffs, log_n_in, d_ext, log_n_out, batch):
        self._check(self.lib.zkfhe_coeff_to_extended_dev(self.h, d_coeffs, log_n_in, d_ext, log_n_out, batch))

    # -- stage (2) ----------------------------------------------------------
    def load_srs(self, k, g=None, g_lagrange=None):
        self._check(self.lib.zkfhe_load_srs(self.h, k, _addr(g), _addr(g_lagrange)))
        self.srs_k = k

    def srs_setup(self, k, tau, want_host_copy=False):
        """Insecure test SRS from an explicit tau (canonical int); returns (g, g_lagrange) bytes if asked."""
        t = fr_mont_bytes(tau)
        g = bytearray(64 << k) if want_host_copy else None
        gl = bytearray(64 << k) if want_host_copy else None
        self._check(self.lib.zkfhe_srs_setup(self.h, k, _addr(t), _addr(g), _addr(gl)))
        self.srs_k = k
        return (bytes(g), bytes(gl)) if want_host_copy else None

    def srs_g2(self, tau):
        """[tau]_2 of the test SRS (128 bytes, Montgomery): what `verify` pairs against."""
        t = fr_mont_bytes(tau)
        out = bytearray(128)
        self._check(self.lib.zkfhe_srs_g2(_addr(t), _addr(out)))
        return bytes(out)

    def share_srs(self, other):
        """Use `other`'s resident commitment-key tables (same GPU); `other` must stay alive."""
        self._check(self.lib.zkfhe_share_srs(self.h, other.h))
        self.srs_k = other.srs_k
        self._srs_owner = other

    def fr_convert_dev(self, d_ptr, count, to_montgomery=True):
        self._check(self.lib.zkfhe_fr_convert_dev(self.h, d_ptr, count, int(to_montgomery)))

    def msm_g1(self, scalars, batch, basis=1):
        out = bytearray(64 * batch)
        self._check(self.lib.zkfhe_msm_g1(self.h, _addr(scalars), batch, basis, _addr(out)))
        return bytes(out)

    def msm_g1_dev(self, d_scalars, batch, basis, d_out, small_values=False):
        if small_values:
            self._check(self.lib.zkfhe_msm_g1_dev_ex(self.h, d_scalars, batch, basis, 1, d_out))
        else:
            self._check(self.lib.zkfhe_msm_g1_dev(self.h, d_scalars, batch, basis, d_out))
