"""CPU-side checks of the drop-in boundary: the library builds, loads and exports
every symbol include/zkfhe_b200.h declares; without a GPU it fails loudly."""
import ctypes

import pytest


def test_library_exports_every_declared_symbol():
    import zk_fhe_b200
    lib = zk_fhe_b200.load_library()
    names = zk_fhe_b200.declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/zkfhe_b200.h but not exported"
    assert b"sm_100a" in lib.zkfhe_version()


def test_binding_types_every_declared_symbol():
    import zk_fhe_b200
    from zk_fhe_b200 import capi
    assert sorted(capi._SIGNATURES) == zk_fhe_b200.declared_symbols()


def test_no_cpu_fallback_without_gpu():
    import torch
    import zk_fhe_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zk_fhe_b200.ZkfheError) as e:
        zk_fhe_b200.Context(0)
    assert e.value.code == zk_fhe_b200.capi.ERR_CUDA
    lib = zk_fhe_b200.load_library()
    h = ctypes.c_void_p()
    assert lib.zkfhe_init(0, ctypes.byref(h)) == zk_fhe_b200.capi.ERR_CUDA and not h.value


def test_binary_euclid_inversion_on_the_host(tmp_path):
    """csrc/inv_bin.cuh is `__host__ __device__`: the exact code the kernels run (field inversion off the IMAD
    pipe) compiled for the host and checked against Python big-int inverses for both BN254 moduli."""
    import os
    import random
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "inv_host_test"
    subprocess.run([nvcc, "-O2", "-Wno-deprecated-gpu-targets", "-o", str(exe), os.path.join(root, "tools", "inv_host_test.cu")],
                   check=True, capture_output=True)
    R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
    rng = random.Random(7)
    cases = [(a, p) for p in (R, P)
             for a in [1, 2, 3, p - 1, p - 2, (p + 1) // 2, 1 << 253, (1 << 128) + 1] + [rng.randrange(1, p) for _ in range(500)]]
    cases.append((0, R))
    out = subprocess.run([str(exe)], input="\n".join(f"{a:x} {p:x}" for a, p in cases), capture_output=True, text=True,
                         check=True).stdout.split()
    assert len(out) == len(cases)
    for (a, p), o in zip(cases, out):
        assert int(o, 16) == (pow(a, -1, p) if a else 0)
