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
