"""zk-fhe_b200: B200-native `prove` hot path for zk-fhe's BFV encryption circuit.

The directory name carries a hyphen (it mirrors the reference's name), so the
importable alias is `zk_fhe_b200` (a three-line package at the repo root whose
__path__ points here).

  capi     ctypes binding of the C ABI (include/zkfhe_b200.h)
  build    nvcc build of lib/libzkfhe_b200.so for sm_100a
  csrc/    hand-written CUDA: field/curve arithmetic, NTT, MSM, witness kernels
"""
from .capi import Context, ZkfheError, declared_symbols, load_library, reference_test_tau  # noqa: F401
