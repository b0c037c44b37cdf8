#!/usr/bin/env python3
"""Host arithmetic behind the transcript on this machine: ns per Fr product, us per Poseidon permutation (no GPU needed)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_fhe_b200  # noqa: E402

lib = zk_fhe_b200.load_library()
for kind, name, iters in ((1, "dependent Fr product", 2000000), (0, "Poseidon permutation (optimised)", 20000), (2, "Poseidon permutation (plain)", 10000)):
    ns, buf = ctypes.c_double(), ctypes.create_string_buffer(64)
    assert lib.zkfhe_host_microbench(kind, iters, ctypes.byref(ns), buf, 64) == 0
    print(f"{name:48s} {ns.value:10.1f} ns   [{buf.value.decode()}]")
