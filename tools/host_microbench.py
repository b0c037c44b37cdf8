#!/usr/bin/env python3
"""Host arithmetic behind the transcript on this machine: ns per Fr product, us per Poseidon permutation (no GPU needed)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_fhe_b200  # noqa: E402

lib = zk_fhe_b200.load_library()
KINDS = ((1, "dependent Fr product", 2000000), (0, "Poseidon permutation (as the transcript runs it)", 20000),
         (3, "Poseidon permutation (scalar, optimised form)", 20000), (2, "Poseidon permutation (scalar, plain form)", 10000),
         (4, "Poseidon permutation (AVX-512 IFMA form)", 20000))
for rep in range(3):
    for kind, name, iters in KINDS:
        ns, buf = ctypes.c_double(), ctypes.create_string_buffer(64)
        rc = lib.zkfhe_host_microbench(kind, iters, ctypes.byref(ns), buf, 64)
        if rc != 0:
            print(f"{name:52s} unavailable on this CPU")
            continue
        print(f"{name:52s} {ns.value:10.1f} ns   [{buf.value.decode()}]")
