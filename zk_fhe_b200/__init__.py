"""Importable alias of the `zk-fhe_b200/` package (a hyphen cannot be imported)."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "zk-fhe_b200"))
from .capi import Context, ZkfheError, declared_symbols, load_library, reference_test_tau  # noqa: E402,F401
