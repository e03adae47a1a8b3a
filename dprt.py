"""Alias: ``import dprt`` == the package ``pg2024-data-parallel-ray-tracing_b200`` (its name is not an identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
sys.modules[__name__] = _pkg
