"""Importable alias of the `pb-llm_b200/` package directory (a hyphen cannot be written in an
import statement)."""
import importlib
import sys

_real = importlib.import_module("pb-llm_b200")
sys.modules[__name__] = _real
