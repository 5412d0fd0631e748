"""Importable alias for the `physics-based-rl_b200/` package (the directory name has a hyphen)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("physics-based-rl_b200")
