"""Importable alias of the package directory `robust-audio-deepfake-evolution_b200/` (hyphens
cannot appear in an `import` statement):  `import bimamba_b200 as bm; bm.Mamba(144, 16)`."""
import importlib
import sys

_pkg = importlib.import_module("robust-audio-deepfake-evolution_b200")
sys.modules[__name__] = _pkg
