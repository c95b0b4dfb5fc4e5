"""python compat/taichi_pathtracer/10_final — drop-in for the reference's taichi_pathtracer/10_final (same idiom through the shim:
compat/taichi_pathtracer/_shim_driver.py; the reference's own script also runs unmodified with PYTHONPATH=.../_shim)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _shim_driver import main  # noqa: E402

main("10_final")
