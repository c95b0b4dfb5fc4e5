"""Drop-in for taichi_pathtracer/3_adding_a_sphere/__main__.py: run as `python compat/taichi_pathtracer/3_adding_a_sphere` from the repo root."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _driver import main  # noqa: E402

main("3_adding_a_sphere")
