"""Drop-in for taichi_pathtracer/4_objects/__main__.py: run as `python compat/taichi_pathtracer/4_objects` from the repo root."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _driver import main  # noqa: E402

main("4_objects")
