"""`from world import World, Sphere` (taichi_pathtracer/10_final/world.py:5-60)."""
from learn_path_tracing_b200 import Sphere, World  # noqa: F401
