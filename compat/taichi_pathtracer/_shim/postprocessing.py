"""`from postprocessing import ACES_tonemapping, gamma_correction` (10_final/postprocessing.py:5-29): host versions (the
device versions run inside pt_postprocess when the script's post_processing() kernel is called)."""
from learn_path_tracing_b200 import ACES_tonemapping, gamma_correction  # noqa: F401
