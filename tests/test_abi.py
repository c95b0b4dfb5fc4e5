"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/pt_api.h declares, struct layouts match the header, and there is no CPU fallback."""
import ctypes
import os
import re

import pytest

import learn_path_tracing_b200 as L
from learn_path_tracing_b200 import _lib
from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pt_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.API_SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libb200pt.so not built (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert _lib.load().pt_version() == 1


def test_struct_sizes_match_header():
    assert ctypes.sizeof(_lib.PtMaterial) == 32
    assert ctypes.sizeof(_lib.PtCamera) == 64
    assert ctypes.sizeof(_lib.PtRenderParams) == 64
    assert ctypes.sizeof(_lib.PtStats) == 80
    assert _lib.MATERIAL_DTYPE.itemsize == 32


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(L.PtError):
        _lib.Context(0)
    with pytest.raises(L.PtError):
        L.default_context()


def test_product_package_never_touches_the_oracle():
    """The package AND the drop-in layer above it (compat/: shim + drivers) never import, load or execute the oracle."""
    for pkg in (os.path.join(ROOT, "learn_path_tracing_b200"), os.path.join(ROOT, "compat")):
        for dirpath, _, files in os.walk(pkg):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "ptoracle" not in text and "libptoracle" not in text, f
                    assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
