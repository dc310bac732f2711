"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol
include/mitgcm_b200.h declares (no compute calls without a GPU)."""
import os

import pytest

from mitgcm_b200 import _lib, build


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    L = _lib.lib()
    missing = [f for f in _lib.declared_functions() if not hasattr(L, f)]
    assert not missing, missing


def test_enum_ids_are_unique_per_class():
    e = _lib.ENUMS
    for prefix in ("MG_", "MP_", "MI_"):
        vals = [v for k, v in e.items() if k.startswith(prefix) and not k.endswith("_END") and k not in ("MG_N2D", "MP_ND")]
        assert len(vals) == len(set(vals))


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mitgcm_b200 import runtime
    from mitgcm_b200.grid import Dims
    with pytest.raises(runtime.B200Error):
        runtime.init(Dims(8, 8, 2, 2))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under mitgcm_b200/ may import, load or execute it."""
    import os
    import re
    root = os.path.join(os.path.dirname(__file__), "..", "mitgcm_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt, f
