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
