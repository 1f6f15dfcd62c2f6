"""The C-ABI library loads without a GPU and exports every symbol include/vggheads_b200.h declares;
the product refuses to run without a CUDA device instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

from head_detector_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    header = open(os.path.join(ROOT, "include", "vggheads_b200.h")).read()
    declared = set(re.findall(r"\b(vgh_[a-z_0-9]+)\s*\(", header))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTS) == declared


def test_version_and_error_string():
    assert _lib.lib().vgh_version() == 1
    assert _lib.lib().vgh_last_error() is not None


def test_struct_sizes_match_c_layout():
    assert ctypes.sizeof(_lib.BufDesc) == 20
    assert ctypes.sizeof(_lib.OpDesc) == 18 * 4 + 2 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.NetDesc) % 8 == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_silent_cpu_fallback():
    import head_detector_b200
    from head_detector_b200.flame import FLAMELayer

    with pytest.raises(RuntimeError):
        head_detector_b200.HeadDetector()
    with pytest.raises(RuntimeError):
        FLAMELayer().handle()
