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


def test_ctypes_mirror_matches_the_header_field_by_field(tmp_path):
    """The three descriptor structs of include/vggheads_b200.h, compiled by gcc, against the ctypes mirror in _lib.py:
    size and the offset of every field (a field added on one side only would silently shift everything after it)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = {"vgh_buf_desc": _lib.BufDesc, "vgh_op_desc": _lib.OpDesc, "vgh_net_desc": _lib.NetDesc}
    lines = []
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vggheads_b200.h"\nint main(void) {\n' + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, field, value = line.split()
        cls = structs[cname]
        want = ctypes.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(value) == want, (cname, field, int(value), want)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_silent_cpu_fallback():
    import head_detector_b200
    from head_detector_b200.flame import FLAMELayer

    with pytest.raises(RuntimeError):
        head_detector_b200.HeadDetector()
    with pytest.raises(RuntimeError):
        FLAMELayer().handle()
