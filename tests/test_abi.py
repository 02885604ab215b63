"""The C-ABI library loads on a CPU-only box, exports every symbol include/mpx.h declares, mirrors the header's
structs, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "mpx.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mpx_[a-z_0-9]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(libmpx):
    from mpopt_b200 import _lib

    names = _declared()
    assert len(names) >= 20
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes and include/mpx.h disagree"
    for n in names:
        assert getattr(libmpx, n) is not None
    assert libmpx.mpx_version() == 100


def test_struct_layout_matches_header(tmp_path):
    from mpopt_b200 import _lib

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mpx.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   "sizeof(mpx_phase_desc), sizeof(mpx_problem_desc), offsetof(mpx_problem_desc, scale_t),"
                   "offsetof(mpx_problem_desc, program_source), offsetof(mpx_phase_desc, cost_t));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.PhaseDesc), ctypes.sizeof(_lib.ProblemDesc), _lib.ProblemDesc.scale_t.offset,
            _lib.ProblemDesc.program_source.offset, _lib.PhaseDesc.cost_t.offset]
    assert got == want


def test_no_cpu_fallback(libmpx):
    """Without a CUDA device plan creation and table computation fail loudly instead of computing on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mpopt_b200 import _lib
    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import moon_lander

    with pytest.raises(_lib.MpxError) as e:
        Transcription(moon_lander(), 4, 3, "LGR")
    assert e.value.code == _lib.MPX_ENODEVICE
    r = np.zeros(4)
    assert libmpx.mpx_collocation_tables(0, 3, -1.0, 1.0, 0, _lib.ptr(r), None, None, None) == _lib.MPX_ENODEVICE


def test_argument_validation_precedes_device_use(libmpx):
    from mpopt_b200 import _lib

    assert libmpx.mpx_plan_create(None, None) == _lib.MPX_EINVAL
    assert libmpx.mpx_collocation_tables(7, 3, -1.0, 1.0, 0, None, None, None, None) == _lib.MPX_EINVAL
    assert libmpx.mpx_collocation_tables(0, 0, -1.0, 1.0, 0, None, None, None, None) == _lib.MPX_ELIMIT
    assert b"degree" in libmpx.mpx_last_error()
    assert libmpx.mpx_sizes(None, None, None, None, None) == _lib.MPX_EINVAL


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from mpopt_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()
