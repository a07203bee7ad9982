"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/jaadb200.h declares, and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jaadb200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jaadb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from jaadec_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), "libjaadb200.so does not export %s" % s
    assert sorted(_lib.SYMBOLS) == syms, "jaadec_b200/_lib.py and include/jaadb200.h disagree"
    assert lib.jaadb_abi_version() == 1


def test_header_compiles_as_c():
    """The boundary is a plain C header: no C++ or torch types in the signatures."""
    code = '#include "jaadb200.h"\nint main(void){ jaadb_options o; (void)o; return sizeof(jaadb_frame_desc) == 16 ? 0 : 1; }\n'
    exe = os.path.join(ROOT, "jaadec_b200", "_build", "abi_c_check")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-x", "c", "-",
                        "-o", exe], input=code.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert subprocess.run([exe]).returncode == 0


def test_struct_layouts_match_the_python_binding():
    from jaadec_b200 import FRAME_DESC_DTYPE, FRAME_RESULT_DTYPE, _lib
    assert FRAME_DESC_DTYPE.itemsize == C.sizeof(_lib.FrameDesc) == 16
    assert FRAME_RESULT_DTYPE.itemsize == C.sizeof(_lib.FrameResult) == 16
    assert C.sizeof(_lib.Options) == 32 and C.sizeof(_lib.StreamInfo) == 32 and C.sizeof(_lib.Timings) == 32


def test_status_strings_mirror_jaad_messages():
    from jaadec_b200 import _lib
    lib = _lib.load()
    assert lib.jaadb_status_string(0) == b"ok"
    assert lib.jaadb_status_string(2) == b"invalid huffman codebook: 12"     # ICStream.java:129
    assert lib.jaadb_status_string(3) == b"too many bands"                   # ICStream.java:138
    assert lib.jaadb_status_string(4) == b"scalefactor out of range"         # ICStream.java:213
    assert lib.jaadb_status_string(7) == b"reserved MS mask type used"       # CPE.java:114
    assert lib.jaadb_status_string(8) == b"TNS filter out of range"          # TNS.java:47


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="checks the behaviour of a box without a GPU")
def test_no_gpu_means_loud_failure_not_a_cpu_fallback():
    from jaadec_b200 import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(device=0, max_streams=4)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jaadec_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, os.path.join(root, f)


def test_python_mirror_constants_equal_the_header():
    """The ctypes mirror (jaadec_b200/engine.py) re-states the header's option values; they must be the header's."""
    import re
    import jaadec_b200 as jb
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "jaadb200.h")).read()

    def define(name):
        m = re.search(r"#define\s+%s\s+(\d+)u?\b" % name, hdr)
        assert m, name
        return int(m.group(1))

    pairs = {"JAADB_FLAG_PROFILE": jb.FLAG_PROFILE, "JAADB_FLAG_DEBUG_TAPS": jb.FLAG_DEBUG_TAPS, "JAADB_FLAG_PULSE_ISO": jb.FLAG_PULSE_ISO,
             "JAADB_TNS_JAAD": jb.TNS_JAAD, "JAADB_TNS_ISO": jb.TNS_ISO, "JAADB_PCM_S16LE": jb.PCM_S16LE, "JAADB_PCM_S16BE": jb.PCM_S16BE,
             "JAADB_PCM_F32_PLANAR": jb.PCM_F32_PLANAR}
    for name, value in pairs.items():
        assert define(name) == value, name
