"""The C-ABI shared library builds for sm_100a, loads without a GPU, and exports exactly the entry
points include/wsb.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "wsb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wsb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from whisperseg_b200.build import build
    from whisperseg_b200 import _lib
    path = build()
    lib = ctypes.CDLL(path)
    declared = _declared()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(_lib.EXPORTS) == declared, "python binding table and header disagree"
    lib.wsb_abi_version.restype = ctypes.c_int
    assert lib.wsb_abi_version() == 1


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    import shutil
    import subprocess
    import pytest
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    from whisperseg_b200.build import build
    sass = subprocess.run(["cuobjdump", "-sass", build()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


def test_product_path_has_no_cpu_fallback():
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from whisperseg_b200.segmenter import WhisperSegmenter
    with pytest.raises(RuntimeError):
        WhisperSegmenter("/nonexistent", device="cpu")
    # and nothing under whisperseg_b200/ imports the oracle
    pkg = os.path.join(ROOT, "whisperseg_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), fn
