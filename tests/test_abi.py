"""CPU tests of the drop-in boundary: libdct3d.so builds for sm_100a without a GPU, loads, exports
every function include/dct3d.h declares, and fails loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import pytest

from conftest import pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "dct3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dct3d_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    lib = pkg("_lib")
    names = declared_functions()
    assert len(names) >= 30
    assert sorted(lib.SYMBOLS) == names          # the ctypes binding covers exactly the header


def test_library_exports_every_declared_symbol():
    pkg("build").build()
    lib = pkg("_lib")
    L = lib.load()
    for name in declared_functions():
        assert getattr(L, name) is not None      # AttributeError if the .so lacks the symbol


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = pkg("_lib")
    L = lib.load()
    assert L.dct3d_device_count() < 0
    h = C.c_void_p()
    assert L.dct3d_create(C.byref(h), 0, 64, 48, 8) == lib.E_CUDA      # fails loudly, no silent CPU path
    assert b"CUDA" in L.dct3d_last_error(None)
    codec = pkg("codec")
    with pytest.raises(codec.Dct3dError):
        codec.Codec(64, 48, 8)


def test_argument_validation_precedes_device_use():
    lib = pkg("_lib")
    L = lib.load()
    h = C.c_void_p()
    assert L.dct3d_create(C.byref(h), 0, 64, 48, 5) == lib.E_INVALID      # cube edge must be 8 or 4
    assert L.dct3d_create(C.byref(h), 0, 100, 48, 8) == lib.E_INVALID     # width not a multiple of the cube
    assert L.dct3d_create(None, 0, 64, 48, 8) == lib.E_INVALID


def test_sass_is_blackwell_native():
    """The shipped cubin is sm_100a and uses TMA (UTMALDG) and mbarrier (SYNCS) instructions."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    pkg("build").build()
    out = subprocess.run(["cuobjdump", "-sass", pkg("_lib").lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UTMALDG" in out and "SYNCS" in out
