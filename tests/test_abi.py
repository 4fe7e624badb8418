"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/mpet_b200.h
declares (no compute calls without a GPU); the product path refuses to run without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    src = open(os.path.join(ROOT, "include", "mpet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpet_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from waterscapes_b200 import build, _lib
    lib_path = build.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), "symbol %s declared in mpet_b200.h is not exported" % name
    # the ctypes prototypes cover exactly the header
    assert sorted(_lib.PROTOTYPES) == names
    assert lib.mpet_abi_version() == 1


def test_sass_is_sm100():
    import subprocess
    from waterscapes_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from waterscapes_b200 import _lib
    from waterscapes_b200.engine import Engine
    with pytest.raises(_lib.MpetLibraryError):
        Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "waterscapes_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)
