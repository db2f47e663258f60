"""The C-ABI library builds for sm_100a, loads, and exports every symbol the header declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from mobrob_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mobrob_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mr_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mobrob_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == syms, "ctypes table and header are out of sync"


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.mr_version() >= 100
    assert isinstance(lib.mr_last_error(), bytes)
    assert lib.mr_launch_count() >= 0


def test_bad_arguments_fail_loudly_without_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.mr_env_create(7, 4, 0, 1000, 1, ctypes.byref(h)) != 0
    assert b"kind" in lib.mr_last_error()
    assert lib.mr_env_create(0, 0, 0, 1000, 1, ctypes.byref(h)) != 0
