"""The C-ABI shared library loads on a CPU-only host and exports exactly what include/polarcub_b200.h declares.
No compute entry point is called here (there is no GPU); argument validation that needs no device is exercised."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "polarcub_b200.h")


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pc_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from polarcub_b200 import build, _lib
    build.build()  # nvcc cross-compiles sm_100a without a GPU; no-op when the .so is current
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header(lib):
    from polarcub_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_argument_validation_without_a_device(lib):
    assert lib.pc_version() >= 100
    out = ctypes.c_void_p(0)
    rc = lib.pc_plan_create(1, 3, None, None, ctypes.byref(out))  # q = 1 is not an alphabet
    assert rc == -1 and out.value is None
    assert b"alphabet" in lib.pc_last_error()
    assert lib.pc_plan_k(None) == -1
    assert lib.pc_kernel_launch_count() == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from polarcub_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PolarcubError, match="no CPU fallback"):
        _lib.lib()
