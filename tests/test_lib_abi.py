"""The C-ABI library loads without a GPU and exports every symbol include/maua_b200.h declares."""
import ctypes
import os
import re

from maua_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "maua_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_built_in_tree():
    path = _build.build()
    assert os.path.exists(path) and path.startswith(ROOT)


def test_exports_every_declared_symbol():
    lib = ctypes.CDLL(_build.build())
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/maua_b200.h but not exported"


def test_python_binding_covers_header():
    declared = set(_declared_functions())
    bound = set(_lib.signatures().keys())
    assert declared <= bound | {"mb_debug_read"}, declared - bound


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.mb_abi_version() >= 1
    assert isinstance(lib.mb_last_error(), (bytes, type(None)))


def test_invalid_arguments_return_codes_not_crashes():
    lib = _lib.load()
    cfg = _lib.SG3Cfg()
    lib.mb_sg3_default_cfg(ctypes.byref(cfg), 0)
    cfg.conv_kernel = 5
    layers = (_lib.SG3Layer * 15)()
    rc = lib.mb_sg3_geometry(ctypes.byref(cfg), layers, None, None, None, None)
    assert rc == -1 and b"conv_kernel" in lib.mb_last_error()
