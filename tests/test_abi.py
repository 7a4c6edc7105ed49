"""C-ABI surface checks that need no GPU: libdn4gl.so loads, exports exactly what include/dn4gl.h declares, and the
entry points that do no device work behave (version, error text).  No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

from dummynode4graphlearning_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_names():
    src = open(os.path.join(ROOT, "include", "dn4gl.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    return sorted(set(re.findall(r"\b(dn4gl_\w+)\s*\(", src)))


def _exported_names():
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH], text=True)
    return sorted({ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("dn4gl_")})


def test_library_exists_and_loads():
    assert os.path.exists(_lib.SO_PATH), "build with `python -c 'import __graft_entry__ as g; g.build()'`"
    lib = _lib.lib()
    assert lib.raw("dn4gl_version")() == 1


def test_every_declared_symbol_is_exported_and_bound():
    declared = _header_names()
    assert len(declared) >= 60
    dll = ctypes.CDLL(_lib.SO_PATH)
    missing = [n for n in declared if not hasattr(dll, n)]
    assert not missing, "declared in include/dn4gl.h but not exported: %s" % missing
    # the ctypes binding parses the same header: it must see every prototype, each with a resolvable signature
    protos = _lib.parse_header()
    assert sorted(protos) == declared


def test_no_undeclared_entry_points():
    """everything the library exports under the dn4gl_ prefix is part of the documented ABI (no private back doors)."""
    extra = sorted(set(_exported_names()) - set(_header_names()))
    assert not extra, "exported but not declared in include/dn4gl.h: %s" % extra


def test_signatures_are_plain_c():
    """no torch / C++ types cross the boundary: only pointers, fixed-width ints, float, size_t."""
    for name, (restype, argtypes) in _lib.parse_header().items():
        for t in argtypes:
            assert t in (ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_float,
                         ctypes.c_size_t, ctypes.c_uint32), (name, t)
        assert restype in (ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_void_p), name


def test_workspace_queries_are_pure_host_functions():
    """*_workspace_bytes entry points are size arithmetic only (callable without a device) and monotone."""
    lib = _lib.lib()
    small = lib.size("dn4gl_csr_workspace_bytes", 10, 20)
    big = lib.size("dn4gl_csr_workspace_bytes", 1000, 2000)
    assert 0 < small <= big
    assert lib.size("dn4gl_scan_workspace_bytes", 1 << 20) >= lib.size("dn4gl_scan_workspace_bytes", 1)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """the product path has no CPU fallback: without the shared library the binding raises."""
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "libdn4gl.so"))
    with pytest.raises(_lib.Dn4glError):
        _lib._Lib()
