"""The C-ABI library loads without a GPU, exports every symbol include/*.h declares, and refuses to
compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from lingodb_b200 import capi, datagen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ldb_[a-z0-9_]+|ldbgen_[a-z0-9_]+)\s*\(", src)))


def test_gpu_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared("ldb_gpu.h") + declared("ldb_tpch.h")
    assert len(names) > 40
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported by libldb_gpu.so"
    # and the python binding covers exactly the declared surface
    assert sorted(capi.SIGNATURES) == sorted(names)


def test_host_datagen_library_exports_every_declared_symbol():
    L = datagen.lib()
    for n in declared("ldb_datagen.h"):
        assert hasattr(L, n), n


def test_struct_layouts_match_the_reference_abi():
    # LdbArrayView must have the field layout of lingodb::runtime::ArrayView (ArrowView.h:8-21): 5 x i64 + 2 pointers
    assert C.sizeof(capi.ArrayView) == 56
    assert capi.ArrayView.offset.offset == 16 and capi.ArrayView.buffers.offset == 40
    assert C.sizeof(capi.I128) == 16 and C.sizeof(capi.GroupRow) == 8 + 16 * 8
    assert C.sizeof(capi.Error) == 256


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback_without_a_device():
    L = capi.lib()
    h, e = C.c_void_p(), capi.Error()
    rc = L.ldb_gpu_context_create(0, C.byref(h), C.byref(e))
    assert rc == capi.LDB_ERR_NO_DEVICE
    assert b"no CPU fallback" in e.message
    from lingodb_b200 import runtime
    with pytest.raises(capi.LdbRuntimeError):
        runtime.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lingo-db_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("the oracle's HashBuilder", "").replace("CPU oracle", "").replace("the oracle", ""), f
