"""The reference-side binding (integration/GPUPipeline.{h,cpp}) must compile against the reference's OWN
headers: its static_asserts pin LdbArrayView == lingodb::runtime::ArrayView and LdbFilterOp == lingodb::runtime::FilterOp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include", "lingodb")), reason="needs /root/reference (build container only)")
def test_reference_side_shim_compiles_against_reference_headers():
    import pyarrow
    inc = os.path.join(os.path.dirname(pyarrow.__file__), "include")
    cmd = ["/usr/bin/g++", "-std=c++20", "-fsyntax-only", "-Wall", "-Werror", f"-I{REF}/include", f"-I{REF}/vendored", f"-I{inc}", f"-I{ROOT}/include",
           os.path.join(ROOT, "integration", "GPUPipeline.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_shim_uses_only_declared_entry_points():
    import re
    src = open(os.path.join(ROOT, "integration", "GPUPipeline.cpp")).read()
    hdr = open(os.path.join(ROOT, "include", "ldb_gpu.h")).read()
    for fn in set(re.findall(r"\b(ldb_gpu_[a-z0-9_]+)\s*\(", src)):
        assert re.search(r"\b%s\s*\(" % fn, hdr), fn
