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


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include", "lingodb")), reason="needs /root/reference (build container only)")
def test_shim_links_against_the_reference_runtime_and_surfaces_errors_as_exceptions(tmp_path):
    """integration/selftest.cpp: the shim LINKED with the reference's own runtime objects (oracle/_ref: ExecutionContext, VarLen32 arenas) and
    the product library; a serialised step handed to lingodb::runtime::GPUPipeline::run(VarLen32) reaches the C-ABI, and the library's error
    (no device here; an unknown table on a GPU box) comes back as the std::runtime_error the reference's runtime functions throw."""
    import pyarrow
    from oracle import oracle as O
    if not os.path.exists(O.REF_LIB):
        pytest.skip("oracle/_ref not built")
    pa = os.path.dirname(pyarrow.__file__)
    exe = str(tmp_path / "shim_selftest")
    cmd = ["/usr/bin/g++", "-std=c++20", "-O1", "-Wall", "-Werror", "-DENABLE_REFCOUNT=1", f"-I{REF}/include", f"-I{REF}/vendored", f"-I{pa}/include", f"-I{ROOT}/include",
           os.path.join(ROOT, "integration", "GPUPipeline.cpp"), os.path.join(ROOT, "integration", "selftest.cpp"), "-o", exe,
           f"-L{os.path.dirname(O.REF_LIB)}", "-l:liboracle_ref.so", f"-L{ROOT}/lingo-db_b200", "-l:libldb_gpu.so", f"-L{pa}", "-l:libarrow.so.2400",
           f"-Wl,-rpath,{os.path.dirname(O.REF_LIB)}", f"-Wl,-rpath,{ROOT}/lingo-db_b200", f"-Wl,-rpath,{pa}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("runtime_error: "), (r.returncode, r.stdout, r.stderr[-500:])
