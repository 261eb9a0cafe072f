"""Host half of compressed staging (csrc/pack_host.cpp): frame-of-reference packing of one column block.
The device half (unpack kernel) is covered by the -m gpu parity tests, which stage every HOST batch through it."""
import ctypes as C

import numpy as np
import pytest

from lingodb_b200 import capi


def _pack(src: np.ndarray, kind: int, n: int):
    L = capi.lib()
    L.ldb_pack_block.restype = C.c_size_t
    L.ldb_pack_block.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    dst = np.zeros(n * 8 + 16, np.uint8)
    mn, w = C.c_int64(), C.c_int32()
    nbytes = L.ldb_pack_block(src.ctypes.data, kind, n, dst.ctypes.data, C.byref(mn), C.byref(w))
    return dst[:nbytes], mn.value, w.value


def _unpack(packed, mn, w, n):
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w]
    return (packed.view(dt)[:n].astype(np.uint64) + np.uint64(mn & 0xFFFFFFFFFFFFFFFF)).astype(np.int64)  # wrapping add


@pytest.mark.parametrize("lo,hi,width", [(5, 200, 1), (-3, 250, 1), (0, 256, 2), (100, 50_000, 2), (-70_000, 70_000, 4), (0, 2**31 - 1, 4)])
def test_int32_blocks_round_trip(lo, hi, width):
    rng = np.random.default_rng(1)
    n = 65536
    v = rng.integers(lo, hi + 1, n, dtype=np.int64).astype(np.int32)
    v[0], v[1] = lo, hi
    packed, mn, w = _pack(v, 0, n)
    assert (mn, w) == (lo, width) and packed.size == n * width
    assert np.array_equal(_unpack(packed, mn, w, n), v.astype(np.int64))


def test_decimal128_cells_use_their_low_8_bytes_and_ragged_length():
    rng = np.random.default_rng(2)
    n = 12345  # last block of a batch
    lo64 = rng.integers(90_000, 10_500_000, n, dtype=np.int64)
    cells = np.zeros((n, 2), np.int64)
    cells[:, 0] = lo64
    packed, mn, w = _pack(cells, 2, n)
    assert w == 4 and mn == lo64.min()
    assert np.array_equal(_unpack(packed, mn, w, n), lo64)
    neg = cells.copy()
    neg[:, 0] = -lo64
    neg[:, 1] = -1  # sign extension
    packed, mn, w = _pack(neg, 2, n)
    assert np.array_equal(_unpack(packed, mn, w, n), -lo64)


def test_int64_full_range_needs_8_bytes():
    v = np.array([np.iinfo(np.int64).min, -1, 0, np.iinfo(np.int64).max], np.int64)
    packed, mn, w = _pack(v, 1, 4)
    assert w == 8 and mn == np.iinfo(np.int64).min
    assert np.array_equal(_unpack(packed, mn, w, 4), v)


def test_speculative_single_pass_matches_two_pass_and_falls_back():
    """ldb_pack_block_hinted: a hint that covers the block packs in one pass against a base <= min; a hint that does not is
    detected and the block is re-packed exactly.  Either way the decoded values are the source values."""
    L = capi.lib()
    L.ldb_pack_block_hinted.restype = C.c_size_t
    L.ldb_pack_block_hinted.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    rng = np.random.default_rng(7)
    n = 65536
    for lo, hi, hint, width in [(1000, 1200, (1010, 1190), 1), (1000, 1200, (5000, 5100), 1), (-50, 60000, (0, 50000), 2), (0, 10**7, (100, 9 * 10**6), 4),
                                (5, 9, (2**40, 2**40 + 3), 1), (-2**62, 2**62, (0, 10), 8), (0, 100, (-2**63, -2**63 + 50), 1)]:
        v = rng.integers(lo, hi + 1, n, dtype=np.int64)
        v[0], v[1] = lo, hi
        dst = np.zeros(n * 8, np.uint8)
        mn, w, hl, hh = C.c_int64(), C.c_int32(), C.c_int64(hint[0]), C.c_int64(hint[1])
        nbytes = L.ldb_pack_block_hinted(v.ctypes.data, 1, n, dst.ctypes.data, C.byref(mn), C.byref(w), C.byref(hl), C.byref(hh))
        assert (hl.value, hh.value) == (lo, hi)  # the hint now describes this block
        assert w.value == width and nbytes == n * width and mn.value <= lo
        assert np.array_equal(_unpack(dst[:nbytes], mn.value, w.value, n), v), (lo, hi, hint)
