// pack_host.cpp — host side of compressed staging: frame-of-reference packing of one column block.
//
// HOST Arrow batches reach HBM over ONE PCIe 5 x16 link (≈55 GB/s) while the CPU the reference runs on scans the same
// buffers at several times that; so the bytes that cross the link are cut: per block of 65 536 values a column is stored as
// (value − block minimum) in the narrowest of 1/2/4/8 bytes.  TPC-H Q1's seven columns shrink from 76 B/row (Arrow) to
// ≈12 B/row.  decimal128(p<19) contributes its low 8 bytes only — the reference's JIT truncates it to i64 the same way
// (DBToStd/LowerToStd.cpp:111-209).  The GPU undoes the packing in a tiny kernel right after the copy (staging.cu), so
// every pipeline kernel still sees the plain staged layout.  Compiled by g++ (not nvcc) for function multi-versioning.
#include <cstddef>
#include <cstdint>
#include <cstring>

#define LDB_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))

namespace {
template <class T>
inline int64_t loadAs64(const uint8_t* p) {
   T v;
   memcpy(&v, p, sizeof(T));
   return (int64_t) v;
}
template <class T, int STRIDE>
LDB_CLONES void minMax(const uint8_t* src, int64_t n, int64_t* mn, int64_t* mx) {
   int64_t lo = INT64_MAX, hi = INT64_MIN;
   for (int64_t i = 0; i < n; i++) {
      const int64_t v = loadAs64<T>(src + (size_t) i * STRIDE);
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
   }
   *mn = lo;
   *mx = hi;
}
template <class T, int STRIDE, class OUT>
LDB_CLONES void packAs(const uint8_t* src, int64_t n, int64_t base, uint8_t* dst) {
   OUT* out = reinterpret_cast<OUT*>(dst);
   for (int64_t i = 0; i < n; i++) out[i] = (OUT) ((uint64_t) loadAs64<T>(src + (size_t) i * STRIDE) - (uint64_t) base);
}
// one pass: pack against a GUESSED base while tracking the block's real range
template <class T, int STRIDE, class OUT>
LDB_CLONES void packTracking(const uint8_t* src, int64_t n, int64_t base, uint8_t* dst, int64_t* mn, int64_t* mx) {
   OUT* out = reinterpret_cast<OUT*>(dst);
   int64_t lo = INT64_MAX, hi = INT64_MIN;
   for (int64_t i = 0; i < n; i++) {
      const int64_t v = loadAs64<T>(src + (size_t) i * STRIDE);
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
      out[i] = (OUT) ((uint64_t) v - (uint64_t) base);
   }
   *mn = lo;
   *mx = hi;
}
inline int widthFor(uint64_t range) { return range < (1ull << 8) ? 1 : range < (1ull << 16) ? 2 : range < (1ull << 32) ? 4 : 8; }
template <class T, int STRIDE>
size_t packBlock(const uint8_t* src, int64_t n, uint8_t* dst, int64_t* minOut, int32_t* widthOut, int64_t* hintLo, int64_t* hintHi) {
   int64_t lo, hi;
   // Speculation: the previous block of this column had the range [*hintLo, *hintHi]; centre a window of the same width on it
   // and pack in ONE pass over the source (the two-pass form reads every value twice and leaves the memory pipeline idle during
   // the second pass).  If the block's real range leaves the window the exact two-pass form below redoes it from cache.
   if (hintLo && *hintLo <= *hintHi) {
      const uint64_t prevRange = (uint64_t) *hintHi - (uint64_t) *hintLo;
      const int w = widthFor(prevRange);
      if (w < 8) {
         const uint64_t window = (1ull << (8 * w)) - 1;
         const uint64_t margin = (window - prevRange) / 2;
         int64_t base = (uint64_t) *hintLo - (uint64_t) INT64_MIN >= margin ? (int64_t) ((uint64_t) *hintLo - margin) : INT64_MIN; // no wrap below INT64_MIN
         switch (w) {
            case 1: packTracking<T, STRIDE, uint8_t>(src, n, base, dst, &lo, &hi); break;
            case 2: packTracking<T, STRIDE, uint16_t>(src, n, base, dst, &lo, &hi); break;
            default: packTracking<T, STRIDE, uint32_t>(src, n, base, dst, &lo, &hi); break;
         }
         *hintLo = lo;
         *hintHi = hi;
         if (lo >= base && (uint64_t) hi - (uint64_t) base <= window) {
            *minOut = base;
            *widthOut = w;
            return (size_t) n * w;
         }
      } else {
         minMax<T, STRIDE>(src, n, &lo, &hi);
      }
   } else {
      minMax<T, STRIDE>(src, n, &lo, &hi);
   }
   if (hintLo) {
      *hintLo = lo;
      *hintHi = hi;
   }
   const uint64_t range = (uint64_t) hi - (uint64_t) lo;
   int w = range < (1ull << 8) ? 1 : range < (1ull << 16) ? 2 : range < (1ull << 32) ? 4 : 8;
   switch (w) {
      case 1: packAs<T, STRIDE, uint8_t>(src, n, lo, dst); break;
      case 2: packAs<T, STRIDE, uint16_t>(src, n, lo, dst); break;
      case 4: packAs<T, STRIDE, uint32_t>(src, n, lo, dst); break;
      default: packAs<T, STRIDE, uint64_t>(src, n, lo, dst); break;
   }
   *minOut = lo;
   *widthOut = w;
   return (size_t) n * w;
}
} // namespace

// src_kind: 0 = int32 cells (int32/date32/fixed_size_binary(4)), 1 = int64 cells, 2 = decimal128 cells (low 8 bytes used).
// Writes n * width bytes to dst (dst must hold n * 8) and returns that size.  *min_out is the base the values are stored
// against (<= the block minimum).  hint_lo/hint_hi (may be NULL): in = value range of the previous block of this column
// (lo > hi: none), out = this block's range.
extern "C" size_t ldb_pack_block_hinted(const uint8_t* src, int32_t src_kind, int64_t n, uint8_t* dst, int64_t* min_out, int32_t* width_out, int64_t* hint_lo, int64_t* hint_hi) {
   if (n <= 0) {
      *min_out = 0;
      *width_out = 1;
      return 0;
   }
   switch (src_kind) {
      case 0: return packBlock<int32_t, 4>(src, n, dst, min_out, width_out, hint_lo, hint_hi);
      case 1: return packBlock<int64_t, 8>(src, n, dst, min_out, width_out, hint_lo, hint_hi);
      default: return packBlock<int64_t, 16>(src, n, dst, min_out, width_out, hint_lo, hint_hi);
   }
}
extern "C" size_t ldb_pack_block(const uint8_t* src, int32_t src_kind, int64_t n, uint8_t* dst, int64_t* min_out, int32_t* width_out) {
   return ldb_pack_block_hinted(src, src_kind, n, dst, min_out, width_out, nullptr, nullptr);
}
