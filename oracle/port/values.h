// oracle/port/values.h — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product path).
//
// CPU restatement of the reference's value-level semantics for the hot path: hashing, exact
// decimal arithmetic, date constants.  These are *generated inline by the JIT* in the reference
// (there is no runtime function to call), so they are restated here from the lowering patterns
// and pinned against the reference's own known-answer tests (tests/test_oracle_kat.py):
//   test/lit/DB/hash.mlir:27-34, test/unittests/storage/TestStorage.cpp:289,411,
//   test/sqlite-datasets/tpchSf1.test:25-28 (avg columns of Q1).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <string_view>

namespace oracle {

using i128 = __int128;
using u128 = unsigned __int128;

// ---- util.hash_64 / util.hash_combine  (src/compiler/Conversion/UtilToLLVM/LowerToLLVM.cpp:493-514;
//      runtime twin src/runtime/Hash.cpp:25-33)
inline uint64_t hash64(uint64_t v) {
   uint64_t m = v * 11400714819323198549ull; // = 0x9E3779B97F4A7C55 (not the golden-ratio ...7C15; pinned by the KATs)
   return m ^ __builtin_bswap64(m);
}
// combineHashes(hash1 = new piece, totalHash) = hash1 ^ bswap(totalHash)
// (LowerToStd.cpp:1066-1072 + LowerToLLVM.cpp:505-514)
inline uint64_t hashCombine(uint64_t newPiece, uint64_t total) { return newPiece ^ __builtin_bswap64(total); }

// db.hash over a value list: the first piece becomes the running hash, every further piece is
// combined in (LowerToStd.cpp:1076-1096, 1139-1150).
struct HashBuilder {
   uint64_t total = 0;
   bool any = false;
   void addPiece(uint64_t h) {
      total = any ? hashCombine(h, total) : h;
      any = true;
   }
   void addInt(int64_t v) { addPiece(hash64((uint64_t) v)); } // i1/i8/i16/i32/i64 are sign-extended to index
   void addBool(bool b) { addInt(b ? -1 : 0); } // i1 true sign-extends to -1 (arith.index_cast; KAT hash.mlir:29)
   void addI128(i128 v) { // high half first, then low half (LowerToStd.cpp:1078-1090)
      addPiece(hash64((uint64_t) ((u128) v >> 64)));
      addPiece(hash64((uint64_t) v));
   }
   void addNull() { // NULL contributes nothing; an all-NULL hash is 0 (LowerToStd.cpp:1119-1131)
      if (!any) {
         total = 0;
         any = true;
      }
   }
};

// ---- XXH64 (seed 0): llvm::xxHash64 of LLVM 20.1 (third-party, not under /root/reference; call site
//      src/runtime/Hash.cpp:13-16).  Published algorithm (Yann Collet, xxHash spec), restated.
inline uint64_t xxh64(const uint8_t* p, size_t len, uint64_t seed = 0) {
   constexpr uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                      P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
   auto rotl = [](uint64_t x, int r) { return (x << r) | (x >> (64 - r)); };
   auto rd64 = [](const uint8_t* q) { uint64_t v; memcpy(&v, q, 8); return v; };
   auto rd32 = [](const uint8_t* q) { uint32_t v; memcpy(&v, q, 4); return v; };
   auto round = [&](uint64_t acc, uint64_t in) { return rotl(acc + in * P2, 31) * P1; };
   auto mergeRound = [&](uint64_t acc, uint64_t val) { return (acc ^ round(0, val)) * P1 + P4; };
   const uint8_t* end = p + len;
   uint64_t h;
   if (len >= 32) {
      uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
      const uint8_t* limit = end - 32;
      do {
         v1 = round(v1, rd64(p));
         v2 = round(v2, rd64(p + 8));
         v3 = round(v3, rd64(p + 16));
         v4 = round(v4, rd64(p + 24));
         p += 32;
      } while (p <= limit);
      h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
      h = mergeRound(h, v1);
      h = mergeRound(h, v2);
      h = mergeRound(h, v3);
      h = mergeRound(h, v4);
   } else {
      h = seed + P5;
   }
   h += (uint64_t) len;
   while (p + 8 <= end) {
      h ^= round(0, rd64(p));
      h = rotl(h, 27) * P1 + P4;
      p += 8;
   }
   if (p + 4 <= end) {
      h ^= (uint64_t) rd32(p) * P1;
      h = rotl(h, 23) * P2 + P3;
      p += 4;
   }
   while (p < end) {
      h ^= (*p) * P5;
      h = rotl(h, 11) * P1;
      p++;
   }
   h ^= h >> 33;
   h *= P2;
   h ^= h >> 29;
   h *= P3;
   h ^= h >> 32;
   return h;
}

// ---- VarLen32: 16-byte string handle (include/lingodb/runtime/helpers.h:82-240).
//   word0 low 32 bits = length; len <= 12: the bytes are stored inline right after the length
//   (12 bytes, zero padded); len > 12: 4-byte prefix follows the length and the upper 8 bytes hold
//   the data pointer.
struct VarLen32 {
   uint32_t len;
   uint8_t first4[4];
   union {
      uint8_t last8[8];
      const uint8_t* ptr;
   };
   VarLen32(const uint8_t* data, uint32_t l) {
      len = l;
      memset(first4, 0, 4);
      memset(last8, 0, 8);
      if (l <= 12) {
         memcpy(first4, data, l < 4 ? l : 4);
         if (l > 4) memcpy(last8, data + 4, l - 4);
      } else {
         memcpy(first4, data, 4);
         ptr = data;
      }
   }
   const uint8_t* data() const { return len <= 12 ? first4 : ptr; }
   std::string_view view() const { return std::string_view((const char*) data(), len); }
   uint64_t lo64() const { uint64_t v; memcpy(&v, this, 8); return v; }
   uint64_t hi64() const { uint64_t v; memcpy(&v, (const uint8_t*) this + 8, 8); return v; }
};
static_assert(sizeof(VarLen32) == 16);

// util.varlen_try_cheap_hash + util.hash_varlen (LowerToLLVM.cpp:372-391, LowerToStd.cpp:1097-1106;
// runtime twin src/runtime/Hash.cpp:47-58): len < 13 → combine(h64(lo), h64(hi)) = h64(lo) ^ bswap(h64(hi)),
// else xxHash64 of the bytes.
inline uint64_t hashVarLen(const VarLen32& v) {
   if (v.len < 13) return hashCombine(hash64(v.lo64()), hash64(v.hi64()));
   return xxh64(v.data(), v.len);
}

// ---- decimals: exact integers.  decimal(p,s) is computed in i64 when p < 19 and in
//      i128 otherwise (LowerToStd.cpp:1479-1486); all arithmetic is two's-complement wrapping like LLVM's.
inline i128 pow10_128(int k) {
   i128 r = 1;
   while (k-- > 0) r *= 10;
   return r;
}
inline i128 wrapAdd(i128 a, i128 b) { return (i128) ((u128) a + (u128) b); }
inline i128 wrapSub(i128 a, i128 b) { return (i128) ((u128) a - (u128) b); }
inline i128 wrapMul(i128 a, i128 b) { return (i128) ((u128) a * (u128) b); }
inline int64_t wrapAdd64(int64_t a, int64_t b) { return (int64_t) ((uint64_t) a + (uint64_t) b); }

// decimal128 Arrow value → i128 (ArrowToStd.cpp:67-85 fixed-size load); trunc to i64 when p < 19
// (LowerToStd.cpp:111-209).
inline i128 loadDec128(const void* base, int64_t row) {
   i128 v;
   memcpy(&v, (const uint8_t*) base + 16 * row, 16);
   return v;
}

// avg(x decimal(p,s)) is rewritten to sum(x) / cast(count -> decimal(19,0))
// (Dialect/RelAlg/Transforms/SimplifyAggregations.cpp:160-181); DecimalDiv pre-scales the numerator by
// 10^(s_res + s_r - s_l) with s_res = 21 for decimal(12,2)/decimal(19,0) → 10^19, then sdiv
// (LowerToStd.cpp:651-700; typing DBOps.cpp:221-262).  Result: decimal(31,21), truncating division.
inline i128 avgDec12_2(int64_t sumRaw, int64_t count) { return wrapMul((i128) sumRaw, pow10_128(19)) / (i128) count; }

// "YYYY-MM-DD" → date32 (days since 1970-01-01); the reference uses arrow::internal::ParseValue
// <Date32Type> (src/runtime/storage/Restrictions.cpp:17-25).  Civil-from-days per H. Hinnant.
inline int32_t parseDate32(const std::string& s) {
   int y, m, d;
   if (sscanf(s.c_str(), "%d-%d-%d", &y, &m, &d) != 3) throw std::runtime_error("could not parse date");
   y -= m <= 2;
   int era = (y >= 0 ? y : y - 399) / 400;
   unsigned yoe = (unsigned) (y - era * 400);
   unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
   unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
   return era * 146097 + (int) doe - 719468;
}
// date32 → i64 nanoseconds (generated loads multiply by 86 400 000 000 000; LowerToStd.cpp:111-209)
inline int64_t dateToNs(int32_t days) { return (int64_t) days * 86400000000000ll; }

// decimal constant "0.05" → raw i128 at the column's scale (Restrictions.cpp:455-467 uses
// arrow::Decimal128::FromString + Rescale; rescale up = multiply by 10^k, down must be exact).
inline i128 parseDecimal(const std::string& s, int targetScale) {
   bool neg = false;
   size_t i = 0;
   if (i < s.size() && (s[i] == '-' || s[i] == '+')) neg = s[i++] == '-';
   i128 v = 0;
   int scale = 0;
   bool dot = false;
   for (; i < s.size(); i++) {
      if (s[i] == '.') {
         dot = true;
         continue;
      }
      if (s[i] < '0' || s[i] > '9') throw std::runtime_error("could not parse decimal const");
      v = v * 10 + (s[i] - '0');
      if (dot) scale++;
   }
   while (scale < targetScale) {
      v *= 10;
      scale++;
   }
   while (scale > targetScale) {
      if (v % 10 != 0) throw std::runtime_error("decimal rescale would lose data");
      v /= 10;
      scale--;
   }
   return neg ? -v : v;
}

// ---- scalar runtime used by Q9 (restated; the _ref build calls the reference's own functions instead, runtime_ref.h)
// DateRuntime::extractYear (src/runtime/DateRuntime.cpp:99-101): year of floor<days>(ns) through
// arrow_vendored::date::year_month_day — the published civil-from-days algorithm (H. Hinnant, "chrono-Compatible
// Low-Level Date Algorithms"), restated.
inline int64_t yearOfDays(int64_t days) {
   int64_t z = days + 719468;
   int64_t era = (z >= 0 ? z : z - 146096) / 146097;
   int64_t doe = z - era * 146097;
   int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
   int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
   int64_t mp = (5 * doy + 2) / 153;
   return yoe + era * 400 + (mp >= 10 ? 1 : 0); // months Jan/Feb belong to the next civil year
}
inline int64_t extractYearPort(int64_t ns) {
   constexpr int64_t dayNs = 86400000000000ll;
   int64_t days = ns / dayNs - (ns % dayNs < 0 ? 1 : 0); // floor
   return yearOfDays(days);
}
// `x like '%needle%'` with a constant pattern: ConstLike lowering (RuntimeFunctions.cpp:87-170, matchPart :60-86)
// reduces to one StringRuntime::findMatch(str, needle, 0, len) (StringRuntime.cpp:337-345) != invalidPos.
inline bool containsPort(std::string_view str, std::string_view needle) {
   if (needle.size() > str.size()) return false;
   return str.find(needle, 0) != std::string_view::npos;
}

} // namespace oracle
