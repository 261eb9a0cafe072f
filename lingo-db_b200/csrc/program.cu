// program.cu — the GENERIC pipeline kernel: scan → register program → {hash aggregation | join-table build | materialize},
// the large-domain hash aggregation table, its export, and a radix sort for ORDER BY.  See program.h.
#include "device_utils.cuh"
#include "program.h"
#include "../../include/ldb_gpu.h"

#include <algorithm>

namespace ldb {

typedef __int128 s128;
// twins of the join-table primitives of kernels.cu (plain single-key tables and direct-address tables)
__device__ __forceinline__ int32_t directLoadProg(const JoinTableDev& t, int32_t key) {
   const uint32_t idx = (uint32_t) key - (uint32_t) t.keyMin;
   return idx < t.range ? __ldg((const int32_t*) t.base + idx) : kDirectEmpty;
}
__device__ bool joinInsertProg(const JoinTableDev& t, int32_t key, int32_t payload) {
   const unsigned long long packed = ((unsigned long long) (uint32_t) payload << 32) | (uint32_t) key;
   if (packed == ~0ull) {
      atomicExch(t.error, 3);
      return false;
   }
   const uint64_t h = hashI32(key);
   uint64_t s = h & t.mask;
   const uint64_t limit = t.mask < 16384 ? t.mask + 1 : 16384;
   for (uint64_t probes = 0; probes < limit; probes++) {
      const unsigned long long old = atomicCAS((unsigned long long*) (t.base + s * t.stride), ~0ull, packed);
      if (old == ~0ull) {
         if (t.bloom) atomicOr(&t.bloom[(uint32_t) (h >> 32) & t.bloomMask], bloomBits(h));
         return true;
      }
      if ((int32_t) (uint32_t) old == key && t.unique) { // a set of keys (semi-join build side): duplicates are dropped, not an error
         return false;
      }
      s = (s + 1) & t.mask;
   }
   atomicExch(t.error, 1);
   return false;
}
struct Val {
   s128 v;    // integers, decimals (raw), dates (days), booleans (0/1); doubles live in the low 64 bits
   bool null;
};
__device__ __forceinline__ double asF64(const Val& x) { return __longlong_as_double((long long) (uint64_t) x.v); }
__device__ __forceinline__ Val fromF64(double d, bool null) { return Val{(s128) (uint64_t) __double_as_longlong(d), null}; }

__device__ __forceinline__ bool colIsNull(const ProgCol& c, int64_t row) {
   if (c.validBytes) return c.validBytes[row] == 0;
   if (!c.validity) return false;
   const int64_t bit = c.bitOffset + row;
   return !((c.validity[bit >> 3] >> (bit & 7)) & 1u);
}
// LoadArrowOp lowering (ArrowToStd.cpp:67-173, LowerToStd.cpp:111-209): physical cell → value register
__device__ __forceinline__ Val loadCol(const ProgCol& c, int64_t row) {
   Val r;
   r.null = colIsNull(c, row);
   switch (c.type) {
      case LDB_INT32:
      case LDB_DATE32:
      case LDB_FSB4: r.v = (s128) ((const int32_t*) c.data)[row]; break;
      case LDB_INT64: r.v = (s128) ((const int64_t*) c.data)[row]; break;
      case LDB_INT8: r.v = (s128) ((const int8_t*) c.data)[row]; break;
      case LDB_INT16: r.v = (s128) ((const int16_t*) c.data)[row]; break;
      case LDB_FLOAT32: return fromF64((double) ((const float*) c.data)[row], r.null);
      case LDB_FLOAT64: return fromF64(((const double*) c.data)[row], r.null);
      case LDB_DECIMAL128:
         if (c.elemBytes == 16) {
            const ulonglong2 cell = ((const ulonglong2*) c.data)[row];
            r.v = (s128) (((unsigned __int128) cell.y << 64) | cell.x);
         } else {
            r.v = (s128) ((const int64_t*) c.data)[row]; // narrowed HOST batch (p < 19): sign-extend
         }
         break;
      default: r.v = 0; r.null = true;
   }
   return r;
}
__device__ __forceinline__ bool cmpI(s128 a, s128 b, int op) {
   switch (op) {
      case LDB_EQ: return a == b;
      case LDB_NEQ: return a != b;
      case LDB_LT: return a < b;
      case LDB_LTE: return a <= b;
      case LDB_GT: return a > b;
      default: return a >= b;
   }
}
__device__ __forceinline__ bool cmpF(double a, double b, int op) {
   switch (op) {
      case LDB_EQ: return a == b;
      case LDB_NEQ: return a != b;
      case LDB_LT: return a < b;
      case LDB_LTE: return a <= b;
      case LDB_GT: return a > b;
      default: return a >= b;
   }
}
// VarLen32 ordering (VarLen32Filter<CMP>, Restrictions.cpp:234-325): bytewise, shorter string first on a common prefix
__device__ int strCompare(const ProgCol& c, int64_t row, const uint8_t* k, int klen) {
   const int32_t* off = (const int32_t*) c.data + row;
   const int32_t b = off[0], n = off[1] - off[0];
   const uint8_t* s = c.bytes + b;
   const int m = n < klen ? n : klen;
   for (int i = 0; i < m; i++)
      if (s[i] != k[i]) return s[i] < k[i] ? -1 : 1;
   return n == klen ? 0 : (n < klen ? -1 : 1);
}
__device__ bool strLike(const ProgCol& c, int64_t row, const uint8_t* k, int klen, int kind) {
   const int32_t* off = (const int32_t*) c.data + row;
   const int32_t b = off[0], n = off[1] - off[0];
   const uint8_t* s = c.bytes + b;
   if (n < klen) return false;
   if (kind == 0 || kind == 1) { // prefix% / %suffix
      const uint8_t* p = kind == 0 ? s : s + (n - klen);
      for (int i = 0; i < klen; i++)
         if (p[i] != k[i]) return false;
      return true;
   }
   for (int i = 0; i + klen <= n; i++) { // %contains%
      int j = 0;
      while (j < klen && s[i + j] == k[j]) j++;
      if (j == klen) return true;
   }
   return false;
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
   x ^= x >> 33;
   x *= 0xff51afd7ed558ccdull;
   x ^= x >> 33;
   x *= 0xc4ceb9fe1a85ec53ull;
   x ^= x >> 33;
   return x;
}

// ---------------------------------------------------------------- hash aggregation table
constexpr uint32_t kSeenBit = 1u, kClaimBit = 1u << 8, kKeyNullBit = 1u << 16;
__device__ __forceinline__ uint8_t* entryAt(const HashAggDev& t, uint64_t s) { return t.base + s * t.entryBytes; }
// lookup-or-insert by key tuple (subop.lookup_or_insert, SubOpToControlFlow.cpp:3065-3157; NULL keys form a group of their own)
__device__ uint8_t* hashAggFind(const ProgramParams& p, const int64_t* keys, uint32_t keyNulls) {
   const HashAggDev& t = p.agg;
   if (t.nKeys == 0) return t.base; // keyless: the table is one pre-initialised entry
   uint64_t h = 0x9E3779B97F4A7C55ull ^ keyNulls;
   for (int k = 0; k < t.nKeys; k++) h = mix64(h ^ (uint64_t) keys[k]) + 0x632BE59BD9B4E019ull * (k + 1);
   uint64_t s = h & t.mask;
   const uint64_t limit = t.mask + 1 < 65536 ? t.mask + 1 : 65536;
   for (uint64_t probes = 0; probes < limit; probes++) {
      uint8_t* e = entryAt(t, s);
      uint32_t* state = (uint32_t*) e;
      uint32_t st = *((volatile uint32_t*) state);
      if (st == 0) {
         st = atomicCAS(state, 0u, 1u);
         if (st == 0) { // this thread creates the group: keys, key-null bits, aggregate identities, then publish
            uint32_t* flags = state + 1;
            *flags = keyNulls * kKeyNullBit;
            int64_t* ek = (int64_t*) (e + 16);
            for (int k = 0; k < t.nKeys; k++) ek[k] = keys[k];
            unsigned long long* ea = (unsigned long long*) (e + 16 + 8 * kProgMaxKeys);
            for (int a = 0; a < t.nAggs; a++) {
               unsigned long long lo = 0, hi = 0;
               switch (p.aggs[a].kind) {
                  case LDB_AGG_MIN: lo = (unsigned long long) INT64_MAX; break;
                  case LDB_AGG_MAX: lo = (unsigned long long) INT64_MIN; break;
                  case LDB_AGG_MIN_F64: lo = (unsigned long long) __double_as_longlong(INFINITY); break;
                  case LDB_AGG_MAX_F64: lo = (unsigned long long) __double_as_longlong(-INFINITY); break;
                  default: break;
               }
               ea[2 * a] = lo;
               ea[2 * a + 1] = hi;
            }
            __threadfence();
            atomicExch(state, 2u);
            atomicAdd(t.count, 1ull);
            return e;
         }
      }
      while (st == 1) st = *((volatile uint32_t*) state);
      __threadfence();
      const uint32_t fl = *((volatile uint32_t*) (state + 1));
      bool eq = ((fl >> 16) & 0xfu) == keyNulls;
      const volatile int64_t* ek = (const volatile int64_t*) (e + 16);
      for (int k = 0; k < t.nKeys && eq; k++) eq = ((keyNulls >> k) & 1u) || ek[k] == keys[k];
      if (eq) return e;
      s = (s + 1) & t.mask;
   }
   atomicExch(t.error, 1);
   return nullptr;
}
// in-place aggregate update (subop.reduce lowering + combine functions, SubOpToControlFlow.cpp:3540-3769, RelAlgToSubOp.cpp:1809-2025):
// NULL inputs are skipped, an aggregate that never saw a value stays NULL (its "seen" bit)
__device__ void hashAggUpdate(const ProgramParams& p, uint8_t* e, const Val* regs) {
   uint32_t* flags = (uint32_t*) e + 1;
   unsigned long long* ea = (unsigned long long*) (e + 16 + 8 * kProgMaxKeys);
   for (int a = 0; a < p.nAggs; a++) {
      const int kind = p.aggs[a].kind;
      unsigned long long* lo = ea + 2 * a;
      if (kind == LDB_AGG_COUNT_STAR) {
         atomicAdd(lo, 1ull);
         continue;
      }
      const Val x = regs[p.aggs[a].reg];
      if (x.null) continue;
      const uint32_t seen = kSeenBit << a;
      switch (kind) {
         case LDB_AGG_SUM: atomicAdd128(lo, lo + 1, i128{(uint64_t) x.v, (int64_t) (x.v >> 64)}); break;
         case LDB_AGG_SUM_F64: atomicAdd((double*) lo, asF64(x)); break;
         case LDB_AGG_COUNT: atomicAdd(lo, 1ull); break;
         case LDB_AGG_MIN: atomicMin((long long*) lo, (long long) x.v); break;
         case LDB_AGG_MAX: atomicMax((long long*) lo, (long long) x.v); break;
         case LDB_AGG_MIN_F64:
         case LDB_AGG_MAX_F64: {
            const double d = asF64(x);
            unsigned long long cur = *((volatile unsigned long long*) lo);
            while (kind == LDB_AGG_MIN_F64 ? d < __longlong_as_double((long long) cur) : d > __longlong_as_double((long long) cur)) {
               const unsigned long long prev = atomicCAS(lo, cur, (unsigned long long) __double_as_longlong(d));
               if (prev == cur) break;
               cur = prev;
            }
            break;
         }
         case LDB_AGG_ANY: {
            if (*((volatile uint32_t*) flags) & (kClaimBit << a)) continue;
            const uint32_t old = atomicOr(flags, kClaimBit << a);
            if (old & (kClaimBit << a)) continue; // somebody else's value is the group's "any"
            lo[0] = (unsigned long long) (uint64_t) x.v;
            lo[1] = (unsigned long long) (uint64_t) (x.v >> 64);
            __threadfence();
            break;
         }
         default: break;
      }
      if (!(*((volatile uint32_t*) flags) & seen)) atomicOr(flags, seen);
   }
}

// ---------------------------------------------------------------- the interpreter
__global__ void __launch_bounds__(256) programKernel(const __grid_constant__ ProgramParams p) {
   unsigned long long inserted = 0;
   for (int64_t base = (int64_t) blockIdx.x * blockDim.x; base < p.nRows; base += (int64_t) gridDim.x * blockDim.x) {
      const int64_t row = base + threadIdx.x;
      const bool valid = row < p.nRows;
      Val regs[kProgMaxRegs] = {};
      bool pass = valid;
      if (valid) {
         for (int pc = 0; pc < p.nInstr; pc++) {
            const ProgInstr in = p.instr[pc];
            const Val a = regs[in.a], b = regs[in.b];
            Val r;
            r.v = 0;
            r.null = false;
            switch (in.op) {
               case LDB_OP_LOAD: r = loadCol(p.cols[in.arg], row); break;
               case LDB_OP_CONST: r.v = (s128) (((unsigned __int128) (uint64_t) p.constHi[in.arg] << 64) | p.constLo[in.arg]); break;
               case LDB_OP_ADD: r.v = (s128) ((unsigned __int128) a.v + (unsigned __int128) b.v); r.null = a.null | b.null; break;
               case LDB_OP_SUB: r.v = (s128) ((unsigned __int128) a.v - (unsigned __int128) b.v); r.null = a.null | b.null; break;
               case LDB_OP_MUL: r.v = (s128) ((unsigned __int128) a.v * (unsigned __int128) b.v); r.null = a.null | b.null; break;
               case LDB_OP_DIV:
                  r.null = a.null | b.null | (b.v == 0);
                  if (!r.null) r.v = a.v / b.v; // sdiv: truncating (DecimalDiv lowering, LowerToStd.cpp:651-700)
                  break;
               case LDB_OP_NEG: r.v = (s128) (0 - (unsigned __int128) a.v); r.null = a.null; break;
               case LDB_OP_CMP: r.v = cmpI(a.v, b.v, in.arg); r.null = a.null | b.null; break;
               case LDB_OP_AND: // three-valued: false dominates NULL
                  if ((!a.null && a.v == 0) || (!b.null && b.v == 0)) r.v = 0;
                  else if (a.null | b.null) r.null = true;
                  else r.v = 1;
                  break;
               case LDB_OP_OR: // true dominates NULL
                  if ((!a.null && a.v != 0) || (!b.null && b.v != 0)) r.v = 1;
                  else if (a.null | b.null) r.null = true;
                  else r.v = 0;
                  break;
               case LDB_OP_NOT: r.v = a.v == 0; r.null = a.null; break;
               case LDB_OP_ISNULL: r.v = a.null; break;
               case LDB_OP_SELECT: {
                  const Val c = regs[in.arg];
                  r = (!c.null && c.v != 0) ? a : b;
                  break;
               }
               case LDB_OP_I2F: r = fromF64((double) a.v, a.null); break;
               case LDB_OP_FADD: r = fromF64(asF64(a) + asF64(b), a.null | b.null); break;
               case LDB_OP_FSUB: r = fromF64(asF64(a) - asF64(b), a.null | b.null); break;
               case LDB_OP_FMUL: r = fromF64(asF64(a) * asF64(b), a.null | b.null); break;
               case LDB_OP_FDIV: r = fromF64(asF64(a) / asF64(b), a.null | b.null); break;
               case LDB_OP_FCMP: r.v = cmpF(asF64(a), asF64(b), in.arg); r.null = a.null | b.null; break;
               case LDB_OP_STRCMP: {
                  const ProgCol& c = p.cols[in.a];
                  r.null = colIsNull(c, row);
                  if (!r.null) r.v = cmpI((s128) strCompare(c, row, p.strings[in.arg], p.stringLen[in.arg]), 0, in.b);
                  break;
               }
               case LDB_OP_STRLIKE: {
                  const ProgCol& c = p.cols[in.a];
                  r.null = colIsNull(c, row);
                  if (!r.null) r.v = strLike(c, row, p.strings[in.arg], p.stringLen[in.arg], in.b);
                  break;
               }
               case LDB_OP_YEAR: r.v = (s128) yearOfDays((int32_t) a.v); r.null = a.null; break;
               case LDB_OP_STRKEY8: {
                  const ProgCol& c = p.cols[in.a];
                  r.null = colIsNull(c, row);
                  if (!r.null) {
                     const int32_t* off = (const int32_t*) c.data + row;
                     const int32_t b0 = off[0], n = off[1] - off[0];
                     uint64_t k = 0;
                     for (int i = 0; i < 8; i++) k = (k << 8) | (i < n ? c.bytes[b0 + i] : 0);
                     r.v = (s128) (int64_t) k;
                  }
                  break;
               }
               case LDB_OP_PROBE: { // key → payload of a unique/multimap single-key table; absent key = NULL (semi / anti / mark / outer)
                  const JoinTableDev& t = p.tables[in.arg];
                  r.null = true;
                  if (!a.null && a.v == (s128) (int32_t) a.v) {
                     const int32_t key = (int32_t) a.v;
                     if (t.direct) {
                        const int32_t pay = directLoadProg(t, key);
                        if (pay != kDirectEmpty) {
                           r.v = pay;
                           r.null = false;
                        }
                     } else {
                        const uint64_t h = hashI32(key);
                        bool maybe = true;
                        if (t.bloom) {
                           const uint32_t bits = bloomBits(h);
                           maybe = (__ldg(&t.bloom[(uint32_t) (h >> 32) & t.bloomMask]) & bits) == bits;
                        }
                        uint64_t s = h & t.mask;
                        for (uint64_t probes = 0; maybe && probes <= t.mask && probes < 16384; probes++) {
                           const unsigned long long e = __ldg((const unsigned long long*) (t.base + s * t.stride));
                           if (e == ~0ull) break;
                           if ((int32_t) (uint32_t) e == key) {
                              r.v = (s128) (int32_t) ((uint32_t) (e >> 32) & (t.stride == 32 ? 0x7fffffffu : 0xffffffffu));
                              r.null = false;
                              break;
                           }
                           s = (s + 1) & t.mask;
                        }
                     }
                  }
                  break;
               }
               default: r.null = true;
            }
            regs[in.dst] = r;
         }
         if (p.filterReg >= 0) { // WHERE: NULL is not true
            const Val f = regs[p.filterReg];
            pass = !f.null && f.v != 0;
         }
      }
      if (p.sinkKind == 1) {
         if (pass) {
            int64_t keys[kProgMaxKeys];
            uint32_t nulls = 0;
            for (int k = 0; k < p.nKeys; k++) {
               const Val kv = regs[p.keyReg[k]];
               keys[k] = kv.null ? 0 : (int64_t) kv.v;
               nulls |= (kv.null ? 1u : 0u) << k;
            }
            uint8_t* e = hashAggFind(p, keys, nulls);
            if (e) hashAggUpdate(p, e, regs);
         }
      } else if (p.sinkKind == 2) {
         if (pass) {
            const Val k = regs[p.buildKeyReg];
            if (!k.null) { // NULL keys never match (SQL join semantics)
               const int32_t pay = p.buildPayloadReg >= 0 && !regs[p.buildPayloadReg].null ? (int32_t) regs[p.buildPayloadReg].v : 0;
               if (joinInsertProg(p.build, (int32_t) k.v, pay)) inserted++;
            }
         }
      } else if (p.sinkKind == 3) { // warp-aggregated append of the selected registers
         const unsigned m = __ballot_sync(0xffffffffu, pass);
         if (m) {
            const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
            unsigned long long pos = 0;
            if (lane == leader) pos = atomicAdd(p.outCount, (unsigned long long) __popc(m));
            pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(m & ((1u << lane) - 1));
            if (pass && pos < (unsigned long long) p.outCapacity) {
               for (int c = 0; c < p.nOut; c++) {
                  const Val x = regs[p.outReg[c]];
                  ulonglong2 cell;
                  cell.x = (unsigned long long) (uint64_t) x.v;
                  cell.y = (unsigned long long) (uint64_t) (x.v >> 64);
                  ((ulonglong2*) p.outValues[c])[pos] = cell;
                  p.outValid[c][pos] = x.null ? 0 : 1;
               }
            }
         }
      }
   }
   if (p.sinkKind == 2) {
      unsigned long long total = warpSum64(inserted);
      if ((threadIdx.x & 31) == 0 && total) atomicAdd(p.build.count, total);
   }
}
void launchProgram(const ProgramParams& p, int smCount, cudaStream_t s) {
   int grid = (int) std::min<int64_t>(std::max<int64_t>((p.nRows + 255) / 256, 1), (int64_t) smCount * 8);
   programKernel<<<grid, 256, 0, s>>>(p);
}

__global__ void hashAggInitKernel(HashAggDev t) {
   const uint64_t words = (t.mask + 1) * (uint64_t) t.entryBytes / 8;
   unsigned long long* w = (unsigned long long*) t.base;
   for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (uint64_t) gridDim.x * blockDim.x) w[i] = 0;
}
void launchHashAggInit(const HashAggDev& t, int smCount, cudaStream_t s) {
   const uint64_t words = (t.mask + 1) * (uint64_t) t.entryBytes / 8;
   int grid = (int) std::min<uint64_t>((words + 255) / 256, (uint64_t) smCount * 16);
   hashAggInitKernel<<<grid < 1 ? 1 : grid, 256, 0, s>>>(t);
}
struct ExportPtrs {
   int64_t* keyCols[kProgMaxKeys];
   uint8_t* keyValid[kProgMaxKeys];
   uint8_t* aggCols[kProgMaxAggs];
   uint8_t* aggValid[kProgMaxAggs];
};
// the scan over the hash table's entries that starts the reference's next pipeline (createIterator, PreAggregationHashtable.cpp:160-170),
// as a compaction into columns
__global__ void __launch_bounds__(256) hashAggExportKernel(HashAggDev t, ExportPtrs o, unsigned long long* counter, uint32_t countAggMask) {
   const uint64_t cap = t.mask + 1;
   for (uint64_t base = (uint64_t) blockIdx.x * blockDim.x; base < cap; base += (uint64_t) gridDim.x * blockDim.x) {
      const uint64_t s = base + threadIdx.x;
      const uint8_t* e = s < cap ? entryAt(t, s) : nullptr;
      const bool occ = e && *((const uint32_t*) e) == 2u;
      const unsigned m = __ballot_sync(0xffffffffu, occ);
      if (!m) continue;
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      unsigned long long pos = 0;
      if (lane == leader) pos = atomicAdd(counter, (unsigned long long) __popc(m));
      pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(m & ((1u << lane) - 1));
      if (!occ) continue;
      const uint32_t fl = ((const uint32_t*) e)[1];
      const int64_t* ek = (const int64_t*) (e + 16);
      for (int k = 0; k < t.nKeys; k++) {
         o.keyCols[k][pos] = ek[k];
         o.keyValid[k][pos] = ((fl >> (16 + k)) & 1u) ? 0 : 1;
      }
      const ulonglong2* ea = (const ulonglong2*) (e + 16 + 8 * kProgMaxKeys);
      for (int a = 0; a < t.nAggs; a++) {
         ((ulonglong2*) o.aggCols[a])[pos] = ea[a];
         o.aggValid[a][pos] = (((fl >> a) & 1u) || ((countAggMask >> a) & 1u)) ? 1 : 0; // counts are never NULL
      }
   }
}
void launchHashAggExport(const HashAggDev& t, int64_t* const* keyCols, uint8_t* const* keyValid, uint8_t* const* aggCols, uint8_t* const* aggValid, unsigned long long* counter, uint32_t countAggMask, int smCount, cudaStream_t s) {
   ExportPtrs o{};
   for (int k = 0; k < t.nKeys; k++) {
      o.keyCols[k] = keyCols[k];
      o.keyValid[k] = keyValid[k];
   }
   for (int a = 0; a < t.nAggs; a++) {
      o.aggCols[a] = aggCols[a];
      o.aggValid[a] = aggValid[a];
   }
   const uint64_t cap = t.mask + 1;
   int grid = (int) std::min<uint64_t>((cap + 255) / 256, (uint64_t) smCount * 8);
   hashAggExportKernel<<<grid < 1 ? 1 : grid, 256, 0, s>>>(t, o, counter, countAggMask);
}

// ---------------------------------------------------------------- radix sort (64-bit keys, 32-bit values), 8 passes of 8 bits
// (GrowingBuffer::sort → parallel sort of the materialised tuples, GrowingBuffer.cpp:54-78, Sorting.cpp; here LSD radix on an
//  order-preserving 64-bit key the host builds from the ORDER BY columns).  Stable; HBM-bound: 2 x (12 B read + 12 B write) per pass.
constexpr int kSortThreads = 256;
constexpr int kSortItemsPerCta = 256 * 16;
__global__ void __launch_bounds__(kSortThreads) sortHistKernel(const unsigned long long* keys, int64_t n, int shift, unsigned int* hist /* [256][gridDim.x] */) {
   __shared__ unsigned int h[256];
   h[threadIdx.x] = 0;
   __syncthreads();
   const int64_t begin = (int64_t) blockIdx.x * kSortItemsPerCta, end = begin + kSortItemsPerCta < n ? begin + kSortItemsPerCta : n;
   for (int64_t i = begin + threadIdx.x; i < end; i += kSortThreads) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
   __syncthreads();
   hist[(size_t) threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}
// exclusive scan over digit-major [256][nCtas] counts (single CTA; nCtas * 256 entries)
__global__ void __launch_bounds__(1024) sortScanKernel(unsigned int* hist, int64_t total) {
   __shared__ unsigned long long carry;
   __shared__ unsigned int warpSums[32];
   if (threadIdx.x == 0) carry = 0;
   __syncthreads();
   for (int64_t base = 0; base < total; base += 1024) {
      const int64_t i = base + threadIdx.x;
      unsigned int v = i < total ? hist[i] : 0u, x = v;
      for (int o = 1; o < 32; o <<= 1) {
         unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
         if ((threadIdx.x & 31) >= o) x += y;
      }
      if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = x;
      __syncthreads();
      if (threadIdx.x < 32) {
         unsigned int w = warpSums[threadIdx.x], ws = w;
         for (int o = 1; o < 32; o <<= 1) {
            unsigned int y = __shfl_up_sync(0xffffffffu, ws, o);
            if (threadIdx.x >= o) ws += y;
         }
         warpSums[threadIdx.x] = ws - w; // exclusive
      }
      __syncthreads();
      const unsigned long long excl = carry + warpSums[threadIdx.x >> 5] + (x - v);
      if (i < total) hist[i] = (unsigned int) excl;
      __syncthreads();
      if (threadIdx.x == 1023) carry = excl + v;
      __syncthreads();
   }
}
__global__ void __launch_bounds__(kSortThreads) sortScatterKernel(const unsigned long long* keys, const uint32_t* vals, unsigned long long* keysOut, uint32_t* valsOut, int64_t n, int shift, const unsigned int* hist) {
   __shared__ unsigned int running[256];      // global base + items of this digit already placed by this CTA
   __shared__ unsigned int warpCnt[8][256];   // per-warp digit counts of the current tile
   running[threadIdx.x] = hist[(size_t) threadIdx.x * gridDim.x + blockIdx.x];
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int64_t begin = (int64_t) blockIdx.x * kSortItemsPerCta, end = begin + kSortItemsPerCta < n ? begin + kSortItemsPerCta : n;
   for (int64_t tile = begin; tile < end; tile += kSortThreads) {
      for (int i = threadIdx.x; i < 8 * 256; i += kSortThreads) (&warpCnt[0][0])[i] = 0;
      __syncthreads();
      const int64_t i = tile + threadIdx.x;
      const bool valid = i < end;
      const unsigned long long k = valid ? keys[i] : 0ull;
      const unsigned d = valid ? (unsigned) ((k >> shift) & 255u) : 256u;
      const unsigned active = __ballot_sync(0xffffffffu, valid);
      unsigned rankInWarp = 0;
      if (valid) {
         const unsigned peers = __match_any_sync(active, d);
         rankInWarp = __popc(peers & ((1u << lane) - 1));
         if (rankInWarp == 0) warpCnt[warp][d] = __popc(peers);
      }
      __syncthreads();
      if (valid) {
         unsigned before = 0;
         for (int w = 0; w < warp; w++) before += warpCnt[w][d];
         const unsigned pos = running[d] + before + rankInWarp;
         keysOut[pos] = k;
         valsOut[pos] = vals[i];
      }
      __syncthreads();
      unsigned tot = 0;
      for (int w = 0; w < 8; w++) tot += warpCnt[w][threadIdx.x];
      running[threadIdx.x] += tot;
      __syncthreads();
   }
}
__global__ void buildSortKeysKernel(const uint8_t* col, int elemBytes, int64_t n, int descending, unsigned long long* keys, uint32_t* ids) {
   for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
      const int64_t v = elemBytes == 4 ? (int64_t) ((const int32_t*) col)[i] : *(const int64_t*) (col + (size_t) i * elemBytes);
      const unsigned long long k = (unsigned long long) v ^ 0x8000000000000000ull;
      keys[i] = descending ? ~k : k;
      ids[i] = (uint32_t) i;
   }
}
void launchBuildSortKeys(const uint8_t* col, int elemBytes, int64_t n, int descending, unsigned long long* keys, uint32_t* ids, int smCount, cudaStream_t s) {
   int grid = (int) std::min<int64_t>(std::max<int64_t>((n + 255) / 256, 1), (int64_t) smCount * 8);
   buildSortKeysKernel<<<grid, 256, 0, s>>>(col, elemBytes, n, descending, keys, ids);
}
void launchRadixSortPairs(unsigned long long* keys, uint32_t* vals, unsigned long long* keysTmp, uint32_t* valsTmp, int64_t n, unsigned int* histScratch, int smCount, cudaStream_t s) {
   if (n <= 0) return;
   const int ctas = (int) ((n + kSortItemsPerCta - 1) / kSortItemsPerCta);
   unsigned long long* kin = keys;
   unsigned long long* kout = keysTmp;
   uint32_t* vin = vals;
   uint32_t* vout = valsTmp;
   for (int pass = 0; pass < 8; pass++) {
      sortHistKernel<<<ctas, kSortThreads, 0, s>>>(kin, n, pass * 8, histScratch);
      sortScanKernel<<<1, 1024, 0, s>>>(histScratch, (int64_t) ctas * 256);
      sortScatterKernel<<<ctas, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, pass * 8, histScratch);
      std::swap(kin, kout);
      std::swap(vin, vout);
   }
   // 8 passes: the result is back in `keys` / `vals`
}

} // namespace ldb
