// program.h — device-side structures of the GENERIC pipeline (program.cu): scan → register program → sink.
//
// The hand-specialised kernels of kernels.cu cover the TPC-H hot shapes at HBM speed; everything else the sub-operator
// dialect can put into a scan pipeline — arbitrary expressions (db.add/sub/mul/div/cmp/and/or/not/between/case,
// LowerToStd.cpp:612-700,851-910), nullable inputs (validity bits, Restrictions.cpp:67-162), i8…i64/float/decimal(38)/string
// operands, SUM/COUNT/MIN/MAX/ANY with SQL null semantics (RelAlgToSubOp.cpp:1809-2025), group-by over millions of groups
// (PreAggregationHashtable.cpp:76-170), semi/anti/mark probes (RelAlgToSubOp.cpp:1129-1206,1340-1588) — runs through ONE
// kernel that interprets a small register program per row.  It is the GPU stand-in for "whatever the JIT would have emitted".
#pragma once
#include "kernels.h"

namespace ldb {

constexpr int kProgMaxInstr = 96;
constexpr int kProgMaxRegs = 48;
constexpr int kProgMaxCols = 12;
constexpr int kProgMaxConsts = 24;
constexpr int kProgMaxStrings = 12;
constexpr int kProgStringBytes = 32;
constexpr int kProgMaxTables = 4;
constexpr int kProgMaxKeys = 4;
constexpr int kProgMaxAggs = 8;

struct ProgCol {
   const uint8_t* data;     // values, or utf8 offsets (int32)
   const uint8_t* bytes;    // utf8 data
   const uint8_t* validity; // Arrow validity bitmap (LSB first) or null = no nulls
   const uint8_t* validBytes; // one validity BYTE per row (columns produced by this library: exported groups, materialised rows)
   int64_t bitOffset;       // bit index of row 0 inside `validity`
   int32_t type;            // LdbPhysType
   int32_t elemBytes;       // as staged (decimal128: 16, or 8 when the HOST batch was narrowed)
};
struct ProgInstr {
   uint8_t op, dst, a, b;
   int32_t arg;
};
struct ProgAgg {
   int32_t kind; // LdbAggKind
   int32_t reg;
};
// large-domain hash aggregation table (rt::PreAggregationHashtable after merge / rt::Hashtable): open addressing in HBM
//   entry = { state:u32 (0 empty, 1 being written, 2 ready), flags:u32 (bit a: aggregate a has seen a non-null input; bit 8+a: claimed
//             by an ANY; bit 16+k: key k is NULL), pad:u64, keys[4]:i64, aggs[nAggs] x {lo:u64, hi:u64} } = 48 + 16 nAggs bytes
struct HashAggDev {
   uint8_t* base;
   uint64_t mask; // capacity - 1
   uint32_t entryBytes;
   int32_t nKeys, nAggs;
   unsigned long long* count; // groups
   int32_t* error;            // 1 = table full
};
struct ProgramParams {
   int64_t nRows;
   int32_t nCols, nInstr, nTables;
   ProgCol cols[kProgMaxCols];
   ProgInstr instr[kProgMaxInstr];
   unsigned long long constLo[kProgMaxConsts];
   long long constHi[kProgMaxConsts];
   uint8_t strings[kProgMaxStrings][kProgStringBytes];
   int32_t stringLen[kProgMaxStrings];
   JoinTableDev tables[kProgMaxTables];
   int32_t filterReg; // -1: every row passes
   int32_t sinkKind;  // 1 hash aggregation, 2 join-table build, 3 materialize
   // sink 1
   int32_t nKeys, nAggs;
   int32_t keyReg[kProgMaxKeys];
   ProgAgg aggs[kProgMaxAggs];
   HashAggDev agg;
   // sink 2
   int32_t buildKeyReg, buildPayloadReg;
   JoinTableDev build;
   // sink 3: compacted output columns, 16 bytes per value (i128 / double bits in lo) + 1 validity byte
   int32_t nOut;
   int32_t outReg[kProgMaxAggs];
   uint8_t* outValues[kProgMaxAggs];
   uint8_t* outValid[kProgMaxAggs];
   int64_t outCapacity;
   unsigned long long* outCount;
};

void launchProgram(const ProgramParams& p, int smCount, cudaStream_t s);
void launchHashAggInit(const HashAggDev& t, int smCount, cudaStream_t s);
// compacts the occupied entries into columnar buffers: keys (int64 + validity byte), aggregates (16 B + validity byte)
void launchHashAggExport(const HashAggDev& t, int64_t* const* keyCols, uint8_t* const* keyValid, uint8_t* const* aggCols, uint8_t* const* aggValid, unsigned long long* counter, uint32_t countAggMask, int smCount, cudaStream_t s);
// LSD radix sort of (64-bit key, 32-bit row id) pairs — ORDER BY / top-k over materialised rows (GrowingBuffer::sort, Sorting.cpp)
// order-preserving 64-bit sort keys + row ids from one fixed-width column (low 8 bytes of a cell, sign bit flipped; inverted for DESC)
void launchBuildSortKeys(const uint8_t* col, int elemBytes, int64_t n, int descending, unsigned long long* keys, uint32_t* ids, int smCount, cudaStream_t s);
void launchRadixSortPairs(unsigned long long* keys, uint32_t* vals, unsigned long long* keysTmp, uint32_t* valsTmp, int64_t n, unsigned int* histScratch, int smCount, cudaStream_t s);

} // namespace ldb
