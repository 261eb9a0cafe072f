// tpch_gen.h — deterministic, counter-based TPC-H-shaped table generator (host + device).
//
// Every value is a pure function of (seed, table, column, row index), so any row range of any
// table can be produced on any GPU or on the host and is bit-identical everywhere.  This is the
// "generator-synthesised tables" input of BASELINE.json / SURVEY.md §8(d); it is NOT dbgen (the
// reference's tools/generate/tpch.sh downloads dbgen, which is impossible here), but it follows
// the TPC-H value distributions the survey lists so that selectivities and cardinalities match:
//   - sparse o_orderkey (8 of every 32), o_custkey never a multiple of 3
//   - 1..7 lineitems per order (every 7 consecutive orders hold a permutation of 1..7 → exactly
//     28 lineitems per 7 orders, which gives a closed-form order→first-row prefix, so lineitem
//     rows are randomly addressable without a scan)
//   - l_extendedprice = l_quantity * p_retailprice(l_partkey), dates/flags per the spec
// Physical types are the reference's Arrow types (src/runtime/storage/LingoDBTable.cpp:122-195):
//   integer→int32, decimal(12,2)→decimal128 (16 B little-endian two's complement),
//   date→date32 (days), char(1)→fixed_size_binary(4) (byte 0 = char), char(n)/varchar→utf8.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LDB_HD __host__ __device__ __forceinline__
#else
#define LDB_HD inline
#endif

namespace ldbgen {

enum Table : uint32_t { T_ORDERS = 1, T_LINEITEM = 2, T_CUSTOMER = 3, T_SUPPLIER = 4, T_PART = 5, T_PARTSUPP = 6 };

// Column tags (only used to decorrelate the random streams).
enum Col : uint32_t {
   C_O_CUSTKEY = 1, C_O_ORDERDATE = 2, C_O_PERM = 3,
   C_L_PARTKEY = 10, C_L_SUPPJ = 11, C_L_QUANTITY = 12, C_L_DISCOUNT = 13, C_L_TAX = 14,
   C_L_SHIPDELTA = 15, C_L_COMMITDELTA = 16, C_L_RECEIPTDELTA = 17, C_L_RFLAG = 18,
   C_C_NATIONKEY = 30, C_C_MKTSEGMENT = 31,
   C_S_NATIONKEY = 40,
   C_P_NAME = 50,
   C_PS_SUPPLYCOST = 60,
};

constexpr int32_t DATE_1992_01_01 = 8035;  // days since 1970-01-01
constexpr int32_t DATE_1995_06_17 = 9298;  // dbgen CURRENTDATE
constexpr int32_t ORDERDATE_SPAN = 2406;   // 1992-01-01 .. 1998-08-02 inclusive

struct Scale {
   uint64_t seed;
   int64_t nOrders, nCustomer, nSupplier, nPart;
   LDB_HD int64_t nLineitem() const;
};

LDB_HD uint64_t mix64(uint64_t z) { // splitmix64 finaliser
   z += 0x9E3779B97F4A7C15ull;
   z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
   z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
   return z ^ (z >> 31);
}
LDB_HD uint64_t rnd(uint64_t seed, uint32_t table, uint32_t col, uint64_t idx) {
   return mix64(mix64(seed ^ (uint64_t(table) << 56) ^ (uint64_t(col) << 48)) + idx);
}
LDB_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
   return __umul64hi(a, b);
#else
   return (uint64_t) (((unsigned __int128) a * b) >> 64);
#endif
}
// uniform integer in [lo, hi] (inclusive), division-free
LDB_HD int64_t uniform(uint64_t r, int64_t lo, int64_t hi) {
   return lo + (int64_t) mulhi64(r, (uint64_t) (hi - lo + 1));
}

// ---------------------------------------------------------------- orders
LDB_HD int32_t orderKey(int64_t orderIdx) { // dbgen mk_sparse: keep 3 low bits, insert 2 zero bits
   uint64_t i = (uint64_t) orderIdx + 1;
   return (int32_t) (((i >> 3) << 5) | (i & 7));
}
LDB_HD int32_t orderCustKey(const Scale& s, int64_t orderIdx) {
   int64_t nonMult3 = s.nCustomer - s.nCustomer / 3;
   int64_t k = uniform(rnd(s.seed, T_ORDERS, C_O_CUSTKEY, orderIdx), 0, nonMult3 - 1);
   return (int32_t) (k + k / 2 + 1); // 0→1, 1→2, 2→4, 3→5 … never a multiple of 3
}
LDB_HD int32_t orderDate(const Scale& s, int64_t orderIdx) {
   return DATE_1992_01_01 + (int32_t) uniform(rnd(s.seed, T_ORDERS, C_O_ORDERDATE, orderIdx), 0, ORDERDATE_SPAN - 1);
}
LDB_HD int32_t orderShipPriority(const Scale&, int64_t) { return 0; }

// lineitems per order: block b = orderIdx/7 holds a hash-chosen permutation of {1..7}
LDB_HD void linePerm(const Scale& s, int64_t block, int32_t perm[7]) {
   uint64_t r = rnd(s.seed, T_ORDERS, C_O_PERM, block);
   for (int i = 0; i < 7; i++) perm[i] = i + 1;
   for (int i = 6; i > 0; i--) { // Fisher–Yates with 8 hash bits per step
      int j = (int) (((r & 0xFF) * (uint64_t) (i + 1)) >> 8);
      r >>= 8;
      int32_t t = perm[i];
      perm[i] = perm[j];
      perm[j] = t;
   }
}
// first lineitem row of an order and its line count
LDB_HD int64_t orderFirstLine(const Scale& s, int64_t orderIdx, int32_t* count) {
   int32_t perm[7];
   int64_t block = orderIdx / 7;
   int within = (int) (orderIdx % 7);
   linePerm(s, block, perm);
   int64_t first = block * 28;
   for (int j = 0; j < 7; j++) {
      if (j < within) first += perm[j];
   }
   *count = perm[within];
   return first;
}
LDB_HD int64_t Scale::nLineitem() const {
   if (nOrders == 0) return 0;
   int32_t c;
   int64_t f = orderFirstLine(*this, nOrders - 1, &c);
   return f + c;
}
// inverse: lineitem row → (order index, 0-based line number)
LDB_HD int64_t lineToOrder(const Scale& s, int64_t row, int32_t* lineNo) {
   int32_t perm[7];
   int64_t block = row / 28;
   int32_t q = (int32_t) (row % 28);
   linePerm(s, block, perm);
   int j = 0;
   for (int k = 0; k < 6; k++) {
      if (q >= perm[j]) {
         q -= perm[j];
         j++;
      }
   }
   *lineNo = q;
   return block * 7 + j;
}

// ---------------------------------------------------------------- lineitem
struct LineItem {
   int32_t orderkey, partkey, suppkey, linenumber;
   int64_t quantity, extendedprice, discount, tax; // decimal(12,2) raw (scale 2)
   int32_t shipdate, commitdate, receiptdate;
   int32_t returnflag, linestatus; // fixed_size_binary(4) as little-endian int32 (byte 0 = char)
};
LDB_HD int64_t partRetailPrice(int64_t pk) { // cents, TPC-H spec 4.2.3
   return 90000 + ((pk / 10) % 20001) + 100 * (pk % 1000);
}
LDB_HD LineItem lineItem(const Scale& s, int64_t row) {
   LineItem l;
   int32_t ln;
   int64_t o = lineToOrder(s, row, &ln);
   uint64_t id = (uint64_t) row;
   l.orderkey = orderKey(o);
   l.linenumber = ln + 1;
   int64_t pk = uniform(rnd(s.seed, T_LINEITEM, C_L_PARTKEY, id), 1, s.nPart);
   int64_t j = uniform(rnd(s.seed, T_LINEITEM, C_L_SUPPJ, id), 0, 3);
   int64_t S = s.nSupplier;
   l.partkey = (int32_t) pk;
   l.suppkey = (int32_t) ((pk + j * (S / 4 + (pk - 1) / S)) % S + 1);
   int64_t qty = uniform(rnd(s.seed, T_LINEITEM, C_L_QUANTITY, id), 1, 50);
   l.quantity = qty * 100;
   l.extendedprice = qty * partRetailPrice(pk);
   l.discount = uniform(rnd(s.seed, T_LINEITEM, C_L_DISCOUNT, id), 0, 10);
   l.tax = uniform(rnd(s.seed, T_LINEITEM, C_L_TAX, id), 0, 8);
   int32_t od = orderDate(s, o);
   l.shipdate = od + (int32_t) uniform(rnd(s.seed, T_LINEITEM, C_L_SHIPDELTA, id), 1, 121);
   l.commitdate = od + (int32_t) uniform(rnd(s.seed, T_LINEITEM, C_L_COMMITDELTA, id), 30, 90);
   l.receiptdate = l.shipdate + (int32_t) uniform(rnd(s.seed, T_LINEITEM, C_L_RECEIPTDELTA, id), 1, 30);
   if (l.receiptdate <= DATE_1995_06_17) {
      l.returnflag = (rnd(s.seed, T_LINEITEM, C_L_RFLAG, id) & 1) ? 'R' : 'A';
   } else {
      l.returnflag = 'N';
   }
   l.linestatus = l.shipdate > DATE_1995_06_17 ? 'O' : 'F';
   return l;
}

// ---------------------------------------------------------------- customer / supplier
LDB_HD int32_t customerNationKey(const Scale& s, int64_t i) {
   return (int32_t) uniform(rnd(s.seed, T_CUSTOMER, C_C_NATIONKEY, i), 0, 24);
}
LDB_HD int32_t customerSegment(const Scale& s, int64_t i) { // index into SEGMENTS
   return (int32_t) uniform(rnd(s.seed, T_CUSTOMER, C_C_MKTSEGMENT, i), 0, 4);
}
LDB_HD int32_t supplierNationKey(const Scale& s, int64_t i) {
   return (int32_t) uniform(rnd(s.seed, T_SUPPLIER, C_S_NATIONKEY, i), 0, 24);
}
LDB_HD int32_t segmentLen(int32_t seg) {
   // AUTOMOBILE BUILDING FURNITURE MACHINERY HOUSEHOLD
   return seg == 0 ? 10 : (seg == 1 ? 8 : 9);
}
LDB_HD char segmentChar(int32_t seg, int32_t pos) {
   const char* names = "AUTOMOBILEBUILDING  FURNITURE MACHINERY HOUSEHOLD ";
   return names[seg * 10 + pos];
}

// ---------------------------------------------------------------- part / partsupp (Q9)
// p_name = five DISTINCT words of dbgen's 92-word colour list, blank separated (TPC-H spec 4.2.3, P_NAME);
// `p_name like '%green%'` therefore selects 5/92 = 5.4 percent of the parts.
constexpr int32_t N_COLORS = 92;
LDB_HD const char* colorBlob() {
   return "almondantiqueaquamarineazurebeigebisqueblackblanchedblueblushbrownburlywoodburnishedchartreusechiffonchocolatecoralcornflowercornsilkcreamcyandarkdeepdimdodgerdrabfirebrickfloralforestfrostedgainsboroghostgoldenrodgreengreyhoneydewhotindianivorykhakilacelavenderlawnlemonlightlimelinenmagentamaroonmediummetallicmidnightmintmistymoccasinnavajonavyoliveorangeorchidpalepapayapeachperupinkplumpowderpuffpurpleredroserosyroyalsaddlesalmonsandyseashellsiennaskyslatesmokesnowspringsteeltanthistletomatoturquoisevioletwheatwhiteyellow";
}
LDB_HD int32_t colorOffset(int32_t i) {
   // prefix sums of the word lengths (93 entries)
   const uint16_t offs[93] = {0, 6, 13, 23, 28, 33, 39, 44, 52, 56, 61, 66, 75, 84, 94, 101, 110, 115, 125, 133, 138, 142, 146, 150, 153, 159, 163, 172, 178, 184, 191, 200, 205, 214, 219, 223, 231, 234, 240, 245, 250, 254, 262, 266, 271, 276, 280, 285, 292, 298, 304, 312, 320, 324, 329, 337, 343, 347, 352, 358, 364, 368, 374, 379, 383, 387, 391, 397, 401, 407, 410, 414, 418, 423, 429, 435, 440, 448, 454, 457, 462, 467, 471, 477, 482, 485, 492, 498, 507, 513, 518, 523, 529};
   return offs[i];
}
LDB_HD void partNameWords(const Scale& s, int64_t partIdx, int32_t w[5]) {
   uint64_t r = rnd(s.seed, T_PART, C_P_NAME, (uint64_t) partIdx);
   // draw without replacement: x_k uniform over the (92-k) words not yet taken, mapped past the taken ones in ascending order
   int32_t taken[5];
   for (int k = 0; k < 5; k++) {
      int32_t x = (int32_t) (((r & 0xFFF) * (uint64_t) (N_COLORS - k)) >> 12);
      r >>= 12;
      for (int j = 0; j < k; j++)
         if (x >= taken[j]) x++;
      w[k] = x;
      int j = k; // keep `taken` sorted
      while (j > 0 && taken[j - 1] > x) {
         taken[j] = taken[j - 1];
         j--;
      }
      taken[j] = x;
   }
}
LDB_HD int32_t partNameLen(const Scale& s, int64_t partIdx) {
   int32_t w[5];
   partNameWords(s, partIdx, w);
   int32_t len = 4;
   for (int k = 0; k < 5; k++) len += colorOffset(w[k] + 1) - colorOffset(w[k]);
   return len;
}
LDB_HD void partNameWrite(const Scale& s, int64_t partIdx, uint8_t* out) {
   int32_t w[5];
   partNameWords(s, partIdx, w);
   const char* blob = colorBlob();
   for (int k = 0; k < 5; k++) {
      if (k) *out++ = ' ';
      for (int32_t c = colorOffset(w[k]); c < colorOffset(w[k] + 1); c++) *out++ = (uint8_t) blob[c];
   }
}
// partsupp row r: part r/4 + 1, its i-th supplier (spec 4.2.3 PS_SUPPKEY; the same formula picks l_suppkey above)
LDB_HD int32_t partSuppPartKey(int64_t r) { return (int32_t) (r / 4 + 1); }
LDB_HD int32_t partSuppSuppKey(const Scale& s, int64_t r) {
   int64_t pk = r / 4 + 1, i = r % 4, S = s.nSupplier;
   return (int32_t) ((pk + i * (S / 4 + (pk - 1) / S)) % S + 1);
}
LDB_HD int64_t partSuppSupplyCost(const Scale& s, int64_t r) { // cents, U[1.00, 1000.00]
   return uniform(rnd(s.seed, T_PARTSUPP, C_PS_SUPPLYCOST, (uint64_t) r), 100, 100000);
}

} // namespace ldbgen
