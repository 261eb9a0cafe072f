// dbgen_gen.h — dbgen-faithful TPC-H value functions (host + device), the compiled twin of lingo-db_b200/dbgen.py.
//
// Restates the data-generation algorithm of the TPC's dbgen (not under /root/reference: the reference downloads it,
// tools/generate/tpch.sh) for the columns TPC-H Q1/Q3/Q5/Q6/Q9 read.  One Park-Miller stream per column
// (seed' = seed * 16807 mod (2^31 - 1)) with dbgen's start seeds; every stream advances a FIXED number of draws per row
// (1 for orders/customer/supplier, 7 for lineitem columns, 4 for partsupp, 92 for p_name), so the k-th draw of row r is
// stream element r * perRow + k and can be reached directly with a modular power ("jump-ahead") — which makes the
// generator random-access and therefore usable from a CUDA grid.  The numpy version is validated against the
// reference's own SF1 answers (tests/test_reference_answers_sf1.py); this header is validated against the numpy version
// (tests/test_datagen.py::test_compiled_dbgen_twin_equals_numpy).
#pragma once
#include "tpch_gen.h" // LDB_HD

namespace ldbdbgen {

constexpr uint64_t kModulus = 2147483647ull, kMultiplier = 16807ull;
constexpr int32_t kEpochOffset = 83966; // dbgen day counter → days since 1970-01-01 (92001 = 1992-01-01 = 8035)
constexpr int32_t kMinDate = 92001, kCurrentEpochDay = 9298 /* 1995-06-17 */;

enum Seed : uint32_t {
   S_O_ORDERDATE = 1066728069u, S_O_CUSTKEY = 851767375u, S_O_LINECOUNT = 1434868289u,
   S_L_QUANTITY = 209208115u, S_L_DISCOUNT = 554590007u, S_L_TAX = 721958466u, S_L_PARTKEY = 1808217256u, S_L_SUPPNUM = 2095021727u,
   S_L_SHIPDATE = 1769349045u, S_L_COMMITDATE = 904914315u, S_L_RECEIPTDATE = 373135028u, S_L_RETURNFLAG = 717419739u,
   S_C_MKTSEGMENT = 1140279430u, S_C_NATIONKEY = 1489529863u, S_S_NATIONKEY = 110356601u, S_P_NAME = 709314158u, S_PS_SUPPLYCOST = 1051288424u,
};

LDB_HD uint64_t step(uint64_t seed) { return seed * kMultiplier % kModulus; }
// seed after n steps: seed * 16807^n mod (2^31 - 1), square-and-multiply
LDB_HD uint64_t jump(uint64_t seed, uint64_t n) {
   uint64_t base = kMultiplier, acc = seed % kModulus;
   while (n) {
      if (n & 1) acc = acc * base % kModulus;
      base = base * base % kModulus;
      n >>= 1;
   }
   return acc;
}
// dbgen UnifInt on an already advanced seed
LDB_HD int64_t unif(uint64_t seed, int64_t lo, int64_t hi) { return lo + (int64_t) (((double) seed / 2147483647.0) * (double) (hi - lo + 1)); }
// element k (0-based) of a stream = seed after k + 1 steps
LDB_HD uint64_t element(uint32_t startSeed, uint64_t k) { return jump(startSeed, k + 1); }

struct Scale {
   int64_t nOrders, nCustomer, nSupplier, nPart;
};

// ---------------------------------------------------------------- orders
LDB_HD int32_t orderKey(int64_t orderIdx) { // 0-based row → sparse key (keep 3 low bits, insert 2 zero bits)
   uint64_t i = (uint64_t) orderIdx + 1;
   return (int32_t) (((i >> 3) << 5) | (i & 7));
}
LDB_HD int32_t orderDateRaw(int64_t orderIdx) { return kMinDate + (int32_t) unif(element(S_O_ORDERDATE, (uint64_t) orderIdx), 0, 2557 - 151 - 1); }
LDB_HD int32_t orderCustKey(const Scale& s, int64_t orderIdx) {
   int64_t k = unif(element(S_O_CUSTKEY, (uint64_t) orderIdx), 1, s.nCustomer), delta = 1;
   while (k % 3 == 0) { // customer "mortality": a third of the customers never order
      k += delta;
      if (k > s.nCustomer) k = s.nCustomer;
      delta = -delta;
   }
   return (int32_t) k;
}
LDB_HD int32_t orderLineCount(int64_t orderIdx) { return (int32_t) unif(element(S_O_LINECOUNT, (uint64_t) orderIdx), 1, 7); }

// ---------------------------------------------------------------- lineitem: all lines of one order
struct Line {
   int32_t partkey, suppkey;
   int64_t quantity, extendedprice, discount, tax; // decimal(12,2) raw
   int32_t shipdate, commitdate, receiptdate;      // epoch days
   int32_t returnflag, linestatus;                 // fixed_size_binary(4): byte 0 = char
};
LDB_HD int64_t retailPrice(int64_t pk) { return 90000 + (pk / 10) % 20001 + 100 * (pk % 1000); }
LDB_HD int32_t partSupplier(const Scale& s, int64_t pk, int64_t j) { return (int32_t) ((pk + j * (s.nSupplier / 4 + (pk - 1) / s.nSupplier)) % s.nSupplier + 1); }
// fills out[0 .. count) and returns count
LDB_HD int32_t orderLines(const Scale& s, int64_t orderIdx, Line out[7]) {
   const int32_t count = orderLineCount(orderIdx), od = orderDateRaw(orderIdx);
   const uint64_t first = (uint64_t) orderIdx * 7; // the order's block in every 7-per-row stream
   uint64_t q = jump(S_L_QUANTITY, first), d = jump(S_L_DISCOUNT, first), t = jump(S_L_TAX, first), p = jump(S_L_PARTKEY, first), n = jump(S_L_SUPPNUM, first),
            sh = jump(S_L_SHIPDATE, first), cm = jump(S_L_COMMITDATE, first), rc = jump(S_L_RECEIPTDATE, first), fl = jump(S_L_RETURNFLAG, first);
   for (int k = 0; k < count; k++) {
      q = step(q), d = step(d), t = step(t), p = step(p), n = step(n), sh = step(sh), cm = step(cm), rc = step(rc);
      Line& l = out[k];
      const int64_t qty = unif(q, 1, 50), pk = unif(p, 1, s.nPart);
      l.quantity = qty * 100;
      l.discount = unif(d, 0, 10);
      l.tax = unif(t, 0, 8);
      l.partkey = (int32_t) pk;
      l.suppkey = partSupplier(s, pk, unif(n, 0, 3));
      l.extendedprice = qty * retailPrice(pk);
      const int32_t ship = od + (int32_t) unif(sh, 1, 121), commit = od + (int32_t) unif(cm, 30, 90), receipt = ship + (int32_t) unif(rc, 1, 30);
      l.shipdate = ship - kEpochOffset;
      l.commitdate = commit - kEpochOffset;
      l.receiptdate = receipt - kEpochOffset;
      if (l.receiptdate <= kCurrentEpochDay) { // the flag stream is drawn only for lines already received
         fl = step(fl);
         l.returnflag = unif(fl, 0, 1) == 0 ? 'R' : 'A';
      } else {
         l.returnflag = 'N';
      }
      l.linestatus = l.shipdate <= kCurrentEpochDay ? 'F' : 'O';
   }
   return count;
}

// ---------------------------------------------------------------- customer / supplier / part / partsupp
LDB_HD int32_t customerNationKey(int64_t i) { return (int32_t) unif(element(S_C_NATIONKEY, (uint64_t) i), 0, 24); }
LDB_HD int32_t customerSegment(int64_t i) { return (int32_t) unif(element(S_C_MKTSEGMENT, (uint64_t) i), 0, 4); } // AUTOMOBILE BUILDING FURNITURE HOUSEHOLD MACHINERY
LDB_HD int32_t supplierNationKey(int64_t i) { return (int32_t) unif(element(S_S_NATIONKEY, (uint64_t) i), 0, 24); }
LDB_HD int32_t segmentLen(int32_t seg) { return seg == 0 ? 10 : (seg == 1 ? 8 : 9); }
LDB_HD char segmentChar(int32_t seg, int32_t pos) {
   const char* names = "AUTOMOBILEBUILDING  FURNITURE HOUSEHOLD MACHINERY ";
   return names[seg * 10 + pos];
}
// p_name: the first five positions of a Fisher-Yates pass over a fresh identity permutation of the 92 colours
LDB_HD void partNameWords(int64_t partIdx, int32_t w[5]) {
   uint64_t seed = jump(S_P_NAME, (uint64_t) partIdx * 92);
   // only the first five positions are ever read, so track just the entries a swap has touched (at most 10)
   int32_t idx[10], val[10];
   int n = 0;
   auto get = [&](int32_t i) {
      for (int k = n - 1; k >= 0; k--)
         if (idx[k] == i) return val[k];
      return i;
   };
   auto set = [&](int32_t i, int32_t v) {
      for (int k = 0; k < n; k++)
         if (idx[k] == i) {
            val[k] = v;
            return;
         }
      idx[n] = i;
      val[n++] = v;
   };
   for (int pos = 0; pos < 5; pos++) {
      seed = step(seed);
      const int32_t sw = pos + (int32_t) (((double) seed / 2147483647.0) * (double) (92 - pos));
      const int32_t a = get(pos), b = get(sw);
      set(pos, b);
      set(sw, a);
   }
   for (int pos = 0; pos < 5; pos++) w[pos] = get(pos);
}
LDB_HD int32_t partNameLen(int64_t partIdx) {
   int32_t w[5];
   partNameWords(partIdx, w);
   int32_t len = 4;
   for (int k = 0; k < 5; k++) len += ldbgen::colorOffset(w[k] + 1) - ldbgen::colorOffset(w[k]);
   return len;
}
LDB_HD void partNameWrite(int64_t partIdx, uint8_t* out) {
   int32_t w[5];
   partNameWords(partIdx, w);
   const char* blob = ldbgen::colorBlob();
   for (int k = 0; k < 5; k++) {
      if (k) *out++ = ' ';
      for (int32_t c = ldbgen::colorOffset(w[k]); c < ldbgen::colorOffset(w[k] + 1); c++) *out++ = (uint8_t) blob[c];
   }
}
LDB_HD int32_t partSuppSuppKey(const Scale& s, int64_t r) { return partSupplier(s, r / 4 + 1, r % 4); }
LDB_HD int64_t partSuppSupplyCost(int64_t r) { return unif(element(S_PS_SUPPLYCOST, (uint64_t) r), 100, 100000); }

} // namespace ldbdbgen
