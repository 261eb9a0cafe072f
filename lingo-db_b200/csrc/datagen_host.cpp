// datagen_host.cpp — host side of the deterministic TPC-H-shaped generator (tpch_gen.h).
// Fills caller-provided column buffers in the reference's Arrow physical layout
// (src/runtime/storage/LingoDBTable.cpp:122-195).  No CUDA dependency: this library feeds the
// CPU oracle and the host-buffer (e2e) path of the C-ABI; the device twin lives in datagen.cu.
#include "tpch_gen.h"
#include "dbgen_gen.h"
#include "../../include/ldb_datagen.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

using namespace ldbgen;

namespace {
Scale toScale(const LdbGenScale* g) {
   Scale s;
   s.seed = g->seed;
   s.nOrders = g->n_orders;
   s.nCustomer = g->n_customer;
   s.nSupplier = g->n_supplier;
   s.nPart = g->n_part;
   return s;
}
inline void storeDec(uint8_t* col, int64_t i, int64_t v) { // decimal128 little endian, sign-extended
   int64_t* p = reinterpret_cast<int64_t*>(col + 16 * i);
   p[0] = v;
   p[1] = v < 0 ? -1 : 0;
}
template <class Fn>
void parallelFor(int64_t n, const Fn& fn) {
   unsigned hw = std::max(1u, std::thread::hardware_concurrency());
   int64_t nThreads = std::min<int64_t>(hw, std::max<int64_t>(1, n / 65536));
   if (nThreads <= 1) {
      fn(0, n);
      return;
   }
   std::vector<std::thread> ts;
   int64_t per = (n + nThreads - 1) / nThreads;
   for (int64_t t = 0; t < nThreads; t++) {
      int64_t b = t * per, e = std::min(n, b + per);
      if (b >= e) break;
      ts.emplace_back([=, &fn] { fn(b, e); });
   }
   for (auto& t : ts) t.join();
}
} // namespace

extern "C" {

void ldbgen_scale(double sf, uint64_t seed, LdbGenScale* out) {
   auto r = [&](double base, int64_t mn) { return std::max<int64_t>(mn, (int64_t) (base * sf + 0.5)); };
   out->seed = seed;
   out->n_orders = r(1500000.0, 7);
   out->n_customer = r(150000.0, 30);
   out->n_supplier = r(10000.0, 8);
   out->n_part = r(200000.0, 40);
   Scale s = toScale(out);
   out->n_lineitem = s.nLineitem();
}

int64_t ldbgen_order_first_line(const LdbGenScale* g, int64_t order_idx) {
   Scale s = toScale(g);
   if (order_idx >= s.nOrders) return s.nLineitem();
   int32_t c;
   return orderFirstLine(s, order_idx, &c);
}

void ldbgen_lineitem_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenLineitemCols* c) {
   Scale s = toScale(g);
   parallelFor(n_rows, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; i++) {
         LineItem l = lineItem(s, row_begin + i);
         if (c->l_orderkey) c->l_orderkey[i] = l.orderkey;
         if (c->l_partkey) c->l_partkey[i] = l.partkey;
         if (c->l_suppkey) c->l_suppkey[i] = l.suppkey;
         if (c->l_quantity) storeDec(c->l_quantity, i, l.quantity);
         if (c->l_extendedprice) storeDec(c->l_extendedprice, i, l.extendedprice);
         if (c->l_discount) storeDec(c->l_discount, i, l.discount);
         if (c->l_tax) storeDec(c->l_tax, i, l.tax);
         if (c->l_returnflag) c->l_returnflag[i] = l.returnflag;
         if (c->l_linestatus) c->l_linestatus[i] = l.linestatus;
         if (c->l_shipdate) c->l_shipdate[i] = l.shipdate;
         if (c->l_commitdate) c->l_commitdate[i] = l.commitdate;
         if (c->l_receiptdate) c->l_receiptdate[i] = l.receiptdate;
      }
   });
}

void ldbgen_orders_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenOrdersCols* c) {
   Scale s = toScale(g);
   parallelFor(n_rows, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; i++) {
         int64_t o = row_begin + i;
         if (c->o_orderkey) c->o_orderkey[i] = orderKey(o);
         if (c->o_custkey) c->o_custkey[i] = orderCustKey(s, o);
         if (c->o_orderdate) c->o_orderdate[i] = orderDate(s, o);
         if (c->o_shippriority) c->o_shippriority[i] = orderShipPriority(s, o);
      }
   });
}

// utf8 column: offsets has n_rows+1 entries, relative to this batch (offsets[0] = 0).
// Returns the number of data bytes; pass data == NULL to size the buffer first.
int64_t ldbgen_customer_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenCustomerCols* c) {
   Scale s = toScale(g);
   parallelFor(n_rows, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; i++) {
         int64_t r = row_begin + i;
         if (c->c_custkey) c->c_custkey[i] = (int32_t) (r + 1);
         if (c->c_nationkey) c->c_nationkey[i] = customerNationKey(s, r);
      }
   });
   int64_t bytes = 0;
   for (int64_t i = 0; i < n_rows; i++) {
      int32_t seg = customerSegment(s, row_begin + i);
      int32_t len = segmentLen(seg);
      if (c->c_mktsegment_offsets) c->c_mktsegment_offsets[i] = (int32_t) bytes;
      if (c->c_mktsegment_data) {
         for (int32_t k = 0; k < len; k++) c->c_mktsegment_data[bytes + k] = (uint8_t) segmentChar(seg, k);
      }
      bytes += len;
   }
   if (c->c_mktsegment_offsets) c->c_mktsegment_offsets[n_rows] = (int32_t) bytes;
   return bytes;
}

void ldbgen_supplier_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenSupplierCols* c) {
   Scale s = toScale(g);
   for (int64_t i = 0; i < n_rows; i++) {
      int64_t r = row_begin + i;
      if (c->s_suppkey) c->s_suppkey[i] = (int32_t) (r + 1);
      if (c->s_nationkey) c->s_nationkey[i] = supplierNationKey(s, r);
   }
}

int64_t ldbgen_part_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenPartCols* c) {
   Scale s = toScale(g);
   int64_t bytes = 0;
   for (int64_t i = 0; i < n_rows; i++) {
      int64_t r = row_begin + i;
      if (c->p_partkey) c->p_partkey[i] = (int32_t) (r + 1);
      if (c->p_name_offsets) c->p_name_offsets[i] = (int32_t) bytes;
      if (c->p_name_data) partNameWrite(s, r, c->p_name_data + bytes);
      bytes += partNameLen(s, r);
   }
   if (c->p_name_offsets) c->p_name_offsets[n_rows] = (int32_t) bytes;
   return bytes;
}

void ldbgen_partsupp_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenPartsuppCols* c) {
   Scale s = toScale(g);
   parallelFor(n_rows, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; i++) {
         int64_t r = row_begin + i;
         if (c->ps_partkey) c->ps_partkey[i] = partSuppPartKey(r);
         if (c->ps_suppkey) c->ps_suppkey[i] = partSuppSuppKey(s, r);
         if (c->ps_supplycost) storeDec(c->ps_supplycost, i, partSuppSupplyCost(s, r));
      }
   });
}

// ------------------------------------------------------------------------------------------------ dbgen-faithful variant
static ldbdbgen::Scale toDbgenScale(const LdbGenScale* g) { return ldbdbgen::Scale{g->n_orders, g->n_customer, g->n_supplier, g->n_part}; }

void ldbgen_dbgen_scale(double sf, int32_t count_lines, LdbGenScale* out) {
   out->seed = 0;
   out->n_orders = (int64_t) (1500000.0 * sf);
   out->n_customer = (int64_t) (150000.0 * sf);
   out->n_supplier = (int64_t) (10000.0 * sf);
   out->n_part = (int64_t) (200000.0 * sf);
   out->n_lineitem = 0;
   if (count_lines) {
      std::vector<int64_t> partial(256, 0);
      const int64_t n = out->n_orders, per = (n + 255) / 256;
      parallelFor(256, [&](int64_t b, int64_t e) {
         for (int64_t t = b; t < e; t++) {
            const int64_t lo = t * per, hi = std::min(n, lo + per);
            if (lo >= hi) continue;
            uint64_t seed = ldbdbgen::jump(ldbdbgen::S_O_LINECOUNT, (uint64_t) lo);
            int64_t sum = 0;
            for (int64_t o = lo; o < hi; o++) {
               seed = ldbdbgen::step(seed);
               sum += ldbdbgen::unif(seed, 1, 7);
            }
            partial[t] = sum;
         }
      });
      for (auto v : partial) out->n_lineitem += v;
   }
}
void ldbgen_dbgen_line_counts_host(const LdbGenScale*, int64_t order_begin, int64_t n_orders, int32_t* counts) {
   parallelFor(n_orders, [&](int64_t b, int64_t e) {
      uint64_t seed = ldbdbgen::jump(ldbdbgen::S_O_LINECOUNT, (uint64_t) (order_begin + b));
      for (int64_t i = b; i < e; i++) {
         seed = ldbdbgen::step(seed);
         counts[i] = (int32_t) ldbdbgen::unif(seed, 1, 7);
      }
   });
}
void ldbgen_dbgen_lineitem_host(const LdbGenScale* g, int64_t order_begin, int64_t n_orders, const int64_t* first_row, const LdbGenLineitemCols* c) {
   const ldbdbgen::Scale s = toDbgenScale(g);
   parallelFor(n_orders, [&](int64_t b, int64_t e) {
      ldbdbgen::Line lines[7];
      for (int64_t o = b; o < e; o++) {
         const int32_t n = ldbdbgen::orderLines(s, order_begin + o, lines);
         const int32_t key = ldbdbgen::orderKey(order_begin + o);
         for (int32_t k = 0; k < n; k++) {
            const int64_t i = first_row[o] + k;
            const ldbdbgen::Line& l = lines[k];
            if (c->l_orderkey) c->l_orderkey[i] = key;
            if (c->l_partkey) c->l_partkey[i] = l.partkey;
            if (c->l_suppkey) c->l_suppkey[i] = l.suppkey;
            if (c->l_quantity) storeDec(c->l_quantity, i, l.quantity);
            if (c->l_extendedprice) storeDec(c->l_extendedprice, i, l.extendedprice);
            if (c->l_discount) storeDec(c->l_discount, i, l.discount);
            if (c->l_tax) storeDec(c->l_tax, i, l.tax);
            if (c->l_returnflag) c->l_returnflag[i] = l.returnflag;
            if (c->l_linestatus) c->l_linestatus[i] = l.linestatus;
            if (c->l_shipdate) c->l_shipdate[i] = l.shipdate;
            if (c->l_commitdate) c->l_commitdate[i] = l.commitdate;
            if (c->l_receiptdate) c->l_receiptdate[i] = l.receiptdate;
         }
      }
   });
}
void ldbgen_dbgen_orders_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenOrdersCols* c) {
   const ldbdbgen::Scale s = toDbgenScale(g);
   parallelFor(n_rows, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; i++) {
         const int64_t o = row_begin + i;
         if (c->o_orderkey) c->o_orderkey[i] = ldbdbgen::orderKey(o);
         if (c->o_custkey) c->o_custkey[i] = ldbdbgen::orderCustKey(s, o);
         if (c->o_orderdate) c->o_orderdate[i] = ldbdbgen::orderDateRaw(o) - ldbdbgen::kEpochOffset;
         if (c->o_shippriority) c->o_shippriority[i] = 0;
      }
   });
}
int64_t ldbgen_dbgen_customer_host(const LdbGenScale*, int64_t row_begin, int64_t n_rows, const LdbGenCustomerCols* c) {
   int64_t bytes = 0;
   for (int64_t i = 0; i < n_rows; i++) {
      const int64_t r = row_begin + i;
      if (c->c_custkey) c->c_custkey[i] = (int32_t) (r + 1);
      if (c->c_nationkey) c->c_nationkey[i] = ldbdbgen::customerNationKey(r);
      const int32_t seg = ldbdbgen::customerSegment(r), len = ldbdbgen::segmentLen(seg);
      if (c->c_mktsegment_offsets) c->c_mktsegment_offsets[i] = (int32_t) bytes;
      if (c->c_mktsegment_data)
         for (int32_t k = 0; k < len; k++) c->c_mktsegment_data[bytes + k] = (uint8_t) ldbdbgen::segmentChar(seg, k);
      bytes += len;
   }
   if (c->c_mktsegment_offsets) c->c_mktsegment_offsets[n_rows] = (int32_t) bytes;
   return bytes;
}
void ldbgen_dbgen_supplier_host(const LdbGenScale*, int64_t row_begin, int64_t n_rows, const LdbGenSupplierCols* c) {
   for (int64_t i = 0; i < n_rows; i++) {
      if (c->s_suppkey) c->s_suppkey[i] = (int32_t) (row_begin + i + 1);
      if (c->s_nationkey) c->s_nationkey[i] = ldbdbgen::supplierNationKey(row_begin + i);
   }
}
int64_t ldbgen_dbgen_part_host(const LdbGenScale*, int64_t row_begin, int64_t n_rows, const LdbGenPartCols* c) {
   int64_t bytes = 0;
   for (int64_t i = 0; i < n_rows; i++) {
      const int64_t r = row_begin + i;
      if (c->p_partkey) c->p_partkey[i] = (int32_t) (r + 1);
      if (c->p_name_offsets) c->p_name_offsets[i] = (int32_t) bytes;
      if (c->p_name_data) ldbdbgen::partNameWrite(r, c->p_name_data + bytes);
      bytes += ldbdbgen::partNameLen(r);
   }
   if (c->p_name_offsets) c->p_name_offsets[n_rows] = (int32_t) bytes;
   return bytes;
}
void ldbgen_dbgen_partsupp_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenPartsuppCols* c) {
   const ldbdbgen::Scale s = toDbgenScale(g);
   parallelFor(n_rows, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; i++) {
         const int64_t r = row_begin + i;
         if (c->ps_partkey) c->ps_partkey[i] = (int32_t) (r / 4 + 1);
         if (c->ps_suppkey) c->ps_suppkey[i] = ldbdbgen::partSuppSuppKey(s, r);
         if (c->ps_supplycost) storeDec(c->ps_supplycost, i, ldbdbgen::partSuppSupplyCost(r));
      }
   });
}

} // extern "C"
