// oracle/port/capi.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle).  extern "C" surface for ctypes
// (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).
#include "pipelines.h"
#include "rt_select.h"
#include "scan.h"
#include "scheduler.h"

#include <atomic>
#include <cstring>
#include <mutex>

using namespace oracle;

namespace {
thread_local std::string lastError;
inline void split(i128 v, int64_t* lo, int64_t* hi) {
   *lo = (int64_t) (uint64_t) v;
   *hi = (int64_t) ((u128) v >> 64);
}
template <class Fn>
int guarded(const Fn& fn) {
   try {
      fn();
      return 0;
   } catch (const std::exception& e) {
      lastError = e.what();
      return 1;
   }
}
} // namespace

extern "C" {

struct OracleQ1Row {
   int32_t returnflag, linestatus;
   int64_t sum_qty, sum_base_price;
   int64_t sum_disc_price[2], sum_charge[2], avg_qty[2], avg_price[2], avg_disc[2]; // {lo, hi}
   int64_t count;
};
struct OracleQ3Row {
   int32_t orderkey, orderdate, shippriority, pad;
   int64_t revenue[2];
};
struct OracleQ5Row {
   char name[32];
   int64_t revenue[2];
};
struct OracleCountRow { // Q4: {name, a}; Q12: {name, a, b}
   char name[32];
   int64_t a, b;
};
struct OracleQ18Row {
   char name[32];
   int32_t custkey, orderkey, orderdate, pad;
   int64_t totalprice, sum_quantity;
};
struct OracleQ9Row {
   char nation[32];
   int64_t year;
   int64_t sum_profit[2];
};

const char* oracle_last_error() { return lastError.c_str(); }
const char* oracle_runtime_kind() { return rt::runtimeKind; } // "port" | "reference"
void oracle_set_workers(int n) { sched::start((size_t) n); }
int oracle_num_workers() { return (int) sched::getNumWorkers(); }

// ---- tables
void* oracle_table_create(const char* name) {
   auto* t = new HostTable;
   t->name = name;
   return t;
}
void oracle_table_free(void* t) { delete (HostTable*) t; }
void oracle_table_add_column(void* t, const char* name, int physType, int precision, int scale) {
   ((HostTable*) t)->schema.push_back(ColumnSchema{name, (PhysType) physType, precision, scale});
}
// buffers: 3 pointers per column {validity or NULL, data/offsets, utf8 bytes or NULL}; borrowed
void oracle_table_add_chunk(void* tp, int64_t nRows, const void** buffers) {
   auto* t = (HostTable*) tp;
   HostChunk c;
   c.numRows = nRows;
   c.buffers.assign(buffers, buffers + 3 * t->schema.size());
   t->chunks.push_back(std::move(c));
   t->numRows += nRows;
}

// ---- value-level KAT hooks (values.h)
uint64_t oracle_hash_i64(int64_t v) { return hash64((uint64_t) v); }
uint64_t oracle_hash_bool(int v) {
   HashBuilder hb;
   hb.addBool(v != 0);
   return hb.total;
}
uint64_t oracle_hash_i128(int64_t lo, int64_t hi) {
   HashBuilder hb;
   hb.addI128((i128) (((u128) (uint64_t) hi << 64) | (uint64_t) lo));
   return hb.total;
}
uint64_t oracle_hash_date_days(int32_t days) { return hash64((uint64_t) dateToNs(days)); }
uint64_t oracle_hash_string(const uint8_t* p, uint32_t len) { return hashVarLen(VarLen32(p, len)); }
uint64_t oracle_hash_combine(uint64_t newPiece, uint64_t total) { return hashCombine(newPiece, total); }
uint64_t oracle_xxh64(const uint8_t* p, uint64_t len) { return xxh64(p, len); }
int32_t oracle_parse_date(const char* s) { return parseDate32(s); }
void oracle_parse_decimal(const char* s, int scale, int64_t* lo, int64_t* hi) { split(parseDecimal(s, scale), lo, hi); }
void oracle_avg_dec12_2(int64_t sum, int64_t count, int64_t* lo, int64_t* hi) { split(avgDec12_2(sum, count), lo, hi); }
void oracle_mul_i128(int64_t alo, int64_t ahi, int64_t blo, int64_t bhi, int64_t* lo, int64_t* hi) {
   i128 a = (i128) (((u128) (uint64_t) ahi << 64) | (uint64_t) alo), b = (i128) (((u128) (uint64_t) bhi << 64) | (uint64_t) blo);
   split(wrapMul(a, b), lo, hi);
}

// ---- a scan with pushed-down filters over (possibly nullable) columns: count(*) and sum(<int32 | decimal(p<19) column>) of the rows the
// reference's Restrictions::applyFilters lets through (NOTNULL filters consult the validity bitmap, Restrictions.cpp:67-162); NULL cells of
// the summed column are skipped like the JIT'd aggregate skips them.  ops: FilterOp order (EQ, NEQ, LT, LTE, GT, GTE, NOTNULL).
int oracle_scan_count_sum(void* table, int nFilters, const char* const* columns, const int* ops, const int* isInt, const char* const* strValues, const int64_t* intValues,
                          const char* sumColumn, int64_t* count, int64_t* sumLo, int64_t* sumHi) {
   return guarded([&] {
      HostTable& t = *(HostTable*) table;
      std::vector<FilterDescription> fds;
      std::vector<std::string> cols;
      for (int i = 0; i < nFilters; i++) {
         FilterDescription fd;
         fd.columnName = columns[i];
         int id = t.colIndex(columns[i]);
         if (id < 0) throw std::runtime_error("unknown filter column");
         fd.columnId = (size_t) id;
         fd.op = (FilterOp) ops[i];
         if (isInt[i]) fd.value = intValues[i];
         else fd.value = std::string(strValues[i] ? strValues[i] : "");
         fds.push_back(fd);
      }
      int sumId = sumColumn ? t.colIndex(sumColumn) : -1;
      if (sumColumn && sumId < 0) throw std::runtime_error("unknown sum column");
      if (sumColumn) cols.push_back(sumColumn);
      else cols.push_back(t.schema[0].name);
      rt::QueryContextScope scope;
      std::atomic<int64_t> n{0};
      std::mutex m;
      i128 total = 0;
      const bool dec = sumId >= 0 && t.schema[sumId].type == PhysType::DECIMAL128;
      scanTable(t, cols, fds, [&](rt::BatchView* b) {
         int64_t localN = 0;
         i128 local = 0;
         ColReader r(b, 0);
         const rt::ArrayView* av = b->arrays[0];
         const uint8_t* valid = av->nullCount ? (const uint8_t*) av->buffers[0] : nullptr;
         for (size_t i = 0; i < (size_t) b->length; i++) {
            const int64_t idx = b->selectionVector[i];
            localN++;
            if (sumId < 0) continue;
            const int64_t abs = av->offset + b->offset + idx;
            if (valid && !((valid[abs / 8] >> (abs % 8)) & 1)) continue;
            local += dec ? (i128) r.dec64(idx) : (i128) r.i32(idx);
         }
         n += localN;
         std::lock_guard<std::mutex> l(m);
         total += local;
      });
      *count = n.load();
      split(total, sumLo, sumHi);
   });
}

// ---- queries; *seconds = wall time of the pipelines only (tables resident), like the reference's
// executionTime (src/execution/LLVMBackends.cpp:856-865)
int oracle_q6(void* lineitem, const char* dateGe, const char* dateLt, const char* discGe, const char* discLe, int64_t qtyLt, int64_t* revLo, int64_t* revHi, double* seconds) {
   return guarded([&] {
      auto r = runQ6(*(HostTable*) lineitem, Q6Params{dateGe, dateLt, discGe, discLe, qtyLt});
      split(r.revenue, revLo, revHi);
      *seconds = r.seconds;
   });
}
int oracle_q1(void* lineitem, const char* dateLe, OracleQ1Row* out, int maxRows, int* nRows, double* seconds) {
   return guarded([&] {
      auto rows = runQ1(*(HostTable*) lineitem, Q1Params{dateLe}, seconds);
      *nRows = (int) rows.size();
      for (int i = 0; i < (int) rows.size() && i < maxRows; i++) {
         auto& r = rows[i];
         auto& o = out[i];
         o.returnflag = r.returnflag;
         o.linestatus = r.linestatus;
         o.sum_qty = r.sumQty;
         o.sum_base_price = r.sumBasePrice;
         split(r.sumDiscPrice, &o.sum_disc_price[0], &o.sum_disc_price[1]);
         split(r.sumCharge, &o.sum_charge[0], &o.sum_charge[1]);
         split(r.avgQty, &o.avg_qty[0], &o.avg_qty[1]);
         split(r.avgPrice, &o.avg_price[0], &o.avg_price[1]);
         split(r.avgDisc, &o.avg_disc[0], &o.avg_disc[1]);
         o.count = r.count;
      }
   });
}
int oracle_q3(void* customer, void* orders, void* lineitem, const char* segment, const char* date, OracleQ3Row* out, int maxRows, int* nRows, double* seconds) {
   return guarded([&] {
      auto rows = runQ3(*(HostTable*) customer, *(HostTable*) orders, *(HostTable*) lineitem, Q3Params{segment, date}, seconds);
      *nRows = (int) rows.size();
      for (int i = 0; i < (int) rows.size() && i < maxRows; i++) {
         out[i].orderkey = rows[i].orderkey;
         out[i].orderdate = rows[i].orderdate;
         out[i].shippriority = rows[i].shippriority;
         out[i].pad = 0;
         split(rows[i].revenue, &out[i].revenue[0], &out[i].revenue[1]);
      }
   });
}
int oracle_q5(void* customer, void* orders, void* lineitem, void* supplier, void* nation, void* region, const char* regionName, const char* dateGe, const char* dateLt, OracleQ5Row* out, int maxRows, int* nRows, double* seconds) {
   return guarded([&] {
      auto rows = runQ5(*(HostTable*) customer, *(HostTable*) orders, *(HostTable*) lineitem, *(HostTable*) supplier, *(HostTable*) nation, *(HostTable*) region, Q5Params{regionName, dateGe, dateLt}, seconds);
      *nRows = (int) rows.size();
      for (int i = 0; i < (int) rows.size() && i < maxRows; i++) {
         memset(out[i].name, 0, sizeof(out[i].name));
         strncpy(out[i].name, rows[i].name.c_str(), sizeof(out[i].name) - 1);
         split(rows[i].revenue, &out[i].revenue[0], &out[i].revenue[1]);
      }
   });
}
int oracle_q9(void* part, void* supplier, void* lineitem, void* partsupp, void* orders, void* nation, const char* needle, OracleQ9Row* out, int maxRows, int* nRows, double* seconds) {
   return guarded([&] {
      auto rows = runQ9(*(HostTable*) part, *(HostTable*) supplier, *(HostTable*) lineitem, *(HostTable*) partsupp, *(HostTable*) orders, *(HostTable*) nation, Q9Params{needle}, seconds);
      *nRows = (int) rows.size();
      for (int i = 0; i < (int) rows.size() && i < maxRows; i++) {
         memset(out[i].nation, 0, sizeof(out[i].nation));
         strncpy(out[i].nation, rows[i].nation.c_str(), sizeof(out[i].nation) - 1);
         out[i].year = rows[i].year;
         split(rows[i].sumProfit, &out[i].sum_profit[0], &out[i].sum_profit[1]);
      }
   });
}
int oracle_q4(void* orders, void* lineitem, const char* dateGe, const char* dateLt, OracleCountRow* out, int maxRows, int* nRows, double* seconds) {
   return guarded([&] {
      auto rows = runQ4(*(HostTable*) orders, *(HostTable*) lineitem, Q4Params{dateGe, dateLt}, seconds);
      *nRows = (int) rows.size();
      for (int i = 0; i < (int) rows.size() && i < maxRows; i++) {
         memset(out[i].name, 0, sizeof(out[i].name));
         strncpy(out[i].name, rows[i].priority.c_str(), sizeof(out[i].name) - 1);
         out[i].a = rows[i].orderCount;
         out[i].b = 0;
      }
   });
}
int oracle_q12(void* orders, void* lineitem, const char* mode1, const char* mode2, const char* dateGe, const char* dateLt, OracleCountRow* out, int maxRows, int* nRows, double* seconds) {
   return guarded([&] {
      auto rows = runQ12(*(HostTable*) orders, *(HostTable*) lineitem, Q12Params{mode1, mode2, dateGe, dateLt}, seconds);
      *nRows = (int) rows.size();
      for (int i = 0; i < (int) rows.size() && i < maxRows; i++) {
         memset(out[i].name, 0, sizeof(out[i].name));
         strncpy(out[i].name, rows[i].shipmode.c_str(), sizeof(out[i].name) - 1);
         out[i].a = rows[i].highLineCount;
         out[i].b = rows[i].lowLineCount;
      }
   });
}
int oracle_q18(void* customer, void* orders, void* lineitem, int64_t quantityGt, OracleQ18Row* out, int maxRows, int* nRows, double* seconds) {
   return guarded([&] {
      auto rows = runQ18(*(HostTable*) customer, *(HostTable*) orders, *(HostTable*) lineitem, quantityGt, seconds);
      *nRows = (int) rows.size();
      for (int i = 0; i < (int) rows.size() && i < maxRows; i++) {
         memset(&out[i], 0, sizeof(out[i]));
         strncpy(out[i].name, rows[i].name.c_str(), sizeof(out[i].name) - 1);
         out[i].custkey = rows[i].custkey;
         out[i].orderkey = rows[i].orderkey;
         out[i].orderdate = rows[i].orderdate;
         out[i].totalprice = rows[i].totalprice;
         out[i].sum_quantity = rows[i].sumQuantity;
      }
   });
}
int64_t oracle_extract_year(int64_t ns) { return rt::extractYear(ns); }
int oracle_const_like_contains(const char* str, uint32_t len, const char* needle) { return rt::constLikeContains(VarLen32((const uint8_t*) str, len), needle) ? 1 : 0; }

} // extern "C"
