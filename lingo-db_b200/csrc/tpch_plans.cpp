// tpch_plans.cpp — query drivers (include/ldb_tpch.h): the pipeline sequence of the reference's
// JIT'd main() for TPC-H Q1/Q3/Q5/Q6, expressed as ldb_gpu_run_pipeline descriptors.  Host C++,
// like src/execution; every data-parallel step runs in a CUDA kernel, the tiny result projections
// (avg division, ORDER BY over <= 25 rows) stay on the host like the reference's result side
// ("BOUNDARY (tiny outputs; keep on host)").
#include "../../include/ldb_tpch.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace {
struct PlanError : std::runtime_error {
   LdbError e;
   explicit PlanError(const LdbError& e) : std::runtime_error(e.message), e(e) {}
};
inline void check(int rc, const LdbError& e) {
   if (rc != LDB_OK) throw PlanError(e);
}
template <class Fn>
int guarded(LdbError* err, const Fn& fn) {
   try {
      fn();
      if (err) {
         err->code = LDB_OK;
         err->message[0] = 0;
      }
      return LDB_OK;
   } catch (const PlanError& p) {
      if (err) *err = p.e;
      return p.e.code;
   } catch (const std::exception& ex) {
      if (err) {
         err->code = LDB_ERR_INVALID;
         snprintf(err->message, sizeof(err->message), "%s", ex.what());
      }
      return LDB_ERR_INVALID;
   }
}
LdbFilterDesc strFilter(const char* col, int op, const char* v) {
   LdbFilterDesc f{};
   f.column = col;
   f.op = op;
   f.str_value = v;
   return f;
}
LdbFilterDesc intFilter(const char* col, int op, int64_t v) {
   LdbFilterDesc f{};
   f.column = col;
   f.op = op;
   f.value_is_int = 1;
   f.int_value = v;
   return f;
}
LdbAggDesc agg(int expr, const char* a = nullptr, const char* b = nullptr, const char* c = nullptr) { return LdbAggDesc{expr, {a, b, c}}; }

struct StateGuard { // the query's ExecutionContext: frees every state it registered
   std::vector<LdbState*> states;
   ~StateGuard() {
      for (auto* s : states) ldb_gpu_state_destroy(s);
   }
   LdbState* own(LdbState* s) {
      states.push_back(s);
      return s;
   }
};
using i128 = __int128;
i128 toI128(const LdbI128& v) { return (i128) (((unsigned __int128) (uint64_t) v.hi << 64) | v.lo); }
LdbI128 fromI128(i128 v) { return LdbI128{(uint64_t) v, (int64_t) ((unsigned __int128) v >> 64)}; }
// avg(x decimal(12,2)) = (sum * 10^19) sdiv count  → decimal(31,21)
// (SimplifyAggregations.cpp:160-181; DecimalDiv lowering LowerToStd.cpp:651-700)
LdbI128 avgDec(int64_t sum, int64_t count) {
   i128 p = 1;
   for (int i = 0; i < 19; i++) p *= 10;
   return fromI128((i128) ((unsigned __int128) (i128) sum * (unsigned __int128) p) / (i128) count);
}

// Build a join table with the pipeline `d`, growing it if the cardinality estimate was too low
// (the reference sizes after materialisation, LazyJoinHashtable.cpp:16; a GPU build must pre-size).
LdbState* buildJoin(LdbContext* ctx, StateGuard& g, LdbPipelineDesc d, int64_t estimate, int unique, int nSide, int nAggs, bool pairKey = false) {
   LdbError e;
   for (int attempt = 0; attempt < 6; attempt++) {
      LdbState* s = nullptr;
      if (pairKey) check(ldb_gpu_join_table_create_pair(ctx, estimate, unique, &s, &e), e);
      else check(ldb_gpu_join_table_create(ctx, estimate, unique, nSide, nAggs, &s, &e), e);
      d.sink = s;
      int rc = ldb_gpu_run_pipeline(ctx, &d, &e);
      int64_t n = 0;
      if (rc == LDB_OK) rc = ldb_gpu_join_table_count(s, &n, &e);
      if (rc == LDB_OK) return g.own(s);
      ldb_gpu_state_destroy(s);
      if (rc != LDB_ERR_CAPACITY) throw PlanError(e);
      estimate *= 4;
   }
   throw PlanError(e);
}
// Foreign-key build side {key → payload} over an unfiltered primary-key column: a direct-address table when the keys are
// dense ((max - min + 1) <= 8 x rows: TPC-H surrogate keys are; o_orderkey uses 8 of every 32 values), else a hash table
// without Bloom filter (FK probes always hit).  LDB_DIRECT_TABLES=0 forces the hash table (tests cover both).
LdbState* buildForeignKeyTable(LdbContext* ctx, StateGuard& g, LdbPipelineDesc d, LdbTable* table, const char* keyColumn) {
   LdbError e;
   const int64_t n = ldb_gpu_table_num_rows(table);
   const char* env = getenv("LDB_DIRECT_TABLES");
   if (n > 0 && !(env && env[0] == '0')) {
      int32_t lo = 0, hi = -1;
      check(ldb_gpu_table_column_range(table, keyColumn, &lo, &hi, &e), e);
      if (hi >= lo && (int64_t) hi - lo + 1 <= 8 * n) {
         LdbState* s = nullptr;
         check(ldb_gpu_join_table_create_direct(ctx, lo, hi, &s, &e), e);
         g.own(s);
         d.sink = s;
         check(ldb_gpu_run_pipeline(ctx, &d, &e), e);
         int64_t cnt = 0;
         check(ldb_gpu_join_table_count(s, &cnt, &e), e); // surfaces duplicate / out-of-range keys
         return s;
      }
   }
   return buildJoin(ctx, g, d, n + 1024, LDB_JOIN_UNIQUE | LDB_JOIN_NO_BLOOM, 0, 0);
}
} // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------ Q6
int ldb_tpch_q6_partial(LdbContext* ctx, const LdbTpchTables* t, const char* dateGe, const char* dateLt, const char* discGe, const char* discLe, int64_t qtyLt, LdbState** state, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      LdbState* s = nullptr;
      check(ldb_gpu_simple_state_create(ctx, 1, &s, &e), e);
      LdbFilterDesc f[5] = {strFilter("l_shipdate", LDB_GTE, dateGe), strFilter("l_shipdate", LDB_LT, dateLt), strFilter("l_discount", LDB_GTE, discGe),
                            strFilter("l_discount", LDB_LTE, discLe), intFilter("l_quantity", LDB_LT, qtyLt)};
      LdbPipelineDesc d{};
      d.kind = LDB_PIPE_SCAN_REDUCE;
      d.source = t->lineitem;
      d.n_filters = 5;
      d.filters = f;
      d.n_aggs = 1;
      d.aggs[0] = agg(LDB_EXPR_MUL, "l_extendedprice", "l_discount");
      d.sink = s;
      int rc = ldb_gpu_run_pipeline(ctx, &d, &e);
      if (rc != LDB_OK) {
         ldb_gpu_state_destroy(s);
         throw PlanError(e);
      }
      *state = s;
   });
}
int ldb_tpch_q6(LdbContext* ctx, const LdbTpchTables* t, const char* dateGe, const char* dateLt, const char* discGe, const char* discLe, int64_t qtyLt, LdbI128* revenue, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      StateGuard g;
      LdbState* s = nullptr;
      check(ldb_tpch_q6_partial(ctx, t, dateGe, dateLt, discGe, discLe, qtyLt, &s, &e), e);
      g.own(s);
      check(ldb_gpu_simple_state_read(s, revenue, &e), e);
   });
}

// ------------------------------------------------------------------------------------------------ Q1
int ldb_tpch_q1_partial(LdbContext* ctx, const LdbTpchTables* t, const char* dateLe, LdbState** state, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      LdbState* s = nullptr;
      check(ldb_gpu_groupby_create(ctx, 2, 6, 64, &s, &e), e);
      LdbFilterDesc f[1] = {strFilter("l_shipdate", LDB_LTE, dateLe)};
      LdbPipelineDesc d{};
      d.kind = LDB_PIPE_SCAN_GROUPBY;
      d.source = t->lineitem;
      d.n_filters = 1;
      d.filters = f;
      d.n_keys = 2;
      d.key_columns[0] = "l_returnflag";
      d.key_columns[1] = "l_linestatus";
      // avg(x) is sum(x)/count (SimplifyAggregations.cpp:160-181); equal sums are shared
      d.n_aggs = 6;
      d.aggs[0] = agg(LDB_EXPR_COL, "l_quantity");
      d.aggs[1] = agg(LDB_EXPR_COL, "l_extendedprice");
      d.aggs[2] = agg(LDB_EXPR_MUL_1MINUS, "l_extendedprice", "l_discount");
      d.aggs[3] = agg(LDB_EXPR_MUL_1MINUS_1PLUS, "l_extendedprice", "l_discount", "l_tax");
      d.aggs[4] = agg(LDB_EXPR_COL, "l_discount");
      d.aggs[5] = agg(LDB_EXPR_ONE);
      d.sink = s;
      int rc = ldb_gpu_run_pipeline(ctx, &d, &e);
      if (rc != LDB_OK) {
         ldb_gpu_state_destroy(s);
         throw PlanError(e);
      }
      *state = s;
   });
}
int ldb_tpch_q1_finish(LdbState* state, LdbQ1Row* rows, int32_t maxRows, int32_t* nRows, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      std::vector<LdbGroupRow> g(64);
      int32_t n = 0;
      check(ldb_gpu_groupby_read(state, g.data(), 64, &n, &e), e);
      std::vector<LdbQ1Row> out;
      for (int i = 0; i < n && i < 64; i++) {
         LdbQ1Row r{};
         r.l_returnflag = g[i].keys[0];
         r.l_linestatus = g[i].keys[1];
         r.sum_qty = (int64_t) g[i].aggs[0].lo; // sum(decimal(12,2)) wraps at 64 bits like the reference's i64 accumulator
         r.sum_base_price = (int64_t) g[i].aggs[1].lo;
         r.sum_disc_price = g[i].aggs[2];
         r.sum_charge = g[i].aggs[3];
         int64_t sumDisc = (int64_t) g[i].aggs[4].lo;
         r.count_order = (int64_t) g[i].aggs[5].lo;
         r.avg_qty = avgDec(r.sum_qty, r.count_order);
         r.avg_price = avgDec(r.sum_base_price, r.count_order);
         r.avg_disc = avgDec(sumDisc, r.count_order);
         out.push_back(r);
      }
      std::sort(out.begin(), out.end(), [](const LdbQ1Row& a, const LdbQ1Row& b) { return a.l_returnflag != b.l_returnflag ? a.l_returnflag < b.l_returnflag : a.l_linestatus < b.l_linestatus; });
      *nRows = (int32_t) out.size();
      for (int i = 0; i < (int) out.size() && i < maxRows; i++) rows[i] = out[i];
   });
}
int ldb_tpch_q1(LdbContext* ctx, const LdbTpchTables* t, const char* dateLe, LdbQ1Row* rows, int32_t maxRows, int32_t* nRows, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      StateGuard g;
      LdbState* s = nullptr;
      check(ldb_tpch_q1_partial(ctx, t, dateLe, &s, &e), e);
      g.own(s);
      check(ldb_tpch_q1_finish(s, rows, maxRows, nRows, &e), e);
   });
}

// ------------------------------------------------------------------------------------------------ Q3
int ldb_tpch_q3(LdbContext* ctx, const LdbTpchTables* t, const char* segment, const char* date, LdbQ3Row* rows, int32_t* nRows, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      StateGuard g;
      int64_t nCust = ldb_gpu_table_num_rows(t->customer), nOrd = ldb_gpu_table_num_rows(t->orders);
      // P1: customer(c_mktsegment = X) → key set on c_custkey
      LdbFilterDesc fc[1] = {strFilter("c_mktsegment", LDB_EQ, segment)};
      LdbPipelineDesc d1{};
      d1.kind = LDB_PIPE_SCAN_BUILD;
      d1.source = t->customer;
      d1.n_filters = 1;
      d1.filters = fc;
      d1.build_key_column = "c_custkey";
      LdbState* cust = buildJoin(ctx, g, d1, nCust / 4 + 1024, 1, 0, 0);
      // P2: orders(o_orderdate < D) ⋈ customer → group-join map keyed by o_orderkey {o_orderdate, o_shippriority | revenue}
      LdbFilterDesc fo[1] = {strFilter("o_orderdate", LDB_LT, date)};
      LdbPipelineDesc d2{};
      d2.kind = LDB_PIPE_SCAN_BUILD;
      d2.source = t->orders;
      d2.n_filters = 1;
      d2.filters = fo;
      d2.n_probes = 1;
      d2.probe_states[0] = cust;
      d2.probe_key_columns[0] = "o_custkey";
      d2.build_key_column = "o_orderkey";
      d2.n_side = 2;
      d2.side_columns[0] = "o_orderdate";
      d2.side_columns[1] = "o_shippriority";
      LdbState* map = buildJoin(ctx, g, d2, nOrd / 10 + 1024, 1, 2, 1);
      // P3: lineitem(l_shipdate > D) pure lookup + SUM(l_extendedprice * (1 - l_discount)) into the entry
      LdbFilterDesc fl[1] = {strFilter("l_shipdate", LDB_GT, date)};
      LdbPipelineDesc d3{};
      d3.kind = LDB_PIPE_SCAN_PROBE_AGG;
      d3.source = t->lineitem;
      d3.n_filters = 1;
      d3.filters = fl;
      d3.n_probes = 1;
      d3.probe_states[0] = map;
      d3.probe_key_columns[0] = "l_orderkey";
      d3.n_aggs = 1;
      d3.aggs[0] = agg(LDB_EXPR_MUL_1MINUS, "l_extendedprice", "l_discount");
      d3.sink = map;
      check(ldb_gpu_run_pipeline(ctx, &d3, &e), e);
      // P4: marked groups, ORDER BY revenue desc, o_orderdate LIMIT 10
      LdbTopKRow top[10];
      int32_t n = 0;
      check(ldb_gpu_join_table_topk(map, 10, top, &n, &e), e);
      for (int i = 0; i < n; i++) rows[i] = LdbQ3Row{top[i].key, top[i].side[0], top[i].side[1], 0, top[i].agg};
      *nRows = n;
   });
}

// ------------------------------------------------------------------------------------------------ Q5
int ldb_tpch_q5(LdbContext* ctx, const LdbTpchTables* t, const char* regionName, const char* dateGe, const char* dateLt, LdbQ5Row* rows, int32_t* nRows, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      StateGuard g;
      int64_t nCust = ldb_gpu_table_num_rows(t->customer), nOrd = ldb_gpu_table_num_rows(t->orders), nSupp = ldb_gpu_table_num_rows(t->supplier);
      // region(r_name = X) → {r_regionkey}
      LdbFilterDesc fr[1] = {strFilter("r_name", LDB_EQ, regionName)};
      LdbPipelineDesc dr{};
      dr.kind = LDB_PIPE_SCAN_BUILD;
      dr.source = t->region;
      dr.n_filters = 1;
      dr.filters = fr;
      dr.build_key_column = "r_regionkey";
      LdbState* region = buildJoin(ctx, g, dr, 16, 1, 0, 0);
      // nation ⋈ region → {n_nationkey → n_nationkey}
      LdbPipelineDesc dn{};
      dn.kind = LDB_PIPE_SCAN_BUILD;
      dn.source = t->nation;
      dn.n_probes = 1;
      dn.probe_states[0] = region;
      dn.probe_key_columns[0] = "n_regionkey";
      dn.build_key_column = "n_nationkey";
      dn.build_payload_column = "n_nationkey";
      LdbState* nation = buildJoin(ctx, g, dn, 64, 1, 0, 0);
      // customer ⋈ nation → {c_custkey → c_nationkey}
      LdbPipelineDesc dc{};
      dc.kind = LDB_PIPE_SCAN_BUILD;
      dc.source = t->customer;
      dc.n_probes = 1;
      dc.probe_states[0] = nation;
      dc.probe_key_columns[0] = "c_nationkey";
      dc.build_key_column = "c_custkey";
      dc.build_payload_column = "c_nationkey";
      LdbState* cust = buildJoin(ctx, g, dc, nCust / 4 + 1024, 1, 0, 0);
      // orders(date range) ⋈ customer → {o_orderkey → c_nationkey}
      LdbFilterDesc fo[2] = {strFilter("o_orderdate", LDB_GTE, dateGe), strFilter("o_orderdate", LDB_LT, dateLt)};
      LdbPipelineDesc dor{};
      dor.kind = LDB_PIPE_SCAN_BUILD;
      dor.source = t->orders;
      dor.n_filters = 2;
      dor.filters = fo;
      dor.n_probes = 1;
      dor.probe_states[0] = cust;
      dor.probe_key_columns[0] = "o_custkey";
      dor.build_key_column = "o_orderkey";
      LdbState* ord = buildJoin(ctx, g, dor, nOrd / 24 + 1024, 1, 0, 0);
      // supplier ⋈ nation → {s_suppkey → s_nationkey}
      LdbPipelineDesc ds{};
      ds.kind = LDB_PIPE_SCAN_BUILD;
      ds.source = t->supplier;
      ds.n_probes = 1;
      ds.probe_states[0] = nation;
      ds.probe_key_columns[0] = "s_nationkey";
      ds.build_key_column = "s_suppkey";
      ds.build_payload_column = "s_nationkey";
      LdbState* supp = buildJoin(ctx, g, ds, nSupp / 4 + 1024, 1, 0, 0);
      // lineitem ⋈ orders ⋈ supplier on (l_suppkey, c_nationkey) → group by nation → SUM(ext * (1 - disc))
      LdbState* groups = nullptr;
      check(ldb_gpu_groupby_create(ctx, 1, 1, 64, &groups, &e), e);
      g.own(groups);
      LdbPipelineDesc dl{};
      dl.kind = LDB_PIPE_SCAN_PROBE2_GROUPBY;
      dl.source = t->lineitem;
      dl.n_probes = 2;
      // cheapest filter first: the supplier table (and its Bloom filter) is ~20x smaller than the orders table
      dl.probe_states[0] = supp;
      dl.probe_key_columns[0] = "l_suppkey";
      dl.probe_states[1] = ord;
      dl.probe_key_columns[1] = "l_orderkey";
      dl.n_aggs = 1;
      dl.aggs[0] = agg(LDB_EXPR_MUL_1MINUS, "l_extendedprice", "l_discount");
      dl.sink = groups;
      check(ldb_gpu_run_pipeline(ctx, &dl, &e), e);
      std::vector<LdbGroupRow> gr(64);
      int32_t n = 0;
      check(ldb_gpu_groupby_read(groups, gr.data(), 64, &n, &e), e);
      std::vector<LdbQ5Row> out;
      for (int i = 0; i < n; i++) out.push_back(LdbQ5Row{gr[i].keys[0], 0, gr[i].aggs[0]});
      std::sort(out.begin(), out.end(), [](const LdbQ5Row& a, const LdbQ5Row& b) {
         i128 x = toI128(a.revenue), y = toI128(b.revenue);
         return x != y ? x > y : a.n_nationkey < b.n_nationkey;
      });
      for (int i = 0; i < (int) out.size() && i < 25; i++) rows[i] = out[i];
      *nRows = (int32_t) std::min<size_t>(out.size(), 25);
   });
}

// ------------------------------------------------------------------------------------------------ Q5, repartitioned across GPUs
namespace {
struct Q5Layout {
   int64_t capOrd, capLi, bloomBytes;
   int64_t cursorsA, cursorsB, countsA, countsB, ordRecv, liRecv, bloom, total;
   int64_t expectedOrders;
};
Q5Layout q5Layout(int64_t nOrdTotal, int64_t nLiTotal, int world) {
   Q5Layout l{};
   const int64_t w2 = (int64_t) world * world;
   // qualifying orders ≈ 3 % of orders (one year of seven, one region of five), lineitem survivors of the Bloom semi-join ≈ 4-5 %:
   // every (source, destination) sub-region gets twice its expected share
   l.capOrd = nOrdTotal / 16 / w2 + 8192;
   l.capLi = nLiTotal / 10 / w2 + 8192;
   l.expectedOrders = nOrdTotal / 24 + 4096;
   ldb_gpu_join_table_create_shared_bloom(nullptr, l.expectedOrders, LDB_JOIN_UNIQUE, nullptr, 0, &l.bloomBytes, nullptr, nullptr);
   int64_t off = 0;
   auto take = [&](int64_t bytes) {
      int64_t at = off;
      off = (off + bytes + 255) & ~int64_t(255);
      return at;
   };
   l.cursorsA = take(16 * 8);
   l.cursorsB = take(16 * 8);
   l.countsA = take(8 * 8);
   l.countsB = take(8 * 8);
   l.ordRecv = take((int64_t) world * l.capOrd * 8);
   l.liRecv = take((int64_t) world * l.capLi * 24);
   l.bloom = take(l.bloomBytes);
   l.total = off;
   return l;
}
} // namespace
int64_t ldb_tpch_q5_repartitioned_heap_bytes(int64_t n_orders_total, int64_t n_lineitem_total, int32_t world) { return q5Layout(n_orders_total, n_lineitem_total, world).total; }
int ldb_tpch_q5_repartitioned(LdbContext* ctx, const LdbTpchTables* t, LdbComm* comm, const char* regionName, const char* dateGe, const char* dateLt, int64_t nOrdTotal, int64_t nLiTotal,
                              LdbQ5Row* rows, int32_t* nRows, LdbQ5ShuffleStats* stats, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      StateGuard g;
      if (!comm) throw std::runtime_error("repartitioned Q5 needs a comm");
      const int world = ldb_gpu_comm_world(comm);
      const Q5Layout L = q5Layout(nOrdTotal, nLiTotal, world);
      int64_t heapBytes = 0;
      ldb_gpu_comm_heap(comm, &heapBytes);
      if (heapBytes < L.total) {
         LdbError he{LDB_ERR_CAPACITY, ""};
         snprintf(he.message, sizeof(he.message), "comm user heap holds %lld bytes, repartitioned Q5 needs %lld (ldb_tpch_q5_repartitioned_heap_bytes)", (long long) heapBytes, (long long) L.total);
         throw PlanError(he);
      }
      int64_t nCust = ldb_gpu_table_num_rows(t->customer), nSupp = ldb_gpu_table_num_rows(t->supplier);
      check(ldb_gpu_comm_heap_zero(comm, 0, L.ordRecv, &e), e); // cursors + counts
      // replicated build sides (as in ldb_tpch_q5)
      LdbFilterDesc fr[1] = {strFilter("r_name", LDB_EQ, regionName)};
      LdbPipelineDesc dr{};
      dr.kind = LDB_PIPE_SCAN_BUILD;
      dr.source = t->region;
      dr.n_filters = 1;
      dr.filters = fr;
      dr.build_key_column = "r_regionkey";
      LdbState* region = buildJoin(ctx, g, dr, 16, 1, 0, 0);
      LdbPipelineDesc dn{};
      dn.kind = LDB_PIPE_SCAN_BUILD;
      dn.source = t->nation;
      dn.n_probes = 1;
      dn.probe_states[0] = region;
      dn.probe_key_columns[0] = "n_regionkey";
      dn.build_key_column = "n_nationkey";
      dn.build_payload_column = "n_nationkey";
      LdbState* nation = buildJoin(ctx, g, dn, 64, 1, 0, 0);
      LdbPipelineDesc dc{};
      dc.kind = LDB_PIPE_SCAN_BUILD;
      dc.source = t->customer;
      dc.n_probes = 1;
      dc.probe_states[0] = nation;
      dc.probe_key_columns[0] = "c_nationkey";
      dc.build_key_column = "c_custkey";
      dc.build_payload_column = "c_nationkey";
      LdbState* cust = buildJoin(ctx, g, dc, nCust / 4 + 1024, 1, 0, 0);
      LdbPipelineDesc ds{};
      ds.kind = LDB_PIPE_SCAN_BUILD;
      ds.source = t->supplier;
      ds.n_probes = 1;
      ds.probe_states[0] = nation;
      ds.probe_key_columns[0] = "s_nationkey";
      ds.build_key_column = "s_suppkey";
      ds.build_payload_column = "s_nationkey";
      LdbState* supp = buildJoin(ctx, g, ds, nSupp / 4 + 1024, 1, 0, 0);
      // this rank's hash partition of orders ⋈ customer, sized for the GLOBAL key set so that every rank's Bloom filter has the same
      // geometry (their OR keeps the false-positive rate of a single-GPU build)
      LdbState* ordp = nullptr;
      check(ldb_gpu_join_table_create_shared_bloom(ctx, L.expectedOrders, LDB_JOIN_UNIQUE, comm, L.bloom, nullptr, &ordp, &e), e);
      g.own(ordp);
      check(ldb_gpu_comm_barrier(comm, &e), e); // every rank cleared its cursors, counts and filter
      // ---- orders shard → {o_orderkey, c_nationkey} tuples → owner rank (K10: partition fused with the NVLink store)
      LdbFilterDesc fo[2] = {strFilter("o_orderdate", LDB_GTE, dateGe), strFilter("o_orderdate", LDB_LT, dateLt)};
      LdbPipelineDesc so{};
      so.kind = LDB_PIPE_SCAN_PARTITION_SEND;
      so.source = t->orders;
      so.n_filters = 2;
      so.filters = fo;
      so.n_probes = 1;
      so.probe_states[0] = cust;
      so.probe_key_columns[0] = "o_custkey";
      so.n_out_cols = 2;
      so.out_columns[0] = "o_orderkey";
      so.out_columns[1] = "$payload";
      so.comm = comm;
      so.send_offset = L.ordRecv;
      so.send_capacity = L.capOrd;
      so.send_cursors_offset = L.cursorsA;
      check(ldb_gpu_run_pipeline(ctx, &so, &e), e);
      check(ldb_gpu_comm_publish_counts(comm, L.cursorsA, L.countsA, &e), e);
      check(ldb_gpu_comm_barrier(comm, &e), e); // all tuples and counts landed
      check(ldb_gpu_join_table_insert_received(ordp, comm, L.ordRecv, L.capOrd, L.countsA, &e), e);
      check(ldb_gpu_comm_barrier(comm, &e), e); // every partition built
      check(ldb_gpu_comm_or_reduce(comm, L.bloom, L.bloomBytes, &e), e);
      check(ldb_gpu_comm_barrier(comm, &e), e); // nobody still reads a filter that is about to be probed / reused
      // ---- lineitem shard → Bloom semi-join → {l_orderkey, l_suppkey, l_extendedprice, l_discount} tuples → owner rank
      LdbPipelineDesc sl{};
      sl.kind = LDB_PIPE_SCAN_PARTITION_SEND;
      sl.source = t->lineitem;
      sl.n_probes = 1;
      sl.probe_states[0] = ordp;
      sl.probe_key_columns[0] = "l_orderkey";
      sl.probe_bloom_only = 1;
      sl.n_out_cols = 4;
      sl.out_columns[0] = "l_orderkey";
      sl.out_columns[1] = "l_suppkey";
      sl.out_columns[2] = "l_extendedprice";
      sl.out_columns[3] = "l_discount";
      sl.comm = comm;
      sl.send_offset = L.liRecv;
      sl.send_capacity = L.capLi;
      sl.send_cursors_offset = L.cursorsB;
      check(ldb_gpu_run_pipeline(ctx, &sl, &e), e);
      check(ldb_gpu_comm_publish_counts(comm, L.cursorsB, L.countsB, &e), e);
      check(ldb_gpu_comm_barrier(comm, &e), e);
      // ---- probes + aggregation where the partition lives, then the all-merge of the 5-group tables
      LdbState* groups = nullptr;
      check(ldb_gpu_groupby_create(ctx, 1, 1, 64, &groups, &e), e);
      g.own(groups);
      check(ldb_gpu_probe_received_groupby(ordp, supp, groups, comm, L.liRecv, L.capLi, L.countsB, 2, &e), e);
      check(ldb_gpu_groupby_allmerge(groups, comm, &e), e);
      // ---- the only host synchronisation of the data path: result + bookkeeping read-back
      std::vector<LdbGroupRow> gr(64);
      int32_t n = 0;
      check(ldb_gpu_groupby_read(groups, gr.data(), 64, &n, &e), e);
      uint64_t hdr[48]; // cursorsA[16] | cursorsB[16] | countsA[8] | countsB[8]
      check(ldb_gpu_comm_heap_read(comm, L.cursorsA, 16 * 8, hdr, &e), e);
      check(ldb_gpu_comm_heap_read(comm, L.cursorsB, 16 * 8, hdr + 16, &e), e);
      check(ldb_gpu_comm_heap_read(comm, L.countsA, 8 * 8, hdr + 32, &e), e);
      check(ldb_gpu_comm_heap_read(comm, L.countsB, 8 * 8, hdr + 40, &e), e);
      check(ldb_gpu_comm_check(comm, &e), e);
      if ((uint32_t) hdr[8] || (uint32_t) hdr[16 + 8]) {
         LdbError ce{LDB_ERR_CAPACITY, "a repartition receive sub-region overflowed (skewed partitions): raise the comm heap / capacities"};
         throw PlanError(ce);
      }
      int64_t cnt = 0;
      check(ldb_gpu_join_table_count(ordp, &cnt, &e), e); // surfaces table-full / duplicate-key errors of the partition build
      if (stats) {
         *stats = LdbQ5ShuffleStats{};
         for (int r = 0; r < world; r++) {
            stats->orders_tuples_sent += (int64_t) hdr[r];
            stats->lineitem_tuples_sent += (int64_t) hdr[16 + r];
            stats->orders_tuples_received += (int64_t) hdr[32 + r];
            stats->lineitem_tuples_received += (int64_t) hdr[40 + r];
         }
         stats->shuffle_bytes_out = stats->orders_tuples_sent * 8 + stats->lineitem_tuples_sent * 24;
         stats->heap_bytes = L.total;
      }
      std::vector<LdbQ5Row> out;
      for (int i = 0; i < n; i++) out.push_back(LdbQ5Row{gr[i].keys[0], 0, gr[i].aggs[0]});
      std::sort(out.begin(), out.end(), [](const LdbQ5Row& a, const LdbQ5Row& b) {
         i128 x = toI128(a.revenue), y = toI128(b.revenue);
         return x != y ? x > y : a.n_nationkey < b.n_nationkey;
      });
      for (int i = 0; i < (int) out.size() && i < 25; i++) rows[i] = out[i];
      *nRows = (int32_t) std::min<size_t>(out.size(), 25);
   });
}

// ------------------------------------------------------------------------------------------------ Q9
// part(p_name like '%X%') → partsupp ⋈ part keyed (ps_partkey, ps_suppkey) → supplier, orders(→ year) → lineitem star probe.
// p_partkey = l_partkey follows from ps_partkey = l_partkey and p_partkey = ps_partkey (the partsupp table only holds
// parts that passed the LIKE), so the lineitem pipeline probes three tables, not four.
// _partial runs every pipeline over the tables it is given and returns the (nation, year) group state; with lineitem and
// orders sharded by the same order range (the join is co-partitioned) and the small sides replicated, each GPU of a
// multi-GPU run calls it on its shard, the group images are all-gathered + merged (K7), then _finish materialises.
int ldb_tpch_q9_partial(LdbContext* ctx, const LdbTpchTables* t, const char* nameContains, LdbState** state, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      StateGuard g;
      if (!t->part || !t->partsupp) throw std::runtime_error("Q9 needs the part and partsupp tables");
      int64_t nPart = ldb_gpu_table_num_rows(t->part), nPs = ldb_gpu_table_num_rows(t->partsupp);
      LdbFilterDesc fp[1] = {strFilter("p_name", LDB_CONTAINS, nameContains)};
      LdbPipelineDesc dp{};
      dp.kind = LDB_PIPE_SCAN_BUILD;
      dp.source = t->part;
      dp.n_filters = 1;
      dp.filters = fp;
      dp.build_key_column = "p_partkey";
      LdbState* part = buildJoin(ctx, g, dp, nPart / 12 + 1024, 1, 0, 0);
      LdbPipelineDesc dps{};
      dps.kind = LDB_PIPE_SCAN_BUILD;
      dps.source = t->partsupp;
      dps.n_probes = 1;
      dps.probe_states[0] = part;
      dps.probe_key_columns[0] = "ps_partkey";
      dps.build_key_column = "ps_partkey";
      dps.build_key2_column = "ps_suppkey";
      dps.build_payload_column = "ps_supplycost";
      // (ps_partkey, ps_suppkey) is partsupp's primary key: build a unique table (a probe stops at its first match); if the
      // data disagrees (the tiny-scale generator repeats suppliers) the build reports the duplicate and a multimap is built
      LdbState* ps = nullptr;
      try {
         ps = buildJoin(ctx, g, dps, nPs / 12 + 1024, LDB_JOIN_UNIQUE, 0, 0, true);
      } catch (const PlanError& pe) {
         if (pe.e.code != LDB_ERR_INVALID) throw;
         ps = buildJoin(ctx, g, dps, nPs / 12 + 1024, 0, 0, 0, true);
      }
      LdbPipelineDesc ds{};
      ds.kind = LDB_PIPE_SCAN_BUILD;
      ds.source = t->supplier;
      ds.build_key_column = "s_suppkey";
      ds.build_payload_column = "s_nationkey";
      LdbState* supp = buildForeignKeyTable(ctx, g, ds, t->supplier, "s_suppkey");
      LdbPipelineDesc dor{};
      dor.kind = LDB_PIPE_SCAN_BUILD;
      dor.source = t->orders;
      dor.build_key_column = "o_orderkey";
      dor.build_payload_column = "o_orderdate";
      dor.build_payload_expr = LDB_PAYLOAD_YEAR;
      LdbState* ord = buildForeignKeyTable(ctx, g, dor, t->orders, "o_orderkey");
      LdbState* groups = nullptr;
      check(ldb_gpu_groupby_create(ctx, 2, 1, 1024, &groups, &e), e);
      LdbPipelineDesc dl{};
      dl.kind = LDB_PIPE_SCAN_STAR_PROBE_GROUPBY;
      dl.source = t->lineitem;
      dl.n_probes = 3;
      dl.probe_states[0] = ps;
      dl.probe_key_columns[0] = "l_partkey";
      dl.probe_key2_columns[0] = "l_suppkey";
      dl.probe_states[1] = supp;
      dl.probe_key_columns[1] = "l_suppkey";
      dl.probe_states[2] = ord;
      dl.probe_key_columns[2] = "l_orderkey";
      dl.n_aggs = 1;
      dl.aggs[0] = agg(LDB_EXPR_MUL_1MINUS_MINUS_PAYMUL, "l_extendedprice", "l_discount", "l_quantity");
      dl.sink = groups;
      int rc = ldb_gpu_run_pipeline(ctx, &dl, &e);
      if (rc != LDB_OK) {
         ldb_gpu_state_destroy(groups);
         throw PlanError(e);
      }
      *state = groups; // the join tables die here (StateGuard); the caller owns the group state
   });
}
// ------------------------------------------------------------------------------------------------ Q9, orders hash-partitioned across GPUs
namespace {
struct Q9Layout {
   int64_t capOrd, capLi, cursorsA, cursorsB, countsA, countsB, ordRecv, liRecv, total;
};
Q9Layout q9Layout(int64_t nOrdTotal, int64_t nLiTotal, int world) {
   Q9Layout l{};
   const int64_t w2 = (int64_t) world * world;
   l.capOrd = nOrdTotal / w2 + nOrdTotal / w2 / 4 + 8192; // every order is shipped: expected share + 25 %
   l.capLi = nLiTotal / 10 / w2 + 8192;                   // ~5.4 % of lineitem survives the partsupp probe: twice the expected share
   int64_t off = 0;
   auto take = [&](int64_t bytes) {
      int64_t at = off;
      off = (off + bytes + 255) & ~int64_t(255);
      return at;
   };
   l.cursorsA = take(16 * 8);
   l.cursorsB = take(16 * 8);
   l.countsA = take(8 * 8);
   l.countsB = take(8 * 8);
   l.ordRecv = take((int64_t) world * l.capOrd * 8);
   l.liRecv = take((int64_t) world * l.capLi * 24);
   l.total = off;
   return l;
}
} // namespace
int64_t ldb_tpch_q9_repartitioned_heap_bytes(int64_t n_orders_total, int64_t n_lineitem_total, int32_t world) { return q9Layout(n_orders_total, n_lineitem_total, world).total; }
int ldb_tpch_q9_repartitioned(LdbContext* ctx, const LdbTpchTables* t, LdbComm* comm, const char* nameContains, int64_t nOrdTotal, int64_t nLiTotal, LdbQ9Row* rows, int32_t maxRows,
                              int32_t* nRows, LdbQ5ShuffleStats* stats, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      StateGuard g;
      if (!comm) throw std::runtime_error("repartitioned Q9 needs a comm");
      if (!t->part || !t->partsupp) throw std::runtime_error("Q9 needs the part and partsupp tables");
      const int world = ldb_gpu_comm_world(comm);
      const Q9Layout L = q9Layout(nOrdTotal, nLiTotal, world);
      int64_t heapBytes = 0;
      ldb_gpu_comm_heap(comm, &heapBytes);
      if (heapBytes < L.total) {
         LdbError he{LDB_ERR_CAPACITY, ""};
         snprintf(he.message, sizeof(he.message), "comm user heap holds %lld bytes, repartitioned Q9 needs %lld (ldb_tpch_q9_repartitioned_heap_bytes)", (long long) heapBytes, (long long) L.total);
         throw PlanError(he);
      }
      check(ldb_gpu_comm_heap_zero(comm, 0, L.ordRecv, &e), e);
      int64_t nPart = ldb_gpu_table_num_rows(t->part), nPs = ldb_gpu_table_num_rows(t->partsupp);
      // replicated build sides, as in ldb_tpch_q9_partial
      LdbFilterDesc fp[1] = {strFilter("p_name", LDB_CONTAINS, nameContains)};
      LdbPipelineDesc dp{};
      dp.kind = LDB_PIPE_SCAN_BUILD;
      dp.source = t->part;
      dp.n_filters = 1;
      dp.filters = fp;
      dp.build_key_column = "p_partkey";
      LdbState* part = buildJoin(ctx, g, dp, nPart / 12 + 1024, 1, 0, 0);
      LdbPipelineDesc dps{};
      dps.kind = LDB_PIPE_SCAN_BUILD;
      dps.source = t->partsupp;
      dps.n_probes = 1;
      dps.probe_states[0] = part;
      dps.probe_key_columns[0] = "ps_partkey";
      dps.build_key_column = "ps_partkey";
      dps.build_key2_column = "ps_suppkey";
      dps.build_payload_column = "ps_supplycost";
      LdbState* ps = nullptr;
      try {
         ps = buildJoin(ctx, g, dps, nPs / 12 + 1024, LDB_JOIN_UNIQUE, 0, 0, true);
      } catch (const PlanError& pe) {
         if (pe.e.code != LDB_ERR_INVALID) throw;
         ps = buildJoin(ctx, g, dps, nPs / 12 + 1024, 0, 0, 0, true);
      }
      LdbPipelineDesc ds{};
      ds.kind = LDB_PIPE_SCAN_BUILD;
      ds.source = t->supplier;
      ds.build_key_column = "s_suppkey";
      ds.build_payload_column = "s_nationkey";
      LdbState* supp = buildForeignKeyTable(ctx, g, ds, t->supplier, "s_suppkey");
      // this rank's hash partition of orders: o_orderkey → year (foreign-key probes always hit: no Bloom filter)
      LdbState* ordp = nullptr;
      check(ldb_gpu_join_table_create(ctx, nOrdTotal / world + nOrdTotal / world / 4 + 4096, LDB_JOIN_UNIQUE | LDB_JOIN_NO_BLOOM, 0, 0, &ordp, &e), e);
      g.own(ordp);
      check(ldb_gpu_comm_barrier(comm, &e), e); // every rank cleared its cursors and counts
      LdbPipelineDesc so{};
      so.kind = LDB_PIPE_SCAN_PARTITION_SEND;
      so.source = t->orders;
      so.n_out_cols = 2;
      so.out_columns[0] = "o_orderkey";
      so.out_columns[1] = "o_orderdate";
      so.build_payload_expr = LDB_PAYLOAD_YEAR;
      so.comm = comm;
      so.send_offset = L.ordRecv;
      so.send_capacity = L.capOrd;
      so.send_cursors_offset = L.cursorsA;
      check(ldb_gpu_run_pipeline(ctx, &so, &e), e);
      check(ldb_gpu_comm_publish_counts(comm, L.cursorsA, L.countsA, &e), e);
      // the lineitem side does not depend on the orders partition: its send kernel runs before the barrier that publishes the orders
      LdbPipelineDesc sl{};
      sl.kind = LDB_PIPE_SCAN_STAR_PROBE_SEND;
      sl.source = t->lineitem;
      sl.n_probes = 2;
      sl.probe_states[0] = ps;
      sl.probe_key_columns[0] = "l_partkey";
      sl.probe_key2_columns[0] = "l_suppkey";
      sl.probe_states[1] = supp;
      sl.probe_key_columns[1] = "l_suppkey";
      sl.n_aggs = 1;
      sl.aggs[0] = agg(LDB_EXPR_MUL_1MINUS_MINUS_PAYMUL, "l_extendedprice", "l_discount", "l_quantity");
      sl.n_out_cols = 1;
      sl.out_columns[0] = "l_orderkey";
      sl.comm = comm;
      sl.send_offset = L.liRecv;
      sl.send_capacity = L.capLi;
      sl.send_cursors_offset = L.cursorsB;
      check(ldb_gpu_run_pipeline(ctx, &sl, &e), e);
      check(ldb_gpu_comm_publish_counts(comm, L.cursorsB, L.countsB, &e), e);
      check(ldb_gpu_comm_barrier(comm, &e), e); // all tuples and counts of both shuffles landed
      check(ldb_gpu_join_table_insert_received(ordp, comm, L.ordRecv, L.capOrd, L.countsA, &e), e);
      LdbState* groups = nullptr;
      check(ldb_gpu_groupby_create(ctx, 2, 1, 1024, &groups, &e), e);
      g.own(groups);
      check(ldb_gpu_probe_received_groupby2(ordp, groups, comm, L.liRecv, L.capLi, L.countsB, &e), e);
      check(ldb_gpu_groupby_allmerge(groups, comm, &e), e);
      check(ldb_tpch_q9_finish(groups, rows, maxRows, nRows, &e), e);
      uint64_t hdr[48];
      check(ldb_gpu_comm_heap_read(comm, L.cursorsA, 16 * 8, hdr, &e), e);
      check(ldb_gpu_comm_heap_read(comm, L.cursorsB, 16 * 8, hdr + 16, &e), e);
      check(ldb_gpu_comm_heap_read(comm, L.countsA, 8 * 8, hdr + 32, &e), e);
      check(ldb_gpu_comm_heap_read(comm, L.countsB, 8 * 8, hdr + 40, &e), e);
      check(ldb_gpu_comm_check(comm, &e), e);
      if ((uint32_t) hdr[8] || (uint32_t) hdr[16 + 8]) {
         LdbError ce{LDB_ERR_CAPACITY, "a repartition receive sub-region overflowed (skewed partitions): raise the comm heap / capacities"};
         throw PlanError(ce);
      }
      int64_t cnt = 0;
      check(ldb_gpu_join_table_count(ordp, &cnt, &e), e);
      if (stats) {
         *stats = LdbQ5ShuffleStats{};
         for (int r = 0; r < world; r++) {
            stats->orders_tuples_sent += (int64_t) hdr[r];
            stats->lineitem_tuples_sent += (int64_t) hdr[16 + r];
            stats->orders_tuples_received += (int64_t) hdr[32 + r];
            stats->lineitem_tuples_received += (int64_t) hdr[40 + r];
         }
         stats->shuffle_bytes_out = stats->orders_tuples_sent * 8 + stats->lineitem_tuples_sent * 24;
         stats->heap_bytes = L.total;
      }
   });
}

int ldb_tpch_q9_finish(LdbState* groups, LdbQ9Row* rows, int32_t maxRows, int32_t* nRows, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      std::vector<LdbGroupRow> gr(1024);
      int32_t n = 0;
      check(ldb_gpu_groupby_read(groups, gr.data(), 1024, &n, &e), e);
      std::vector<LdbQ9Row> out;
      for (int i = 0; i < n; i++) out.push_back(LdbQ9Row{gr[i].keys[0], gr[i].keys[1], gr[i].aggs[0]});
      // the host orders by n_name once it resolved the names; here: nationkey, year desc (stable and complete)
      std::sort(out.begin(), out.end(), [](const LdbQ9Row& a, const LdbQ9Row& b) { return a.n_nationkey != b.n_nationkey ? a.n_nationkey < b.n_nationkey : a.o_year > b.o_year; });
      if ((int) out.size() > maxRows) throw std::runtime_error("Q9 result has more rows than max_rows");
      for (size_t i = 0; i < out.size(); i++) rows[i] = out[i];
      *nRows = (int32_t) out.size();
   });
}
int ldb_tpch_q9(LdbContext* ctx, const LdbTpchTables* t, const char* nameContains, LdbQ9Row* rows, int32_t maxRows, int32_t* nRows, LdbError* err) {
   return guarded(err, [&] {
      LdbError e;
      StateGuard g;
      LdbState* s = nullptr;
      check(ldb_tpch_q9_partial(ctx, t, nameContains, &s, &e), e);
      g.own(s);
      check(ldb_tpch_q9_finish(s, rows, maxRows, nRows, &e), e);
   });
}

} // extern "C"
