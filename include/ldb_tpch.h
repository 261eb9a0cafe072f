/* ldb_tpch.h — query drivers: the sequence of pipeline calls the reference's JIT'd `main()` performs
 * for TPC-H Q1/Q3/Q5/Q6/Q9 (DefaultCPULLVMBackend::execute → mainFunc(), src/execution/LLVMBackends.cpp:795-867),
 * issued against the GPU C-ABI (ldb_gpu.h).  This is what a `GPUPatternList` lowering
 * (SubOpToControlFlow.cpp:4254-4394) would emit; it lives in C++ like src/execution.
 * Results are exact integers: decimal raw values with the scale the reference's typing gives them.
 */
#ifndef LDB_TPCH_H
#define LDB_TPCH_H
#include "ldb_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct LdbTpchTables {
   LdbTable* lineitem;
   LdbTable* orders;
   LdbTable* customer;
   LdbTable* supplier;
   LdbTable* nation;
   LdbTable* region;
   LdbTable* part;     /* Q9 only */
   LdbTable* partsupp; /* Q9 only */
} LdbTpchTables;

/* Q6: revenue = sum(l_extendedprice * l_discount), decimal(24,4) */
int ldb_tpch_q6(LdbContext* ctx, const LdbTpchTables* t, const char* shipdate_ge, const char* shipdate_lt, const char* discount_ge, const char* discount_le, int64_t quantity_lt,
                LdbI128* revenue, LdbError* err);
/* the per-GPU half of Q6/Q1: run the scan pipeline into a fresh state and return it (rows of other
 * GPUs are folded in with ldb_gpu_groupby_merge_rows before the finish call) */
int ldb_tpch_q6_partial(LdbContext* ctx, const LdbTpchTables* t, const char* shipdate_ge, const char* shipdate_lt, const char* discount_ge, const char* discount_le, int64_t quantity_lt,
                        LdbState** state, LdbError* err);

typedef struct LdbQ1Row {
   int32_t l_returnflag, l_linestatus;      /* fixed_size_binary(4) cells */
   int64_t sum_qty, sum_base_price;         /* decimal(12,2) */
   LdbI128 sum_disc_price;                  /* decimal(33,4) */
   LdbI128 sum_charge;                      /* decimal(38,6) */
   LdbI128 avg_qty, avg_price, avg_disc;    /* decimal(31,21): (sum * 10^19) sdiv count */
   int64_t count_order;
} LdbQ1Row;
int ldb_tpch_q1(LdbContext* ctx, const LdbTpchTables* t, const char* shipdate_le, LdbQ1Row* rows, int32_t max_rows, int32_t* n_rows, LdbError* err);
int ldb_tpch_q1_partial(LdbContext* ctx, const LdbTpchTables* t, const char* shipdate_le, LdbState** state, LdbError* err);
int ldb_tpch_q1_finish(LdbState* state, LdbQ1Row* rows, int32_t max_rows, int32_t* n_rows, LdbError* err);

typedef struct LdbQ3Row {
   int32_t l_orderkey, o_orderdate, o_shippriority, pad;
   LdbI128 revenue; /* decimal(33,4) */
} LdbQ3Row;
int ldb_tpch_q3(LdbContext* ctx, const LdbTpchTables* t, const char* segment, const char* date, LdbQ3Row* rows /* 10 */, int32_t* n_rows, LdbError* err);

typedef struct LdbQ5Row {
   int32_t n_nationkey, pad; /* n_name is resolved from the nation table at materialisation (host) */
   LdbI128 revenue;          /* decimal(33,4) */
} LdbQ5Row;
int ldb_tpch_q5(LdbContext* ctx, const LdbTpchTables* t, const char* region_name, const char* date_ge, const char* date_lt, LdbQ5Row* rows /* 25 */, int32_t* n_rows, LdbError* err);

/* Q9 (resources/sql/tpch/9.sql): group by (nation, extract(year from o_orderdate)); ordered by n_name, o_year desc on the host */
typedef struct LdbQ9Row {
   int32_t n_nationkey, o_year; /* n_name is resolved from the nation table at materialisation (host) */
   LdbI128 sum_profit;          /* decimal(34,4) raw: l_extendedprice * (1 - l_discount) - ps_supplycost * l_quantity, summed */
} LdbQ9Row;
int ldb_tpch_q9_partial(LdbContext* ctx, const LdbTpchTables* t, const char* name_contains, LdbState** group_state, LdbError* err);
int ldb_tpch_q9_finish(LdbState* group_state, LdbQ9Row* rows /* max_rows */, int32_t max_rows, int32_t* n_rows, LdbError* err);
int ldb_tpch_q9(LdbContext* ctx, const LdbTpchTables* t, const char* name_contains, LdbQ9Row* rows /* max_rows */, int32_t max_rows, int32_t* n_rows, LdbError* err);

/* Q5 with the orders ⋈ lineitem join radix-partitioned across the ranks of `comm` (BASELINE.json config 3; no reference
 * counterpart, the reference is single-process).  `t` holds THIS rank's shard of orders and lineitem (any split) and full copies
 * of customer/supplier/nation/region (small build sides are replicated).  Every data-path step is a kernel on the context's
 * stream: qualifying orders are partitioned by h64(o_orderkey) and stored straight into the owning rank's receive region over
 * NVLink (K10), each rank builds its hash partition, the partitions' Bloom filters are OR-ed through peer loads so that every
 * rank pre-filters its lineitem shard before the second partition-send, probes + the 5-group aggregation run where the
 * partition lives, and the group tables are merged by the peer all-merge kernel.  No NCCL call, no host synchronisation
 * between the first send and the result read.  The comm's user heap must hold ldb_tpch_q5_repartitioned_heap_bytes(). */
typedef struct LdbQ5ShuffleStats {
   int64_t orders_tuples_sent, orders_tuples_received, lineitem_tuples_sent, lineitem_tuples_received;
   int64_t shuffle_bytes_out; /* 8 B per orders tuple + 24 B per lineitem tuple, tuples that stay on this rank included */
   int64_t heap_bytes;
} LdbQ5ShuffleStats;
int64_t ldb_tpch_q5_repartitioned_heap_bytes(int64_t n_orders_total, int64_t n_lineitem_total, int32_t world);
int ldb_tpch_q5_repartitioned(LdbContext* ctx, const LdbTpchTables* t, struct LdbComm* comm, const char* region_name, const char* date_ge, const char* date_lt,
                              int64_t n_orders_total, int64_t n_lineitem_total, LdbQ5Row* rows /* 25 */, int32_t* n_rows, LdbQ5ShuffleStats* stats, LdbError* err);

/* Q9 with orders HASH-PARTITIONED across the ranks of `comm` (BASELINE.json config 4: "large build side … NVLink shuffle"): `t` holds
 * this rank's shard of orders and of lineitem (ANY split — the join is not assumed co-partitioned) and full copies of part, partsupp,
 * supplier (the composite-key and the supplier tables are built on every rank).  {o_orderkey, year(o_orderdate)} tuples go to the
 * owner of h64(o_orderkey) (K10), every rank builds its orders partition; lineitem rows that survive the partsupp probe ship their
 * contribution {l_orderkey | nation, amount} to the same owner (K11), which probes its partition for the year and aggregates; the
 * 175-group tables are all-merged.  stats: orders_* = orders tuples, lineitem_* = lineitem tuples (24 B each). */
int64_t ldb_tpch_q9_repartitioned_heap_bytes(int64_t n_orders_total, int64_t n_lineitem_total, int32_t world);
int ldb_tpch_q9_repartitioned(LdbContext* ctx, const LdbTpchTables* t, struct LdbComm* comm, const char* name_contains, int64_t n_orders_total, int64_t n_lineitem_total,
                              LdbQ9Row* rows /* max_rows */, int32_t max_rows, int32_t* n_rows, LdbQ5ShuffleStats* stats, LdbError* err);

#ifdef __cplusplus
}
#endif
#endif
