/* ldb_datagen.h — C-ABI of the deterministic TPC-H-shaped table generator.
 *
 * Not part of the reference's runtime surface: the reference gets its tables from dbgen + COPY
 * (tools/generate/tpch.sh, resources/sql/tpch/initialize.sql).  BASELINE.json asks for
 * "generator-synthesised tables"; this generator produces them directly in the reference's Arrow
 * physical column layout (src/runtime/storage/LingoDBTable.cpp:122-195) on the host
 * (libldb_datagen_host.so) and on the device (ldb_gpu_datagen_* in libldb_gpu.so), bit-identically.
 */
#ifndef LDB_DATAGEN_H
#define LDB_DATAGEN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct LdbGenScale {
   uint64_t seed;
   int64_t n_orders, n_customer, n_supplier, n_part;
   int64_t n_lineitem; /* derived: closed-form prefix of the per-order line counts */
} LdbGenScale;

/* Column buffers; NULL = skip the column.  decimal columns are 16 bytes/value. */
typedef struct LdbGenLineitemCols {
   int32_t* l_orderkey;
   int32_t* l_partkey;
   int32_t* l_suppkey;
   uint8_t* l_quantity;      /* decimal128(12,2) */
   uint8_t* l_extendedprice; /* decimal128(12,2) */
   uint8_t* l_discount;      /* decimal128(12,2) */
   uint8_t* l_tax;           /* decimal128(12,2) */
   int32_t* l_returnflag;    /* fixed_size_binary(4) */
   int32_t* l_linestatus;    /* fixed_size_binary(4) */
   int32_t* l_shipdate;      /* date32 */
   int32_t* l_commitdate;
   int32_t* l_receiptdate;
} LdbGenLineitemCols;

typedef struct LdbGenOrdersCols {
   int32_t* o_orderkey;
   int32_t* o_custkey;
   int32_t* o_orderdate; /* date32 */
   int32_t* o_shippriority;
} LdbGenOrdersCols;

typedef struct LdbGenCustomerCols {
   int32_t* c_custkey;
   int32_t* c_nationkey;
   int32_t* c_mktsegment_offsets; /* utf8 offsets, n_rows + 1 */
   uint8_t* c_mktsegment_data;
} LdbGenCustomerCols;

typedef struct LdbGenSupplierCols {
   int32_t* s_suppkey;
   int32_t* s_nationkey;
} LdbGenSupplierCols;

typedef struct LdbGenPartCols {
   int32_t* p_partkey;
   int32_t* p_name_offsets; /* utf8 offsets, n_rows + 1 */
   uint8_t* p_name_data;
} LdbGenPartCols;

typedef struct LdbGenPartsuppCols {
   int32_t* ps_partkey;
   int32_t* ps_suppkey;
   uint8_t* ps_supplycost; /* decimal128(12,2) */
} LdbGenPartsuppCols;

/* host side (libldb_datagen_host.so) */
void ldbgen_scale(double sf, uint64_t seed, LdbGenScale* out);
int64_t ldbgen_order_first_line(const LdbGenScale* g, int64_t order_idx);
void ldbgen_lineitem_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenLineitemCols* cols);
void ldbgen_orders_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenOrdersCols* cols);
int64_t ldbgen_customer_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenCustomerCols* cols);
void ldbgen_supplier_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenSupplierCols* cols);
int64_t ldbgen_part_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenPartCols* cols); /* returns utf8 bytes, like customer */
void ldbgen_partsupp_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenPartsuppCols* cols); /* 4 * n_part rows */

/* ---- dbgen-faithful variant (csrc/dbgen_gen.h): the TPC generator's own random streams for the same columns, so that
 * tables are the ones `dbgen -s SF` writes (validated against the reference's SF1 answers, tests/test_reference_answers_sf1.py).
 * Line counts per order are random (1..7), so lineitem is addressed by ORDER: the caller takes the counts, prefix-sums them
 * and passes each order's first output row.  LdbGenScale.seed is unused (dbgen's seeds are fixed); n_lineitem is filled by
 * ldbgen_dbgen_scale when count_lines != 0 (one pass over the line-count stream). */
void ldbgen_dbgen_scale(double sf, int32_t count_lines, LdbGenScale* out);
void ldbgen_dbgen_line_counts_host(const LdbGenScale* g, int64_t order_begin, int64_t n_orders, int32_t* counts);
void ldbgen_dbgen_lineitem_host(const LdbGenScale* g, int64_t order_begin, int64_t n_orders, const int64_t* first_row /* per order, into cols */, const LdbGenLineitemCols* cols);
void ldbgen_dbgen_orders_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenOrdersCols* cols);
int64_t ldbgen_dbgen_customer_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenCustomerCols* cols);
void ldbgen_dbgen_supplier_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenSupplierCols* cols);
int64_t ldbgen_dbgen_part_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenPartCols* cols);
void ldbgen_dbgen_partsupp_host(const LdbGenScale* g, int64_t row_begin, int64_t n_rows, const LdbGenPartsuppCols* cols);

#ifdef __cplusplus
}
#endif
#endif
