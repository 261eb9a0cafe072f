/* ldb_arrow_io.h — C++ storage → HBM path (libldb_arrow_io.so).
 * Replaces: LingoDBTable::loadTable (src/runtime/storage/LingoDBTable.cpp:27-54: arrow::ipc::RecordBatchFileReader over
 * `<dbDir>/<table>.arrow`) + TableChunk::getArrayView (:200-225).  The file is memory-mapped; every record batch is appended to
 * a backend table as LdbArrayViews over the mapped buffers (zero copy on the host, validity bitmaps and offsets respected).  The
 * staged table is the column cache: it stays resident until ldb_gpu_table_clear (ownership as LingoDBTable.cpp:294-305). */
#ifndef LDB_ARROW_IO_H
#define LDB_ARROW_IO_H
#include "ldb_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct LdbArrowFile LdbArrowFile;
/* columns == NULL / n_columns <= 0: every column of the file */
int ldb_arrow_file_open(const char* path, const char* const* columns, int32_t n_columns, LdbArrowFile** out, LdbError* err);
int32_t ldb_arrow_file_num_columns(const LdbArrowFile* f);
const LdbColumnSchema* ldb_arrow_file_schema(const LdbArrowFile* f); /* valid until close: pass to ldb_gpu_table_create */
int32_t ldb_arrow_file_num_batches(const LdbArrowFile* f);
typedef int (*LdbAppendBatchFn)(LdbTable*, int64_t, const LdbArrayView*, const int64_t*, int32_t, LdbError*); /* = ldb_gpu_table_append_batch */
/* appends every record batch (split at max_rows_per_batch rows; <= 0: whole batches) to `table`.  The file must stay open
 * until the table was cleared or destroyed: the views point into the mapping. */
int ldb_arrow_file_load(LdbArrowFile* f, LdbTable* table, LdbAppendBatchFn append, int64_t max_rows_per_batch, int64_t* rows_out, LdbError* err);
void ldb_arrow_file_close(LdbArrowFile* f);
#ifdef __cplusplus
}
#endif
#endif
