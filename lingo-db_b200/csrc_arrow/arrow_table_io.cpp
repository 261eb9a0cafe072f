// arrow_table_io.cpp — C++ storage → HBM path: LingoDB persists a table as one Arrow IPC *file*
// (`<dbDir>/<table>.arrow`; storeTable / loadTable use arrow::ipc::MakeFileWriter / RecordBatchFileReader,
// src/runtime/storage/LingoDBTable.cpp:27-54) and scans ArrayViews that point INTO the Arrow buffers (:200-225).  This library
// opens such a file memory-mapped with the same Arrow C++ API and hands every record batch to the GPU backend's C-ABI as
// LdbArrayViews over the mapped buffers — no copy on the host, validity bitmaps and array offsets respected; the backend's
// compressed staging does the rest.  The staged table IS the column cache: it stays in HBM until ldb_gpu_table_clear, later
// pipelines read it without touching the file again (ownership as in LingoDBTable.cpp:294-305, where the table object owns
// its batches).  Built as its own shared library (libldb_arrow_io.so) so that libldb_gpu.so does not depend on Arrow.
#include "../../include/ldb_arrow_io.h"

#include <arrow/api.h>
#include <arrow/io/file.h>
#include <arrow/ipc/reader.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {
int fail(LdbError* err, int code, const std::string& m) {
   if (err) {
      err->code = code;
      snprintf(err->message, sizeof(err->message), "%s", m.c_str());
   }
   return code;
}
bool physType(const arrow::DataType& t, LdbColumnSchema* out) {
   out->precision = out->scale = 0;
   switch (t.id()) {
      case arrow::Type::INT8: out->type = LDB_INT8; return true;
      case arrow::Type::INT16: out->type = LDB_INT16; return true;
      case arrow::Type::INT32: out->type = LDB_INT32; return true;
      case arrow::Type::INT64: out->type = LDB_INT64; return true;
      case arrow::Type::DATE32: out->type = LDB_DATE32; return true;
      case arrow::Type::FLOAT: out->type = LDB_FLOAT32; return true;
      case arrow::Type::DOUBLE: out->type = LDB_FLOAT64; return true;
      case arrow::Type::STRING: out->type = LDB_UTF8; return true;
      case arrow::Type::DECIMAL128: {
         const auto& d = static_cast<const arrow::Decimal128Type&>(t);
         out->type = LDB_DECIMAL128;
         out->precision = d.precision();
         out->scale = d.scale();
         return true;
      }
      case arrow::Type::FIXED_SIZE_BINARY:
         if (static_cast<const arrow::FixedSizeBinaryType&>(t).byte_width() == 4) { // char(1) as LingoDBTable.cpp:122-195 stores it
            out->type = LDB_FSB4;
            return true;
         }
         return false;
      default: return false;
   }
}
} // namespace

struct LdbArrowFile {
   std::shared_ptr<arrow::io::MemoryMappedFile> file;
   std::shared_ptr<arrow::ipc::RecordBatchFileReader> reader;
   std::vector<std::string> names;
   std::vector<LdbColumnSchema> schema;
   std::vector<int> fieldIndex;                                   // selected column → field of the file
   std::vector<std::shared_ptr<arrow::RecordBatch>> keepAlive;    // batches whose buffers a table still points into
};

extern "C" {

int ldb_arrow_file_open(const char* path, const char* const* columns, int32_t n_columns, LdbArrowFile** out, LdbError* err) {
   if (!path || !out) return fail(err, LDB_ERR_INVALID, "null argument");
   auto f = std::make_unique<LdbArrowFile>();
   auto mm = arrow::io::MemoryMappedFile::Open(path, arrow::io::FileMode::READ);
   if (!mm.ok()) return fail(err, LDB_ERR_INVALID, "cannot open " + std::string(path) + ": " + mm.status().ToString());
   f->file = *mm;
   auto rd = arrow::ipc::RecordBatchFileReader::Open(f->file);
   if (!rd.ok()) return fail(err, LDB_ERR_INVALID, "not an Arrow IPC file: " + rd.status().ToString());
   f->reader = *rd;
   auto schema = f->reader->schema();
   for (int i = 0; i < schema->num_fields(); i++) {
      const auto& field = schema->field(i);
      bool wanted = n_columns <= 0;
      for (int c = 0; c < n_columns; c++) wanted |= field->name() == columns[c];
      if (!wanted) continue;
      LdbColumnSchema cs{};
      if (!physType(*field->type(), &cs)) return fail(err, LDB_ERR_UNSUPPORTED, "column " + field->name() + ": Arrow type " + field->type()->ToString() + " is not on the GPU hot path");
      f->names.push_back(field->name());
      f->schema.push_back(cs);
      f->fieldIndex.push_back(i);
   }
   for (size_t i = 0; i < f->names.size(); i++) f->schema[i].name = f->names[i].c_str();
   if (n_columns > 0 && (int) f->names.size() != n_columns) return fail(err, LDB_ERR_INVALID, "a requested column is not in the file");
   if (err) {
      err->code = LDB_OK;
      err->message[0] = 0;
   }
   *out = f.release();
   return LDB_OK;
}
int32_t ldb_arrow_file_num_columns(const LdbArrowFile* f) { return f ? (int32_t) f->schema.size() : 0; }
const LdbColumnSchema* ldb_arrow_file_schema(const LdbArrowFile* f) { return f ? f->schema.data() : nullptr; }
int32_t ldb_arrow_file_num_batches(const LdbArrowFile* f) { return f ? f->reader->num_record_batches() : 0; }
void ldb_arrow_file_close(LdbArrowFile* f) { delete f; }

// create the backend table (named after `name`) and append every record batch of the file through `append`
// (= ldb_gpu_table_append_batch; passed as a pointer so that this library needs no link-time dependency on libldb_gpu.so)
int ldb_arrow_file_load(LdbArrowFile* f, LdbTable* table, LdbAppendBatchFn append, int64_t max_rows_per_batch, int64_t* rows_out, LdbError* err) {
   if (!f || !table || !append) return fail(err, LDB_ERR_INVALID, "null argument");
   if (max_rows_per_batch <= 0) max_rows_per_batch = INT64_MAX;
   int64_t total = 0;
   const size_t nc = f->schema.size();
   for (int b = 0; b < f->reader->num_record_batches(); b++) {
      auto rb = f->reader->ReadRecordBatch(b);
      if (!rb.ok()) return fail(err, LDB_ERR_INVALID, "reading record batch: " + rb.status().ToString());
      std::shared_ptr<arrow::RecordBatch> batch = *rb;
      f->keepAlive.push_back(batch); // zero copy: the views below point into the mapped file
      for (int64_t r0 = 0; r0 < batch->num_rows(); r0 += max_rows_per_batch) {
         const int64_t n = std::min<int64_t>(max_rows_per_batch, batch->num_rows() - r0);
         std::vector<LdbArrayView> views(nc);
         std::vector<std::vector<const void*>> bufs(nc, std::vector<const void*>(3, nullptr));
         std::vector<int64_t> utf8(nc, 0);
         for (size_t c = 0; c < nc; c++) {
            const auto& data = batch->column_data(f->fieldIndex[c]);
            LdbArrayView& v = views[c];
            v.length = n;
            v.offset = data->offset + r0;
            v.null_count = data->GetNullCount() ? -1 : 0; // "look at the bitmap" (a slice may still be null-free)
            v.n_buffers = (int64_t) data->buffers.size();
            v.n_children = 0;
            v.children = nullptr;
            for (size_t k = 0; k < data->buffers.size() && k < 3; k++) bufs[c][k] = data->buffers[k] ? data->buffers[k]->data() : nullptr;
            if (!v.null_count) bufs[c][0] = nullptr;
            v.buffers = bufs[c].data();
            if (f->schema[c].type == LDB_UTF8) utf8[c] = data->buffers[2] ? data->buffers[2]->size() : 0;
         }
         int rc = append(table, n, views.data(), utf8.data(), LDB_MEM_HOST, err);
         if (rc != LDB_OK) return rc;
         total += n;
      }
   }
   if (rows_out) *rows_out = total;
   if (err) {
      err->code = LDB_OK;
      err->message[0] = 0;
   }
   return LDB_OK;
}

} // extern "C"
