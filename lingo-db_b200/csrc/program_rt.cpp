// program_rt.cpp — host side of the generic program pipeline (include/ldb_gpu.h "program pipelines"): validates a program,
// binds it batch by batch and launches the interpreter kernel; hash-aggregation states, their read-back / export as a table,
// ORDER BY … LIMIT through the device radix sort.
#include "context.h"
#include "program.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

using namespace ldb;

namespace {
template <class Fn>
int guardedP(LdbError* err, const Fn& fn) {
   auto set = [&](int code, const char* msg) {
      if (err) {
         err->code = code;
         snprintf(err->message, sizeof(err->message), "%s", msg);
      }
      return code;
   };
   try {
      fn();
      if (err) {
         err->code = LDB_OK;
         err->message[0] = 0;
      }
      return LDB_OK;
   } catch (const CudaError& e) {
      return set(e.code, e.what());
   } catch (const ApiError& e) {
      return set(e.code, e.what());
   } catch (const std::exception& e) {
      return set(LDB_ERR_INVALID, e.what());
   }
}
[[noreturn]] void failP(int code, const std::string& m) { throw ApiError(code, m); }
uint64_t nextPow2P(uint64_t v) {
   v--;
   for (int s = 1; s < 64; s <<= 1) v |= v >> s;
   return v + 1;
}
bool isCountKind(int k) { return k == LDB_AGG_COUNT || k == LDB_AGG_COUNT_STAR; }
size_t cellBytes(int type) {
   switch (type) {
      case LDB_INT8: return 1;
      case LDB_INT16: return 2;
      case LDB_INT32:
      case LDB_DATE32:
      case LDB_FSB4:
      case LDB_FLOAT32:
      case LDB_UTF8: return 4;
      case LDB_INT64:
      case LDB_FLOAT64: return 8;
      default: return 16;
   }
}
} // namespace

extern "C" {

int ldb_gpu_hashagg_create(LdbContext* ctx, int32_t n_keys, int32_t n_aggs, const LdbProgAgg* aggs, int64_t expected_groups, LdbState** out, LdbError* err) {
   return guardedP(err, [&] {
      if (!ctx || !out || (n_aggs > 0 && !aggs)) failP(LDB_ERR_INVALID, "null argument");
      if (n_keys < 0 || n_keys > kProgMaxKeys || n_aggs < 0 || n_aggs > kProgMaxAggs) failP(LDB_ERR_INVALID, "hash aggregation takes 0..4 keys and 0..8 aggregates");
      for (int a = 0; a < n_aggs; a++)
         if (aggs[a].kind < LDB_AGG_SUM || aggs[a].kind > LDB_AGG_ANY) failP(LDB_ERR_INVALID, "unknown aggregate kind");
      LDB_CUDA(cudaSetDevice(ctx->device));
      auto* s = new LdbState;
      s->ctx = ctx;
      s->kind = LDB_STATE_HASHAGG;
      ctx->states.push_back(s);
      auto& h = s->hashagg;
      h.nKeys = n_keys;
      h.nAggs = n_aggs;
      h.entryBytes = (uint32_t) (48 + 16 * n_aggs);
      const uint64_t cap = n_keys == 0 ? 1 : nextPow2P((uint64_t) std::max<int64_t>(expected_groups, 8) * 2);
      h.mask = cap - 1;
      h.base = (uint8_t*) ctx->stagingAlloc(cap * h.entryBytes);
      s->allocations.push_back(h.base);
      ctx->launch("hashagg_init", [&] { launchHashAggInit(h, ctx->smCount, ctx->compute); });
      uint8_t* small = (uint8_t*) ctx->stagingAlloc(256);
      s->allocations.push_back(small);
      LDB_CUDA(cudaMemsetAsync(small, 0, 256, ctx->compute));
      h.count = (unsigned long long*) small;
      h.error = (int32_t*) (small + 8);
      for (int a = 0; a < n_aggs; a++) s->aggKinds[a] = aggs[a].kind;
      if (n_keys == 0) { // the one group exists from the start, with the aggregates' identities
         std::vector<uint8_t> e(h.entryBytes, 0);
         *(uint32_t*) e.data() = 2;
         unsigned long long* ea = (unsigned long long*) (e.data() + 48);
         for (int a = 0; a < n_aggs; a++) {
            double inf = 1.0 / 0.0, ninf = -inf;
            switch (aggs[a].kind) {
               case LDB_AGG_MIN: ea[2 * a] = (unsigned long long) INT64_MAX; break;
               case LDB_AGG_MAX: ea[2 * a] = (unsigned long long) INT64_MIN; break;
               case LDB_AGG_MIN_F64: memcpy(&ea[2 * a], &inf, 8); break;
               case LDB_AGG_MAX_F64: memcpy(&ea[2 * a], &ninf, 8); break;
               default: break;
            }
         }
         LDB_CUDA(cudaMemcpyAsync(h.base, e.data(), e.size(), cudaMemcpyHostToDevice, ctx->compute));
         unsigned long long one = 1;
         LDB_CUDA(cudaMemcpyAsync(h.count, &one, 8, cudaMemcpyHostToDevice, ctx->compute));
         ctx->syncStream(ctx->compute); // the host buffers above are locals
      }
      *out = s;
   });
}
static void checkHashAgg(LdbState* s) {
   if (!s || s->kind != LDB_STATE_HASHAGG) failP(LDB_ERR_INVALID, "not a hash aggregation state");
}
int ldb_gpu_hashagg_count(LdbState* s, int64_t* n_groups, LdbError* err) {
   return guardedP(err, [&] {
      checkHashAgg(s);
      unsigned long long h[2] = {0, 0};
      LDB_CUDA(cudaMemcpyAsync(h, s->hashagg.count, 16, cudaMemcpyDeviceToHost, s->ctx->compute));
      s->ctx->syncStream(s->ctx->compute);
      if ((int32_t) h[1]) failP(LDB_ERR_CAPACITY, "hash aggregation table full: more groups than expected_groups allowed");
      *n_groups = (int64_t) h[0];
   });
}

// export into fresh device columns; returns the row count (synchronises)
struct ExportedGroups {
   int64_t n = 0;
   std::vector<int64_t*> keyCols;
   std::vector<uint8_t*> keyValid, aggCols, aggValid;
};
static ExportedGroups exportGroups(LdbState* s, std::vector<void*>& owned) {
   LdbContext* ctx = s->ctx;
   auto& h = s->hashagg;
   int64_t n = 0;
   LdbError e;
   if (ldb_gpu_hashagg_count(s, &n, &e) != LDB_OK) failP(e.code, e.message);
   ExportedGroups g;
   g.n = n;
   const size_t rows = (size_t) std::max<int64_t>(n, 1);
   auto alloc = [&](size_t bytes) {
      void* p = ctx->stagingAlloc(bytes);
      owned.push_back(p);
      return p;
   };
   for (int k = 0; k < h.nKeys; k++) {
      g.keyCols.push_back((int64_t*) alloc(rows * 8));
      g.keyValid.push_back((uint8_t*) alloc(rows));
   }
   uint32_t countMask = 0;
   for (int a = 0; a < h.nAggs; a++) {
      g.aggCols.push_back((uint8_t*) alloc(rows * 16));
      g.aggValid.push_back((uint8_t*) alloc(rows));
      if (isCountKind(s->aggKinds[a])) countMask |= 1u << a;
   }
   unsigned long long* counter = (unsigned long long*) alloc(8);
   LDB_CUDA(cudaMemsetAsync(counter, 0, 8, ctx->compute));
   ctx->launch("hashagg_export", [&] { launchHashAggExport(h, g.keyCols.data(), g.keyValid.data(), g.aggCols.data(), g.aggValid.data(), counter, countMask, ctx->smCount, ctx->compute); });
   return g;
}
int ldb_gpu_hashagg_read(LdbState* s, LdbHashAggRow* rows, int64_t max_rows, int64_t* n_rows, LdbError* err) {
   return guardedP(err, [&] {
      checkHashAgg(s);
      if (!rows || !n_rows) failP(LDB_ERR_INVALID, "null argument");
      LdbContext* ctx = s->ctx;
      LDB_CUDA(cudaSetDevice(ctx->device));
      std::vector<void*> owned;
      ExportedGroups g = exportGroups(s, owned);
      auto& h = s->hashagg;
      const int64_t n = std::min(g.n, max_rows);
      std::vector<std::vector<int64_t>> keys(h.nKeys, std::vector<int64_t>((size_t) n));
      std::vector<std::vector<uint8_t>> kv(h.nKeys, std::vector<uint8_t>((size_t) n)), av(h.nAggs, std::vector<uint8_t>((size_t) n));
      std::vector<std::vector<LdbI128>> aggs(h.nAggs, std::vector<LdbI128>((size_t) n));
      for (int k = 0; k < h.nKeys && n; k++) {
         LDB_CUDA(cudaMemcpyAsync(keys[k].data(), g.keyCols[k], (size_t) n * 8, cudaMemcpyDeviceToHost, ctx->compute));
         LDB_CUDA(cudaMemcpyAsync(kv[k].data(), g.keyValid[k], (size_t) n, cudaMemcpyDeviceToHost, ctx->compute));
      }
      for (int a = 0; a < h.nAggs && n; a++) {
         LDB_CUDA(cudaMemcpyAsync(aggs[a].data(), g.aggCols[a], (size_t) n * 16, cudaMemcpyDeviceToHost, ctx->compute));
         LDB_CUDA(cudaMemcpyAsync(av[a].data(), g.aggValid[a], (size_t) n, cudaMemcpyDeviceToHost, ctx->compute));
      }
      ctx->syncStream(ctx->compute);
      for (int64_t i = 0; i < n; i++) {
         LdbHashAggRow& r = rows[i];
         memset(&r, 0, sizeof(r));
         for (int k = 0; k < h.nKeys; k++) {
            r.keys[k] = keys[k][(size_t) i];
            if (!kv[k][(size_t) i]) r.key_null_mask |= 1u << k;
         }
         for (int a = 0; a < h.nAggs; a++) {
            r.aggs[a] = aggs[a][(size_t) i];
            if (av[a][(size_t) i]) r.agg_valid_mask |= 1u << a;
         }
      }
      *n_rows = g.n;
      for (void* p : owned) ctx->stagingRelease(p);
   });
}
int ldb_gpu_hashagg_to_table(LdbState* s, const char* name, LdbTable** out, LdbError* err) {
   return guardedP(err, [&] {
      checkHashAgg(s);
      if (!out) failP(LDB_ERR_INVALID, "null argument");
      LdbContext* ctx = s->ctx;
      LDB_CUDA(cudaSetDevice(ctx->device));
      std::vector<void*> owned;
      ExportedGroups g = exportGroups(s, owned);
      ctx->syncStream(ctx->compute);
      auto& h = s->hashagg;
      auto* t = new LdbTable;
      t->ctx = ctx;
      t->name = name ? name : "groups";
      LdbBatch b;
      b.nRows = g.n;
      for (int k = 0; k < h.nKeys; k++) {
         t->columns.push_back({"k" + std::to_string(k), LDB_INT64, 0, 0});
         b.data.push_back(g.keyCols[k]);
         b.bytes.push_back(nullptr);
         b.elemBytes.push_back(8);
         b.validBytes.push_back(g.keyValid[k]);
      }
      for (int a = 0; a < h.nAggs; a++) {
         const int kind = s->aggKinds[a];
         const bool f64 = kind == LDB_AGG_SUM_F64 || kind == LDB_AGG_MIN_F64 || kind == LDB_AGG_MAX_F64;
         t->columns.push_back({"a" + std::to_string(a), f64 ? LDB_FLOAT64 : LDB_DECIMAL128, 38, 0});
         b.data.push_back(g.aggCols[a]);
         b.bytes.push_back(nullptr);
         b.elemBytes.push_back(16); // doubles keep the 16-byte stride (bits in the low 8 bytes)
         b.validBytes.push_back(g.aggValid[a]);
      }
      b.validity.assign(b.data.size(), nullptr);
      b.validityBitOffset.assign(b.data.size(), 0);
      b.owned = owned; // the table owns the exported buffers
      t->numRows = g.n;
      t->batches.push_back(std::move(b));
      ctx->tables.push_back(t);
      *out = t;
   });
}

int ldb_gpu_run_program(LdbContext* ctx, const LdbProgramDesc* d, LdbError* err) {
   return guardedP(err, [&] {
      if (!ctx || !d || !d->source) failP(LDB_ERR_INVALID, "null argument");
      LdbTable* t = d->source;
      if (t->ctx != ctx) failP(LDB_ERR_INVALID, "table belongs to another context");
      if (d->n_columns < 0 || d->n_columns > kProgMaxCols || d->n_instr < 0 || d->n_instr > kProgMaxInstr || d->n_consts < 0 || d->n_consts > kProgMaxConsts ||
          d->n_strings < 0 || d->n_strings > kProgMaxStrings || d->n_tables < 0 || d->n_tables > kProgMaxTables)
         failP(LDB_ERR_UNSUPPORTED, "program exceeds the interpreter's limits (12 columns, 96 instructions, 24 constants, 12 strings, 4 tables)");
      LDB_CUDA(cudaSetDevice(ctx->device));
      ProgramParams base{};
      base.nCols = d->n_columns;
      base.nInstr = d->n_instr;
      base.nTables = d->n_tables;
      std::vector<int> colIdx((size_t) d->n_columns);
      for (int c = 0; c < d->n_columns; c++) {
         colIdx[c] = t->colIndex(d->columns[c]);
         if (colIdx[c] < 0) failP(LDB_ERR_INVALID, std::string("unknown column ") + (d->columns[c] ? d->columns[c] : "(null)"));
      }
      // static validation: every register read was written before, every index is in range, types fit the opcode
      bool written[kProgMaxRegs] = {};
      auto wantReg = [&](int r, const char* what) {
         if (r < 0 || r >= kProgMaxRegs || !written[r]) failP(LDB_ERR_INVALID, std::string("program reads an unwritten or out-of-range register (") + what + ")");
      };
      for (int i = 0; i < d->n_instr; i++) {
         const LdbInstr& in = d->instr[i];
         if (in.dst >= kProgMaxRegs) failP(LDB_ERR_INVALID, "destination register out of range");
         switch (in.op) {
            case LDB_OP_LOAD:
               if (in.arg < 0 || in.arg >= d->n_columns) failP(LDB_ERR_INVALID, "LOAD: column index out of range");
               if (t->columns[colIdx[in.arg]].type == LDB_UTF8) failP(LDB_ERR_UNSUPPORTED, "LOAD of a string column (strings are operands of STRCMP / STRLIKE only)");
               break;
            case LDB_OP_CONST:
               if (in.arg < 0 || in.arg >= d->n_consts) failP(LDB_ERR_INVALID, "CONST: constant index out of range");
               break;
            case LDB_OP_ADD: case LDB_OP_SUB: case LDB_OP_MUL: case LDB_OP_DIV: case LDB_OP_AND: case LDB_OP_OR:
            case LDB_OP_FADD: case LDB_OP_FSUB: case LDB_OP_FMUL: case LDB_OP_FDIV:
               wantReg(in.a, "a");
               wantReg(in.b, "b");
               break;
            case LDB_OP_CMP: case LDB_OP_FCMP:
               wantReg(in.a, "a");
               wantReg(in.b, "b");
               if (in.arg < LDB_EQ || in.arg > LDB_GTE) failP(LDB_ERR_INVALID, "CMP: unknown comparison");
               break;
            case LDB_OP_NEG: case LDB_OP_NOT: case LDB_OP_ISNULL: case LDB_OP_I2F: case LDB_OP_YEAR: wantReg(in.a, "a"); break;
            case LDB_OP_SELECT:
               wantReg(in.a, "a");
               wantReg(in.b, "b");
               wantReg(in.arg, "condition");
               break;
            case LDB_OP_STRKEY8:
               if (in.a >= d->n_columns || t->columns[colIdx[in.a]].type != LDB_UTF8) failP(LDB_ERR_INVALID, "STRKEY8 needs a utf8 column");
               break;
            case LDB_OP_STRCMP: case LDB_OP_STRLIKE:
               if (in.a >= d->n_columns || t->columns[colIdx[in.a]].type != LDB_UTF8) failP(LDB_ERR_INVALID, "string op needs a utf8 column");
               if (in.arg < 0 || in.arg >= d->n_strings) failP(LDB_ERR_INVALID, "string constant index out of range");
               if (in.op == LDB_OP_STRCMP ? in.b > LDB_GTE : in.b > 2) failP(LDB_ERR_INVALID, "string op: unknown comparison / pattern kind");
               break;
            case LDB_OP_PROBE:
               wantReg(in.a, "key");
               if (in.arg < 0 || in.arg >= d->n_tables) failP(LDB_ERR_INVALID, "PROBE: table index out of range");
               break;
            default: failP(LDB_ERR_UNSUPPORTED, "unknown opcode " + std::to_string(in.op));
         }
         written[in.dst] = true;
         base.instr[i] = ProgInstr{in.op, in.dst, in.a, in.b, in.arg};
      }
      for (int c = 0; c < d->n_consts; c++) {
         base.constLo[c] = d->consts[c].lo;
         base.constHi[c] = d->consts[c].hi;
      }
      for (int c = 0; c < d->n_strings; c++) {
         const size_t n = d->strings[c] ? strlen(d->strings[c]) : 0;
         if (n > (size_t) kProgStringBytes) failP(LDB_ERR_UNSUPPORTED, "string constant longer than 32 bytes");
         memcpy(base.strings[c], d->strings[c], n);
         base.stringLen[c] = (int32_t) n;
      }
      for (int k = 0; k < d->n_tables; k++) {
         LdbState* js = d->tables[k];
         if (!js || js->kind != LDB_STATE_JOIN_TABLE || js->join.stride == 16) failP(LDB_ERR_INVALID, "PROBE tables are single-key join tables");
         base.tables[k] = js->join;
      }
      base.filterReg = d->filter_reg;
      if (d->filter_reg >= 0) wantReg(d->filter_reg, "filter");
      base.sinkKind = d->sink_kind;
      LdbState* sink = d->sink;
      std::vector<void*> outOwned;
      std::vector<uint8_t*> outVals, outValid;
      unsigned long long* outCount = nullptr;
      if (d->sink_kind == LDB_SINK_HASHAGG) {
         checkHashAgg(sink);
         if (sink->ctx != ctx || d->n_keys != sink->hashagg.nKeys || d->n_aggs != sink->hashagg.nAggs) failP(LDB_ERR_INVALID, "key / aggregate count differs from the state's");
         base.nKeys = d->n_keys;
         base.nAggs = d->n_aggs;
         for (int k = 0; k < d->n_keys; k++) {
            wantReg(d->key_regs[k], "group key");
            base.keyReg[k] = d->key_regs[k];
         }
         for (int a = 0; a < d->n_aggs; a++) {
            if (d->aggs[a].kind != sink->aggKinds[a]) failP(LDB_ERR_INVALID, "aggregate kind differs from the state's");
            if (d->aggs[a].kind != LDB_AGG_COUNT_STAR) wantReg(d->aggs[a].reg, "aggregate input");
            base.aggs[a] = ProgAgg{d->aggs[a].kind, d->aggs[a].reg};
         }
         base.agg = sink->hashagg;
      } else if (d->sink_kind == LDB_SINK_JOIN_BUILD) {
         if (!sink || sink->kind != LDB_STATE_JOIN_TABLE || sink->join.stride != 8 || sink->join.direct) failP(LDB_ERR_INVALID, "build sink must be a plain single-key join table");
         wantReg(d->build_key_reg, "build key");
         if (d->build_payload_reg >= 0) wantReg(d->build_payload_reg, "build payload");
         base.buildKeyReg = d->build_key_reg;
         base.buildPayloadReg = d->build_payload_reg;
         base.build = sink->join;
      } else if (d->sink_kind == LDB_SINK_MATERIALIZE) {
         if (d->n_out < 1 || d->n_out > kProgMaxAggs || !d->out_table) failP(LDB_ERR_INVALID, "materialize needs 1..8 output registers and out_table");
         const size_t rows = (size_t) std::max<int64_t>(t->numRows, 1);
         base.nOut = d->n_out;
         for (int c = 0; c < d->n_out; c++) {
            wantReg(d->out_regs[c], "output");
            base.outReg[c] = d->out_regs[c];
            outVals.push_back((uint8_t*) ctx->stagingAlloc(rows * 16));
            outValid.push_back((uint8_t*) ctx->stagingAlloc(rows));
            outOwned.push_back(outVals.back());
            outOwned.push_back(outValid.back());
            base.outValues[c] = outVals.back();
            base.outValid[c] = outValid.back();
         }
         outCount = (unsigned long long*) ctx->stagingAlloc(8);
         outOwned.push_back(outCount);
         LDB_CUDA(cudaMemsetAsync(outCount, 0, 8, ctx->compute));
         base.outCapacity = (int64_t) rows;
         base.outCount = outCount;
      } else {
         failP(LDB_ERR_INVALID, "unknown sink kind");
      }
      for (auto& b : t->batches) {
         if (b.nRows == 0) continue;
         ProgramParams p = base;
         p.nRows = b.nRows;
         for (int c = 0; c < d->n_columns; c++) {
            const int ci = colIdx[c];
            ProgCol& pc = p.cols[c];
            pc.data = (const uint8_t*) b.data[ci];
            pc.bytes = (const uint8_t*) b.bytes[ci];
            pc.type = t->columns[ci].type;
            pc.elemBytes = b.elemBytes[ci];
            pc.validity = ci < (int) b.validity.size() ? (const uint8_t*) b.validity[ci] : nullptr;
            pc.bitOffset = ci < (int) b.validityBitOffset.size() ? b.validityBitOffset[ci] : 0;
            pc.validBytes = ci < (int) b.validBytes.size() ? b.validBytes[ci] : nullptr;
         }
         ldb_gpu_wait_batch_internal(ctx, &b);
         ctx->launch("program", [&] { launchProgram(p, ctx->smCount, ctx->compute); });
      }
      if (d->sink_kind == LDB_SINK_MATERIALIZE) {
         unsigned long long n = 0;
         LDB_CUDA(cudaMemcpyAsync(&n, outCount, 8, cudaMemcpyDeviceToHost, ctx->compute));
         ctx->syncStream(ctx->compute);
         auto* ot = new LdbTable;
         ot->ctx = ctx;
         ot->name = t->name + "_out";
         LdbBatch ob;
         ob.nRows = (int64_t) n;
         for (int c = 0; c < d->n_out; c++) {
            ot->columns.push_back({"c" + std::to_string(c), LDB_DECIMAL128, 38, 0});
            ob.data.push_back(outVals[c]);
            ob.bytes.push_back(nullptr);
            ob.elemBytes.push_back(16);
            ob.validBytes.push_back(outValid[c]);
         }
         ob.validity.assign(ob.data.size(), nullptr);
         ob.validityBitOffset.assign(ob.data.size(), 0);
         ob.owned = outOwned;
         ot->numRows = (int64_t) n;
         ot->batches.push_back(std::move(ob));
         ctx->tables.push_back(ot);
         *d->out_table = ot;
      }
   });
}

// ---------------------------------------------------------------- ORDER BY … LIMIT and result gather
int ldb_gpu_table_order_by(LdbTable* t, const char* column, int32_t descending, int64_t limit, int64_t* row_ids, int64_t* n_out, LdbError* err) {
   return guardedP(err, [&] {
      if (!t || !row_ids || !n_out) failP(LDB_ERR_INVALID, "null argument");
      LdbContext* ctx = t->ctx;
      const int c = t->colIndex(column);
      if (c < 0) failP(LDB_ERR_INVALID, "unknown column");
      const int type = t->columns[c].type;
      if (type == LDB_UTF8 || type == LDB_FLOAT32 || type == LDB_FLOAT64 || type == LDB_INT8 || type == LDB_INT16) failP(LDB_ERR_UNSUPPORTED, "ORDER BY column must be int32/date32/char(1)/int64/decimal");
      if (t->batches.size() != 1) failP(LDB_ERR_UNSUPPORTED, "ORDER BY runs over single-batch tables (materialised results, exported groups)");
      LdbBatch& b = t->batches[0];
      const int64_t n = b.nRows;
      if (n >= (int64_t) 1 << 32) failP(LDB_ERR_UNSUPPORTED, "ORDER BY handles up to 2^32 - 1 rows");
      LDB_CUDA(cudaSetDevice(ctx->device));
      ldb_gpu_wait_batch_internal(ctx, &b);
      const size_t rows = (size_t) std::max<int64_t>(n, 1);
      unsigned long long* dk = (unsigned long long*) ctx->stagingAlloc(rows * 8);
      unsigned long long* dk2 = (unsigned long long*) ctx->stagingAlloc(rows * 8);
      uint32_t* dv = (uint32_t*) ctx->stagingAlloc(rows * 4);
      uint32_t* dv2 = (uint32_t*) ctx->stagingAlloc(rows * 4);
      const int ctas = (int) ((n + 4095) / 4096);
      unsigned int* hist = (unsigned int*) ctx->stagingAlloc((size_t) std::max(ctas, 1) * 256 * 4);
      if (n) {
         ctx->launch("radix_sort", [&] {
            launchBuildSortKeys((const uint8_t*) b.data[c], b.elemBytes[c], n, descending, dk, dv, ctx->smCount, ctx->compute);
            launchRadixSortPairs(dk, dv, dk2, dv2, n, hist, ctx->smCount, ctx->compute);
         });
      }
      const int64_t m = std::min<int64_t>(n, limit < 0 ? n : limit);
      std::vector<uint32_t> top((size_t) m);
      if (m) LDB_CUDA(cudaMemcpyAsync(top.data(), dv, (size_t) m * 4, cudaMemcpyDeviceToHost, ctx->compute));
      ctx->syncStream(ctx->compute);
      for (int64_t i = 0; i < m; i++) row_ids[i] = top[(size_t) i];
      *n_out = m;
      for (void* p : {(void*) dk, (void*) dk2, (void*) dv, (void*) dv2, (void*) hist}) ctx->stagingRelease(p);
   });
}
int ldb_gpu_table_gather(LdbTable* t, const char* column, const int64_t* row_ids, int64_t n, void* host_dst, uint8_t* host_valid, LdbError* err) {
   return guardedP(err, [&] {
      if (!t || !row_ids || !host_dst) failP(LDB_ERR_INVALID, "null argument");
      LdbContext* ctx = t->ctx;
      const int c = t->colIndex(column);
      if (c < 0) failP(LDB_ERR_INVALID, "unknown column");
      if (t->columns[c].type == LDB_UTF8) failP(LDB_ERR_UNSUPPORTED, "gather reads fixed-width columns");
      if (t->batches.size() != 1) failP(LDB_ERR_UNSUPPORTED, "gather runs over single-batch tables");
      LdbBatch& b = t->batches[0];
      LDB_CUDA(cudaSetDevice(ctx->device));
      ldb_gpu_wait_batch_internal(ctx, &b);
      const size_t w = (size_t) b.elemBytes[c];
      for (int64_t i = 0; i < n; i++) {
         if (row_ids[i] < 0 || row_ids[i] >= b.nRows) failP(LDB_ERR_INVALID, "row id out of range");
         LDB_CUDA(cudaMemcpyAsync((uint8_t*) host_dst + (size_t) i * w, (const uint8_t*) b.data[c] + (size_t) row_ids[i] * w, w, cudaMemcpyDeviceToHost, ctx->compute));
         if (host_valid) {
            if (c < (int) b.validBytes.size() && b.validBytes[c]) LDB_CUDA(cudaMemcpyAsync(host_valid + i, b.validBytes[c] + row_ids[i], 1, cudaMemcpyDeviceToHost, ctx->compute));
            else host_valid[i] = 1;
         }
      }
      ctx->syncStream(ctx->compute);
   });
}

} // extern "C"
