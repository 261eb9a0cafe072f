// integration/GPUPipeline.cpp — reference-side shim (would live at src/runtime/GPUPipeline.cpp); see GPUPipeline.h.
#include "GPUPipeline.h"

#include <cstddef>
#include <stdexcept>
#include <type_traits>

namespace lingodb::runtime {

// ---- the ABI facts the C boundary relies on, checked against the reference's own headers at compile time
static_assert(sizeof(LdbArrayView) == sizeof(ArrayView), "LdbArrayView must be layout-identical to lingodb::runtime::ArrayView");
static_assert(offsetof(LdbArrayView, length) == offsetof(ArrayView, length) && offsetof(LdbArrayView, null_count) == offsetof(ArrayView, nullCount) &&
                 offsetof(LdbArrayView, offset) == offsetof(ArrayView, offset) && offsetof(LdbArrayView, n_buffers) == offsetof(ArrayView, nBuffers) &&
                 offsetof(LdbArrayView, n_children) == offsetof(ArrayView, nChildren) && offsetof(LdbArrayView, buffers) == offsetof(ArrayView, buffers) &&
                 offsetof(LdbArrayView, children) == offsetof(ArrayView, children),
              "field offsets of LdbArrayView differ from ArrayView (ArrowView.h:8-21)");
static_assert((int) FilterOp::EQ == LDB_EQ && (int) FilterOp::NEQ == LDB_NEQ && (int) FilterOp::LT == LDB_LT && (int) FilterOp::LTE == LDB_LTE &&
                 (int) FilterOp::GT == LDB_GT && (int) FilterOp::GTE == LDB_GTE && (int) FilterOp::NOTNULL == LDB_NOTNULL && (int) FilterOp::IN == LDB_IN,
              "LdbFilterOp must keep the order of lingodb::runtime::FilterOp (TableStorage.h:14-24)");

namespace {
void check(int rc, const LdbError& e) {
   if (rc != LDB_OK) throw std::runtime_error(e.message); // the reference's convention: exceptions through JIT frames (Hashtable.cpp:106)
}
LdbState* own(LdbState* s) {
   getCurrentExecutionContext()->registerState({s, [](void* p) { ldb_gpu_state_destroy((LdbState*) p); }});
   return s;
}
} // namespace

LdbContext* GPUPipeline::context() {
   static LdbContext* ctx = [] {
      LdbContext* c = nullptr;
      LdbError e;
      check(ldb_gpu_context_create(0, &c, &e), e);
      return c;
   }();
   return ctx;
}
LdbState* GPUPipeline::createSimpleState(int32_t nAggs) {
   LdbState* s = nullptr;
   LdbError e;
   check(ldb_gpu_simple_state_create(context(), nAggs, &s, &e), e);
   return own(s);
}
LdbState* GPUPipeline::createGroupBy(int32_t nKeys, int32_t nAggs, int32_t capacity) {
   LdbState* s = nullptr;
   LdbError e;
   check(ldb_gpu_groupby_create(context(), nKeys, nAggs, capacity, &s, &e), e);
   return own(s);
}
LdbState* GPUPipeline::createJoinTable(int64_t expectedRows, int32_t flags, int32_t nSide, int32_t nAggs) {
   LdbState* s = nullptr;
   LdbError e;
   check(ldb_gpu_join_table_create(context(), expectedRows, flags, nSide, nAggs, &s, &e), e);
   return own(s);
}
LdbState* GPUPipeline::createJoinTablePair(int64_t expectedRows, int32_t flags) {
   LdbState* s = nullptr;
   LdbError e;
   check(ldb_gpu_join_table_create_pair(context(), expectedRows, flags, &s, &e), e);
   return own(s);
}
LdbState* GPUPipeline::createJoinTableDirect(int32_t keyMin, int32_t keyMax) {
   LdbState* s = nullptr;
   LdbError e;
   check(ldb_gpu_join_table_create_direct(context(), keyMin, keyMax, &s, &e), e);
   return own(s);
}
void GPUPipeline::run(const LdbPipelineDesc& desc) {
   LdbError e;
   check(ldb_gpu_run_pipeline(context(), &desc, &e), e);
}
LdbState* GPUPipeline::createHashAggregation(int32_t nKeys, int32_t nAggs, const LdbProgAgg* aggs, int64_t expectedGroups) {
   LdbState* s = nullptr;
   LdbError e;
   check(ldb_gpu_hashagg_create(context(), nKeys, nAggs, aggs, expectedGroups, &s, &e), e);
   return own(s);
}
void GPUPipeline::run(VarLen32 description) {
   const std::string step = description.str();
   LdbError e;
   // plain JSON starts with '{'; anything else is the hex form utility::serializeToHexString produces for DataSource descriptions
   if (!step.empty() && step.front() == '{') check(ldb_gpu_run_step(context(), step.c_str(), &e), e);
   else check(ldb_gpu_run_step_hex(context(), step.c_str(), &e), e);
}
void GPUPipeline::run(const LdbProgramDesc& program) {
   LdbError e;
   check(ldb_gpu_run_program(context(), &program, &e), e);
}
void GPUPipeline::registerState(VarLen32 name, LdbState* state) {
   LdbError e;
   check(ldb_gpu_register_state(context(), name.str().c_str(), state, &e), e);
}
LdbState* GPUPipeline::findState(VarLen32 name) {
   LdbState* s = ldb_gpu_find_state(context(), name.str().c_str());
   if (!s) throw std::runtime_error("no GPU state named " + name.str());
   return s;
}
void GPUPipeline::appendChunk(LdbTable* table, int64_t numRows, const ArrayView* const* columns, size_t nColumns, const int64_t* utf8Bytes) {
   std::vector<LdbArrayView> views(nColumns);
   for (size_t c = 0; c < nColumns; c++) views[c] = *reinterpret_cast<const LdbArrayView*>(columns[c]);
   LdbError e;
   check(ldb_gpu_table_append_batch(table, numRows, views.data(), utf8Bytes, LDB_MEM_HOST, &e), e);
}

LdbFilterDesc toLdbFilter(const FilterDescription& f, std::deque<std::string>& keep) {
   LdbFilterDesc d{};
   keep.push_back(f.columnName);
   d.column = keep.back().c_str();
   d.op = (int32_t) f.op;
   auto scalar = [&](const auto& v, const char** str, int64_t* i) {
      using T = std::decay_t<decltype(v)>;
      if constexpr (std::is_same_v<T, std::string>) {
         keep.push_back(v);
         *str = keep.back().c_str();
      } else if constexpr (std::is_same_v<T, int64_t>) {
         d.value_is_int = 1;
         *i = v;
      } else {
         throw std::runtime_error("floating-point filter constants are not on the GPU path"); // → the step keeps its CPU lowering
      }
   };
   if (f.op == FilterOp::IN) {
      std::visit(
         [&](const auto& vec) {
            if (vec.size() > LDB_MAX_IN_VALUES) throw std::runtime_error("IN list longer than the GPU path supports");
            d.n_values = (int32_t) vec.size();
            for (size_t k = 0; k < vec.size(); k++) scalar(vec[k], &d.str_values[k], &d.int_values[k]);
         },
         f.values);
   } else if (f.op != FilterOp::NOTNULL) {
      std::visit([&](const auto& v) { scalar(v, &d.str_value, &d.int_value); }, f.value);
   }
   return d;
}

} // namespace lingodb::runtime
