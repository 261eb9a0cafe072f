// integration/GPUPipeline.h — the REFERENCE-SIDE binding of the GPU backend (would live at
// include/lingodb/runtime/GPUPipeline.h of lingo-db/lingo-db and be listed for runtime-header-tool next to
// Hashtable.h, CMakeLists.txt:164-190).  Not part of the product library: it is the shim a maintainer adds, kept here
// so that it is COMPILED against the reference's own headers (tests/test_integration_shim.py) instead of being prose.
#pragma once
#include "ldb_gpu.h"

#include "lingodb/runtime/ArrowView.h"
#include "lingodb/runtime/ExecutionContext.h"
#include "lingodb/runtime/helpers.h"
#include "lingodb/runtime/storage/TableStorage.h"

#include <deque>
#include <string>
#include <vector>

namespace lingodb::runtime {
class GPUPipeline {
   public:
   // state objects behind the reference's names; every handle is registered with the current ExecutionContext
   // (ExecutionContext.h:111-113) and dies with the query, exactly like rt::Hashtable / rt::GrowingBuffer objects
   static LdbState* createSimpleState(int32_t nAggs);                                              // rt::SimpleState
   static LdbState* createGroupBy(int32_t nKeys, int32_t nAggs, int32_t capacity);                   // rt::PreAggregationHashtable
   static LdbState* createJoinTable(int64_t expectedRows, int32_t flags, int32_t nSide, int32_t nAggs); // rt::GrowingBuffer + rt::HashIndexedView
   static LdbState* createJoinTablePair(int64_t expectedRows, int32_t flags);                        // composite (i32, i32) key
   static LdbState* createJoinTableDirect(int32_t keyMin, int32_t keyMax);                           // dense surrogate keys: slot = key - min
   static LdbState* createHashAggregation(int32_t nKeys, int32_t nAggs, const LdbProgAgg* aggs, int64_t expectedGroups);  // rt::Hashtable / large-domain rt::PreAggregationHashtable
   // one execution step (scan → pushed-down filters → probes → sink) instead of DataSourceIteration::iterate(scan_func)
   static void run(const LdbPipelineDesc& desc);
   // the same step SERIALISED, the way DataSource::get receives its description (DataSourceIteration.cpp:57-88: a hex string the lowering
   // baked into the module as a VarLen32 constant): JSON, plain or hex-encoded, naming tables and states (include/ldb_gpu.h "serialised steps")
   static void run(VarLen32 description);
   // a register program (expression evaluation, NULLs, hash aggregation over any number of groups, semi/anti/mark/outer probes) for the steps
   // the specialised pipelines do not cover — the GPU stand-in for the JIT'd scan body (SubOpToControlFlow.cpp:1123-1203)
   static void run(const LdbProgramDesc& program);
   // states by name: a step's sink under the name later steps read it by
   static void registerState(VarLen32 name, LdbState* state);
   static LdbState* findState(VarLen32 name);
   // TableChunk::getArrayView() results (LingoDBTable.cpp:200-225) go through unchanged: LdbArrayView IS ArrayView
   static void appendChunk(LdbTable* table, int64_t numRows, const ArrayView* const* columns, size_t nColumns, const int64_t* utf8Bytes);
   static LdbContext* context(); // one per process and device; owned by the Session in a full integration
};

// FilterDescription (TableStorage.h:25-31) → LdbFilterDesc; `keep` owns the strings the descriptor points to
LdbFilterDesc toLdbFilter(const FilterDescription& f, std::deque<std::string>& keep);
} // namespace lingodb::runtime
