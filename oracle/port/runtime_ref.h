// oracle/port/runtime_ref.h — TEST INFRASTRUCTURE ONLY (CPU oracle, "reference" flavour).
//
// Binds pipelines.cpp to the reference's OWN runtime objects, compiled verbatim from
// /root/reference/src/runtime/*.cpp into oracle/_ref/ (recipe: oracle/Makefile, target `ref`).
// Nothing from the reference is copied into this repo: the sources are compiled where they lie.
// Objects used: Buffer/FlexibleBuffer, GrowingBuffer, ThreadLocal, SimpleState, HashIndexedView
// (LazyJoinHashtable.cpp), PreAggregationHashtable(+Fragment), EntryLock, ExecutionContext,
// storage/Restrictions, helpers (bloomMasks, tag/untag).  Not buildable and therefore restated in
// the port (scan.h / values.h): the morsel driver of LingoDBTable.cpp, Hash.cpp, the scheduler.
#pragma once
#include "scheduler.h"
#include "table.h"
#include "values.h"

#include "lingodb/runtime/ArrowView.h"
#include "lingodb/runtime/Buffer.h"
#include "lingodb/runtime/EntryLock.h"
#include "lingodb/runtime/ExecutionContext.h"
#include "lingodb/runtime/GrowingBuffer.h"
#include "lingodb/runtime/LazyJoinHashtable.h"
#include "lingodb/runtime/PreAggregationHashtable.h"
#include "lingodb/runtime/DateRuntime.h"
#include "lingodb/runtime/SimpleState.h"
#include "lingodb/runtime/StringRuntime.h"
#include "lingodb/runtime/ThreadLocal.h"
#include "lingodb/runtime/helpers.h"
#include "lingodb/runtime/storage/Restrictions.h"

#include <arrow/type.h>

namespace oracle::refrt {
namespace lr = lingodb::runtime;
using lr::ArrayView;
using lr::BatchView;
using lr::Buffer;
using lr::BufferIterator;
using lr::EntryLock;
using lr::FlexibleBuffer;
using lr::GrowingBuffer;
using lr::GrowingBufferAllocator;
using lr::HashIndexedView;
using lr::PreAggregationHashtable;
using lr::PreAggregationHashtableFragment;
using lr::Restrictions;
using lr::SimpleState;
using lr::ThreadLocal;
using lr::matchesTag;
using lr::untag;

inline uint16_t* defaultSelVec() { return BatchView::defaultSelectionVector.data(); }
inline const uint8_t* allValidBitmap() { return ArrayView::validData.data(); }

inline std::unique_ptr<Restrictions> makeRestrictions(const std::vector<FilterDescription>& descs, const std::vector<ColumnSchema>& schema) {
   arrow::FieldVector fields;
   for (auto& c : schema) {
      std::shared_ptr<arrow::DataType> t;
      switch (c.type) {
         case PhysType::INT32: t = arrow::int32(); break;
         case PhysType::INT64: t = arrow::int64(); break;
         case PhysType::DATE32: t = arrow::date32(); break;
         case PhysType::DECIMAL128: t = arrow::decimal128(c.precision, c.scale); break;
         case PhysType::FSB4: t = arrow::fixed_size_binary(4); break;
         case PhysType::STRING: t = arrow::utf8(); break;
      }
      fields.push_back(arrow::field(c.name, t));
   }
   arrow::Schema arrowSchema(fields);
   std::vector<lr::FilterDescription> fds;
   for (auto& d : descs) {
      lr::FilterDescription fd{};
      fd.columnName = d.columnName;
      fd.columnId = d.columnId;
      fd.op = static_cast<lr::FilterOp>(static_cast<uint8_t>(d.op));
      fd.value = d.value;
      fds.push_back(fd);
   }
   return Restrictions::create(fds, arrowSchema);
}

// The reference's ExecutionContext wants a Session (catalog) it never touches on this path.
struct QueryContextScope {
   alignas(16) unsigned char fakeSession[256];
   std::unique_ptr<lr::ExecutionContext> ctx;
   QueryContextScope() {
      memset(fakeSession, 0, sizeof(fakeSession));
      ctx = std::make_unique<lr::ExecutionContext>(*reinterpret_cast<lr::Session*>(fakeSession));
      lr::setCurrentExecutionContext(ctx.get());
   }
   ~QueryContextScope() {
      ctx.reset();
      lr::setCurrentExecutionContext(nullptr);
   }
};
// currentExecutionContext is thread_local in the reference (ExecutionContext.cpp:48-60; set by
// Task::setup, include/lingodb/scheduler/Tasks.h:10-17): the oracle's own tasks bind it the same way.
struct WorkerContextBinder {
   lr::ExecutionContext* ctx = lr::getCurrentExecutionContext();
   void bind() { lr::setCurrentExecutionContext(ctx); }
   void unbind() {
      if (sched::currentWorkerId() != 0) lr::setCurrentExecutionContext(nullptr);
   }
};
// scalar runtime of Q9: the reference's own functions (DateRuntime.cpp:99-101, StringRuntime.cpp:337-345)
inline int64_t extractYear(int64_t ns) { return lr::DateRuntime::extractYear(ns); }
inline bool constLikeContains(const oracle::VarLen32& str, std::string_view needle) {
   lr::VarLen32 s((const uint8_t*) str.data(), str.len), n((const uint8_t*) needle.data(), (uint32_t) needle.size());
   return lr::StringRuntime::findMatch(s, n, 0, str.len) != 0x8000000000000000ull;
}
constexpr const char* runtimeKind = "reference";

} // namespace oracle::refrt
