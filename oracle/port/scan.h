// oracle/port/scan.h — TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// Restatement of the reference's table-scan morsel driver, which cannot be compiled here
// (src/runtime/storage/LingoDBTable.cpp pulls the MLIR-dependent catalog):
//   TableChunk ArrayViews            LingoDBTable.cpp:200-225
//   ScanBatchesTask (split size, per-worker selection vectors, unitRun)   LingoDBTable.cpp:356-470
//   createScanTask (column ids + Restrictions)                            LingoDBTable.cpp:534-548
//   DataSourceIteration::iterate → callback(BatchView*)                    src/runtime/DataSourceIteration.cpp:90-96
// Per unit: fill BatchView{length, offset, selectionVector, arrays}, run the pushed-down filters
// (Restrictions::applyFilters — the reference's own object in the _ref build), then hand the
// batch to the "JIT'd" scan function.  Work hand-out is a shared cursor instead of the reference's
// reserve-and-steal; the set of units is identical.
#pragma once
#include "rt_select.h"

#include <atomic>
#include <functional>
#include <memory>

namespace oracle {

// ArrayViews of one chunk, all columns (buffers[0] falls back to the shared all-valid bitmap)
struct ChunkViews {
   int64_t numRows;
   std::vector<const void*> bufs;
   std::vector<rt::ArrayView> views;
   explicit ChunkViews(const HostTable& t, const HostChunk& c) : numRows(c.numRows) {
      if (c.numRows > (1 << 20)) throw std::runtime_error("LingoDBTable: too many nulls in column"); // LingoDBTable.cpp:214-216
      bufs = c.buffers;
      for (size_t col = 0; col < t.schema.size(); col++)
         if (!bufs[3 * col]) bufs[3 * col] = rt::allValidBitmap();
      for (size_t col = 0; col < t.schema.size(); col++) {
         rt::ArrayView v{};
         v.length = c.numRows;
         v.nullCount = c.buffers[3 * col] ? 1 : 0; // a caller-provided validity bitmap: "has nulls" (the filters test nullCount != 0, Restrictions.cpp:70,112)
         v.offset = 0;
         v.nBuffers = t.schema[col].type == PhysType::STRING ? 3 : 2;
         v.nChildren = 0;
         v.buffers = &bufs[3 * col];
         v.children = nullptr;
         views.push_back(v);
      }
   }
};

class ScanBatchesTask : public sched::TaskIface {
   std::vector<ChunkViews>& chunks;
   std::vector<size_t> colIds;
   std::unique_ptr<rt::Restrictions> restrictions;
   std::function<void(rt::BatchView*)> cb;
   size_t splitSize = 20000;
   struct Unit {
      size_t chunk, begin;
   };
   std::vector<Unit> units;
   std::atomic<size_t> cursor{0};
   struct PerWorker {
      rt::BatchView view;
      std::vector<const rt::ArrayView*> arrays;
      std::vector<uint16_t> sel1, sel2;
      size_t unit;
   };
   std::vector<PerWorker> workers;
   rt::WorkerContextBinder binder;

   public:
   ScanBatchesTask(size_t tableRows, std::vector<ChunkViews>& chunks, std::vector<size_t> colIds, std::unique_ptr<rt::Restrictions> r, std::function<void(rt::BatchView*)> cb)
      : chunks(chunks), colIds(std::move(colIds)), restrictions(std::move(r)), cb(std::move(cb)) {
      size_t nw = sched::getNumWorkers();
      size_t smaller = std::max<size_t>(1000, tableRows / (nw * 2)); // LingoDBTable.cpp:370-373
      if (smaller < splitSize) splitSize = smaller;
      for (size_t c = 0; c < chunks.size(); c++)
         for (size_t b = 0; b < (size_t) chunks[c].numRows; b += splitSize) units.push_back({c, b});
      workers.resize(nw);
      for (auto& w : workers) {
         w.arrays.resize(this->colIds.size());
         w.view.arrays = w.arrays.data();
         w.sel1.resize(splitSize);
         w.sel2.resize(splitSize);
      }
   }
   void setup() override { binder.bind(); }
   void teardown() override { binder.unbind(); }
   bool allocateWork() override {
      size_t i = cursor.fetch_add(1);
      if (i >= units.size()) return false;
      workers[sched::currentWorkerId()].unit = i;
      return true;
   }
   void performWork() override { // unitRun, LingoDBTable.cpp:382-407
      auto& w = workers[sched::currentWorkerId()];
      auto& u = units[w.unit];
      auto& chunk = chunks[u.chunk];
      size_t len = std::min(u.begin + splitSize, (size_t) chunk.numRows) - u.begin;
      w.view.offset = u.begin;
      w.view.selectionVector = rt::defaultSelVec();
      w.view.length = len;
      for (size_t i = 0; i < colIds.size(); i++) w.view.arrays[i] = &chunk.views[colIds[i]];
      auto [newLen, selVec] = restrictions->applyFilters(u.begin, len, w.sel1.data(), w.sel2.data(), [&](size_t colId) { return &chunk.views[colId]; });
      w.view.length = newLen;
      w.view.selectionVector = selVec;
      if (newLen > 0) cb(&w.view);
   }
};

// rt::DataSource::get + DataSourceIteration::init + iterate rolled into one call
// (SubOpToControlFlow.cpp:1342,1147,1200): scan `columns` of `table` under `filters`.
inline void scanTable(const HostTable& table, const std::vector<std::string>& columns, const std::vector<FilterDescription>& filters, const std::function<void(rt::BatchView*)>& cb) {
   std::vector<ChunkViews> chunks;
   chunks.reserve(table.chunks.size());
   for (auto& c : table.chunks) chunks.emplace_back(table, c);
   std::vector<size_t> colIds;
   for (auto& n : columns) {
      int id = table.colIndex(n);
      if (id < 0) throw std::runtime_error("unknown column " + n);
      colIds.push_back((size_t) id);
   }
   auto restrictions = rt::makeRestrictions(filters, table.schema);
   ScanBatchesTask task((size_t) table.numRows, chunks, colIds, std::move(restrictions), cb);
   sched::runTask(task);
}

// ---- value loads as the generated scan loop performs them (ArrowToStd.cpp:67-85, LowerToStd.cpp:111-209)
struct ColReader {
   const uint8_t* data;  // buffers[1] (values or utf8 offsets)
   const uint8_t* extra; // buffers[2] (utf8 bytes)
   int64_t base;         // ArrayView.offset + BatchView.offset
   ColReader(const rt::BatchView* b, size_t i) {
      const rt::ArrayView* a = b->arrays[i];
      data = (const uint8_t*) a->buffers[1];
      extra = a->nBuffers > 2 ? (const uint8_t*) a->buffers[2] : nullptr;
      base = a->offset + b->offset;
   }
   int32_t i32(int64_t idx) const { int32_t v; memcpy(&v, data + 4 * (base + idx), 4); return v; }
   i128 dec128(int64_t idx) const { return loadDec128(data, base + idx); }
   int64_t dec64(int64_t idx) const { return (int64_t) dec128(idx); } // trunc i128→i64 for precision < 19
   int64_t dateNs(int64_t idx) const { return dateToNs(i32(idx)); }
   VarLen32 str(int64_t idx) const {
      const int32_t* off = (const int32_t*) data + base + idx;
      return VarLen32(extra + off[0], (uint32_t) (off[1] - off[0]));
   }
};

} // namespace oracle
