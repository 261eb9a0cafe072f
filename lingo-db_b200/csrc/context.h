// context.h — internal definitions of the opaque C-ABI handles (include/ldb_gpu.h).
#pragma once
#include "../../include/ldb_gpu.h"
#include "kernels.h"
#include "program.h"

#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <string>
#include <vector>

namespace ldb {

struct CudaError : std::runtime_error {
   int code;
   CudaError(int code, const std::string& m) : std::runtime_error(m), code(code) {}
};
inline void cudaCheck(cudaError_t e, const char* what, const char* file, int line) {
   if (e != cudaSuccess) {
      cudaGetLastError();
      throw CudaError(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? LDB_ERR_NO_DEVICE : LDB_ERR_CUDA,
                      std::string(what) + ": " + cudaGetErrorString(e) + " (" + file + ":" + std::to_string(line) + ")");
   }
}
#define LDB_CUDA(x) ::ldb::cudaCheck((x), #x, __FILE__, __LINE__)
struct ApiError : std::runtime_error {
   int code;
   ApiError(int code, const std::string& m) : std::runtime_error(m), code(code) {}
};

// Host worker pool used while staging HOST batches (narrowing decimal128 → the 8 bytes the kernels read).
class HostPool {
   std::vector<std::thread> threads;
   std::mutex m;
   std::condition_variable cvStart, cvDone;
   std::function<void(int, int)> job; // (worker, nWorkers)
   uint64_t generation = 0;
   int running = 0;
   bool stop = false;
   void main(int id);

   public:
   explicit HostPool(int n);
   ~HostPool();
   int size() const { return (int) threads.size() + 1; }
   void run(const std::function<void(int, int)>& fn); // caller is worker 0
};
struct PinnedSlot {
   void* host = nullptr;
   cudaEvent_t done = nullptr;
   bool inFlight = false;
};

struct KernelFamilyTimer {
   std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
   double totalMs = 0;
   int64_t launches = 0;
};

class StagingEngine;
struct PackedBatch;

} // namespace ldb

struct LdbState;
struct LdbTable;

struct LdbGraph;
struct LdbContext {
   LdbGraph* capturing = nullptr; // non-null between ldb_gpu_graph_begin and _end: launches are recorded, not run
   int device = 0;
   int smCount = 0;
   cudaDeviceProp prop{};
   cudaStream_t compute = nullptr, copy = nullptr;
   cudaEvent_t timerStart = nullptr, timerStop = nullptr, computeDone = nullptr;
   int64_t launches = 0;
   bool timing = false;
   std::map<std::string, ldb::KernelFamilyTimer> timers;
   std::vector<cudaEvent_t> eventPool;
   std::vector<LdbState*> states;
   std::vector<LdbTable*> tables;
   std::vector<LdbGraph*> graphs;
   std::map<std::string, LdbState*> namedStates; // states created / registered through serialised steps (step_json.cpp)
   // staging pool for HOST batches: size → free device buffers
   std::multimap<size_t, void*> stagingFree;
   std::map<void*, size_t> stagingSize;
   // narrow staging: decimal128(p<19) HOST columns cross PCIe as 8 bytes/value (the JIT truncates them to i64 anyway)
   bool narrowStaging = true;
   std::unique_ptr<ldb::HostPool> pool;
   static constexpr size_t kPinnedSlotBytes = 32u << 20;
   std::vector<ldb::PinnedSlot> pinned;
   size_t nextPinned = 0;
   std::atomic<int64_t> h2dBytes{0}; // bytes this context copied host→device while staging tables
   // compressed staging (staging.h): HOST batches of >= one block are re-encoded by a pool of independent pipelines
   bool packedStaging = true;
   std::shared_ptr<ldb::StagingEngine> staging;
   std::atomic<uint64_t> stagingGen{0};      // bumped by ldb_gpu_table_clear: workers order their next write after computeDone
   std::atomic<int64_t> stagingLaunches{0};  // unpack kernels launched by the staging workers
   std::atomic<int64_t> rawStagedRows{0};    // rows the raw copiers shipped uncompressed (the rest was packed)

   // Host waits SLEEP instead of spinning (cudaEventBlockingSync): the container's CPU quota is shared with the staging
   // threads and, on a multi-GPU box, with the other ranks — a spinning waiter would burn a whole CPU of it.
   // Short waits (a query's result read: the stream drains within a few hundred microseconds) poll, long ones sleep.
   cudaEvent_t blockingEv = nullptr;
   void syncStream(cudaStream_t s) {
      if (!blockingEv) LDB_CUDA(cudaEventCreateWithFlags(&blockingEv, cudaEventBlockingSync | cudaEventDisableTiming));
      LDB_CUDA(cudaEventRecord(blockingEv, s));
      for (int spin = 0; spin < 20000; spin++) { // ~0.3 ms of polling
         cudaError_t q = cudaEventQuery(blockingEv);
         if (q == cudaSuccess) return;
         if (q != cudaErrorNotReady) LDB_CUDA(q);
      }
      LDB_CUDA(cudaEventSynchronize(blockingEv));
   }
   // pinned scratch for small result reads: a copy into pageable memory would be staged by the driver and serialise with the host
   void* pinnedScratch = nullptr;
   static constexpr size_t kPinnedScratchBytes = 512u << 10;
   void* scratch() {
      if (!pinnedScratch) LDB_CUDA(cudaMallocHost(&pinnedScratch, kPinnedScratchBytes));
      return pinnedScratch;
   }
   void launchCaptured(const char* family, const std::function<void()>& fn); // runtime.cpp
   void* stagingAlloc(size_t bytes);
   void stagingRelease(void* p);
   cudaEvent_t getEvent();
   // wrap one kernel launch: counts it and, when timing is on, brackets it with events on `compute`
   template <class Fn>
   void launch(const char* family, const Fn& fn) {
      if (capturing) {
         launchCaptured(family, [&] { fn(); });
         LDB_CUDA(cudaGetLastError());
         return;
      }
      launches++;
      if (timing) {
         auto& t = timers[family];
         cudaEvent_t a = getEvent(), b = getEvent();
         LDB_CUDA(cudaEventRecord(a, compute));
         fn();
         LDB_CUDA(cudaEventRecord(b, compute));
         t.pending.push_back({a, b});
         t.launches++;
      } else {
         fn();
      }
      LDB_CUDA(cudaGetLastError());
   }
};

struct LdbBatch {
   int64_t nRows = 0;
   std::vector<const void*> data;  // per column: values / utf8 offsets (device)
   std::vector<const void*> bytes; // per column: utf8 bytes (device) or null
   std::vector<int32_t> elemBytes; // per column: bytes per value as staged (decimal128: 16, or 8 when narrowed)
   std::vector<const void*> validity;      // per column: Arrow validity bitmap on the device (null = no nulls in this batch)
   std::vector<int64_t> validityBitOffset; // bit index of row 0 inside the bitmap
   std::vector<uint8_t*> validBytes;       // per column: one validity byte per row (tables this library produced), else empty / null
   std::vector<void*> owned;       // staging buffers to give back on clear
   cudaEvent_t ready = nullptr;    // H2D of this batch finished (null for borrowed device batches)
   std::shared_ptr<ldb::PackedBatch> packed; // columns staged through the compressed staging engine (host wait + worker events)
};
struct LdbColumn {
   std::string name;
   int32_t type, precision, scale;
};
struct LdbTable {
   LdbContext* ctx;
   std::string name;
   std::vector<LdbColumn> columns;
   std::vector<LdbBatch> batches;
   int64_t numRows = 0;
   // column statistics (min, max) of int32/date32 columns, computed on first use and valid while the table has `rows` rows: the planner
   // asks for them on every query (direct-address join tables), the table does not change between two queries
   struct ColumnRange {
      int64_t rows;
      int32_t lo, hi;
   };
   std::map<int, ColumnRange> ranges;
   int colIndex(const char* n) const {
      if (!n) return -1;
      for (size_t i = 0; i < columns.size(); i++)
         if (columns[i].name == n) return (int) i;
      return -1;
   }
};

// A captured query (CUDA graph of everything enqueued on the compute stream between _begin and _end): one cudaGraphLaunch
// replays memsets, pipeline kernels and peer collectives without per-launch host work.
struct LdbGraph {
   LdbContext* ctx = nullptr;
   cudaGraph_t graph = nullptr;
   cudaGraphExec_t exec = nullptr;
   int64_t kernelsPerLaunch = 0;
   struct Timer {
      std::string family;
      cudaEvent_t a, b; // recorded by the graph itself (external event-record nodes)
   };
   std::vector<Timer> timers;
   bool pendingTimes = false; // the last launch's events were not harvested yet
   std::vector<std::function<void()>> onLaunch; // host-side bookkeeping per replay (peer epoch mirrors)
};

// order the compute stream after the staging of one batch (runtime.cpp)
void ldb_gpu_wait_batch_internal(LdbContext* ctx, const struct LdbBatch* b);

struct LdbState {
   LdbContext* ctx;
   int32_t kind;
   ldb::GroupTableDev group{}; // SIMPLE / GROUPBY
   ldb::JoinTableDev join{};   // JOIN_TABLE
   ldb::HashAggDev hashagg{};  // HASHAGG
   int32_t aggKinds[ldb::kProgMaxAggs] = {};
   int32_t nSide = 0, nAggs = 0;
   bool selfTimed = false; // created inside a captured query: its scan kernel's self-measured time is harvested at read
   uint32_t is64Mask = 0; // aggregates that are 64-bit sums (COL / ONE): normalised to a sign-extended i64 on read
   std::vector<void*> allocations;
};
