// staging.h — compressed, multi-threaded staging of HOST Arrow batches to HBM (staging.cu, pack_host.cpp).
// Internal to libldb_gpu.so.  Replaces the reference's "the table already is in the CPU's memory" with the PCIe hop a GPU
// backend has to pay: LingoDBTable keeps Arrow record batches in host RAM (src/runtime/storage/LingoDBTable.cpp:27-54,
// 294-305) and hands pointers to the scan; here every HOST batch is re-encoded by a pool of independent staging pipelines
// (one host thread + one CUDA stream + two pinned slots each) and decoded on the device behind the copy.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct LdbContext;

namespace ldb {

constexpr int kPackBlockRows = 65536;  // frame-of-reference block: one {min, width} per column per 64 Ki values
constexpr int kMaxPackCols = 16;
constexpr size_t kPackSlotBytes = 8u << 20; // one task's packed bytes (header + data) never exceed this

struct PackBlockHdr { // 16 bytes, rides at the head of the slot with the data
   int64_t min;
   uint32_t offset; // of the block's packed values from the slot start (16-byte aligned)
   uint32_t width;  // 1, 2, 4 or 8 bytes per value
};
struct UnpackOuts {
   void* out[kMaxPackCols];        // decoded destination of each column for the task's FIRST row
   int32_t outBytes[kMaxPackCols]; // 4 (int32/date32/fsb4) or 8 (int64, narrowed decimal128)
};
// CPUs this process may actually burn: min(hardware threads, affinity mask, cgroup CPU quota).  Containers on the 128-thread
// box run under `cpu.max 1600000 100000` = 16 CPUs: 64 packing threads there are throttled to a crawl (measured: 1.55 s per
// SF100 pass vs 0.3 s with 14), so the engine sizes itself by the quota, not by hardware_concurrency().
int effectiveCpus();
void launchUnpack(const uint8_t* devSlot, int nCols, int nBlocks, int rows, const UnpackOuts& outs, cudaStream_t s);

// one HOST batch being staged through the engine
struct PackedBatch {
   struct Col {
      const uint8_t* src; // host values (already offset by ArrayView.offset)
      int32_t srcKind;    // 0 int32 cells, 1 int64 cells, 2 decimal128 cells
      int32_t srcStride;  // 4, 8, 16
      uint8_t* out;       // device destination (decoded layout)
      int32_t outBytes;
   };
   std::vector<Col> cols;
   int64_t nRows = 0;
   std::atomic<int> remaining{0};
   std::mutex m;
   std::condition_variable cv;
   std::string error;
   std::vector<uint8_t> used; // workers that staged a task of this batch: the scan waits for their streams (StagingEngine::events)
};

class StagingEngine {
   struct Task {
      std::shared_ptr<PackedBatch> batch;
      int64_t rowBegin;
      int32_t rows;
   };
   LdbContext* ctx;
   int nPack = 0, nRaw = 0; // workers [0, nPack) pack; workers [nPack, nPack + nRaw) ship raw Arrow cells (no CPU work per value)
   std::vector<std::thread> threads;
   std::mutex m;
   std::condition_variable cv;
   std::deque<Task> queue;
   bool stop = false;
   void workerMain(int w);
   void rawMain(int w);

   public:
   // events[w] is re-recorded on worker w's stream after every task it issues: waiting for it covers every earlier task of
   // that (in-order) stream, in particular all tasks of a batch whose remaining count reached zero
   std::vector<cudaEvent_t> events;
   StagingEngine(LdbContext* ctx, int nPackThreads, int nRawThreads);
   ~StagingEngine();
   int workers() const { return (int) threads.size(); }
   // splits the batch into tasks and returns at once; the batch is complete when remaining == 0
   void submit(const std::shared_ptr<PackedBatch>& b);
   static void wait(PackedBatch& b); // host wait until every task of the batch was issued; rethrows a worker error
};

} // namespace ldb
