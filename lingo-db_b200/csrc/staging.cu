// staging.cu — device decode of frame-of-reference packed column blocks + the host staging engine (staging.h).
#include "context.h"
#include "staging.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <sched.h>

extern "C" size_t ldb_pack_block_hinted(const uint8_t* src, int32_t src_kind, int64_t n, uint8_t* dst, int64_t* min_out, int32_t* width_out, int64_t* hint_lo, int64_t* hint_hi); // pack_host.cpp

namespace ldb {

// ---------------------------------------------------------------- device: unpack
// grid = nCols * nBlocks * kSub CTAs; CTA (col, block, sub) decodes kPackBlockRows / kSub values of one block:
// value = min + zero-extended packed word, written as the staged cell (4 or 8 bytes) the pipeline kernels read.
// HBM-bound and tiny: ≈ (packed + decoded) bytes per value, a few percent of the PCIe time it hides behind.
constexpr int kSub = 8;
constexpr int kSubRows = kPackBlockRows / kSub; // 8192
template <int W>
__device__ __forceinline__ void load4(const uint8_t* p, uint64_t (&v)[4]) {
   if constexpr (W == 1) {
      const uint32_t x = *(const uint32_t*) p;
#pragma unroll
      for (int k = 0; k < 4; k++) v[k] = (x >> (8 * k)) & 0xffu;
   } else if constexpr (W == 2) {
      const uint2 x = *(const uint2*) p;
      v[0] = x.x & 0xffffu;
      v[1] = x.x >> 16;
      v[2] = x.y & 0xffffu;
      v[3] = x.y >> 16;
   } else if constexpr (W == 4) {
      const uint4 x = *(const uint4*) p;
      v[0] = x.x;
      v[1] = x.y;
      v[2] = x.z;
      v[3] = x.w;
   } else if constexpr (W == 8) {
      const ulonglong2 a = *(const ulonglong2*) p, b = *(const ulonglong2*) (p + 16);
      v[0] = a.x;
      v[1] = a.y;
      v[2] = b.x;
      v[3] = b.y;
   } else { // W == 16: raw decimal128 cells shipped as they are — keep the low 8 bytes (the JIT's trunc i128 → i64)
#pragma unroll
      for (int k = 0; k < 4; k++) v[k] = ((const ulonglong2*) p)[k].x;
   }
}
template <int W>
__device__ __forceinline__ void unpackRange(const uint8_t* src, int64_t base, int n, uint8_t* out, int outBytes) {
   const int n4 = n / 4;
   for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      uint64_t v[4];
      load4<W>(src + (size_t) i * 4 * W, v);
      if (outBytes == 4) {
         int4 o;
         o.x = (int32_t) (base + (int64_t) v[0]);
         o.y = (int32_t) (base + (int64_t) v[1]);
         o.z = (int32_t) (base + (int64_t) v[2]);
         o.w = (int32_t) (base + (int64_t) v[3]);
         ((int4*) out)[i] = o;
      } else {
         longlong2 a, b;
         a.x = base + (int64_t) v[0];
         a.y = base + (int64_t) v[1];
         b.x = base + (int64_t) v[2];
         b.y = base + (int64_t) v[3];
         ((longlong2*) out)[2 * i] = a;
         ((longlong2*) out)[2 * i + 1] = b;
      }
   }
   for (int i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) { // < 4 tail values
      uint64_t p = 0;
      for (int k = 0; k < (W < 8 ? W : 8); k++) p |= (uint64_t) src[(size_t) i * W + k] << (8 * k);
      if (outBytes == 4) ((int32_t*) out)[i] = (int32_t) (base + (int64_t) p);
      else ((int64_t*) out)[i] = base + (int64_t) p;
   }
}
__global__ void __launch_bounds__(256) unpackKernel(const uint8_t* slot, int nCols, int nBlocks, int rows, const __grid_constant__ UnpackOuts outs) {
   const int sub = blockIdx.x % kSub, block = (blockIdx.x / kSub) % nBlocks, col = blockIdx.x / (kSub * nBlocks);
   const PackBlockHdr h = ((const PackBlockHdr*) slot)[col * nBlocks + block];
   if (h.width == 0) return; // the raw path copied this column straight to its destination
   const int blockRows = min(kPackBlockRows, rows - block * kPackBlockRows);
   const int begin = sub * kSubRows;
   if (begin >= blockRows) return;
   const int n = min(kSubRows, blockRows - begin);
   const int ob = outs.outBytes[col];
   const uint8_t* src = slot + h.offset + (size_t) begin * h.width;
   uint8_t* out = (uint8_t*) outs.out[col] + ((size_t) block * kPackBlockRows + begin) * ob;
   switch (h.width) {
      case 1: unpackRange<1>(src, h.min, n, out, ob); break;
      case 2: unpackRange<2>(src, h.min, n, out, ob); break;
      case 4: unpackRange<4>(src, h.min, n, out, ob); break;
      case 8: unpackRange<8>(src, h.min, n, out, ob); break;
      default: unpackRange<16>(src, h.min, n, out, ob); break;
   }
}
void launchUnpack(const uint8_t* devSlot, int nCols, int nBlocks, int rows, const UnpackOuts& outs, cudaStream_t s) {
   unpackKernel<<<nCols * nBlocks * kSub, 256, 0, s>>>(devSlot, nCols, nBlocks, rows, outs);
}

// ---------------------------------------------------------------- host: staging engine
int effectiveCpus() {
   int n = (int) std::max(1u, std::thread::hardware_concurrency());
   cpu_set_t set;
   if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::min(n, std::max(1, CPU_COUNT(&set)));
   if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) { // cgroup v2: "<quota|max> <period>"
      char quota[32];
      long period = 0;
      if (fscanf(f, "%31s %ld", quota, &period) == 2 && strcmp(quota, "max") != 0 && period > 0) n = std::min(n, (int) std::max(1L, (atol(quota) + period - 1) / period));
      fclose(f);
   } else {
      long quota = -1, period = 0;
      if (FILE* q = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {
         if (fscanf(q, "%ld", &quota) != 1) quota = -1;
         fclose(q);
      }
      if (FILE* p = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) {
         if (fscanf(p, "%ld", &period) != 1) period = 0;
         fclose(p);
      }
      if (quota > 0 && period > 0) n = std::min(n, (int) std::max(1L, (quota + period - 1) / period));
   }
   return n;
}
StagingEngine::StagingEngine(LdbContext* c, int nPackThreads, int nRawThreads) : ctx(c), nPack(nPackThreads), nRaw(nRawThreads) {
   events.assign(nPack + nRaw, nullptr);
   for (auto& e : events) LDB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
   for (int i = 0; i < nPack; i++) threads.emplace_back([this, i] { workerMain(i); });
   for (int i = 0; i < nRaw; i++) threads.emplace_back([this, i] { rawMain(nPack + i); });
}
StagingEngine::~StagingEngine() {
   {
      std::unique_lock<std::mutex> l(m);
      stop = true;
   }
   cv.notify_all();
   for (auto& t : threads) t.join();
   for (auto e : events) cudaEventDestroy(e);
}
void StagingEngine::submit(const std::shared_ptr<PackedBatch>& b) {
   // rows per task: whole 64 Ki-row blocks, as many as the worst case (every column at full width) fits a slot
   size_t fullWidth = 0;
   for (auto& c : b->cols) fullWidth += (size_t) (c.srcKind == 0 ? 4 : 8);
   const size_t perBlock = fullWidth * kPackBlockRows + b->cols.size() * (sizeof(PackBlockHdr) + 16);
   const int blocksPerTask = (int) std::max<size_t>(1, (kPackSlotBytes - 256) / perBlock);
   const int64_t rowsPerTask = (int64_t) blocksPerTask * kPackBlockRows;
   const int nTasks = (int) ((b->nRows + rowsPerTask - 1) / rowsPerTask);
   b->remaining.store(nTasks);
   b->used.assign(events.size(), 0);
   {
      std::unique_lock<std::mutex> l(m);
      for (int t = 0; t < nTasks; t++) {
         const int64_t r0 = (int64_t) t * rowsPerTask;
         queue.push_back(Task{b, r0, (int32_t) std::min<int64_t>(rowsPerTask, b->nRows - r0)});
      }
   }
   cv.notify_all();
}
void StagingEngine::wait(PackedBatch& b) {
   std::unique_lock<std::mutex> l(b.m);
   b.cv.wait(l, [&] { return b.remaining.load() == 0; });
   if (!b.error.empty()) throw ApiError(LDB_ERR_CUDA, "staging worker failed: " + b.error);
}
void StagingEngine::workerMain(int w) {
   struct Slot {
      uint8_t* host = nullptr;
      uint8_t* dev = nullptr;
      cudaEvent_t done = nullptr;
      bool inFlight = false;
   } slots[2];
   cudaStream_t stream = nullptr;
   bool ready = false;
   uint64_t seenGen = 0;
   int next = 0;
   int64_t hintLo[kMaxPackCols], hintHi[kMaxPackCols]; // value range of the last block this worker packed, per column position
   for (int c = 0; c < kMaxPackCols; c++) hintLo[c] = 1, hintHi[c] = 0;
   auto init = [&] {
      LDB_CUDA(cudaSetDevice(ctx->device));
      LDB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      for (auto& s : slots) {
         LDB_CUDA(cudaMallocHost(&s.host, kPackSlotBytes));
         LDB_CUDA(cudaMalloc(&s.dev, kPackSlotBytes));
         LDB_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventBlockingSync | cudaEventDisableTiming)); // waiters sleep
      }
      ready = true;
   };
   while (true) {
      Task t;
      {
         std::unique_lock<std::mutex> l(m);
         cv.wait(l, [&] { return stop || !queue.empty(); });
         if (queue.empty()) break; // stop && drained
         t = std::move(queue.front());
         queue.pop_front();
      }
      PackedBatch& b = *t.batch;
      try {
         if (!ready) init();
         Slot& s = slots[next++ & 1];
         if (s.inFlight) LDB_CUDA(cudaEventSynchronize(s.done));
         // staging buffers released by ldb_gpu_table_clear may still be read by queued pipeline kernels
         const uint64_t gen = ctx->stagingGen.load();
         if (gen != seenGen) {
            LDB_CUDA(cudaStreamWaitEvent(stream, ctx->computeDone, 0));
            seenGen = gen;
         }
         const int nCols = (int) b.cols.size();
         const int nBlocks = (t.rows + kPackBlockRows - 1) / kPackBlockRows;
         PackBlockHdr* hdr = (PackBlockHdr*) s.host;
         size_t off = ((size_t) nCols * nBlocks * sizeof(PackBlockHdr) + 255) & ~size_t(255);
         UnpackOuts outs{};
         for (int c = 0; c < nCols; c++) {
            const auto& col = b.cols[c];
            outs.out[c] = col.out + (size_t) t.rowBegin * col.outBytes;
            outs.outBytes[c] = col.outBytes;
            for (int k = 0; k < nBlocks; k++) {
               const int64_t r0 = t.rowBegin + (int64_t) k * kPackBlockRows;
               const int64_t n = std::min<int64_t>(kPackBlockRows, t.rowBegin + t.rows - r0);
               int64_t mn;
               int32_t width;
               const size_t bytes = ldb_pack_block_hinted(col.src + (size_t) r0 * col.srcStride, col.srcKind, n, s.host + off, &mn, &width, &hintLo[c], &hintHi[c]);
               hdr[c * nBlocks + k] = PackBlockHdr{mn, (uint32_t) off, (uint32_t) width};
               off = (off + bytes + 15) & ~size_t(15);
            }
         }
         LDB_CUDA(cudaMemcpyAsync(s.dev, s.host, off, cudaMemcpyHostToDevice, stream));
         launchUnpack(s.dev, nCols, nBlocks, t.rows, outs, stream);
         LDB_CUDA(cudaGetLastError());
         LDB_CUDA(cudaEventRecord(s.done, stream));
         s.inFlight = true;
         LDB_CUDA(cudaEventRecord(events[w], stream));
         b.used[w] = 1;
         ctx->h2dBytes.fetch_add((int64_t) off);
         ctx->stagingLaunches.fetch_add(1);
      } catch (const std::exception& e) {
         std::unique_lock<std::mutex> l(b.m);
         if (b.error.empty()) b.error = e.what();
      }
      if (b.remaining.fetch_sub(1) == 1) {
         std::unique_lock<std::mutex> l(b.m);
         b.cv.notify_all();
      }
   }
   if (ready) {
      cudaSetDevice(ctx->device);
      cudaStreamSynchronize(stream);
      for (auto& s : slots) {
         cudaFreeHost(s.host);
         cudaFree(s.dev);
         cudaEventDestroy(s.done);
      }
      cudaStreamDestroy(stream);
   }
}

// A raw copier ships Arrow cells as they are, straight from the caller's (pinned) buffers — no CPU work per value, the copy
// engine does the reading: int32/int64 columns land in their destination, decimal128 cells (16 B) go to a device slot and the
// unpack kernel keeps their low 8 bytes.  It takes tasks from the BACK of the queue while the packers take them from the front,
// and only when one of its two slots is free, so the split between "76 B/row, no CPU" and "≈12 B/row, CPU-bound" balances
// itself: with one PCIe link and few CPUs the packers carry most rows, with idle links (8 GPUs share 16 CPUs) the copiers do.
void StagingEngine::rawMain(int w) {
   constexpr size_t kRawSlotBytes = 3 * kPackSlotBytes; // 16 B per decimal cell instead of <= 8
   struct Slot {
      uint8_t* host = nullptr; // header only
      uint8_t* dev = nullptr;
      cudaEvent_t done = nullptr;
      bool inFlight = false;
   } slots[2];
   cudaStream_t stream = nullptr;
   bool ready = false;
   uint64_t seenGen = 0;
   int next = 0;
   auto init = [&] {
      LDB_CUDA(cudaSetDevice(ctx->device));
      LDB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      for (auto& s : slots) {
         LDB_CUDA(cudaMallocHost(&s.host, 65536));
         LDB_CUDA(cudaMalloc(&s.dev, kRawSlotBytes));
         LDB_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventBlockingSync | cudaEventDisableTiming)); // waiters sleep
      }
      ready = true;
   };
   while (true) {
      try {
         if (!ready) init();
      } catch (const std::exception&) {
         return; // no device memory for raw slots: the packers carry everything
      }
      Slot& s = slots[next & 1];
      if (s.inFlight) { // wait for a free slot BEFORE taking work: an idle link is the copier's cue
         if (cudaEventSynchronize(s.done) != cudaSuccess) return;
         s.inFlight = false;
      }
      Task t;
      {
         std::unique_lock<std::mutex> l(m);
         cv.wait(l, [&] { return stop || !queue.empty(); });
         if (queue.empty()) break;
         t = std::move(queue.back());
         queue.pop_back();
      }
      next++;
      PackedBatch& b = *t.batch;
      try {
         const uint64_t gen = ctx->stagingGen.load();
         if (gen != seenGen) {
            LDB_CUDA(cudaStreamWaitEvent(stream, ctx->computeDone, 0));
            seenGen = gen;
         }
         const int nCols = (int) b.cols.size();
         const int nBlocks = (t.rows + kPackBlockRows - 1) / kPackBlockRows;
         PackBlockHdr* hdr = (PackBlockHdr*) s.host;
         size_t off = ((size_t) nCols * nBlocks * sizeof(PackBlockHdr) + 255) & ~size_t(255);
         if (off > 65536) throw ApiError(LDB_ERR_UNSUPPORTED, "raw staging header too large");
         UnpackOuts outs{};
         int64_t bytes = 0;
         bool needKernel = false;
         for (int c = 0; c < nCols; c++) {
            const auto& col = b.cols[c];
            outs.out[c] = col.out + (size_t) t.rowBegin * col.outBytes;
            outs.outBytes[c] = col.outBytes;
            const uint8_t* src = col.src + (size_t) t.rowBegin * col.srcStride;
            if (col.srcKind != 2) { // cells already have their staged width: copy into place
               LDB_CUDA(cudaMemcpyAsync(outs.out[c], src, (size_t) t.rows * col.srcStride, cudaMemcpyHostToDevice, stream));
               for (int k = 0; k < nBlocks; k++) hdr[c * nBlocks + k] = PackBlockHdr{0, 0, 0};
            } else {
               if (off + (size_t) t.rows * 16 > kRawSlotBytes) throw ApiError(LDB_ERR_UNSUPPORTED, "raw staging slot too small");
               LDB_CUDA(cudaMemcpyAsync(s.dev + off, src, (size_t) t.rows * 16, cudaMemcpyHostToDevice, stream));
               for (int k = 0; k < nBlocks; k++) hdr[c * nBlocks + k] = PackBlockHdr{0, (uint32_t) (off + (size_t) k * kPackBlockRows * 16), 16};
               off += ((size_t) t.rows * 16 + 255) & ~size_t(255);
               needKernel = true;
            }
            bytes += (int64_t) t.rows * col.srcStride;
         }
         if (needKernel) {
            const size_t hdrBytes = (size_t) nCols * nBlocks * sizeof(PackBlockHdr);
            LDB_CUDA(cudaMemcpyAsync(s.dev, s.host, hdrBytes, cudaMemcpyHostToDevice, stream));
            launchUnpack(s.dev, nCols, nBlocks, t.rows, outs, stream);
            LDB_CUDA(cudaGetLastError());
            ctx->stagingLaunches.fetch_add(1);
            bytes += (int64_t) hdrBytes;
         }
         LDB_CUDA(cudaEventRecord(s.done, stream));
         s.inFlight = true;
         LDB_CUDA(cudaEventRecord(events[w], stream));
         b.used[w] = 1;
         ctx->h2dBytes.fetch_add(bytes);
         ctx->rawStagedRows.fetch_add(t.rows);
      } catch (const std::exception& e) {
         std::unique_lock<std::mutex> l(b.m);
         if (b.error.empty()) b.error = e.what();
      }
      if (b.remaining.fetch_sub(1) == 1) {
         std::unique_lock<std::mutex> l(b.m);
         b.cv.notify_all();
      }
   }
   if (ready) {
      cudaSetDevice(ctx->device);
      cudaStreamSynchronize(stream);
      for (auto& s : slots) {
         cudaFreeHost(s.host);
         cudaFree(s.dev);
         cudaEventDestroy(s.done);
      }
      cudaStreamDestroy(stream);
   }
}

} // namespace ldb
