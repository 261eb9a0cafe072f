// peer.cu — multi-GPU exchange over peer-mapped HBM (NVLink 5 / NVSwitch), one process per GPU.
//
// No reference counterpart: the reference is single-process (SURVEY §2 "Parallelism strategies").  Every rank owns a
// SYMMETRIC HEAP (one cudaMalloc, same layout everywhere) that its peers map through CUDA IPC; a transfer is a kernel
// that STORES into the peer's heap over NVLink and then publishes a flag (fence.sys + st.release.sys), the receiver's
// kernel spins on the flag in its own memory (ld.acquire.sys).  No NCCL call, no host round trip, no proxy thread on the
// data path — the partial aggregates of Q1/Q6/Q9 (≈9–140 KB) cost one small kernel instead of export + all-gather + merge,
// and the repartition step of a join writes its tuples straight into the receiver's buffer from the partition kernel.
//
//   heap: | barrier flags [world] | gather flags [2][world] | mailbox [2][world][kSlotBytes] | user region … |
//
// Flags are monotonic epoch counters, so nothing is ever reset; the mailbox is double-buffered by epoch parity: a writer
// can be at most one collective ahead of any reader (its next collective waits for that reader's push), so parity p is
// never overwritten while a peer still reads it.
#include "context.h"
#include "device_utils.cuh"
#include "peer.h"

#include <algorithm>
#include <cstring>

namespace ldb {

// ---------------------------------------------------------------- device side
__device__ __forceinline__ void stReleaseSys(unsigned long long* p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long* p) {
   unsigned long long v;
   asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ unsigned long long globalTimerNs() {
   unsigned long long t;
   asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
   return t;
}
// Spin until *flag >= epoch.  Bounded (a peer that died must not hang this GPU): on timeout the comm's error word is set
// and the caller's data is garbage — the host reports LDB_ERR_CUDA at the next check.
__device__ __forceinline__ bool waitFlag(const unsigned long long* flag, unsigned long long epoch, int32_t* error, unsigned long long timeoutNs) {
   const unsigned long long t0 = globalTimerNs();
   unsigned spins = 0;
   while (ldAcquireSys(flag) < epoch) {
      if ((++spins & 1023u) == 0) {
         if (globalTimerNs() - t0 > timeoutNs) {
            atomicExch(error, 1);
            return false;
         }
         __nanosleep(200);
      }
   }
   return true;
}

// all-to-all "I am here": CTA p tells peer p and waits for peer p
enum EpochKind { EPOCH_BARRIER = 0, EPOCH_GATHER = 1, EPOCH_MERGE = 2 };
__device__ __forceinline__ unsigned long long nextEpoch(const PeerView& v, int kind) { return ((const volatile unsigned long long*) (v.heap[v.rank] + kEpochsOff))[kind] + 1; }
// runs behind every collective kernel (stream order): the epoch a kernel reads is stable for all its CTAs
__global__ void peerBumpKernel(PeerView v, int kindA, int kindB) {
   unsigned long long* e = (unsigned long long*) (v.heap[v.rank] + kEpochsOff);
   e[kindA]++;
   if (kindB >= 0) e[kindB]++;
}
__global__ void peerBarrierKernel(PeerView v) {
   const unsigned long long epoch = nextEpoch(v, EPOCH_BARRIER);
   const int p = blockIdx.x;
   if (p == v.rank || threadIdx.x != 0) return;
   __threadfence_system(); // everything this GPU wrote into peer memory before the barrier (earlier kernels of the stream included)
   stReleaseSys((unsigned long long*) (v.heap[p] + kBarrierFlagsOff) + v.rank, epoch);
   waitFlag((const unsigned long long*) (v.heap[v.rank] + kBarrierFlagsOff) + p, epoch, v.error, v.timeoutNs);
}

// all-gather of one small block (<= kSlotBytes, multiple of 16): CTA p copies `src` into peer p's mailbox slot [parity][rank]
// with 128-bit stores, publishes the flag, then waits for peer p's block.  Afterwards mailbox[parity][*] of the own heap is
// complete (the own slot is filled locally by CTA `rank`).
__device__ __forceinline__ void pushBlock(const PeerView& v, int p, const uint8_t* src, size_t bytes, unsigned long long epoch) {
   const int parity = (int) (epoch & 1);
   uint8_t* dst = v.heap[p] + kMailboxOff + ((size_t) parity * v.world + v.rank) * kSlotBytes;
   const int4* s4 = (const int4*) src;
   int4* d4 = (int4*) dst;
   for (size_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) d4[i] = s4[i];
   __syncthreads();
   if (threadIdx.x == 0 && p != v.rank) {
      __threadfence_system();
      stReleaseSys((unsigned long long*) (v.heap[p] + kGatherFlagsOff) + (size_t) parity * v.world + v.rank, epoch);
   }
}
__device__ __forceinline__ bool waitBlock(const PeerView& v, int p, unsigned long long epoch) {
   const int parity = (int) (epoch & 1);
   __shared__ int ok;
   if (threadIdx.x == 0) ok = p == v.rank ? 1 : (waitFlag((const unsigned long long*) (v.heap[v.rank] + kGatherFlagsOff) + (size_t) parity * v.world + p, epoch, v.error, v.timeoutNs) ? 1 : 0);
   __syncthreads();
   return ok != 0;
}
__global__ void __launch_bounds__(256) peerAllGatherKernel(PeerView v, const uint8_t* src, size_t bytes) {
   const unsigned long long epoch = nextEpoch(v, EPOCH_GATHER);
   const int p = blockIdx.x;
   pushBlock(v, p, src, bytes, epoch);
   waitBlock(v, p, epoch);
}

// K7 over NVLink: all-gather of the group-table image fused with the merge.  CTA p pushes this rank's table (the table IS
// the image: state | keys | acc) to peer p, waits for peer p's image and folds it into the local table with the same
// lookup-or-insert + two-word atomic adds the single-GPU flush uses (rt::PreAggregationHashtable::merge semantics: sums per key).
__device__ int peerGroupLookupOrInsert(const GroupTableDev& t, const int32_t* k); // kernels.cu twin, defined below
__global__ void __launch_bounds__(256) peerGroupAllMergeKernel(PeerView v, GroupTableDev t, size_t imageBytes) {
   const unsigned long long epoch = nextEpoch(v, EPOCH_GATHER);
   const unsigned long long localTarget = nextEpoch(v, EPOCH_MERGE) * (unsigned long long) (v.world - 1);
   const int p = blockIdx.x;
   if (p == v.rank) return; // the own table is merged into, not from
   pushBlock(v, p, (const uint8_t*) t.state, imageBytes, epoch);
   // every CTA must have READ the table (its push) before any CTA starts folding a peer's image INTO it: rendezvous of the
   // world-1 co-resident CTAs on a monotonic counter in this GPU's own memory
   if (threadIdx.x == 0) {
      unsigned long long* sync = (unsigned long long*) (v.heap[v.rank] + kLocalSyncOff);
      __threadfence();
      atomicAdd(sync, 1ull);
      while (*((volatile unsigned long long*) sync) < localTarget) {}
   }
   __syncthreads();
   if (!waitBlock(v, p, epoch)) return;
   const int parity = (int) (epoch & 1);
   const uint8_t* img = v.heap[v.rank] + kMailboxOff + ((size_t) parity * v.world + p) * kSlotBytes;
   const size_t cap = (size_t) t.capacity;
   const int32_t* st = (const int32_t*) img;
   const int32_t* keys = (const int32_t*) (img + cap * 4);
   const unsigned long long* acc = (const unsigned long long*) (img + cap * 4 + cap * kMaxKeys * 4);
   const int total = t.capacity * t.nAggs; // <= 2048 * 8
   for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int slotIdx = i / t.nAggs, a = i % t.nAggs;
      if (t.nKeys == 0 ? slotIdx != 0 : st[slotIdx] != 2) continue;
      int32_t kk[2] = {keys[(size_t) slotIdx * kMaxKeys], keys[(size_t) slotIdx * kMaxKeys + 1]};
      const int slot = peerGroupLookupOrInsert(t, kk);
      if (slot < 0) continue;
      const unsigned long long* src = acc + ((size_t) slotIdx * kMaxAggs + a) * 2;
      unsigned long long* dst = t.acc + ((size_t) slot * kMaxAggs + a) * 2;
      atomicAdd128(dst, dst + 1, i128{src[0], (int64_t) src[1]});
   }
}
// same open-addressing protocol as kernels.cu groupLookupOrInsert (state 0 empty / 1 being written / 2 ready)
__device__ int peerGroupLookupOrInsert(const GroupTableDev& t, const int32_t* k) {
   if (t.nKeys == 0) return 0;
   const uint32_t mask = (uint32_t) t.capacity - 1;
   uint64_t h = hashI32(k[0]);
   if (t.nKeys > 1) h = hashCombine(hashI32(k[1]), h);
   uint32_t s = (uint32_t) h & mask;
   for (int probes = 0; probes < t.capacity; probes++) {
      int st = atomicCAS(&t.state[s], 0, 1);
      if (st == 0) {
         t.keys[s * kMaxKeys + 0] = k[0];
         t.keys[s * kMaxKeys + 1] = t.nKeys > 1 ? k[1] : 0;
         __threadfence();
         atomicExch(&t.state[s], 2);
         return (int) s;
      }
      while (st == 1) st = *((volatile int32_t*) &t.state[s]);
      __threadfence();
      const volatile int32_t* tk = t.keys + s * kMaxKeys;
      if (tk[0] == k[0] && (t.nKeys < 2 || tk[1] == k[1])) return (int) s;
      s = (s + 1) & mask;
   }
   atomicExch(t.error, 1);
   return -1;
}

// OR-all-reduce of a bit array that every rank keeps at the SAME heap offset (the Bloom filters of the hash partitions of one
// logical build side): every rank reads its peers' words over NVLink (P2P loads) and ORs them into its own copy.  A peer that
// already folded some ranks in only contributes bits the result contains anyway, so no second buffer is needed; barriers
// before (all builds done) and after (nobody is still reading) are the caller's.
__global__ void __launch_bounds__(256) peerOrReduceKernel(PeerView v, size_t heapOff, size_t words4 /* number of uint4 */) {
   uint4* own = (uint4*) (v.heap[v.rank] + heapOff);
   for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < words4; i += (size_t) gridDim.x * blockDim.x) {
      uint4 a = own[i];
      for (int p = 0; p < v.world; p++) {
         if (p == v.rank) continue;
         uint4 b; // peer words change while they are read (the peer ORs too): a coherent (non-nc) 128-bit load
         asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"((const uint4*) (v.heap[p] + heapOff) + i) : "memory");
         a.x |= b.x;
         a.y |= b.y;
         a.z |= b.z;
         a.w |= b.w;
      }
      own[i] = a;
   }
}

__global__ void peerPublishCountsKernel(PeerView v, size_t cursorsOff, size_t countsOff) {
   const int d = threadIdx.x;
   if (d >= v.world) return;
   const unsigned long long n = *((const unsigned long long*) (v.heap[v.rank] + cursorsOff) + d);
   *((unsigned long long*) (v.heap[d] + countsOff) + v.rank) = n; // plain store into the peer; the barrier that follows publishes it
}

} // namespace ldb

using namespace ldb;

// ---------------------------------------------------------------- host side
namespace {
template <class Fn>
int guardedPeer(LdbError* err, const Fn& fn) {
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
[[noreturn]] void failPeer(int code, const std::string& m) { throw ApiError(code, m); }
} // namespace

PeerView LdbComm::view() const {
   PeerView v{};
   v.rank = rank;
   v.world = world;
   for (int i = 0; i < world; i++) v.heap[i] = peerHeap[i];
   v.error = error;
   v.timeoutNs = timeoutNs;
   return v;
}

extern "C" {

int64_t ldb_gpu_comm_reserved_bytes(void) { return (int64_t) kUserOff; }

int ldb_gpu_comm_create(LdbContext* ctx, int32_t rank, int32_t world, int64_t user_bytes, LdbComm** out, uint8_t* handle_out, LdbError* err) {
   return guardedPeer(err, [&] {
      if (!ctx || !out || !handle_out) failPeer(LDB_ERR_INVALID, "null argument");
      if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) failPeer(LDB_ERR_INVALID, "rank/world out of range (1..8 ranks)");
      if (user_bytes < 0) failPeer(LDB_ERR_INVALID, "negative heap size");
      LDB_CUDA(cudaSetDevice(ctx->device));
      auto c = std::make_unique<LdbComm>();
      c->ctx = ctx;
      c->rank = rank;
      c->world = world;
      c->userBytes = ((size_t) user_bytes + 255) & ~size_t(255);
      c->heapBytes = kUserOff + c->userBytes + 256;
      LDB_CUDA(cudaMalloc((void**) &c->heap, c->heapBytes)); // plain cudaMalloc: legacy IPC handles cannot export pool / VMM allocations
      LDB_CUDA(cudaMemsetAsync(c->heap, 0, kUserOff, ctx->compute));
      c->error = (int32_t*) (c->heap + kUserOff + c->userBytes);
      LDB_CUDA(cudaMemsetAsync(c->error, 0, 256, ctx->compute));
      ctx->syncStream(ctx->compute);
      c->peerHeap[rank] = c->heap;
      if (const char* e = getenv("LDB_PEER_TIMEOUT_MS")) c->timeoutNs = (unsigned long long) std::max(1, atoi(e)) * 1000000ull;
      // CUDA loads kernels lazily, and loading one may have to wait for running kernels: a collective kernel that spins on a peer
      // while the next launch (of this or another rank in the same process) is still being loaded would stall until the
      // timeout — so every collective kernel is loaded now
      cudaFuncAttributes fa;
      LDB_CUDA(cudaFuncGetAttributes(&fa, peerBarrierKernel));
      LDB_CUDA(cudaFuncGetAttributes(&fa, peerBumpKernel));
      LDB_CUDA(cudaFuncGetAttributes(&fa, peerAllGatherKernel));
      LDB_CUDA(cudaFuncGetAttributes(&fa, peerGroupAllMergeKernel));
      LDB_CUDA(cudaFuncGetAttributes(&fa, peerOrReduceKernel));
      LDB_CUDA(cudaFuncGetAttributes(&fa, peerPublishCountsKernel));
      cudaIpcMemHandle_t h;
      LDB_CUDA(cudaIpcGetMemHandle(&h, c->heap));
      static_assert(sizeof(h) == LDB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
      memcpy(handle_out, &h, sizeof(h));
      *out = c.release();
   });
}

// all_handles: world x 64 bytes in rank order (exchanged by the caller: torch.distributed all_gather, MPI, a file …)
int ldb_gpu_comm_connect(LdbComm* c, const uint8_t* all_handles, LdbError* err) {
   return guardedPeer(err, [&] {
      if (!c || !all_handles) failPeer(LDB_ERR_INVALID, "null argument");
      LDB_CUDA(cudaSetDevice(c->ctx->device));
      for (int p = 0; p < c->world; p++) {
         if (p == c->rank) continue;
         cudaIpcMemHandle_t h;
         memcpy(&h, all_handles + (size_t) p * LDB_IPC_HANDLE_BYTES, sizeof(h));
         void* ptr = nullptr;
         LDB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
         c->peerHeap[p] = (uint8_t*) ptr;
         c->ipcOpened[p] = true;
      }
      c->connected = true;
   });
}

// single-process variant (tests, one process driving several devices): peers are other LdbComm objects of this process
int ldb_gpu_comm_connect_local(LdbComm** comms, int32_t n, LdbError* err) {
   return guardedPeer(err, [&] {
      if (!comms || n < 1 || n > kMaxPeers) failPeer(LDB_ERR_INVALID, "bad comm list");
      for (int i = 0; i < n; i++) {
         if (!comms[i] || comms[i]->world != n || comms[i]->rank != i) failPeer(LDB_ERR_INVALID, "comm list must hold ranks 0..n-1 of one world");
         LDB_CUDA(cudaSetDevice(comms[i]->ctx->device));
         for (int p = 0; p < n; p++) {
            if (p == i) continue;
            if (comms[p]->ctx->device != comms[i]->ctx->device) {
               int can = 0;
               LDB_CUDA(cudaDeviceCanAccessPeer(&can, comms[i]->ctx->device, comms[p]->ctx->device));
               if (!can) failPeer(LDB_ERR_UNSUPPORTED, "devices cannot access each other's memory");
               cudaError_t e = cudaDeviceEnablePeerAccess(comms[p]->ctx->device, 0);
               if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) LDB_CUDA(e);
               cudaGetLastError();
            }
            comms[i]->peerHeap[p] = comms[p]->heap;
         }
         comms[i]->connected = true;
      }
   });
}

void ldb_gpu_comm_destroy(LdbComm* c) {
   if (!c) return;
   cudaSetDevice(c->ctx->device);
   cudaStreamSynchronize(c->ctx->compute);
   for (int p = 0; p < c->world; p++)
      if (c->ipcOpened[p]) cudaIpcCloseMemHandle(c->peerHeap[p]);
   cudaFree(c->heap);
   delete c;
}

void* ldb_gpu_comm_heap(LdbComm* c, int64_t* user_bytes) {
   if (!c) return nullptr;
   if (user_bytes) *user_bytes = (int64_t) c->userBytes;
   return c->heap + kUserOff;
}
int32_t ldb_gpu_comm_rank(LdbComm* c) { return c ? c->rank : -1; }
int32_t ldb_gpu_comm_world(LdbComm* c) { return c ? c->world : 0; }

static void wantConnected(LdbComm* c) {
   if (!c) failPeer(LDB_ERR_INVALID, "null comm");
   if (!c->connected && c->world > 1) failPeer(LDB_ERR_INVALID, "comm is not connected to its peers yet");
}

int ldb_gpu_comm_barrier(LdbComm* c, LdbError* err) {
   return guardedPeer(err, [&] {
      wantConnected(c);
      if (c->world == 1) return;
      LdbContext* ctx = c->ctx;
      LDB_CUDA(cudaSetDevice(ctx->device));
      ctx->launch("peer_barrier", [&] {
         peerBarrierKernel<<<c->world, 32, 0, ctx->compute>>>(c->view());
         peerBumpKernel<<<1, 1, 0, ctx->compute>>>(c->view(), EPOCH_BARRIER, -1);
      });
   });
}

// all-gather of `bytes` (multiple of 16, <= slot size) from DEVICE memory `src`; returns the device address of the gathered
// blocks of this collective: block of rank r at result + r * ldb_gpu_comm_slot_bytes().  Valid until the next-but-one gather.
int64_t ldb_gpu_comm_slot_bytes(void) { return (int64_t) kSlotBytes; }
int ldb_gpu_comm_allgather_small(LdbComm* c, const void* src, int64_t bytes, void** result, LdbError* err) {
   return guardedPeer(err, [&] {
      wantConnected(c);
      if (bytes <= 0 || bytes > (int64_t) kSlotBytes || bytes % 16) failPeer(LDB_ERR_INVALID, "all-gather blocks are 16..262144 bytes, multiples of 16");
      LdbContext* ctx = c->ctx;
      LDB_CUDA(cudaSetDevice(ctx->device));
      if (ctx->capturing) failPeer(LDB_ERR_UNSUPPORTED, "the small all-gather returns a parity-dependent address and cannot be captured");
      const unsigned long long epoch = ++c->gatherEpochHost; // mirror of the device counter (every rank issues the same collectives)
      ctx->launch("peer_allgather", [&] {
         peerAllGatherKernel<<<c->world, 256, 0, ctx->compute>>>(c->view(), (const uint8_t*) src, (size_t) bytes);
         peerBumpKernel<<<1, 1, 0, ctx->compute>>>(c->view(), EPOCH_GATHER, -1);
      });
      if (result) *result = c->heap + kMailboxOff + (size_t) (epoch & 1) * c->world * kSlotBytes;
   });
}

int ldb_gpu_groupby_allmerge(LdbState* s, LdbComm* c, LdbError* err) {
   return guardedPeer(err, [&] {
      if (!s || (s->kind != LDB_STATE_GROUPBY && s->kind != LDB_STATE_SIMPLE)) failPeer(LDB_ERR_INVALID, "not a group state");
      wantConnected(c);
      if (s->ctx != c->ctx) failPeer(LDB_ERR_INVALID, "state and comm belong to different contexts");
      if (c->world == 1) return;
      const size_t image = (groupImageBytes(s->group.capacity) + 15) & ~size_t(15); // the table allocation carries 16 spare bytes (error word)
      if (image > kSlotBytes) failPeer(LDB_ERR_UNSUPPORTED, "group table image larger than a mailbox slot (capacity <= 1024 groups)");
      LdbContext* ctx = c->ctx;
      LDB_CUDA(cudaSetDevice(ctx->device));
      ++c->gatherEpochHost;
      if (ctx->capturing) ctx->capturing->onLaunch.push_back([c] { ++c->gatherEpochHost; }); // every replay bumps the device epoch once more
      ctx->launch("peer_group_allmerge", [&] {
         peerGroupAllMergeKernel<<<c->world, 256, 0, ctx->compute>>>(c->view(), s->group, image);
         peerBumpKernel<<<1, 1, 0, ctx->compute>>>(c->view(), EPOCH_GATHER, EPOCH_MERGE);
      });
   });
}

int ldb_gpu_comm_or_reduce(LdbComm* c, int64_t user_offset, int64_t bytes, LdbError* err) {
   return guardedPeer(err, [&] {
      wantConnected(c);
      if (user_offset < 0 || bytes < 0 || user_offset % 16 || bytes % 16 || (size_t) (user_offset + bytes) > c->userBytes) failPeer(LDB_ERR_INVALID, "OR-reduce range outside the heap or not 16-byte aligned");
      if (c->world == 1 || bytes == 0) return;
      LdbContext* ctx = c->ctx;
      LDB_CUDA(cudaSetDevice(ctx->device));
      const size_t n4 = (size_t) bytes / 16;
      const int grid = (int) std::min<size_t>((n4 + 255) / 256, (size_t) ctx->smCount * 8);
      ctx->launch("peer_or_reduce", [&] { peerOrReduceKernel<<<grid, 256, 0, ctx->compute>>>(c->view(), kUserOff + (size_t) user_offset, n4); });
   });
}

int ldb_gpu_comm_heap_zero(LdbComm* c, int64_t user_offset, int64_t bytes, LdbError* err) {
   return guardedPeer(err, [&] {
      if (!c || user_offset < 0 || bytes < 0 || (size_t) (user_offset + bytes) > c->userBytes) failPeer(LDB_ERR_INVALID, "range outside the comm's user heap");
      LDB_CUDA(cudaSetDevice(c->ctx->device));
      LDB_CUDA(cudaMemsetAsync(c->heap + kUserOff + user_offset, 0, (size_t) bytes, c->ctx->compute));
   });
}
int ldb_gpu_comm_heap_read(LdbComm* c, int64_t user_offset, int64_t bytes, void* host_dst, LdbError* err) {
   return guardedPeer(err, [&] {
      if (!c || !host_dst || user_offset < 0 || bytes < 0 || (size_t) (user_offset + bytes) > c->userBytes) failPeer(LDB_ERR_INVALID, "range outside the comm's user heap");
      LDB_CUDA(cudaSetDevice(c->ctx->device));
      LDB_CUDA(cudaMemcpyAsync(host_dst, c->heap + kUserOff + user_offset, (size_t) bytes, cudaMemcpyDeviceToHost, c->ctx->compute));
      c->ctx->syncStream(c->ctx->compute);
   });
}
static void wantRange(LdbComm* c, int64_t off, int64_t bytes, const char* what) {
   if (off < 0 || bytes < 0 || off % 16 || (size_t) (off + bytes) > c->userBytes) failPeer(LDB_ERR_CAPACITY, std::string(what) + " outside the comm's user heap (create the comm with a larger heap)");
}
int ldb_gpu_comm_publish_counts(LdbComm* c, int64_t cursors_offset, int64_t counts_offset, LdbError* err) {
   return guardedPeer(err, [&] {
      wantConnected(c);
      wantRange(c, cursors_offset, 16 * 8, "cursors");
      wantRange(c, counts_offset, kMaxPeers * 8, "counts");
      LdbContext* ctx = c->ctx;
      LDB_CUDA(cudaSetDevice(ctx->device));
      ctx->launch("peer_publish_counts", [&] { peerPublishCountsKernel<<<1, 32, 0, ctx->compute>>>(c->view(), kUserOff + (size_t) cursors_offset, kUserOff + (size_t) counts_offset); });
   });
}
int ldb_gpu_join_table_insert_received(LdbState* table, LdbComm* c, int64_t recv_offset, int64_t capacity, int64_t counts_offset, LdbError* err) {
   return guardedPeer(err, [&] {
      wantConnected(c);
      if (!table || table->kind != LDB_STATE_JOIN_TABLE || table->join.stride != 8 || table->join.direct) failPeer(LDB_ERR_INVALID, "insert target must be a plain single-key join table");
      wantRange(c, recv_offset, (int64_t) c->world * capacity * 8, "receive region");
      wantRange(c, counts_offset, kMaxPeers * 8, "counts");
      LdbContext* ctx = c->ctx;
      LDB_CUDA(cudaSetDevice(ctx->device));
      uint8_t* user = c->heap + kUserOff;
      ctx->launch("join_build", [&] { launchInsertReceived(table->join, user + recv_offset, c->world, capacity, (const unsigned long long*) (user + counts_offset), ctx->smCount, ctx->compute); });
   });
}
int ldb_gpu_probe_received_groupby(LdbState* ta, LdbState* tb, LdbState* groups, LdbComm* c, int64_t recv_offset, int64_t capacity, int64_t counts_offset, int32_t scale, LdbError* err) {
   return guardedPeer(err, [&] {
      wantConnected(c);
      for (LdbState* t : {ta, tb})
         if (!t || t->kind != LDB_STATE_JOIN_TABLE || t->join.stride != 8 || t->join.direct) failPeer(LDB_ERR_INVALID, "probe tables must be plain single-key join tables");
      if (!groups || groups->kind != LDB_STATE_GROUPBY || groups->group.nKeys != 1 || groups->group.nAggs != 1) failPeer(LDB_ERR_INVALID, "sink must be a group-by state with one key and one aggregate");
      if (scale < 0 || scale > 18) failPeer(LDB_ERR_INVALID, "decimal scale out of range");
      wantRange(c, recv_offset, (int64_t) c->world * capacity * 24, "receive region");
      wantRange(c, counts_offset, kMaxPeers * 8, "counts");
      int64_t one = 1;
      for (int i = 0; i < scale; i++) one *= 10;
      LdbContext* ctx = c->ctx;
      LDB_CUDA(cudaSetDevice(ctx->device));
      uint8_t* user = c->heap + kUserOff;
      ctx->launch("join_probe2_groupby", [&] {
         launchProbeReceivedGroupBy(ta->join, tb->join, groups->group, user + recv_offset, c->world, capacity, (const unsigned long long*) (user + counts_offset), one, ctx->smCount, ctx->compute);
      });
   });
}

int ldb_gpu_probe_received_groupby2(LdbState* table, LdbState* groups, LdbComm* c, int64_t recv_offset, int64_t capacity, int64_t counts_offset, LdbError* err) {
   return guardedPeer(err, [&] {
      wantConnected(c);
      if (!table || table->kind != LDB_STATE_JOIN_TABLE || table->join.stride != 8 || table->join.direct) failPeer(LDB_ERR_INVALID, "probe table must be a plain single-key join table");
      if (!groups || groups->kind != LDB_STATE_GROUPBY || groups->group.nKeys != 2 || groups->group.nAggs != 1) failPeer(LDB_ERR_INVALID, "sink must be a group-by state with two keys and one aggregate");
      wantRange(c, recv_offset, (int64_t) c->world * capacity * 24, "receive region");
      wantRange(c, counts_offset, kMaxPeers * 8, "counts");
      LdbContext* ctx = c->ctx;
      LDB_CUDA(cudaSetDevice(ctx->device));
      uint8_t* user = c->heap + kUserOff;
      ctx->launch("join_probe_received_groupby", [&] {
         launchProbeReceivedGroupBy2(table->join, groups->group, user + recv_offset, c->world, capacity, (const unsigned long long*) (user + counts_offset), ctx->smCount, ctx->compute);
      });
   });
}

// surfaces a timed-out wait (dead or stuck peer); synchronises the compute stream
int ldb_gpu_comm_check(LdbComm* c, LdbError* err) {
   return guardedPeer(err, [&] {
      if (!c) failPeer(LDB_ERR_INVALID, "null comm");
      LDB_CUDA(cudaSetDevice(c->ctx->device));
      int32_t e = 0;
      LDB_CUDA(cudaMemcpyAsync(&e, c->error, 4, cudaMemcpyDeviceToHost, c->ctx->compute));
      c->ctx->syncStream(c->ctx->compute);
      if (e) failPeer(LDB_ERR_CUDA, "a peer did not arrive at a collective within the timeout (LDB_PEER_TIMEOUT_MS)");
   });
}

} // extern "C"
