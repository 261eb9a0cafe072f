// peer.h — symmetric-heap layout and handle of the peer-mapped exchange (peer.cu); internal to libldb_gpu.so.
#pragma once
#include "context.h"

namespace ldb {

constexpr int kMaxPeers = 8;                    // one NVSwitch domain: 8 GPUs of a box
constexpr size_t kSlotBytes = 256u << 10;       // one mailbox block (a 1024-group table image is 140 KB)
constexpr size_t kBarrierFlagsOff = 0;          // u64 [kMaxPeers]           written by peer p at index p
constexpr size_t kGatherFlagsOff = 256;         // u64 [2][kMaxPeers]
constexpr size_t kLocalSyncOff = 512;           // u64: CTA rendezvous counter of this GPU's own collectives
constexpr size_t kEpochsOff = 768;              // u64 [4]: device-side epoch counters {barrier, gather, merge}: a collective kernel reads its epoch
                                                // here and a 1-thread kernel behind it bumps the counter, so a captured CUDA graph replays correctly
constexpr size_t kCountsOff = 1024;             // u64 [2][kMaxPeers][kMaxPeers]  shuffle row counts, [parity][src][dst]
constexpr size_t kMailboxOff = 4096;            // [2][kMaxPeers][kSlotBytes]
constexpr size_t kUserOff = kMailboxOff + 2 * kMaxPeers * kSlotBytes; // 4 MiB + 4 KiB; user region starts here (256-byte aligned)

struct PeerView {
   int32_t rank, world;
   uint8_t* heap[kMaxPeers]; // peer-mapped base of every rank's symmetric heap (own included)
   int32_t* error;           // own error word
   unsigned long long timeoutNs;
};

} // namespace ldb

struct LdbComm {
   LdbContext* ctx = nullptr;
   int32_t rank = 0, world = 1;
   uint8_t* heap = nullptr;
   size_t heapBytes = 0, userBytes = 0;
   uint8_t* peerHeap[ldb::kMaxPeers] = {};
   bool ipcOpened[ldb::kMaxPeers] = {};
   bool connected = false;
   int32_t* error = nullptr;
   unsigned long long gatherEpochHost = 0; // host mirror of the device gather epoch (parity of the mailbox an eager all-gather filled)
   unsigned long long timeoutNs = 20ull * 1000000000ull; // a peer that does not arrive within 20 s is reported, not waited for
   ldb::PeerView view() const;
};
