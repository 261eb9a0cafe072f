// oracle/port/runtime_port.h — TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// Self-contained CPU restatement of the reference's hot-path runtime objects, used when
// /root/reference is not available (GPU box) and cross-checked against the verbatim-compiled
// reference objects (oracle/_ref, runtime_ref.h) where it is.  Same class and method names as the
// reference so that pipelines.cpp — the restated JIT output — is written once against `rt::`.
// Each class cites what it follows.  Only *semantics* (which elements are visited / produced) are
// restated; work-distribution details that cannot change a result (work stealing order) are
// simplified and say so.
#pragma once
#include "scheduler.h"
#include "table.h"
#include "values.h"

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <string>
#include <variant>
#include <vector>

namespace oracle::port {

// ---------------------------------------------------------------------------------------------
// ArrowView ABI structs (include/lingodb/runtime/ArrowView.h:8-29): read directly by generated code.
struct ArrayView {
   int64_t length, nullCount, offset, nBuffers, nChildren;
   const void** buffers;
   const ArrayView** children;
};
struct BatchView {
   static constexpr size_t maxBatchSize = 65536;
   int64_t length, offset;
   uint16_t* selectionVector;
   const ArrayView** arrays;
   static uint16_t* defaultSelectionVector() { // 0,1,2,…,65535 (src/runtime/ArrowView.cpp)
      static std::vector<uint16_t> v = [] {
         std::vector<uint16_t> r(65536);
         for (size_t i = 0; i < 65536; i++) r[i] = (uint16_t) i;
         return r;
      }();
      return v.data();
   }
};

// ---------------------------------------------------------------------------------------------
// ExecutionContext: ownership of every runtime object created during a query
// (include/lingodb/runtime/ExecutionContext.h:111-113; dtor ExecutionContext.cpp:27-40).
struct State {
   void* ptr;
   std::function<void(void*)> freeFn;
};
class ExecutionContext {
   std::vector<std::vector<State>> perWorkerStates;

   public:
   ExecutionContext() { perWorkerStates.resize(sched::getNumWorkers()); }
   void registerState(const State& s) { perWorkerStates[sched::currentWorkerId()].push_back(s); }
   ~ExecutionContext() {
      for (auto& l : perWorkerStates)
         for (auto& s : l) s.freeFn(s.ptr);
   }
};
inline ExecutionContext*& currentContextSlot() {
   static ExecutionContext* ctx = nullptr; // one query at a time in the oracle; shared by all workers
   return ctx;
}
inline ExecutionContext* getCurrentExecutionContext() { return currentContextSlot(); }
struct QueryContextScope {
   ExecutionContext ctx;
   QueryContextScope() { currentContextSlot() = &ctx; }
   ~QueryContextScope() { currentContextSlot() = nullptr; }
};

// ---------------------------------------------------------------------------------------------
// Buffer / FlexibleBuffer (include/lingodb/runtime/Buffer.h:16-105, src/runtime/Buffer.cpp:54-135).
struct Buffer {
   uint64_t numElements;
   uint8_t* ptr;
};
class FlexibleBuffer {
   size_t totalLen = 0, currCapacity, typeSize;
   std::vector<Buffer> buffers;

   public:
   FlexibleBuffer(size_t initialCapacity, size_t typeSize) : currCapacity(initialCapacity), typeSize(typeSize) {
      buffers.push_back(Buffer{0, (uint8_t*) malloc(initialCapacity * typeSize)});
   }
   ~FlexibleBuffer() {
      for (auto& b : buffers) free(b.ptr);
   }
   FlexibleBuffer(const FlexibleBuffer&) = delete;
   uint8_t* insert() { // chunk growth ×1.2, rounded up (Buffer.h:49-67)
      if (buffers.empty() || buffers.back().numElements == currCapacity) {
         size_t next = (size_t) std::ceil(currCapacity * 1.2);
         buffers.push_back(Buffer{0, (uint8_t*) malloc(next * typeSize)});
         currCapacity = next;
      }
      totalLen++;
      Buffer& b = buffers.back();
      return b.ptr + typeSize * (b.numElements++);
   }
   size_t getLen() const { return totalLen; }
   size_t getTypeSize() const { return typeSize; }
   const std::vector<Buffer>& getBuffers() const { return buffers; }
   void merge(FlexibleBuffer& other) { // concatenates the chunk lists (Buffer.h:92-98)
      buffers.insert(buffers.begin(), other.buffers.begin(), other.buffers.end());
      other.buffers.clear();
      totalLen += other.totalLen;
      other.totalLen = 0;
      other.currCapacity = 0;
   }
   template <class Fn>
   void iterate(const Fn& fn) {
      for (auto& b : buffers)
         for (size_t i = 0; i < b.numElements; i++) fn(b.ptr + i * typeSize);
   }
   // Parallel iteration in 20 000-element units (Buffer.cpp:54-135).  The reference reserves units
   // per worker and steals; here a shared cursor hands out the same units — identical coverage.
   void iterateBuffersParallel(const std::function<void(Buffer)>& fn) {
      constexpr size_t splitSize = 20000;
      struct Unit {
         size_t buf, begin, len;
      };
      std::vector<Unit> units;
      for (size_t b = 0; b < buffers.size(); b++)
         for (size_t s = 0; s < buffers[b].numElements; s += splitSize)
            units.push_back({b, s, std::min(splitSize, (size_t) buffers[b].numElements - s)});
      struct T : sched::TaskIface {
         std::vector<Unit>& units;
         FlexibleBuffer& self;
         const std::function<void(Buffer)>& fn;
         std::atomic<size_t> cursor{0};
         std::vector<size_t> resv;
         T(std::vector<Unit>& u, FlexibleBuffer& s, const std::function<void(Buffer)>& fn) : units(u), self(s), fn(fn), resv(sched::getNumWorkers()) {}
         bool allocateWork() override {
            size_t i = cursor.fetch_add(1);
            if (i >= units.size()) return false;
            resv[sched::currentWorkerId()] = i;
            return true;
         }
         void performWork() override {
            auto& u = units[resv[sched::currentWorkerId()]];
            fn(Buffer{u.len, self.buffers[u.buf].ptr + u.begin * std::max<size_t>(1, self.typeSize)});
         }
      } task(units, *this, fn);
      sched::runTask(task);
   }
   template <class Fn>
   void iterateParallel(const Fn& fn) {
      iterateBuffersParallel([&](Buffer b) {
         for (size_t i = 0; i < b.numElements; i++) fn(b.ptr + i * typeSize);
      });
   }
};

// BufferIterator::iterate(it, parallel, forEachChunk, ctx) (Buffer.cpp:188-236): chunk callback gets
// {bytes, ptr}.
struct BufferIterator {
   FlexibleBuffer& fb;
   explicit BufferIterator(FlexibleBuffer& fb) : fb(fb) {}
   static void iterate(BufferIterator* it, bool parallel, void (*forEachChunk)(Buffer, void*), void* ctx) {
      size_t ts = std::max<size_t>(1, it->fb.getTypeSize());
      if (parallel) {
         it->fb.iterateBuffersParallel([&](Buffer b) { forEachChunk(Buffer{b.numElements * ts, b.ptr}, ctx); });
      } else {
         for (auto b : it->fb.getBuffers()) forEachChunk(Buffer{b.numElements * ts, b.ptr}, ctx);
      }
   }
};
inline BufferIterator* createIterator(FlexibleBuffer& fb) {
   auto* it = new BufferIterator(fb);
   getCurrentExecutionContext()->registerState({it, [](void* p) { delete (BufferIterator*) p; }});
   return it;
}

// ---------------------------------------------------------------------------------------------
// ThreadLocal (include/lingodb/runtime/ThreadLocal.h:7-35, src/runtime/ThreadLocal.cpp:8-36)
class ThreadLocal {
   std::vector<uint8_t*> values;
   uint8_t* (*initFn)(uint8_t*);
   uint8_t* arg;
   ThreadLocal(uint8_t* (*f)(uint8_t*), uint8_t* a) : values(sched::getNumWorkers(), nullptr), initFn(f), arg(a) {}

   public:
   static ThreadLocal* create(uint8_t* (*initFn)(uint8_t*), uint8_t* arg) {
      auto* tl = new ThreadLocal(initFn, arg);
      getCurrentExecutionContext()->registerState({tl, [](void* p) { delete (ThreadLocal*) p; }});
      return tl;
   }
   uint8_t* getLocal() {
      auto& v = values[sched::currentWorkerId()];
      if (!v) v = initFn(arg);
      return v;
   }
   // guarantees at least one value exists (ThreadLocal.h:21-29)
   std::vector<uint8_t*>& getThreadLocalValues() {
      bool any = false;
      for (auto* v : values) any |= v != nullptr;
      if (!any) values.back() = initFn(arg);
      return values;
   }
   uint8_t* merge(void (*mergeFn)(uint8_t*, uint8_t*)) {
      uint8_t* first = nullptr;
      for (auto* p : getThreadLocalValues()) {
         if (!p) continue;
         if (!first) first = p;
         else mergeFn(first, p);
      }
      return first;
   }
};

// SimpleState (src/runtime/SimpleState.cpp:8-30): keyless aggregate state per worker.
struct SimpleState {
   static uint8_t* create(size_t size) {
      uint8_t* r = (uint8_t*) malloc(size);
      getCurrentExecutionContext()->registerState({r, [](void* p) { free(p); }});
      return r;
   }
   static uint8_t* merge(ThreadLocal* tl, void (*mergeFn)(uint8_t*, uint8_t*)) { return tl->merge(mergeFn); }
};

// ---------------------------------------------------------------------------------------------
// GrowingBuffer (include/lingodb/runtime/GrowingBuffer.h:16-32, src/runtime/GrowingBuffer.cpp:39-113)
struct GrowingBufferAllocator {
   static GrowingBufferAllocator* getDefaultAllocator() {
      static GrowingBufferAllocator a;
      return &a;
   }
};
class GrowingBuffer {
   FlexibleBuffer values;

   public:
   GrowingBuffer(size_t cap, size_t typeSize) : values(cap, typeSize) {}
   static GrowingBuffer* create(GrowingBufferAllocator*, size_t sizeOfType, size_t initialCapacity) {
      auto* r = new GrowingBuffer(initialCapacity, sizeOfType);
      getCurrentExecutionContext()->registerState({r, [](void* p) { delete (GrowingBuffer*) p; }});
      return r;
   }
   static ThreadLocal* createThreadLocal(size_t sizeOfType) {
      return ThreadLocal::create([](uint8_t* arg) -> uint8_t* { return (uint8_t*) GrowingBuffer::create(GrowingBufferAllocator::getDefaultAllocator(), (size_t) arg, 1024); }, (uint8_t*) sizeOfType);
   }
   uint8_t* insert() { return values.insert(); }
   size_t getLen() const { return values.getLen(); }
   FlexibleBuffer& getValues() { return values; }
   static GrowingBuffer* merge(ThreadLocal* tl) {
      GrowingBuffer* first = nullptr;
      for (auto* p : tl->getThreadLocalValues()) {
         auto* cur = (GrowingBuffer*) p;
         if (!cur) continue;
         if (!first) first = cur;
         else first->values.merge(cur->values);
      }
      return first;
   }
   BufferIterator* createIterator() { return port::createIterator(values); }
};

// ---------------------------------------------------------------------------------------------
// Tagged chain pointers with a 16-bit bloom tag (include/lingodb/runtime/helpers.h:325-346).
// bloomMasks[2048]: the first 1820 entries are all 16-bit patterns with exactly four bits set, in
// increasing order (verified against src/runtime/helpers.cpp:2); the reference fills the remaining
// 228 with a fixed random resample of those.  The tag only decides which probes are rejected before
// the chain walk — it can never change a result — so the port resamples deterministically instead
// of carrying the table; the oracle/_ref build uses the reference's own table.
inline const uint16_t* bloomMasks() {
   static std::vector<uint16_t> m = [] {
      std::vector<uint16_t> r;
      for (uint32_t v = 0; v < 65536; v++)
         if (__builtin_popcount(v) == 4) r.push_back((uint16_t) v);
      for (size_t i = 0; r.size() < 2048; i++) r.push_back(r[(i * 7919 + 13) % 1820]);
      return r;
   }();
   return m.data();
}
template <class T>
T* untag(T* p) { return reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(p) >> 16); }
template <class T>
T* tag(T* p, T* previous, size_t hash) {
   uint16_t prevTag = (uint16_t) reinterpret_cast<uintptr_t>(previous);
   uint16_t curTag = bloomMasks()[hash >> 53];
   return reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(p) << 16 | (uintptr_t) (curTag | prevTag));
}
template <class T>
bool matchesTag(T* p, size_t hash) {
   uint16_t entry = (uint16_t) reinterpret_cast<uintptr_t>(p);
   uint16_t t = bloomMasks()[hash >> 53];
   return !(t & ~entry);
}
inline uint64_t nextPow2(uint64_t v) {
   v--;
   v |= v >> 1;
   v |= v >> 2;
   v |= v >> 4;
   v |= v >> 8;
   v |= v >> 16;
   v |= v >> 32;
   return v + 1;
}

// HashIndexedView (include/lingodb/runtime/LazyJoinHashtable.h:7-32, src/runtime/LazyJoinHashtable.cpp:12-34).
// The first 16 bytes {ht, htMask} are read directly by generated probe code
// (SubOpToControlFlow.cpp:2573-2575).
class HashIndexedView {
   public:
   struct Entry {
      Entry* next;
      uint64_t hashValue;
   };

   private:
   Entry** ht;
   size_t htMask;

   public:
   static HashIndexedView* build(GrowingBuffer* buffer) {
      auto& values = buffer->getValues();
      size_t htSize = std::max<uint64_t>(nextPow2((uint64_t) (values.getLen() * 1.25)), 1);
      auto* v = new HashIndexedView;
      v->ht = (Entry**) calloc(htSize, sizeof(Entry*));
      v->htMask = htSize - 1;
      getCurrentExecutionContext()->registerState({v, [](void* p) {
                                                      auto* h = (HashIndexedView*) p;
                                                      free(h->ht);
                                                      delete h;
                                                   }});
      values.iterateParallel([&](uint8_t* ptr) {
         auto* entry = (Entry*) ptr;
         size_t hash = entry->hashValue;
         std::atomic_ref<Entry*> slot(v->ht[hash & v->htMask]);
         Entry* current = slot.load();
         Entry* tagged;
         do {
            entry->next = untag(current);
            tagged = tag(entry, current, hash);
         } while (!slot.compare_exchange_weak(current, tagged));
      });
      return v;
   }
};

// ---------------------------------------------------------------------------------------------
// PreAggregationHashtableFragment / PreAggregationHashtable
// (include/lingodb/runtime/PreAggregationHashtable.h:8-47, src/runtime/PreAggregationHashtable.cpp:46-170).
// A fragment pointer *is* the Entry* ht[1024] array for generated code (SubOpToControlFlow.cpp:3111-3119).
class PreAggregationHashtableFragment {
   public:
   struct Entry {
      Entry* next;
      size_t hashValue;
      uint8_t content[];
   };
   static constexpr size_t numOutputs = 64, hashtableSize = 1024;
   Entry* ht[hashtableSize];
   size_t typeSize, len;
   FlexibleBuffer* outputs[numOutputs];
   bool withLocks;
   PreAggregationHashtableFragment(size_t typeSize, bool withLocks) : ht(), typeSize(typeSize), len(0), outputs(), withLocks(withLocks) {}
   ~PreAggregationHashtableFragment() {
      for (auto* o : outputs) delete o;
   }
   static PreAggregationHashtableFragment* create(size_t typeSize, bool withLocks) {
      auto* f = new PreAggregationHashtableFragment(typeSize, withLocks);
      getCurrentExecutionContext()->registerState({f, [](void* p) { delete (PreAggregationHashtableFragment*) p; }});
      return f;
   }
   // miss path: append to partition `hash & 63`, overwrite cache slot `(hash >> 6) & 1023`
   Entry* insert(size_t hash) {
      len++;
      size_t out = hash & (numOutputs - 1);
      if (!outputs[out]) outputs[out] = new FlexibleBuffer(256, typeSize);
      auto* e = (Entry*) outputs[out]->insert();
      e->hashValue = hash;
      e->next = nullptr;
      ht[(hash >> 6) & (hashtableSize - 1)] = e;
      return e;
   }
};
class PreAggregationHashtable {
   public:
   using Entry = PreAggregationHashtableFragment::Entry;
   struct PartitionHt {
      Entry** ht;
      size_t hashMask;
   };

   private:
   PartitionHt ht[PreAggregationHashtableFragment::numOutputs]; // offset 0: read by generated lookups (SubOpToControlFlow.cpp:2765-2772)
   FlexibleBuffer buffer;
   PreAggregationHashtable() : ht(), buffer(1, sizeof(Entry*)) {}

   public:
   ~PreAggregationHashtable() {
      for (auto& p : ht) free(p.ht);
   }
   static PreAggregationHashtable* merge(ThreadLocal* tl, bool (*eq)(uint8_t*, uint8_t*), void (*combine)(uint8_t*, uint8_t*)) {
      constexpr size_t numPartitions = PreAggregationHashtableFragment::numOutputs;
      std::vector<FlexibleBuffer*> outputs[numPartitions];
      for (auto* p : tl->getThreadLocalValues()) {
         auto* fragment = (PreAggregationHashtableFragment*) p;
         if (!fragment) continue;
         for (size_t i = 0; i < numPartitions; i++)
            if (fragment->outputs[i]) outputs[i].push_back(fragment->outputs[i]);
      }
      auto* res = new PreAggregationHashtable();
      getCurrentExecutionContext()->registerState({res, [](void* p) { delete (PreAggregationHashtable*) p; }});
      std::mutex mutex;
      struct T : sched::TaskIface {
         std::function<void(size_t)> fn;
         std::atomic<size_t> next{0};
         std::vector<size_t> resv;
         T() : resv(sched::getNumWorkers()) {}
         bool allocateWork() override {
            size_t i = next.fetch_add(1);
            if (i >= numPartitions) return false;
            resv[sched::currentWorkerId()] = i;
            return true;
         }
         void performWork() override { fn(resv[sched::currentWorkerId()]); }
      } task;
      task.fn = [&](size_t id) { // one partition (PreAggregationHashtable.cpp:96-154)
         auto& input = outputs[id];
         size_t total = 0, minValues = 0;
         for (auto* o : input) {
            total += o->getLen();
            minValues = o->getLen(); // PreAggregationHashtable.cpp:103-106: the last fragment's length wins
         }
         FlexibleBuffer local(minValues, sizeof(Entry*));
         size_t htSize = std::max<uint64_t>(nextPow2((uint64_t) (total * 1.25)), 1);
         size_t htMask = htSize - 1;
         Entry** table = (Entry**) calloc(htSize, sizeof(Entry*));
         for (auto* o : input) {
            o->iterate([&](uint8_t* raw) {
               Entry* curr = (Entry*) raw;
               size_t pos = (curr->hashValue >> 6) & htMask;
               Entry* cand = untag(table[pos]);
               bool merged = false;
               while (cand) {
                  if (cand->hashValue == curr->hashValue && eq(cand->content, curr->content)) {
                     combine(cand->content, curr->content);
                     merged = true;
                     break;
                  }
                  cand = cand->next;
               }
               if (!merged) {
                  *(Entry**) local.insert() = curr;
                  Entry* prev = table[pos];
                  table[pos] = tag(curr, prev, curr->hashValue);
                  curr->next = untag(prev);
               }
            });
         }
         res->ht[id] = {table, htMask};
         std::unique_lock<std::mutex> l(mutex);
         res->buffer.merge(local);
      };
      sched::runTask(task);
      return res;
   }
   Entry* lookup(size_t hash) { // PreAggregationHashtable.cpp:162-170
      auto& p = ht[hash & (PreAggregationHashtableFragment::numOutputs - 1)];
      if (!p.ht) return nullptr;
      Entry* e = p.ht[p.hashMask & (hash >> 6)];
      return matchesTag(e, hash) ? untag(e) : nullptr;
   }
   BufferIterator* createIterator() { return port::createIterator(buffer); }
};

// EntryLock (src/runtime/EntryLock.cpp:9-25): one-byte spin lock inside the aggregate entry.
struct EntryLock {
   std::atomic_flag m{};
   static void initialize(EntryLock* l) { new (l) EntryLock(); }
   static void lock(EntryLock* l) {
      while (l->m.test_and_set(std::memory_order_acquire)) {}
   }
   static void unlock(EntryLock* l) { l->m.clear(std::memory_order_release); }
};

// ---------------------------------------------------------------------------------------------
// Pushed-down scan filters (include/lingodb/runtime/storage/TableStorage.h:14-31,
// src/runtime/storage/Restrictions.cpp:67-520).  A conjunction of column-vs-constant predicates,
// evaluated filter by filter; each pass compacts a uint16 selection vector.
// FilterOp / PhysType / FilterDescription / ColumnSchema live in table.h (shared with the _ref adapter).

class Filter {
   public:
   virtual size_t filter(size_t len, const uint16_t* cur, uint16_t* next, const ArrayView* av, size_t offset) = 0;
   virtual ~Filter() {}
};
template <class T>
bool cmpApply(FilterOp op, const T& a, const T& b) {
   switch (op) {
      case FilterOp::EQ: return a == b;
      case FilterOp::NEQ: return a != b;
      case FilterOp::LT: return a < b;
      case FilterOp::LTE: return a <= b;
      case FilterOp::GT: return a > b;
      case FilterOp::GTE: return a >= b;
      default: throw std::runtime_error("unsupported filter op");
   }
}
template <class T>
class SimpleTypeFilter : public Filter { // Restrictions.cpp:163-193: branch-free compare-and-compact
   T value;
   FilterOp op;

   public:
   SimpleTypeFilter(FilterOp op, T v) : value(v), op(op) {}
   size_t filter(size_t len, const uint16_t* cur, uint16_t* next, const ArrayView* av, size_t offset) override {
      const T* data = reinterpret_cast<const T*>(av->buffers[1]) + offset + av->offset;
      uint16_t* w = next;
      for (size_t i = 0; i < len; i++) {
         size_t idx = cur[i];
         *w = (uint16_t) idx;
         w += cmpApply<T>(op, data[idx], value);
      }
      return w - next;
   }
};
class VarLen32Filter : public Filter { // Restrictions.cpp:279-325: utf8 offsets + data
   std::string value;
   FilterOp op;

   public:
   VarLen32Filter(FilterOp op, std::string v) : value(std::move(v)), op(op) {}
   size_t filter(size_t len, const uint16_t* cur, uint16_t* next, const ArrayView* av, size_t offset) override {
      const uint8_t* data = reinterpret_cast<const uint8_t*>(av->buffers[2]);
      const int32_t* offsets = reinterpret_cast<const int32_t*>(av->buffers[1]) + offset + av->offset;
      uint16_t* w = next;
      for (size_t i = 0; i < len; i++) {
         size_t idx = cur[i];
         std::string_view sv((const char*) data + offsets[idx], offsets[idx + 1] - offsets[idx]);
         *w = (uint16_t) idx;
         w += cmpApply<std::string_view>(op, sv, std::string_view(value));
      }
      return w - next;
   }
};
class NotNullFilter : public Filter { // Restrictions.cpp:67-162: keep the positions whose validity bit is set; a column without nulls passes through
   public:
   size_t filter(size_t len, const uint16_t* cur, uint16_t* next, const ArrayView* av, size_t offset) override {
      if (av->nullCount == 0) {
         memcpy(next, cur, len * sizeof(uint16_t));
         return len;
      }
      const uint8_t* bits = reinterpret_cast<const uint8_t*>(av->buffers[0]);
      const size_t first = offset + av->offset;
      uint16_t* w = next;
      for (size_t i = 0; i < len; i++) {
         const size_t bit = first + cur[i];
         *w = cur[i];
         w += (bits[bit >> 3] >> (bit & 7)) & 1;
      }
      return w - next;
   }
};
class Restrictions {
   std::vector<std::pair<std::unique_ptr<Filter>, size_t>> filters;

   public:
   // Restrictions::create (Restrictions.cpp:392-520): constants are parsed per physical column type.
   static std::unique_ptr<Restrictions> create(const std::vector<FilterDescription>& descs, const std::vector<ColumnSchema>& schema) {
      auto r = std::make_unique<Restrictions>();
      for (auto& d : descs) {
         size_t colId = (size_t) -1;
         for (size_t i = 0; i < schema.size(); i++)
            if (schema[i].name == d.columnName) colId = i;
         if (colId == (size_t) -1) throw std::runtime_error("unknown column in filter");
         auto& col = schema[colId];
         if (d.op == FilterOp::NOTNULL) { // Restrictions.cpp:399-405 (value unused)
            r->filters.push_back({std::make_unique<NotNullFilter>(), colId});
            continue;
         }
         switch (col.type) {
            case PhysType::FSB4: { // char(1): compare the 4-byte cell as int32 (Restrictions.cpp:411-424)
               std::string s = std::get<std::string>(d.value);
               int32_t v = 0;
               memcpy(&v, s.data(), std::min<size_t>(4, s.size()));
               r->filters.push_back({std::make_unique<SimpleTypeFilter<int32_t>>(d.op, v), colId});
               break;
            }
            case PhysType::INT32:
               r->filters.push_back({std::make_unique<SimpleTypeFilter<int32_t>>(d.op, (int32_t) std::get<int64_t>(d.value)), colId});
               break;
            case PhysType::INT64:
               r->filters.push_back({std::make_unique<SimpleTypeFilter<int64_t>>(d.op, std::get<int64_t>(d.value)), colId});
               break;
            case PhysType::DATE32:
               r->filters.push_back({std::make_unique<SimpleTypeFilter<int32_t>>(d.op, parseDate32(std::get<std::string>(d.value))), colId});
               break;
            case PhysType::DECIMAL128: { // compared as __int128 at the column's scale (Restrictions.cpp:455-480)
               i128 v;
               if (std::holds_alternative<std::string>(d.value)) {
                  v = parseDecimal(std::get<std::string>(d.value), col.scale);
               } else if (std::holds_alternative<int64_t>(d.value)) {
                  v = (i128) std::get<int64_t>(d.value) * pow10_128(col.scale);
               } else {
                  throw std::runtime_error("unsupported decimal constant type");
               }
               r->filters.push_back({std::make_unique<SimpleTypeFilter<i128>>(d.op, v), colId});
               break;
            }
            case PhysType::STRING:
               r->filters.push_back({std::make_unique<VarLen32Filter>(d.op, std::get<std::string>(d.value)), colId});
               break;
         }
      }
      return r;
   }
   // applyFilters (Restrictions.cpp:365-390): ping-pong between two selection vectors, early exit at 0.
   std::pair<size_t, uint16_t*> applyFilters(size_t offset, size_t length, uint16_t* sel1, uint16_t* sel2, const std::function<const ArrayView*(size_t)>& getArrayView) {
      if (filters.empty()) return {length, BatchView::defaultSelectionVector()};
      uint16_t *cur = sel1, *next = sel2;
      size_t curLen = length;
      bool first = true;
      for (auto& f : filters) {
         curLen = f.first->filter(curLen, first ? BatchView::defaultSelectionVector() : cur, next, getArrayView(f.second), offset);
         std::swap(cur, next);
         if (curLen == 0) return {0, cur};
         first = false;
      }
      return {curLen, cur};
   }
};

// ---- names shared with the _ref adapter (runtime_ref.h) so pipelines.cpp is written once
inline uint16_t* defaultSelVec() { return BatchView::defaultSelectionVector(); }
inline std::unique_ptr<Restrictions> makeRestrictions(const std::vector<FilterDescription>& d, const std::vector<ColumnSchema>& s) { return Restrictions::create(d, s); }
inline const uint8_t* allValidBitmap() { // shared all-valid bitmap for null-free columns (LingoDBTable.cpp:213-218)
   static std::vector<uint8_t> v((1 << 20) / 8, 0xff);
   return v.data();
}
struct WorkerContextBinder { // the port keeps one process-wide context; nothing to bind per worker
   void bind() {}
   void unbind() {}
};
inline int64_t extractYear(int64_t ns) { return extractYearPort(ns); }
inline bool constLikeContains(const oracle::VarLen32& str, std::string_view needle) { return containsPort(str.view(), needle); }
constexpr const char* runtimeKind = "port";

} // namespace oracle::port
