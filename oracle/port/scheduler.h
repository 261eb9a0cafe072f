// oracle/port/scheduler.h — TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-in for the reference's fiber scheduler (src/scheduler/Scheduler.cpp is not buildable
// here: Boost.Context + MLIRContext).  Only the three functions the hot-path runtime needs are
// provided, with the semantics of include/lingodb/scheduler/Scheduler.h:29-40:
//   getNumWorkers(), currentWorkerId(), awaitChildTask(task)
// A task is driven exactly like a scheduler worker drives it (Task.h:7-21): every worker calls
// setup(), then `while (allocateWork()) performWork();`, then teardown().  The calling thread is
// worker 0 and takes part; helper threads are workers 1..N-1.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace oracle::sched {

struct TaskIface { // shape of lingodb::scheduler::Task
   virtual bool allocateWork() = 0;
   virtual void performWork() = 0;
   virtual void setup() {}
   virtual void teardown() {}
   virtual ~TaskIface() {}
};

class Pool {
   std::vector<std::thread> threads;
   std::mutex m;
   std::condition_variable cvStart, cvDone;
   std::function<void()> job; // per-worker body
   uint64_t generation = 0;
   size_t running = 0;
   bool stop = false;
   size_t n;

   void workerMain(size_t id);

   public:
   explicit Pool(size_t n);
   ~Pool();
   size_t size() const { return n; }
   // run body() on every worker (caller = worker 0) and wait for all of them
   void runOnAll(const std::function<void()>& body);
};

void start(size_t numWorkers); // (re)creates the global pool; 0 → hardware_concurrency or ORACLE_PARALLELISM
size_t getNumWorkers();
size_t currentWorkerId();
void runTask(TaskIface& task);

} // namespace oracle::sched
