// oracle/port/scheduler.cpp — TEST INFRASTRUCTURE ONLY.  See scheduler.h.
#include "scheduler.h"

#include <cstdlib>

#ifdef ORACLE_USE_REF
#include "lingodb/runtime/ExecutionContext.h"
#include "lingodb/scheduler/Scheduler.h"
#endif

namespace oracle::sched {
namespace {
std::unique_ptr<Pool> pool;
thread_local size_t workerId = 0;
} // namespace

Pool::Pool(size_t n) : n(n) {
   for (size_t i = 1; i < n; i++) threads.emplace_back([this, i] { workerMain(i); });
}
Pool::~Pool() {
   {
      std::unique_lock<std::mutex> l(m);
      stop = true;
   }
   cvStart.notify_all();
   for (auto& t : threads) t.join();
}
void Pool::workerMain(size_t id) {
   workerId = id;
   uint64_t seen = 0;
   while (true) {
      std::function<void()> body;
      {
         std::unique_lock<std::mutex> l(m);
         cvStart.wait(l, [&] { return stop || generation != seen; });
         if (stop) return;
         seen = generation;
         body = job;
      }
      body();
      {
         std::unique_lock<std::mutex> l(m);
         if (--running == 0) cvDone.notify_all();
      }
   }
}
void Pool::runOnAll(const std::function<void()>& body) {
   {
      std::unique_lock<std::mutex> l(m);
      job = body;
      running = n - 1;
      generation++;
   }
   cvStart.notify_all();
   body(); // worker 0 = caller
   std::unique_lock<std::mutex> l(m);
   cvDone.wait(l, [&] { return running == 0; });
}

void start(size_t numWorkers) {
   if (numWorkers == 0) {
      if (const char* e = std::getenv("ORACLE_PARALLELISM")) numWorkers = std::strtoul(e, nullptr, 10);
   }
   if (numWorkers == 0) numWorkers = std::max(1u, std::thread::hardware_concurrency());
   if (pool && pool->size() == numWorkers) return;
   pool.reset();
   workerId = 0;
   pool = std::make_unique<Pool>(numWorkers);
}
size_t getNumWorkers() {
   if (!pool) start(0);
   return pool->size();
}
size_t currentWorkerId() { return workerId; }
void runTask(TaskIface& task) {
   if (!pool) start(0);
   pool->runOnAll([&] {
      task.setup();
      while (task.allocateWork()) task.performWork();
      task.teardown();
   });
}
} // namespace oracle::sched

#ifdef ORACLE_USE_REF
// The three symbols the verbatim-compiled reference runtime objects leave unresolved.
namespace lingodb::scheduler {
size_t getNumWorkers() { return oracle::sched::getNumWorkers(); }
size_t currentWorkerId() { return oracle::sched::currentWorkerId(); }
void awaitChildTask(std::unique_ptr<Task> task) {
   struct Adapter : oracle::sched::TaskIface {
      Task& t;
      explicit Adapter(Task& t) : t(t) {}
      bool allocateWork() override { return t.allocateWork(); }
      void performWork() override { t.performWork(); }
      void setup() override { t.setup(); }
      void teardown() override { t.teardown(); }
   } a(*task);
   // Task::teardown() clears the thread_local context of every worker, including the caller's
   // (worker 0): the reference's fiber switch restores it, here it is restored by hand.
   auto* callerCtx = lingodb::runtime::getCurrentExecutionContext();
   oracle::sched::runTask(a);
   lingodb::runtime::setCurrentExecutionContext(callerCtx);
}
} // namespace lingodb::scheduler
#endif
