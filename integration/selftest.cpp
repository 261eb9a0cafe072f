// integration/selftest.cpp — links the reference-side shim (GPUPipeline.cpp) against the reference's OWN runtime objects
// (oracle/_ref: ExecutionContext, VarLen32 arenas) and the product library, and drives one serialised step through
// lingodb::runtime::GPUPipeline::run(VarLen32): what a lowering that emits `GPUPipeline::run(descr)` calls would execute.
// Without a CUDA device the step must surface the library's error as the std::runtime_error the reference's runtime
// functions throw (Hashtable.cpp:106) — that is what tests/test_integration_shim.py checks on the CPU box.
#include "GPUPipeline.h"

#include <cstdio>
#include <stdexcept>

using namespace lingodb::runtime;

int main() {
   // the ExecutionContext wants a Session (catalog) this path never touches
   alignas(16) static unsigned char fakeSession[512] = {};
   int rc = 1;
   {
      ExecutionContext ctx(*reinterpret_cast<lingodb::runtime::Session*>(fakeSession));
      setCurrentExecutionContext(&ctx);
      try {
         // longer than 12 bytes: the VarLen32 lives in the ExecutionContext's string arena like a JIT'd constant would
         GPUPipeline::run(VarLen32::fromString("{\"step\": \"pipeline\", \"kind\": \"scan_group_by\", \"source\": \"no such table\"}"));
         std::puts("step ran");
      } catch (const std::runtime_error& e) {
         std::printf("runtime_error: %s\n", e.what());
         rc = 0;
      }
      setCurrentExecutionContext(nullptr);
   }
   return rc;
}
