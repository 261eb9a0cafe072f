// oracle/port/rt_select.h — TEST INFRASTRUCTURE ONLY.  Picks the runtime the restated pipelines run
// on: the reference's own objects (oracle/_ref build, -DORACLE_USE_REF) or the self-contained port.
#pragma once
#ifdef ORACLE_USE_REF
#include "runtime_ref.h"
namespace oracle {
namespace rt = oracle::refrt;
}
#else
#include "runtime_port.h"
namespace oracle {
namespace rt = oracle::port;
}
#endif
