"""Host-side probe for the staging design: CPU quota of the container, NUMA layout, and how the frame-of-reference packer
(csrc/pack_host.cpp) scales with threads over pageable and pinned source memory."""
import ctypes as C
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def sh(cmd):
    import subprocess
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return f"<{e}>"


def main():
    print("cpu.max:", sh("cat /sys/fs/cgroup/cpu.max 2>/dev/null || cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us"))
    print("cpuset:", sh("cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null"), "| affinity:", len(os.sched_getaffinity(0)))
    print(sh("lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|MHz' | head -12"))
    print("cpu.stat:", sh("cat /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\\n' ' '"))
    from lingodb_b200 import capi
    L = capi.lib()
    L.ldb_pack_block.restype = C.c_size_t
    L.ldb_pack_block.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    n_blocks = 4096  # 4096 x 64Ki x 16 B = 4 GiB of decimal128 cells
    rows = n_blocks * 65536
    src = np.empty((rows, 2), np.int64)

    def touch(a, nt=64):
        per = (a.shape[0] + nt - 1) // nt

        def w(i):
            a[i * per:(i + 1) * per, 0] = 1234567
            a[i * per:(i + 1) * per, 1] = 0
        ts = [threading.Thread(target=w, args=(i,)) for i in range(nt)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    touch(src)
    import torch
    pin = None
    if torch.cuda.is_available():
        pin = torch.empty((rows, 2), dtype=torch.int64, pin_memory=True).numpy()
        touch(pin)

    def run(a, nt):
        dsts = [np.empty(65536 * 8, np.uint8) for _ in range(nt)]
        cursor = [0]
        lock = threading.Lock()

        def w(i):
            mn, wd = C.c_int64(), C.c_int32()
            while True:
                with lock:
                    b = cursor[0]
                    cursor[0] += 8
                if b >= n_blocks:
                    return
                for k in range(b, min(b + 8, n_blocks)):
                    L.ldb_pack_block(a.ctypes.data + k * 65536 * 16, 2, 65536, dsts[i].ctypes.data, C.byref(mn), C.byref(wd))
        ts = [threading.Thread(target=w, args=(i,)) for i in range(nt)]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        return rows * 16 / (time.perf_counter() - t0) / 1e9
    for name, a in (("pageable", src), ("pinned", pin)):
        if a is None:
            continue
        for nt in (4, 8, 16, 24, 32, 48, 64, 96, 128):
            run(a, nt)
            print(f"pack {name:8s} threads {nt:3d}: {run(a, nt):7.1f} GB/s of decimal128 source", flush=True)
    print("cpu.stat after:", sh("cat /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\\n' ' '"))


if __name__ == "__main__":
    main()
