"""Bidirectional PCIe rate under the conditions the e2e path can choose: where the process runs (NUMA),
how the pinned memory was allocated (plain / write-combined), and the copy granularity."""
import ctypes as C
import glob
import os
import sys
import time

import torch

n = 1 << 30
rt = C.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]


def nodes():
    out = []
    for d in sorted(glob.glob("/sys/devices/system/node/node*")):
        try:
            out.append((os.path.basename(d), open(d + "/cpulist").read().strip()))
        except OSError:
            pass
    return out


def gpu_cpus(index=0):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    return {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), nbytes, flags)
    assert rc == 0, rc
    C.memset(p, 1, nbytes)            # first touch here, on the CPUs this process is bound to
    return p.value


def run(tag, h_src, h_dst, chunk):
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def go(h2d, d2h, reps=3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for off in range(0, n, chunk):
                m = min(chunk, n - off)
                if h2d:
                    rt.cudaMemcpyAsync(d_a.data_ptr() + off, h_src + off, m, 1, C.c_void_p(s1.cuda_stream))
                if d2h:
                    rt.cudaMemcpyAsync(h_dst + off, d_b.data_ptr() + off, m, 2, C.c_void_p(s2.cuda_stream))
        torch.cuda.synchronize()
        return reps * n / (time.perf_counter() - t0) / 1e9

    go(True, True, 1)
    print(f"{tag:44s} H2D {go(True, False):5.1f}  D2H {go(False, True):5.1f}  both {go(True, True):5.1f} GB/s each way", flush=True)


print("NUMA nodes:", nodes())
print("affinity now:", len(os.sched_getaffinity(0)), "CPUs; GPU 0 local CPUs:", sorted(gpu_cpus(0))[:4], "...", len(gpu_cpus(0)))
torch.cuda.init()
for bound in (False, True):
    if bound:
        cp = gpu_cpus(0) & os.sched_getaffinity(0)
        if not cp:
            print("no GPU-local CPUs in the affinity mask")
            break
        os.sched_setaffinity(0, cp)
    b = "bound" if bound else "unbound"
    for name, fl_src in (("plain", 0), ("write-combined source", 4)):
        src = host_alloc(n, fl_src)
        dst = host_alloc(n, 0)
        run(f"{b}, {name}, 1 GiB copies", src, dst, n)
        run(f"{b}, {name}, 100 MB copies", src, dst, 100_000_000)
        rt.cudaFreeHost(C.c_void_p(src)); rt.cudaFreeHost(C.c_void_p(dst))
