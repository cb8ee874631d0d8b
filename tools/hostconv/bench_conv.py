import ctypes as C, time, torch, numpy as np, threading, os
L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libconv.so"))
L.conv_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int]
n = 25 * 2048 * 2048 * 2 // 2    # half of the job: 105M doubles (839 MB)
src = torch.empty(n, dtype=torch.float64, pin_memory=True); src.normal_()
dst = torch.empty(n, dtype=torch.float32, pin_memory=True)
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for T in (1, 2, 4, 8, 12, 16):
    best = 1e9
    for rep in range(3):
        t = time.perf_counter(); L.conv_rows(dst.data_ptr(), src.data_ptr(), n, 0.5, T); best = min(best, time.perf_counter() - t)
    print("threads %2d: %.1f ms  %.1f GB/s in" % (T, best * 1e3, n * 8 / best / 1e9))
dev = torch.empty(n, dtype=torch.float64, device="cuda")
src2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
for T in (4, 8, 12, 14):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev.copy_(src2, non_blocking=True)
    L.conv_rows(dst.data_ptr(), src.data_ptr(), n, 0.5, T)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("concurrent T=%d: conv %.1f ms (%.1f GB/s), copy done at %.1f ms (%.1f GB/s)" % (T, (t1 - t0) * 1e3, n * 8 / (t1 - t0) / 1e9, (t2 - t0) * 1e3, n * 8 / (t2 - t0) / 1e9))
