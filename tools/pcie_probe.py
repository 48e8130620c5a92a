"""PCIe probe: H2D / D2H rate of a 460 MB copy from page-locked host memory (default vs write-combined allocation),
one direction at a time and both at once — the ceiling of bench.py's `e2e` on this box."""
import ctypes as C
import time

import torch

torch.cuda.init()
rt = C.CDLL("libcudart.so.12")
n = 460 * 1024 * 1024
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
dev2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
for name, flags in (("default pinned", 0), ("write-combined", 4)):
    p, q = C.c_void_p(), C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    assert rt.cudaHostAlloc(C.byref(q), C.c_size_t(n), C.c_uint(0)) == 0
    C.memset(p, 1, n)
    C.memset(q, 0, n)

    def run(h2d, d2h):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            if h2d:
                rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), p, n, 1, C.c_void_p(s1.cuda_stream))
            if d2h:
                rt.cudaMemcpyAsync(q, C.c_void_p(dev2.data_ptr()), n, 2, C.c_void_p(s2.cuda_stream))
        torch.cuda.synchronize()
        return 4 * n / (time.perf_counter() - t0) / 1e9
    run(True, True)
    print(f"{name:16s}: H2D alone {run(True, False):5.1f} GB/s   D2H alone {run(False, True):5.1f} GB/s   "
          f"both at once {run(True, True):5.1f} GB/s per direction", flush=True)
    rt.cudaFreeHost(p)
    rt.cudaFreeHost(q)
