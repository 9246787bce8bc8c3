"""PCIe ceiling of the box (what meepcu_sync_* can hope for): cudaHostAlloc'd memory vs caller memory page-locked in
place (cudaHostRegister, 4 KB pages vs transparent huge pages), whole-buffer copies vs 64 MB pieces, and the rate at
which the host cores can stage a pageable array through pinned buffers (parallel memcpy)."""
import ctypes as C
import mmap
import os
import time

import numpy as np
import torch

rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
rt.cudaHostUnregister.argtypes = [C.c_void_p]
rt.cudaDeviceSynchronize.argtypes = []

n = 1 << 30  # bytes
torch.cuda.init()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
dptr = dev.data_ptr()


def bw(fn, reps=3):
    fn(); rt.cudaDeviceSynchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    rt.cudaDeviceSynchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


def both(name, hptr):
    h2d = bw(lambda: rt.cudaMemcpy(dptr, hptr, n, 1))
    d2h = bw(lambda: rt.cudaMemcpy(hptr, dptr, n, 2))
    piece = 64 << 20
    def pieces(kind):
        for o in range(0, n, piece):
            if kind == 1: rt.cudaMemcpyAsync(dptr + o, hptr + o, piece, 1, None)
            else: rt.cudaMemcpyAsync(hptr + o, dptr + o, piece, 2, None)
    print("%-46s H2D %5.1f  D2H %5.1f   (64 MB pieces: %5.1f / %5.1f) GB/s" % (name, h2d, d2h, bw(lambda: pieces(1)), bw(lambda: pieces(2))), flush=True)


pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
both("cudaHostAlloc (torch pin_memory)", pinned.data_ptr())

a = np.ones(n, dtype=np.uint8)  # touched: 4 KB pages unless THP is 'always'
assert rt.cudaHostRegister(a.ctypes.data, n, 0) == 0
both("numpy array + cudaHostRegister", a.ctypes.data)
rt.cudaHostUnregister(a.ctypes.data)

try:
    print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
except Exception as e:
    print("THP: ?", e)
m = mmap.mmap(-1, n + (2 << 20))
base = C.addressof(C.c_char.from_buffer(m))
al = (base + (2 << 20) - 1) & ~((2 << 20) - 1)
libc = C.CDLL("libc.so.6", use_errno=True)
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
r = libc.madvise(al, n, 14)  # MADV_HUGEPAGE
C.memset(al, 1, n)
assert rt.cudaHostRegister(al, n, 0) == 0
both("mmap + MADV_HUGEPAGE(rc=%d) + cudaHostRegister" % r, al)
rt.cudaHostUnregister(al)

# staging rate: pageable numpy -> pinned buffer with k threads (numpy copies release the GIL)
import threading
src = np.ones(n, dtype=np.uint8)
dst = pinned.numpy()
for k in (1, 4, 8, 16, 32):
    if k > (os.cpu_count() or 1) * 2: break
    def work(i):
        s = slice(i * n // k, (i + 1) * n // k)
        np.copyto(dst[s], src[s])
    t0 = time.perf_counter()
    for _ in range(3):
        th = [threading.Thread(target=work, args=(i,)) for i in range(k)]
        [t.start() for t in th]; [t.join() for t in th]
    print("host memcpy pageable -> pinned, %2d threads: %5.1f GB/s" % (k, 3 * n / (time.perf_counter() - t0) / 1e9), flush=True)
print("cpus:", os.cpu_count())
