"""PCIe ceiling of the box (what meepcu_sync_* can hope for): pinned (cudaHostAlloc) vs registered-in-place host memory."""
import time
import numpy as np
import torch

n = 1 << 28  # 1 GiB of float32
dev = torch.empty(n, dtype=torch.float32, device="cuda")
pinned = torch.empty(n, dtype=torch.float32).pin_memory()
arr = np.zeros(n, dtype=np.float32)
reg = torch.from_numpy(arr)
torch.cuda.cudart().cudaHostRegister(reg.data_ptr(), reg.numel() * 4, 0)


def bw(fn):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    return 3 * n * 4 / (time.perf_counter() - t0) / 1e9


print("H2D cudaHostAlloc  %.1f GB/s" % bw(lambda: dev.copy_(pinned, non_blocking=True)))
print("D2H cudaHostAlloc  %.1f GB/s" % bw(lambda: pinned.copy_(dev, non_blocking=True)))
print("H2D registered     %.1f GB/s" % bw(lambda: torch.cuda.cudart().cudaMemcpy(dev.data_ptr(), reg.data_ptr(), n * 4, 1)))
print("D2H registered     %.1f GB/s" % bw(lambda: torch.cuda.cudart().cudaMemcpy(reg.data_ptr(), dev.data_ptr(), n * 4, 2)))
