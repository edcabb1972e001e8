import torch, time
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("H2D 512 MiB pinned: %.2f ms = %.1f GB/s" % (ms, n / ms / 1e6))
# chunked 32 MiB copies
e0.record()
for _ in range(5):
    for k in range(16):
        d[k * (32 << 20):(k + 1) * (32 << 20)].copy_(h[k * (32 << 20):(k + 1) * (32 << 20)], non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("H2D 16 x 32 MiB pinned: %.2f ms = %.1f GB/s" % (ms, n / ms / 1e6))
