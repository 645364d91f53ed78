import torch, time
for mb in (1, 8, 45, 256):
    x = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    y = torch.empty_like(x, device="cuda")
    for _ in range(3): y.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): y.copy_(x, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("H2D pinned %4d MB: %.3f ms  %.1f GB/s" % (mb, ms, mb / 1024 / (ms * 1e-3)))
x = torch.empty(45 * 1024 * 1024, dtype=torch.uint8)
y = torch.empty_like(x, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter(); 
for _ in range(5): y.copy_(x)
torch.cuda.synchronize(); print("pageable 45MB: %.1f GB/s" % (45 / 1024 * 5 / (time.perf_counter() - t)))
