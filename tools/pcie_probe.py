import torch, time
n = 1643251828
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(973059346, dtype=torch.uint8).pin_memory()
d2 = torch.empty(973059346, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for _ in range(2):
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter() - t
t = time.perf_counter(); d2.copy_(h2, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter() - t
t = time.perf_counter()
with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); t3 = time.perf_counter() - t
print("D2H %.1f GB/s (%.1f ms)  H2D %.1f GB/s (%.1f ms)  both concurrently %.1f ms" % (n / t1 / 1e9, t1 * 1e3, 973059346 / t2 / 1e9, t2 * 1e3, t3 * 1e3))
