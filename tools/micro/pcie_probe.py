"""Developer probe: pinned host <-> device copy bandwidth, flat vs pitched (cudaMemcpy2D-like) rows."""
import torch, time
n = 516 * 516 * 516  # one variable of the 512^3 grid
h = torch.empty(n, dtype=torch.float64, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device="cuda")
dp = torch.empty(516 * 516, 544, dtype=torch.float64, device="cuda")  # pitched rows
def t(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
gb = n * 8 / 1e9
print("flat  H2D %.1f GB/s" % (gb / t(lambda: d.copy_(h, non_blocking=True))))
print("flat  D2H %.1f GB/s" % (gb / t(lambda: h.copy_(d, non_blocking=True))))
h2 = h.view(516 * 516, 516)
print("pitch H2D %.1f GB/s" % (gb / t(lambda: dp[:, 14:530].copy_(h2, non_blocking=True))))
print("pitch D2H %.1f GB/s" % (gb / t(lambda: h2.copy_(dp[:, 14:530], non_blocking=True))))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(n, dtype=torch.float64, device="cuda"); h3 = torch.empty(n, dtype=torch.float64, pin_memory=True)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h3.copy_(d2, non_blocking=True)
print("duplex H2D+D2H %.1f GB/s total" % (2 * gb / t(both)))
