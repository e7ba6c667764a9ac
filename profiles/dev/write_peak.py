"""Pure-write and copy bandwidth of this B200 with torch primitives (context for the copy kernel's write stream)."""
import torch
dev = torch.device("cuda:0")
n = 8 << 30
x = torch.empty(n, dtype=torch.uint8, device=dev)
y = torch.empty(n, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
t = timeit(lambda: x.fill_(46))
print("fill_  (write only)   %.2f ms  %.0f GB/s written" % (t, n / t / 1e6))
t = timeit(lambda: x.zero_())
print("zero_  (memset)       %.2f ms  %.0f GB/s written" % (t, n / t / 1e6))
t = timeit(lambda: y.copy_(x))
print("copy_  (read+write)   %.2f ms  %.0f GB/s read+written, %.0f GB/s written" % (t, 2 * n / t / 1e6, n / t / 1e6))
xv = x.view(torch.int64)
t = timeit(lambda: xv.sum())
print("sum    (read only)    %.2f ms  %.0f GB/s read" % (t, n / t / 1e6))
