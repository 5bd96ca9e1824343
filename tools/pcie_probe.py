"""Host<->device copy ceiling of this box (pinned memory): H2D alone, D2H alone, both at once.  Bounds bench.py's e2e."""
import time
import torch

n = 4 << 30
dev = torch.device("cuda", 0)
h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


t = timed(h2d)
print("H2D alone   %.1f GB/s" % (n / t / 1e9))
t = timed(d2h)
print("D2H alone   %.1f GB/s" % (n / t / 1e9))
t = timed(both)
print("both at once: %.1f GB/s per direction (%.1f total)" % (n / t / 1e9, 2 * n / t / 1e9))
