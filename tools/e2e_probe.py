"""Timeline of the e2e leg (two task slots, host buffers through the C ABI): when does each slot upload / sort+download,
and at what rate -- with and without one-transfer-per-direction locks.  usage: python tools/e2e_probe.py [records]"""
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
import tez_b200 as T
from tez_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda", 0)
d_kv = synth.gen_c2(0, n, seed=2, device=dev)
h_kv = torch.empty(n * 80, dtype=torch.uint8, pin_memory=True)
h_kv.copy_(d_kv)
del d_kv
cap = n * 82 + 64 * 16 + 8192
h_outs = [torch.empty(cap, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
sorters = [T.GpuSorter(64, fixed=(16, 64), device=0) for _ in range(2)]
torch.cuda.synchronize()


def run(mode, steps):
    up, down = threading.Lock(), threading.Lock()
    log = []
    t_origin = time.perf_counter()

    class nolock:
        def __enter__(self): return self
        def __exit__(self, *a): return False

    ev = [threading.Event() for _ in range(3)]
    ev[0].set()
    if mode != "stagger":
        ev[1].set()

    def worker(k):
        s, ho = sorters[k], h_outs[k].numpy()
        ev[k].wait()
        for i in range(steps):
            s.reset()
            with (up if mode == "locked" else nolock()):
                t0 = time.perf_counter()
                s.collect_fixed(h_kv.data_ptr(), n=n)
                t1 = time.perf_counter()
            if i == 0:
                ev[k + 1].set()
            with (down if mode == "locked" else nolock()):
                t2 = time.perf_counter()
                out, _, _, st = s.flush_to_memory(out=ho)
                t3 = time.perf_counter()
            log.append((k, i, t0 - t_origin, t1 - t_origin, t2 - t_origin, t3 - t_origin))

    ths = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    total = time.perf_counter() - t0
    print("mode %s: %.1f ms per step (%.1f GB/s of KV)" % (mode, total / (2 * steps) * 1e3, n * 80 / (total / (2 * steps)) / 1e9))
    for k, i, a, b, c, d in sorted(log, key=lambda r: r[2]):
        print("  slot %d step %d: up %7.1f..%7.1f ms (%.1f GB/s)   flush %7.1f..%7.1f ms (%.1f GB/s incl. sort)" % (
            k, i, a * 1e3, b * 1e3, n * 80 / (b - a) / 1e9, c * 1e3, d * 1e3, n * 82 / (d - c) / 1e9))


run("free", 1)      # warm-up
run("free", 3)
run("stagger", 5)
run("locked", 3)
# single slot alone: the rates without any concurrency
sorters[1].reset()
t0 = time.perf_counter(); sorters[1].collect_fixed(h_kv.data_ptr(), n=n); t1 = time.perf_counter()
out, _, _, st = sorters[1].flush_to_memory(out=h_outs[1].numpy()); t2 = time.perf_counter()
print("alone: up %.1f ms (%.1f GB/s), flush %.1f ms (%.1f GB/s incl. %.1f ms sort)" % ((t1 - t0) * 1e3, n * 80 / (t1 - t0) / 1e9, (t2 - t1) * 1e3, n * 82 / (t2 - t1) / 1e9, st["ms_total"]))
