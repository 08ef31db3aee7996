"""development helper: W worker threads, each committing device-resident random columns (from_values) in a loop —
how close concurrent small commits get to the kernel rates.  usage: _commit_farm.py [log_rows=12] [n_cols=135] [workers=24] [reps=40]"""
import os
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path[:0] = ['.', 'tests', 'tools']
import torch

import city_rollup_b200 as m

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n_cols = int(sys.argv[2]) if len(sys.argv) > 2 else 135
W = int(sys.argv[3]) if len(sys.argv) > 3 else 24
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 40
g = torch.Generator(device="cuda")
g.manual_seed(7)
t = torch.empty((n_cols, 1 << log_n), dtype=torch.int64, device="cuda")
t.random_(0, 2**62, generator=g)
torch.cuda.synchronize()
ctxs = [m.Context(0) for _ in range(W)]
for c in ctxs:
    c.set_blocking_sync(W > 1)
bar = threading.Barrier(W + 1)


def worker(i):
    c = ctxs[i]
    for _ in range(3):
        m.PolynomialBatch.from_values_device(c, t.data_ptr(), n_cols, log_n, 3, 4).free()
    c.synchronize()
    bar.wait()
    for _ in range(reps):
        m.PolynomialBatch.from_values_device(c, t.data_ptr(), n_cols, log_n, 3, 4).free()
    c.synchronize()


th = [threading.Thread(target=worker, args=(i,)) for i in range(W)]
for x in th:
    x.start()
bar.wait()
t0 = time.perf_counter()
for x in th:
    x.join()
dt = time.perf_counter() - t0
n = W * reps
perms = n * ((-(-n_cols // 8)) * (8 << log_n) + (8 << log_n) - 16)
print("workers %d: %.1f commits/s (2^%d x %d), %.3f ms per commit, %.3f Gperm/s" % (W, n / dt, log_n, n_cols, dt / n * 1e3, perms / dt / 1e9))
