#!/usr/bin/env python
"""Host side of the streaming loader on this box: index draw (NumPy PCG64 pipeline) and
the threaded row gather into a staging buffer (sgmc_host_gather_batches), with the source
array in ordinary pages and in transparent huge pages."""
import mmap
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import data, ops  # noqa: E402

N, d, n = int(os.environ.get("N", 1000000)), 1024, 1024
print("cpus", os.cpu_count(), "thp:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(),
      flush=True)
rng = np.random.default_rng(0)
X = rng.random((N, d), dtype=np.float32)
y = rng.random(N, dtype=np.float32)
loader = data.StreamingNumpyDataLoader(x=X, y=y)
init, get, _ = data.random_reference_data(loader, 64, n)
src = get.scan_source(init(), 1000)
t0 = time.perf_counter()
for _ in range(5):
  idx = src["draw"](64)
print(f"draw 64 batches: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms", flush=True)
dst = np.empty(64 * (n * d + n), np.float32)


def run(Xs, label):
  for T in (1, 4, 8, 16, 32, 64):
    if T > (os.cpu_count() or 1):
      break
    ops.host_gather_batches(dst.ctypes.data, Xs, y, idx, 0, n, T)
    t0 = time.perf_counter()
    for _ in range(3):
      ops.host_gather_batches(dst.ctypes.data, Xs, y, idx, 0, n, T)
    dt = (time.perf_counter() - t0) / 3
    print(f"{label}: gather 64 batches, {T} threads: {dt * 1e3:.2f} ms = "
          f"{64 * n * d * 4 / dt / 1e9:.1f} GB/s", flush=True)


run(X, "4K pages")
size = (X.nbytes + (2 << 20) - 1) & ~((2 << 20) - 1)
mm = mmap.mmap(-1, size + (2 << 20))
try:
  mm.madvise(mmap.MADV_HUGEPAGE)
except Exception as e:
  print("madvise failed", e)
Xh = np.frombuffer(mm, np.float32, N * d).reshape(N, d)
Xh[...] = X
run(Xh, "THP")
a = np.empty(67_000_000, np.float32)
b = np.ones_like(a)
np.copyto(a, b)
t0 = time.perf_counter()
np.copyto(a, b)
print(f"contiguous 1-thread copy: {a.nbytes / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
