#!/usr/bin/env python
"""Host link experiment: minibatch rows pulled by a kernel out of a registered (mapped)
host array vs. one contiguous cudaMemcpyAsync of the same bytes from pinned memory.
C2 shape: 1024 rows x 4 KB drawn from a 1M x 1024 f32 array."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import device, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream  # noqa: E402

device.set_device(0)
s = Stream.create()
device.set_current_stream(s)
N, d, n = int(os.environ.get("N", 1_000_000)), 1024, 1024
rng = np.random.default_rng(0)
X = np.empty((N, d), np.float32)
X[:] = np.arange(N, dtype=np.float32)[:, None]
y = np.arange(N, dtype=np.float32)
t0 = time.perf_counter()
Xm = ops.host_register(X)
ym = ops.host_register(y)
print(f"register {X.nbytes / 1e9:.1f} GB: {time.perf_counter() - t0:.2f} s", flush=True)
K = 64
idx_h = rng.integers(0, N, (K, n)).astype(np.int32)
idx = DA.from_numpy(idx_h)
dst = DA((n, d), np.float32)
lab = DA((n,), np.float32)
e0, e1 = Event(), Event()
for ctas in (16, 48):
  for rep in range(2):
    e0.record(s)
    for k in range(K):
      ops.pull_rows(Xm, ym, idx.ptr + k * n * 4, n, 0, n, d, dst, lab, n_ctas=ctas)
    e1.record(s)
    e1.sync()
  ms = e0.elapsed_ms(e1) / K
  print(f"pull ctas={ctas:3d}: {ms * 1e3:7.1f} us per minibatch, {n * d * 4 / ms / 1e6:6.1f} GB/s", flush=True)
got = dst.numpy()
assert np.array_equal(got[:, 0], idx_h[K - 1].astype(np.float32)), "wrong rows"
assert np.array_equal(lab.numpy(), idx_h[K - 1].astype(np.float32)), "wrong labels"
# the DMA reference: one contiguous copy of the same size from pinned memory
from jax_sgmc_b200 import _lib  # noqa: E402
import ctypes as C  # noqa: E402
hp = C.c_void_p()
_lib.call("sgmc_host_alloc", C.byref(hp), n * d * 4)
for rep in range(2):
  e0.record(s)
  for k in range(K):
    _lib.call("sgmc_memcpy_h2d", C.c_void_p(dst.ptr), hp, n * d * 4, s.handle)
  e1.record(s)
  e1.sync()
ms = e0.elapsed_ms(e1) / K
print(f"memcpy H2D 4.2 MB: {ms * 1e3:7.1f} us, {n * d * 4 / ms / 1e6:6.1f} GB/s")
# pull beside the compute kernels: who slows whom?
C_ = 4096
Xd, yd, _ = ops.synth_logistic_data(0, 100_000, d)
spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                    prior_scale=10.0, x_absmax=ops.absmax(Xd))
theta = DA.zeros((C_, d))
vv = DA.full((C_, d), 1.0)
U, var, g = DA((C_,), np.float32), DA((C_,), np.float32), DA((C_, d), np.float32)
di = DA.from_numpy(rng.integers(0, 100_000, n).astype(np.int32))
ws = ops.glm_workspace(C_, n, d, "tc_parity")
kk = [ops.prng_keys(range(C_)), DA((C_, 2), np.uint32)]
s2 = Stream.create()
f0, f1 = Event(), Event()


def potential(k):
  ops.glm_potential_grad(spec, theta, Xd, yd, di, 100_000, U, var, g, workspace=ws, path="tc_parity")


def update(k):
  ops.sgld_update(theta, g, kk[k % 2], kk[(k + 1) % 2], [d], 1e-3, 1.0, v=vv)


def step(k):
  ops.glm_sgld_step(spec, theta, Xd, yd, di, 100_000, U, var, g, kk[k % 2], kk[(k + 1) % 2], 1e-3,
                    1.0, v=vv, workspace=ws, path="tc_parity", write_grad=False,
                    carry=ops.STEP_CARRY_INIT if k == 0 else ops.STEP_CARRY)


for name, fn in (("update kernel", update), ("carried step", step)):
  for ctas in (0, 8, 16, 48):
    for rep in range(2):
      e0.record(s)
      f0.record(s2)
      for k in range(K):
        fn(k)
        if ctas:
          ops.pull_rows(Xm, ym, idx.ptr + k * n * 4, n, 0, n, d, dst, lab, n_ctas=ctas, stream=s2)
      e1.record(s)
      f1.record(s2)
      e1.sync()
      f1.sync()
    print(f"{name}, pull ctas={ctas:2d}: compute {e0.elapsed_ms(e1) / K * 1e3:6.1f} us per call, "
          f"pull stream {f0.elapsed_ms(f1) / K * 1e3:6.1f} us per minibatch", flush=True)
