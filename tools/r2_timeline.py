#!/usr/bin/env python
"""Phase timeline of k_glm_tc_pair at the C2 shape (SGMC_OPT_TC_TIMELINE):
per CTA pair, ns since the earliest kernel start of
  [0] start  [1] epilogue warps done  [2,3] GEMM2 producer: begins to wait for R /
  R available (last GEMM2 tile)  [4,5] MMA issue of tile 0 / last tile finished  [7] kernel entry  [31] kernel exit
  per tile i < 4 (slots 8+5i..12+5i): epilogue ready / accumulator full / chunks
  done / stores complete / published+finalised."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import _lib, device, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA  # noqa: E402

device.set_device(0)
Cc, d, n, N = [int(v) for v in os.environ.get("SHAPE", "4096,1024,1024,100000").split(",")]
X, y, _ = ops.synth_logistic_data(0, N, d)
theta = DA.from_numpy((np.random.default_rng(0).standard_normal((Cc, d)) * 0.3).astype(np.float32))
idx = DA((n,), np.int32)
dk = [DA.from_numpy(ops.prng_key(0)), DA((2,), np.uint32)]
ops.minibatch_draw(dk[0], dk[1], idx, N)
U, var, g = DA((Cc,), np.float32), DA((Cc,), np.float32), DA((Cc, d), np.float32)
spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                    prior_scale=10.0, x_absmax=ops.absmax(X))
lib = _lib.load()
lib.sgmc_debug_pair_timeline.argtypes = [C.c_void_p, C.c_int]
ops.set_option(6, 1)
for path in os.environ.get("PATHS", "tc_parity,tc_throughput").split(","):
  ws = ops.glm_workspace(Cc, n, d, path)
  if os.environ.get("MODE", "potential") == "step":
    # the carried Langevin step (the benchmark's hot path): prior gradient in the update
    v = DA.full((Cc, d), 1.0)
    kk = [ops.prng_keys(range(Cc)), DA((Cc, 2), np.uint32)]
    for k in range(6):
      ops.glm_sgld_step(spec, theta, X, y, idx, N, U, var, g, kk[k % 2], kk[(k + 1) % 2], 1e-3,
                        1.0, v=v, workspace=ws, path=path, write_grad=False,
                        carry=ops.STEP_CARRY_INIT if k == 0 else ops.STEP_CARRY)
      device.synchronize()
  else:
    from jax_sgmc_b200.device import Event, current_stream
    e0, e1, e2 = Event(), Event(), Event()
    for rep in range(3):
      e0.record(current_stream())
      ops.glm_potential_grad(spec, theta, X, y, idx, N, U, var, g, workspace=ws, path=path)
      e1.record(current_stream())
      device.synchronize()
    print(f"CUDA events around the whole op (prepare + pair kernel): {e0.elapsed_ms(e1) * 1e3:.1f} us")
  buf = (C.c_ulonglong * (80 * 32))()
  lib.sgmc_debug_pair_timeline(buf, 80 * 32)
  t = np.array(buf[:], dtype=np.int64).reshape(80, 32)
  used = t[:, 0] > 0
  t0 = t[used, 0].min()
  rel = np.where(t > 0, t - t0, -1)
  print(f"== {path}: {used.sum()} pairs; columns = slots 0..31 (ns since first start, -1 unused)")
  for p in (0, 33, 34, 54, 73):
    if p < 80 and used[p]:
      print(f"pair {p:2d}:", " ".join(f"{v:6d}" for v in rel[p]))
  if os.environ.get("ALL_PAIRS"):
    order = np.argsort(rel[:, 31])
    print("pairs by exit time: pair exit | R-wait-begin R-ready | t1: ready accfull chunksdone | noise-done")
    for p in order:
      if used[p]:
        print(f"  pair {p:2d} exit {rel[p, 31]:6d} | {rel[p, 2]:6d} {rel[p, 3]:6d} | "
              f"{rel[p, 13]:6d} {rel[p, 14]:6d} {rel[p, 15]:6d} | t0: {rel[p, 9]:6d} {rel[p, 10]:6d} | {rel[p, 1]:6d}")
  for sl in range(32):
    col = rel[used, sl]
    col = col[col >= 0]
    if col.size:
      print(f"slot {sl:2d}: min {col.min():6d}  median {int(np.median(col)):6d}  max {col.max():6d}")
