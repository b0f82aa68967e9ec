"""(needs the debug build: SGMC_TC_DEBUG=1 python -c "import __graft_entry__ as g; g.build()")
Phase timers (globaltimer, CTA (0,0), thread 64) of the fused GEMM2 + update kernel."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jax_sgmc_b200 import _lib, device, ops
from jax_sgmc_b200.device import DeviceArray as DA

device.set_device(0)
Cc, d, n, N = 4096, 1024, 1024, 100000
X, y, _ = ops.synth_logistic_data(0, N, d)
theta = DA.from_numpy((np.random.default_rng(0).standard_normal((Cc, d)) * 0.3).astype(np.float32))
v = DA.full((Cc, d), 1.0)
idx = DA((n,), np.int32)
dk = [DA.from_numpy(ops.prng_key(0)), DA((2,), np.uint32)]
ops.minibatch_draw(dk[0], dk[1], idx, N)
U, var, g = DA((Cc,), np.float32), DA((Cc,), np.float32), DA((Cc, d), np.float32)
kk = [ops.prng_keys(range(Cc)), DA((Cc, 2), np.uint32)]
spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                    prior_scale=10.0, x_absmax=ops.absmax(X))
lib = _lib.load()
lib.sgmc_debug_tc_timers.argtypes = [C.c_void_p]
ws = ops.glm_workspace(Cc, n, d, "tc_parity")
for rep in range(4):
  ops.glm_sgld_step(spec, theta, X, y, idx, N, U, var, g, kk[rep % 2], kk[(rep + 1) % 2], 1e-3,
                    1.0, v=v, workspace=ws, path="tc_parity", write_grad=False)
  device.synchronize()
  t = (C.c_ulonglong * 10)()
  lib.sgmc_debug_tc_timers(t)
  a = [int(t[i]) - int(t[0]) for i in range(10)]
print("fused GEMM2+update, ns from start: setup", a[1], "noise_done", a[8], "acc_ready", a[2],
      "tmem_ld", a[4], "staged", a[5], "half0 done", a[6], "epilogue done", a[7], "end", a[3])
