#!/usr/bin/env python
"""Where one carried pSGLD step at the C2 shape spends its time: CUDA events between
the step's launches (SGMC_OPT_STEP_PROFILE; synchronises every step) next to the
free-running step time.  Usage: r2_step_profile.py [SGMC_OPTIONS string]"""
import os
import sys

if len(sys.argv) > 1:
  os.environ["SGMC_OPTIONS"] = sys.argv[1]
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import device, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream  # noqa: E402

device.set_device(0)
stream = Stream.create()
device.set_current_stream(stream)
C, d, n, N = [int(v) for v in os.environ.get("SHAPE", "4096,1024,1024,1000000").split(",")]
path = os.environ.get("TCPATH", "tc_parity")
X, y, _ = ops.synth_logistic_data(0, N, d)
theta, v, grad = DA.zeros((C, d)), DA.full((C, d), 1.0), DA.zeros((C, d))
keys = [ops.prng_keys(range(C)), DA((C, 2), np.uint32)]
dkey = [DA.from_numpy(ops.prng_key(0)), DA((2,), np.uint32)]
idx = DA((n,), np.int32)
U, var = DA((C,), np.float32), DA((C,), np.float32)
spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                    prior_scale=10.0, x_absmax=ops.absmax(X))
ws = ops.glm_workspace(C, n, d, path)
st = {"k": 0}


def step():
  k = st["k"]
  ops.minibatch_draw(dkey[k % 2], dkey[(k + 1) % 2], idx, N)
  ops.glm_sgld_step(spec, theta, X, y, idx, N, U, var, grad, keys[k % 2], keys[(k + 1) % 2],
                    1e-3, 1.0, v=v, workspace=ws, path=path, write_grad=False,
                    carry=ops.STEP_CARRY if k else ops.STEP_CARRY_INIT)
  st["k"] = k + 1


for _ in range(300):
  step()
stream.sync()
e0, e1 = Event(), Event()
K = 1000
e0.record(stream)
for _ in range(K):
  step()
e1.record(stream)
e1.sync()
print(f"free-running: {e0.elapsed_ms(e1) * 1e3 / K:.2f} us per step", flush=True)
ops.set_option(ops.OPT_STEP_PROFILE, 1)
for _ in range(20):
  step()
ops.step_profile(reset=True)
for _ in range(300):
  step()
p = ops.step_profile()
print(f"profiled ({p[3]} steps, synchronised): prepare {p[0]:.2f} us, potential {p[1]:.2f} us, "
      f"update {p[2]:.2f} us", flush=True)
