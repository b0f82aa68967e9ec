"""sgmc_glm_sgld_step: the whole langevin_diffusion step (integrator.py:860-922)
with the SGLD / pSGLD update inside the gradient GEMM's epilogue must give the
same bits as the two-call sequence (potential+gradient, then the stand-alone
fused update) -- same gradient arithmetic, same noise (jax.random threefry /
erf_inv, bit-exact), same update arithmetic -- and must fall back to exactly
that sequence when the shapes do not allow the fusion."""
import numpy as np
import pytest

from oracle import prng
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fused_epilogue(gpu):
  """Turn the in-epilogue update on for these tests (off by default)."""
  from jax_sgmc_b200 import ops
  ops.set_option(ops.OPT_FUSED_STEP_EPILOGUE, 1)
  yield
  ops.set_option(ops.OPT_FUSED_STEP_EPILOGUE, 0)


def _setup(C, n, d, seed=0):
  rng = np.random.default_rng(seed)
  N = 4000
  X = (rng.standard_normal((N, d)) / np.sqrt(d)).astype(np.float32)
  w = rng.standard_normal(d).astype(np.float32)
  y = (rng.random(N) < 1 / (1 + np.exp(-(X @ w)))).astype(np.float32)
  theta = (rng.standard_normal((C, d)) * 0.3).astype(np.float32)
  v = (np.abs(rng.standard_normal((C, d))) * 50 + 1).astype(np.float32)
  idx = rng.integers(0, N, n).astype(np.int32)
  keys = np.stack([prng.PRNGKey(100 + c) for c in range(C)])
  return N, X, y, theta, v, idx, keys


def _two_calls(ops, DA, spec, N, X, y, theta, v, idx, keys, path, eps, T, rms):
  C, d = theta.shape
  dt, dv = DA.from_numpy(theta), DA.from_numpy(v)
  U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
  k0, k1 = DA.from_numpy(keys), DA((C, 2), np.uint32)
  ops.glm_potential_grad(spec, dt, DA.from_numpy(X), DA.from_numpy(y), DA.from_numpy(idx),
                         N, U, var, g, path=path)
  ops.sgld_update(dt, g, k0, k1, [d], eps, T, v=dv if rms else None, alpha=0.9, lmbd=1e-5)
  return dt.numpy(), dv.numpy(), U.numpy(), var.numpy(), g.numpy(), k1.numpy()


def _one_call(ops, DA, spec, N, X, y, theta, v, idx, keys, path, eps, T, rms, write_grad=True):
  C, d = theta.shape
  dt, dv = DA.from_numpy(theta), DA.from_numpy(v)
  U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA.zeros((C, d))
  k0, k1 = DA.from_numpy(keys), DA((C, 2), np.uint32)
  ops.glm_sgld_step(spec, dt, DA.from_numpy(X), DA.from_numpy(y), DA.from_numpy(idx), N, U,
                    var, g, k0, k1, eps, T, v=dv if rms else None, alpha=0.9, lmbd=1e-5,
                    path=path, write_grad=write_grad)
  return dt.numpy(), dv.numpy(), U.numpy(), var.numpy(), g.numpy(), k1.numpy()


def _bits(a):
  return a.view(np.uint32)


@pytest.mark.parametrize("C,n,d", [(128, 64, 256), (256, 200, 512), (384, 1024, 1024)])
@pytest.mark.parametrize("rms", [True, False])
@pytest.mark.parametrize("path", ["tc_parity", "tc_throughput"])
def test_fused_step_equals_two_calls_bitwise(gpu, C, n, d, rms, path):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  N, X, y, theta, v, idx, keys = _setup(C, n, d, seed=C + n)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0)
  a = _two_calls(ops, DA, spec, N, X, y, theta, v, idx, keys, path, 2e-3, 1.3, rms)
  b = _one_call(ops, DA, spec, N, X, y, theta, v, idx, keys, path, 2e-3, 1.3, rms)
  names = ["theta", "v", "U", "var", "grad", "keys"]
  for name, x, z in zip(names, a, b):
    if name == "v" and not rms:
      continue
    assert np.array_equal(_bits(x), _bits(z)), f"{name}: fused step differs from two calls"
  # theta actually moved and the keys advanced as split(key)[0]
  assert not np.array_equal(a[0], theta)
  want_keys = np.stack([prng.split(k, 2)[0] for k in keys])
  assert np.array_equal(b[5], want_keys)
  # without the gradient output the state is the same
  c = _one_call(ops, DA, spec, N, X, y, theta, v, idx, keys, path, 2e-3, 1.3, rms,
                write_grad=False)
  assert np.array_equal(_bits(c[0]), _bits(b[0]))
  if rms:
    assert np.array_equal(_bits(c[1]), _bits(b[1]))


def test_fused_step_matches_oracle_trajectory(gpu):
  """Three pSGLD steps against the oracle's langevin_update (fp tolerance 1e-5 of
  the row scale: the gradient comes from the split-fp16 tensor-core GEMM)."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C, n, d = 128, 96, 256
  N, X, y, theta, v, idx, keys = _setup(C, n, d, seed=5)
  v[:] = 1.0
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0)
  dt, dv = DA.from_numpy(theta), DA.from_numpy(v)
  U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
  kk = [DA.from_numpy(keys), DA((C, 2), np.uint32)]
  dX, dy, di = DA.from_numpy(X), DA.from_numpy(y), DA.from_numpy(idx)
  ws = ops.glm_workspace(C, n, d, "tc_parity")
  pot = osgmc.minibatch_potential(osgmc.Logistic(d, 0), osgmc.Prior("gaussian", 0, d, 10.0))
  st = osgmc.LangevinState(theta.copy(), keys.copy(), v.copy(), np.zeros(C, np.float32),
                           np.ones(C, np.float32))
  for k in range(3):
    ops.glm_sgld_step(spec, dt, dX, dy, di, N, U, var, g, kk[k % 2], kk[(k + 1) % 2], 1e-3,
                      1.0, v=dv, workspace=ws, path="tc_parity")
    st = osgmc.langevin_update(st, lambda th: pot(th, (X[idx], y[idx]), N), [d], 1e-3, 1.0)
  got = dt.numpy()
  scale = np.abs(st.theta).max(axis=1, keepdims=True)
  assert (np.abs(got - st.theta) / scale).max() < 1e-5
  assert np.array_equal(kk[1].numpy(), st.key)
  np.testing.assert_allclose(dv.numpy(), st.v, rtol=2e-5)


@pytest.mark.parametrize("C,n,d,path", [(100, 64, 256, "tc_parity"),     # ragged chain count
                                        (128, 64, 192, "tc_parity"),     # d % 256 != 0
                                        (7, 33, 5, "simt")])             # SIMT path
def test_fused_step_fallback_shapes(gpu, C, n, d, path):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  N, X, y, theta, v, idx, keys = _setup(C, n, d, seed=9)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0)
  a = _two_calls(ops, DA, spec, N, X, y, theta, v, idx, keys, path, 1e-3, 1.0, True)
  b = _one_call(ops, DA, spec, N, X, y, theta, v, idx, keys, path, 1e-3, 1.0, True)
  for x, z in zip(a, b):
    assert np.array_equal(_bits(x), _bits(z))


def test_step_call_default_is_two_kernels_same_bits(gpu):
  """Default options: sgmc_glm_sgld_step == potential/gradient + update kernels."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  ops.set_option(ops.OPT_FUSED_STEP_EPILOGUE, 0)
  C, n, d = 256, 128, 256
  N, X, y, theta, v, idx, keys = _setup(C, n, d, seed=4)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0)
  l0 = ops.launch_count()
  b = _one_call(ops, DA, spec, N, X, y, theta, v, idx, keys, "tc_parity", 1e-3, 1.0, True)
  # minibatch absmax (no data-set bound in this spec), prepare, GEMM1, GEMM2, update
  assert ops.launch_count() - l0 == 5
  a = _two_calls(ops, DA, spec, N, X, y, theta, v, idx, keys, "tc_parity", 1e-3, 1.0, True)
  for x, z in zip(a, b):
    assert np.array_equal(_bits(x), _bits(z))


def test_fused_step_respects_exact_math_option(gpu):
  """With SGMC_OPT_EXACT_UPDATE_MATH the step must use the IEEE update (two
  kernels) and still equal the two-call sequence bit for bit."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C, n, d = 128, 64, 256
  N, X, y, theta, v, idx, keys = _setup(C, n, d, seed=2)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0)
  ops.set_option(ops.OPT_EXACT_UPDATE_MATH, 1)
  try:
    a = _two_calls(ops, DA, spec, N, X, y, theta, v, idx, keys, "tc_parity", 1e-3, 1.0, True)
    b = _one_call(ops, DA, spec, N, X, y, theta, v, idx, keys, "tc_parity", 1e-3, 1.0, True)
  finally:
    ops.set_option(ops.OPT_EXACT_UPDATE_MATH, 0)
  for x, z in zip(a, b):
    assert np.array_equal(_bits(x), _bits(z))
