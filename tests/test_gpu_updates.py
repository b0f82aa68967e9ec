"""GPU parity of the fused integrator / preconditioner updates.

Given the same gradient the CUDA kernels must reproduce the oracle's
 arithmetic BIT FOR BIT: positions, momenta, RMSprop
state and evolved keys.  Per-chain kinetic-energy reductions are compared with
rtol 1e-5 (summation order is not specified by the reference either)."""
import numpy as np
import pytest

from oracle import prng
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu

# [2048, 10] and [6, 64, 3, 136]: float4-capable leaves whose alignment differs from chain to
# chain (P % 4 = 2 and P odd): the per-chain shifted groups of LeafTable::shifted
SIZES = [[1024], [1, 4], [8, 16, 40], [7, 2, 33], [2048, 10], [6, 64, 3, 136]]
LAYOUTS = ["original", "partitionable"]


def _bits(a):
  return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture
def exact_math(gpu):
  """IEEE-exact preconditioner arithmetic (SGMC_OPT_EXACT_UPDATE_MATH=1): the
  mode in which pSGLD is bit-identical to the oracle."""
  from jax_sgmc_b200 import ops
  ops.set_option(ops.OPT_EXACT_UPDATE_MATH, 1)
  yield
  ops.set_option(ops.OPT_EXACT_UPDATE_MATH, 0)


def _setup(C, sizes, seed=0):
  rng = np.random.default_rng(seed)
  P = sum(sizes)
  theta = rng.standard_normal((C, P)).astype(np.float32)
  grad = (rng.standard_normal((C, P)) * 3).astype(np.float32)
  keys = np.stack([prng.PRNGKey(s) for s in range(7, 7 + C)])
  return theta, grad, keys


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("sizes", SIZES)
@pytest.mark.parametrize("rms", [False, True])
def test_sgld_step_bit_exact(gpu, exact_math, sizes, rms, layout):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C = 11
  theta, grad, keys = _setup(C, sizes)
  v = np.abs(np.random.default_rng(1).standard_normal(theta.shape)).astype(np.float32) + 0.1
  eps, T = 0.0123, 1.7
  d_theta, d_grad = DA.from_numpy(theta), DA.from_numpy(grad)
  d_v = DA.from_numpy(v) if rms else None
  d_kin, d_kout = DA.from_numpy(keys), DA((C, 2), np.uint32)
  ops.sgld_update(d_theta, d_grad, d_kin, d_kout, sizes, eps, T, v=d_v,
                  alpha=0.9, lmbd=1e-5, layout=layout)
  ks = prng.split(keys, 2, layout)
  xi = osgmc.random_tree_flat(ks[:, 1], sizes, layout)
  want_theta, want_v = osgmc.sgld_apply(theta, grad, xi, eps, T,
                                        v if rms else None, 0.9, 1e-5)
  assert np.array_equal(d_kout.numpy(), ks[:, 0])
  assert np.array_equal(_bits(d_theta.numpy()), _bits(want_theta))
  if rms:
    assert np.array_equal(_bits(d_v.numpy()), _bits(want_v))


@pytest.mark.parametrize("sizes", SIZES)
def test_psgld_fast_math_within_tolerance(gpu, sizes):
  """Default mode: SFU sqrt/rcp approximations in the RMSprop arithmetic.
  v' stays within 1 ulp, theta' within rtol 2e-6 of the oracle; keys exact."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  assert ops._lib.load().sgmc_get_option(ops.OPT_EXACT_UPDATE_MATH) == 0
  C = 11
  theta, grad, keys = _setup(C, sizes)
  v = np.abs(np.random.default_rng(1).standard_normal(theta.shape)).astype(np.float32) + 0.1
  d_theta, d_v = DA.from_numpy(theta), DA.from_numpy(v)
  d_kout = DA((C, 2), np.uint32)
  ops.sgld_update(d_theta, DA.from_numpy(grad), DA.from_numpy(keys), d_kout,
                  sizes, 0.0123, 1.7, v=d_v)
  ks = prng.split(keys, 2)
  xi = osgmc.random_tree_flat(ks[:, 1], sizes)
  want_theta, want_v = osgmc.sgld_apply(theta, grad, xi, 0.0123, 1.7, v)
  assert np.array_equal(d_kout.numpy(), ks[:, 0])
  np.testing.assert_allclose(d_v.numpy(), want_v, rtol=2.5e-7)
  delta = np.abs(want_theta - theta).max()
  assert np.abs(d_theta.numpy() - want_theta).max() <= 2e-6 * delta + 1e-7


def test_sgld_per_chain_temperature(gpu):
  """reSGLD ladders: temperature per chain (label swap instead of state swap)."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C, sizes = 6, [64]
  theta, grad, keys = _setup(C, sizes, 3)
  temps = np.array([1, 3, 10, 30, 100, 1000], np.float32)
  d_theta = DA.from_numpy(theta)
  d_kout = DA((C, 2), np.uint32)
  ops.sgld_update(d_theta, DA.from_numpy(grad), DA.from_numpy(keys), d_kout,
                  sizes, 0.01, 1.0, temp_per_chain=DA.from_numpy(temps))
  ks = prng.split(keys, 2)
  xi = osgmc.random_tree_flat(ks[:, 1], sizes)
  got = d_theta.numpy()
  for c in range(C):
    want, _ = osgmc.sgld_apply(theta[c:c + 1], grad[c:c + 1], xi[c:c + 1], 0.01,
                               temps[c])
    assert np.array_equal(_bits(got[c:c + 1]), _bits(want))


@pytest.mark.parametrize("rms", [False, True])
def test_sgld_trajectory_bit_exact(gpu, exact_math, rms):
  """200 chained steps with a decaying step size and a gradient that depends on
  the current position (computed on the host from the device state, so both
  sides see identical gradients): the whole trajectory must stay bit-exact."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  from oracle import scheduler as osched
  C, sizes = 5, [1, 4, 19]
  theta, _, keys = _setup(C, sizes, 5)
  a = np.linspace(0.5, 3.0, theta.shape[1]).astype(np.float32)
  st = osgmc.langevin_init(theta, keys, rms=rms)
  d_theta = DA.from_numpy(theta)
  d_v = DA.from_numpy(np.ones_like(theta)) if rms else None
  d_k = [DA.from_numpy(keys), DA((C, 2), np.uint32)]
  d_g = DA(theta.shape, np.float32)
  K = 200
  eps = osched.polynomial_step_size_first_last(K, 0.05, 0.001)
  for k in range(K):
    g = (a * d_theta.numpy()).astype(np.float32)
    d_g.copy_from_host(g)
    ops.sgld_update(d_theta, d_g, d_k[k % 2], d_k[(k + 1) % 2], sizes, eps[k],
                    1.0, v=d_v)
    st = osgmc.langevin_update(
        st, lambda th: (np.zeros(C, np.float32), np.zeros((C, 2), np.float32),
                        (a * th).astype(np.float32)), sizes, eps[k], 1.0)
  assert np.array_equal(d_k[K % 2].numpy(), st.key)
  assert np.array_equal(_bits(d_theta.numpy()), _bits(st.theta))
  if rms:
    assert np.array_equal(_bits(d_v.numpy()), _bits(st.v))


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("sizes", SIZES)
@pytest.mark.parametrize("general", [False, True])
def test_sghmc_integrate_bit_exact(gpu, sizes, general, layout):
  """friction_leapfrog.integrate: resample + 4 inner steps (A.3)."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C, steps, eps = 4, 4, 0.02
  theta, _, keys = _setup(C, sizes, 11)
  P = theta.shape[1]
  rng = np.random.default_rng(2)
  grads = [(rng.standard_normal((C, P)) * 2).astype(np.float32) for _ in range(steps)]
  mass = (rng.random(P) + 0.5).astype(np.float32) if general else None
  fric = (rng.random(P) + 0.1).astype(np.float32) if general else np.float32(0.9)
  st = osgmc.leapfrog_init(theta, keys)
  fns = [(lambda th, g=g: (np.zeros(C, np.float32), None, g)) for g in grads]
  want = osgmc.friction_leapfrog_integrate(st, fns, sizes, eps, fric, mass, layout)

  d_theta, d_p = DA.from_numpy(theta), DA.from_numpy(theta)
  d_k = [DA.from_numpy(keys), DA((C, 2), np.uint32)]
  d_mass = DA.from_numpy(mass) if general else None
  d_fric = DA.from_numpy(fric) if general else None
  ops.sghmc_begin(d_theta, d_p, d_k[0], d_k[1], sizes, eps, d_mass, layout)
  for s in range(steps):
    ops.sghmc_step(d_theta, d_p, DA.from_numpy(grads[s]), d_k[(s + 1) % 2],
                   d_k[s % 2], sizes, eps, friction=0.9, friction_vec=d_fric,
                   mass=d_mass, last=(s == steps - 1), layout=layout)
  assert np.array_equal(d_k[(steps + 1) % 2].numpy(), want.key)
  assert np.array_equal(_bits(d_theta.numpy()), _bits(want.theta))
  assert np.array_equal(_bits(d_p.numpy()), _bits(want.momentum))


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("sizes", SIZES)
@pytest.mark.parametrize("general", [False, True])
def test_obabo_integrate_bit_exact(gpu, sizes, general, layout):
  """obabo.integrate: 3 steps, two gradient evaluations each (A.4)."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C, steps, eps, T, gamma = 4, 3, 0.03, 1.3, 0.8
  theta, _, keys = _setup(C, sizes, 13)
  P = theta.shape[1]
  rng = np.random.default_rng(4)
  g1 = [(rng.standard_normal((C, P)) * 2).astype(np.float32) for _ in range(steps)]
  g2 = [(rng.standard_normal((C, P)) * 2).astype(np.float32) for _ in range(steps)]
  mass = (rng.random(P) + 0.5).astype(np.float32) if general else None
  st = osgmc.obabo_init(theta, keys)
  z = np.zeros(C, np.float32)
  pairs = [((lambda th, g=a: (z, None, g)), (lambda th, g=b: (z, None, g)))
           for a, b in zip(g1, g2)]
  want = osgmc.obabo_integrate(st, pairs, sizes, eps, T, gamma, mass, layout)

  d_theta, d_p = DA.from_numpy(theta), DA.zeros(theta.shape)
  d_ks, d_ke = DA.zeros((C,)), DA.zeros((C,))
  d_k = [DA.from_numpy(keys), DA((C, 2), np.uint32)]
  d_mass = DA.from_numpy(mass) if general else None
  for s in range(steps):
    kin, kout = d_k[s % 2], d_k[(s + 1) % 2]
    ops.obabo_pass_a(d_theta, d_p, DA.from_numpy(g1[s]), d_ks, kin, kout, sizes,
                     eps, T, gamma, d_mass, layout)
    ops.obabo_pass_b(d_p, DA.from_numpy(g2[s]), d_ke, kin, sizes, eps, T, gamma,
                     d_mass, layout)
  assert np.array_equal(d_k[steps % 2].numpy(), want.key)
  assert np.array_equal(_bits(d_theta.numpy()), _bits(want.theta))
  assert np.array_equal(_bits(d_p.numpy()), _bits(want.momentum))
  np.testing.assert_allclose(d_ks.numpy(), want.kinetic_energy_start, rtol=1e-5)
  np.testing.assert_allclose(d_ke.numpy(), want.kinetic_energy_end, rtol=1e-5)


def test_empty_and_tiny_inputs(gpu):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  # zero chains: a no-op, not an error
  ops.sgld_update(DA((0, 8), np.float32), DA((0, 8), np.float32),
                  DA((0, 2), np.uint32), DA((1, 2), np.uint32), [8], 0.1)
  # single scalar parameter, single chain
  theta, grad, keys = _setup(1, [1], 9)
  d_theta, d_kout = DA.from_numpy(theta), DA((1, 2), np.uint32)
  ops.sgld_update(d_theta, DA.from_numpy(grad), DA.from_numpy(keys), d_kout, [1],
                  0.1)
  ks = prng.split(keys, 2)
  want, _ = osgmc.sgld_apply(theta, grad, osgmc.random_tree_flat(ks[:, 1], [1]),
                             0.1, 1.0)
  assert np.array_equal(_bits(d_theta.numpy()), _bits(want))


def test_resgld_decision_and_swap(gpu):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  S = 257
  rng = np.random.default_rng(0)
  U_n = (rng.standard_normal(S) * 50 + 800).astype(np.float32)
  U_h = (rng.standard_normal(S) * 50 + 800).astype(np.float32)
  var = (rng.random(S) * 5).astype(np.float32)
  ssq = (rng.random(S)).astype(np.float32)
  F = np.full(S, 1.0, np.float32)
  keys = np.stack([prng.PRNGKey(s) for s in range(S)])
  for step in (1, 2, 17):
    want_x, want_ssq, want_key, _, _ = osgmc.resgld_swap_decision(
        U_n, U_h, var, ssq, F, step, 1.0, 1000.0, keys)
    d_ssq, d_kout = DA.from_numpy(ssq), DA((S, 2), np.uint32)
    d_x = DA((S,), np.int32)
    ops.resgld_decide(DA.from_numpy(U_n), DA.from_numpy(U_h), DA.from_numpy(var),
                      d_ssq, DA.from_numpy(F), step, 1.0, 1000.0,
                      DA.from_numpy(keys), d_kout, d_x)
    assert np.array_equal(d_x.numpy().astype(bool), want_x)
    assert np.array_equal(_bits(d_ssq.numpy()), _bits(want_ssq))
    assert np.array_equal(d_kout.numpy(), want_key)
    assert 0 < want_x.sum() < S            # both outcomes exercised
  a = rng.standard_normal((S, 33)).astype(np.float32)
  b = rng.standard_normal((S, 33)).astype(np.float32)
  d_a, d_b = DA.from_numpy(a), DA.from_numpy(b)
  ops.swap_rows(d_a, d_b, d_x)
  m = want_x[:, None]
  assert np.array_equal(d_a.numpy(), np.where(m, b, a))
  assert np.array_equal(d_b.numpy(), np.where(m, a, b))
