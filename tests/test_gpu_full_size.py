"""The hot path at BASELINE.json's full sizes, checked through properties that
do not need the oracle to run the whole problem:

* C2 (4096 chains x 1024 features, batch 1024, 10^6 observations): the oracle
  on a handful of chains (chains are independent, so any subset of rows must
  equal the oracle run on those rows alone), the tensor-core potential against
  the fp32 SIMT kernels over ALL chains, chain sharding (running the two halves
  of the batch separately -- what two GPUs do -- gives the same bits), key
  evolution of all chains, determinism, noise moments over 4.2 M normals.
* C3 (256 chains x 669 706 parameters in six leaves): SGHMC step on the whole
  state, the first / a middle / the last chain bit-exact against the oracle.
"""
import numpy as np
import pytest

from oracle import cnative, prng
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu

C2 = dict(C=4096, d=1024, n=1024, N=1_000_000)
MLP_SIZES = [784 * 512, 512, 512 * 512, 512, 512 * 10, 10]      # tree_flatten order


@pytest.fixture(scope="module")
def c2(gpu):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C, d, n, N = C2["C"], C2["d"], C2["n"], C2["N"]
  X, y, _ = ops.synth_logistic_data(0, N, d)
  dkey = [DA.from_numpy(ops.prng_key(0)), DA((2,), np.uint32)]
  idx = DA((n,), np.int32)
  ops.minibatch_draw(dkey[0], dkey[1], idx, N)
  rng = np.random.default_rng(0)
  theta = (rng.standard_normal((C, d)) * 0.05).astype(np.float32)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0, x_absmax=ops.absmax(X))
  hidx = idx.numpy()
  Xb = ops.gather_rows(X, idx).numpy()
  yb = ops.gather_rows(y.reshape(N, 1), idx).numpy().reshape(n)
  return dict(ops=ops, DA=DA, X=X, y=y, idx=idx, hidx=hidx, Xb=Xb, yb=yb, theta=theta,
              spec=spec, **C2)


def _potential(c2, theta, path):
  ops, DA = c2["ops"], c2["DA"]
  C = theta.shape[0]
  U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, c2["d"]), np.float32)
  ops.glm_potential_grad(c2["spec"], DA.from_numpy(theta), c2["X"], c2["y"], c2["idx"],
                         c2["N"], U, var, g, path=path)
  return U.numpy(), var.numpy(), g.numpy()


def test_c2_minibatch_indices_full_size(c2):
  """1024 indices into 10^6 rows: the restated jax.random.randint, bit for bit."""
  from oracle import data as odata
  _, want = odata.device_draw(prng.PRNGKey(0), c2["n"], c2["N"])
  assert np.array_equal(c2["hidx"], want)


def test_c2_potential_all_chains_tc_vs_simt_and_oracle_rows(c2):
  U, var, g = _potential(c2, c2["theta"], "tc_parity")
  U0, var0, g0 = _potential(c2, c2["theta"], "simt")
  np.testing.assert_allclose(U, U0, rtol=1e-5)
  np.testing.assert_allclose(var, var0, rtol=2e-4)
  scale = np.abs(g0).max(axis=1, keepdims=True)
  assert (np.abs(g - g0) / scale).max() < 1e-5
  # the oracle on five of the 4096 chains
  rows = np.array([0, 1, 1234, 2048, 4095])
  pot = osgmc.minibatch_potential(osgmc.Logistic(c2["d"], 0),
                                  osgmc.Prior("gaussian", 0, c2["d"], 10.0))
  wU, well, wg = pot(c2["theta"][rows], (c2["Xb"], c2["yb"]), c2["N"])
  np.testing.assert_allclose(U[rows], wU, rtol=1e-5)
  np.testing.assert_allclose(U0[rows], wU, rtol=1e-5)
  wscale = np.abs(wg).max(axis=1, keepdims=True)
  assert (np.abs(g[rows] - wg) / wscale).max() < 1e-5
  assert (np.abs(g0[rows] - wg) / wscale).max() < 1e-5


def test_c2_chain_sharding_gives_the_same_bits(c2):
  """Chains [0, 2048) and [2048, 4096) evaluated on their own (the two-GPU
  layout) equal the corresponding rows of the full batch, bit for bit -- for the
  potential and for the fused update."""
  ops, DA = c2["ops"], c2["DA"]
  C, d, h = c2["C"], c2["d"], c2["C"] // 2
  full = _potential(c2, c2["theta"], "tc_parity")
  lo = _potential(c2, c2["theta"][:h], "tc_parity")
  hi = _potential(c2, c2["theta"][h:], "tc_parity")
  for f, a, b in zip(full, lo, hi):
    assert np.array_equal(f[:h].view(np.uint32), a.view(np.uint32))
    assert np.array_equal(f[h:].view(np.uint32), b.view(np.uint32))
  keys = np.stack([prng.PRNGKey(c) for c in range(C)])
  v = np.ones((C, d), np.float32)

  def upd(sl):
    t, vv, g = DA.from_numpy(c2["theta"][sl]), DA.from_numpy(v[sl]), DA.from_numpy(full[2][sl])
    k0, k1 = DA.from_numpy(keys[sl]), DA((t.shape[0], 2), np.uint32)
    ops.sgld_update(t, g, k0, k1, [d], 1e-3, 1.0, v=vv)
    return t.numpy(), vv.numpy(), k1.numpy()

  whole = upd(slice(0, C))
  for part, sl in ((upd(slice(0, h)), slice(0, h)), (upd(slice(h, C)), slice(h, C))):
    for w, p in zip(whole, part):
      assert np.array_equal(w[sl].view(np.uint32), p.view(np.uint32))


def test_c2_full_step_keys_rows_determinism(c2):
  """One pSGLD step of all 4096 chains through sgmc_glm_sgld_step: new keys of
  every chain = split(key)[0]; five chains equal the oracle's langevin_update
  (tolerance: the gradient comes from the tensor cores); a second run gives the
  same bits."""
  ops, DA = c2["ops"], c2["DA"]
  C, d, n, N = c2["C"], c2["d"], c2["n"], c2["N"]
  keys = np.stack([prng.PRNGKey(c) for c in range(C)])

  def run():
    t, v = DA.from_numpy(c2["theta"]), DA.full((C, d), 1.0)
    U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
    k0, k1 = DA.from_numpy(keys), DA((C, 2), np.uint32)
    ops.glm_sgld_step(c2["spec"], t, c2["X"], c2["y"], c2["idx"], N, U, var, g, k0, k1,
                      1e-3, 1.0, v=v, path="tc_parity")
    return t.numpy(), v.numpy(), U.numpy(), k1.numpy(), g.numpy()

  t1, v1, U1, k1, g1 = run()
  t2, v2, U2, k2, g2 = run()
  for a, b in ((t1, t2), (v1, v2), (U1, U2), (k1, k2), (g1, g2)):
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
  assert np.array_equal(k1, prng.split(keys, 2)[:, 0])
  rows = np.array([0, 7, 2047, 2048, 4095])
  pot = osgmc.minibatch_potential(osgmc.Logistic(d, 0), osgmc.Prior("gaussian", 0, d, 10.0))
  st = osgmc.LangevinState(c2["theta"][rows], keys[rows], np.ones((len(rows), d), np.float32),
                           np.zeros(len(rows), np.float32), np.ones(len(rows), np.float32))
  want = osgmc.langevin_update(st, lambda th: pot(th, (c2["Xb"], c2["yb"]), N), [d], 1e-3, 1.0)
  scale = np.abs(want.theta).max(axis=1, keepdims=True)
  # Cold start of RMSprop (v = 1 while g^2 ~ 1e5): coordinates with |g| of order one sit
  # on the steep part of g / sqrt(0.9 + 0.1 g^2), so an ABSOLUTE gradient error of 1e-5
  # of the row's max |g| (tensor-core accumulation, asserted above) moves theta' by
  # eps * 1e-5 * max|g| ~ 1e-5 -- 5e-5 of this test's small |theta| ~ 0.15.  Any fp32
  # gradient shows the same amplification against an fp64 one
  # (tools/r2_cold_start_amplification.py); it disappears once v has adapted, which is
  # what the 1 000-step trajectory test below bounds by 1e-5.
  assert (np.abs(t1[rows] - want.theta) / scale).max() < 1e-4
  np.testing.assert_allclose(U1[rows], want.potential, rtol=1e-5)
  # noise + update on the device's own gradient agree to the last bits
  # (default mode: SFU sqrt / rcp, <= 2 ulp)
  same = osgmc.langevin_update(st, lambda th: (want.potential, np.zeros((len(rows), 2), np.float32),
                                               g1[rows]), [d], 1e-3, 1.0)
  assert (np.abs(t1[rows] - same.theta) / scale).max() < 1e-6
  np.testing.assert_allclose(v1[rows], same.v, rtol=1e-6)


def test_c2_noise_full_size(c2):
  """random_tree for 4096 chains x 1024: sampled chains bit-exact against the C
  restatement of the oracle, moments of all 4.2 M normals."""
  ops = c2["ops"]
  C, d = c2["C"], c2["d"]
  keys = np.stack([prng.PRNGKey(1000 + c) for c in range(C)])
  z = ops.normal_like(c2["DA"].from_numpy(keys), [d]).numpy()
  rows = np.array([0, 1, 63, 64, 2047, 4095])
  want = cnative.normal_like(keys[rows], [d])
  assert np.array_equal(z[rows].view(np.uint32), want.view(np.uint32))
  zz = z.astype(np.float64).ravel()
  m = zz.size
  assert abs(zz.mean()) < 5 / np.sqrt(m)
  assert abs(zz.var() - 1) < 5 * np.sqrt(2 / m)
  assert abs((zz ** 3).mean()) < 5 * np.sqrt(15 / m)
  assert abs((zz ** 4).mean() - 3) < 5 * np.sqrt(96 / m)


def test_c3_sghmc_step_full_size(gpu):
  """256 chains x 669 706 parameters (MLP 784-512-512-10 pytree): begin + one
  inner SGHMC step on the whole 686 MB state; three chains against the oracle."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C, P, eps = 256, sum(MLP_SIZES), 0.01
  assert P == 669_706
  rng = np.random.default_rng(5)
  rows = np.array([0, 100, 255])
  theta_rows = (rng.standard_normal((3, P)) * 0.05).astype(np.float32)
  grad_rows = rng.standard_normal((3, P)).astype(np.float32)
  theta = np.zeros((C, P), np.float32)
  grad = np.zeros((C, P), np.float32)
  theta[rows], grad[rows] = theta_rows, grad_rows
  keys = np.stack([prng.PRNGKey(c) for c in range(C)])
  d_t, d_p, d_g = DA.from_numpy(theta), DA.from_numpy(theta), DA.from_numpy(grad)
  k = [DA.from_numpy(keys), DA((C, 2), np.uint32)]
  ops.sghmc_begin(d_t, d_p, k[0], k[1], MLP_SIZES, eps)
  ops.sghmc_step(d_t, d_p, d_g, k[1], k[0], MLP_SIZES, eps, friction=1.0, last=True)
  z = np.zeros(3, np.float32)
  want = osgmc.friction_leapfrog_integrate(
      osgmc.leapfrog_init(theta_rows, keys[rows]), [lambda th: (z, None, grad_rows)],
      MLP_SIZES, eps, 1.0)
  got_t, got_p, got_k = d_t.numpy(), d_p.numpy(), k[0].numpy()
  assert np.array_equal(got_k[rows], want.key)
  assert np.array_equal(got_t[rows].view(np.uint32), want.theta.view(np.uint32))
  assert np.array_equal(got_p[rows].view(np.uint32), want.momentum.view(np.uint32))
  # every other chain moved too (zero gradient: pure noise + friction)
  assert np.all(np.abs(got_p).max(axis=1) > 0)


@pytest.mark.parametrize("carried", [True, False])
def test_c2_trajectory_1000_steps_sampled_chains(gpu, carried):
  """north star: "parameter trajectories ... within rtol 1e-5 (fp32) after 1,000
  steps".  All 4096 chains x 1024 features (batch 1024) run 1 000 pSGLD steps on
  the benchmarked path (tensor-core parity potential, RMSprop, carried operand
  split + shadow noise when `carried`); eight sampled chains are compared with
  the oracle run on those chains alone (chains are independent), same keys, same
  minibatch stream.  N = 65 536 observations so the host copy the oracle gathers
  from stays small; C, d, n are the C2 sizes."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  from oracle import data as odata
  from oracle import scheduler as osched
  C, d, n, N, K = 4096, 1024, 1024, 65536, 1000
  X, y, _ = ops.synth_logistic_data(0, N, d)
  hX, hy = X.numpy(), y.numpy()
  rows = np.array([0, 1, 777, 2047, 2048, 3000, 4094, 4095])
  keys = np.stack([prng.PRNGKey(c) for c in range(C)])
  eps = osched.polynomial_step_size_first_last(K, 1e-2, 1e-3)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0, x_absmax=ops.absmax(X))
  theta, v = DA.zeros((C, d)), DA.full((C, d), 1.0)
  U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
  kk = [DA.from_numpy(keys), DA((C, 2), np.uint32)]
  dk = [DA.from_numpy(prng.PRNGKey(0)), DA((2,), np.uint32)]
  idx = DA((n,), np.int32)
  ws = ops.glm_workspace(C, n, d, "tc_parity")
  for k in range(K):
    ops.minibatch_draw(dk[k % 2], dk[(k + 1) % 2], idx, N)
    carry = 0 if not carried else (ops.STEP_CARRY_INIT if k == 0 else ops.STEP_CARRY)
    ops.glm_sgld_step(spec, theta, X, y, idx, N, U, var, g, kk[k % 2], kk[(k + 1) % 2],
                      float(eps[k]), 1.0, v=v, workspace=ws, path="tc_parity",
                      write_grad=False, carry=carry)
  got, gotU = theta.numpy()[rows], U.numpy()[rows]
  pot = osgmc.minibatch_potential(osgmc.Logistic(d, 0), osgmc.Prior("gaussian", 0, d, 10.0))
  st = osgmc.langevin_init(np.zeros((len(rows), d), np.float32), keys[rows], rms=True)
  dkey = prng.PRNGKey(0)
  for k in range(K):
    dkey, ix = odata.device_draw(dkey, n, N)
    Xb, yb = hX[ix], hy[ix]
    st = osgmc.langevin_update(st, lambda th: pot(th, (Xb, yb), N), [d], eps[k], 1.0)
  scale = np.abs(st.theta).max(axis=1, keepdims=True)
  err = (np.abs(got - st.theta) / scale).max()
  assert err < 1e-5, err
  np.testing.assert_allclose(gotU, st.potential, rtol=1e-5)
  assert np.array_equal(kk[K % 2].numpy()[rows], st.key)        # noise stream: bit-exact
  assert np.array_equal(dk[K % 2].numpy(), dkey)                # minibatch stream: bit-exact
