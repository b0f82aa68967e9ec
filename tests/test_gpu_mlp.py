"""BASELINE.json configs[2]: Bayesian MLP classifier potential + gradient on the GPU
(csrc/mlp.cu) against the oracle's restatement (oracle/sgmc.py::MLPClassifier), then
SGHMC / OBABO driven through the operator API on that potential, compared PER STEP with
the oracle's integrators ("for the MLP/CNN configs compare
per-step, not per-1 000").  Tolerance: rtol 1e-5 on potentials, 1e-5 of the gradient's
scale per chain (fp32 GEMMs with a different summation order than NumPy's)."""
import numpy as np
import pytest

from oracle import data as odata
from oracle import prng
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu


def _problem(sizes, C, N, seed):
  rng = np.random.default_rng(seed)
  w_off, b_off, P = osgmc.mlp_layout(sizes)
  theta = np.zeros((C, P), np.float32)
  for l in range(len(sizes) - 1):
    i, o = sizes[l], sizes[l + 1]
    theta[:, w_off[l]:w_off[l] + i * o] = rng.standard_normal((C, i * o)) * np.sqrt(2.0 / i)
    theta[:, b_off[l]:b_off[l] + o] = rng.standard_normal((C, o)) * 0.1
  X = rng.random((N, sizes[0])).astype(np.float32)
  y = rng.integers(0, sizes[-1], N).astype(np.float32)
  return theta, X, y, (w_off, b_off, P)


def _trees(theta, sizes, layout):
  """Host pytrees {"layer_l": {"b", "w"}} whose ravel is theta's rows."""
  w_off, b_off, _ = layout
  out = []
  for row in theta:
    t = {}
    for l in range(len(sizes) - 1):
      i, o = sizes[l], sizes[l + 1]
      t[f"layer_{l}"] = {"w": row[w_off[l]:w_off[l] + i * o].reshape(i, o),
                         "b": row[b_off[l]:b_off[l] + o]}
    out.append(t)
  return out


def _close_grad(got, want, tol=1e-5):
  scale = np.abs(want).max(axis=1, keepdims=True)
  assert (np.abs(got - want) / scale).max() < tol, (np.abs(got - want) / scale).max()


@pytest.mark.parametrize("sizes,C,n,masked", [
    ((7, 6, 5, 3), 3, 9, False),
    ((20, 33, 17, 10), 5, 70, True),
    ((130, 129, 10), 2, 131, False),          # tile edges on every dimension
    ((16, 4), 4, 32, True),                   # a single layer: multinomial regression
])
def test_mlp_potential_and_gradient_match_oracle(gpu, sizes, C, n, masked):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  N = 500
  theta, X, y, (w_off, b_off, P) = _problem(sizes, C, N, seed=sum(sizes))
  rng = np.random.default_rng(1)
  idx = rng.integers(0, N, n).astype(np.int32)
  mask = (rng.random(n) < 0.7).astype(np.float32) if masked else None
  for prior, T in ((("gaussian", 0, P, 3.0), 1.0), (("flat", 0, 0, 1.0), 2.5)):
    spec = ops.mlp_spec(sizes, w_off, b_off, prior[0], prior[1], prior[2], prior[3], T)
    U, var = DA((C,), np.float32), DA((C,), np.float32)
    g, ell = DA((C, P), np.float32), DA((C, n), np.float32)
    ops.mlp_potential_grad(spec, DA.from_numpy(theta), DA.from_numpy(X), DA.from_numpy(y),
                           DA.from_numpy(idx), N, U, var, g, ell,
                           mask=None if mask is None else DA.from_numpy(mask))
    pot = osgmc.minibatch_potential(osgmc.MLPClassifier(sizes, w_off, b_off),
                                    osgmc.Prior(*prior), T)
    wU, well, wg = pot(theta, (X[idx], y[idx]), N, mask=mask)
    np.testing.assert_allclose(ell.numpy(), well, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(U.numpy(), wU, rtol=1e-5)
    np.testing.assert_allclose(var.numpy(), well.astype(np.float64).var(axis=1), rtol=1e-4)
    _close_grad(g.numpy(), wg)
    # potential only (no gradient buffers): same U
    U2 = DA((C,), np.float32)
    ops.mlp_potential_grad(spec, DA.from_numpy(theta), DA.from_numpy(X), DA.from_numpy(y),
                           DA.from_numpy(idx), N, U2,
                           mask=None if mask is None else DA.from_numpy(mask))
    assert np.array_equal(U2.numpy(), U.numpy())


def test_mlp_c3_shape_sampled_chains(gpu):
  """784-512-512-10, batch 256, 32 chains (the C3 model and batch; the chain count is
  what the oracle finishes in seconds): three chains against the oracle, and the launch
  is deterministic."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  sizes, C, n, N = (784, 512, 512, 10), 32, 256, 2000
  theta, X, y, (w_off, b_off, P) = _problem(sizes, C, N, seed=3)
  assert P == 669_706
  idx = np.random.default_rng(2).integers(0, N, n).astype(np.int32)
  spec = ops.mlp_spec(sizes, w_off, b_off, "gaussian", 0, P, 10.0, 1.0)
  d_t, dX, dy, d_i = DA.from_numpy(theta), DA.from_numpy(X), DA.from_numpy(y), DA.from_numpy(idx)
  ws = ops.mlp_workspace(spec, C, n)
  outs = []
  for _ in range(2):
    U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, P), np.float32)
    ops.mlp_potential_grad(spec, d_t, dX, dy, d_i, 60000, U, var, g, workspace=ws)
    outs.append((U.numpy(), var.numpy(), g.numpy()))
  for a, b in zip(*outs):
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
  rows = np.array([0, 13, 31])
  pot = osgmc.minibatch_potential(osgmc.MLPClassifier(sizes, w_off, b_off),
                                  osgmc.Prior("gaussian", 0, P, 10.0))
  wU, well, wg = pot(theta[rows], (X[idx], y[idx]), 60000)
  U, var, g = outs[0]
  np.testing.assert_allclose(U[rows], wU, rtol=1e-5)
  np.testing.assert_allclose(var[rows], well.astype(np.float64).var(axis=1), rtol=1e-4)
  _close_grad(g[rows], wg)


@pytest.mark.parametrize("which", ["sghmc", "obabo"])
def test_mlp_sghmc_obabo_per_step_against_oracle(gpu, which):
  """integrator.friction_leapfrog / obabo on the MLP potential through the operator API,
  compared with the oracle after EVERY outer step (same keys, same minibatch stream)."""
  from jax_sgmc_b200 import data, glm, integrator, nn, potential, scheduler
  sizes, C, n, N, steps, inner = (12, 16, 8, 4), 4, 24, 300, 6, 3
  theta, X, y, layout = _problem(sizes, C, N, seed=7)
  w_off, b_off, P = layout
  loader = data.DeviceNumpyDataLoader(x=X, y=y)
  pot = potential.minibatch_potential(glm.GaussianPrior(5.0), nn.MLPClassifier(),
                                      strategy="vmap")
  batch_fn = data.random_reference_data(loader, 1, n)
  keys = np.stack([prng.PRNGKey(40 + c) for c in range(C)])
  leaf_sizes = []
  for l in range(len(sizes) - 1):
    leaf_sizes += [sizes[l + 1], sizes[l] * sizes[l + 1]]      # b before w (sorted keys)
  opot = osgmc.minibatch_potential(osgmc.MLPClassifier(sizes, w_off, b_off),
                                   osgmc.Prior("gaussian", 0, P, 5.0))
  if which == "sghmc":
    init, integrate, get = integrator.friction_leapfrog(pot, batch_fn, steps=inner, friction=1.0)
    ost = osgmc.leapfrog_init(theta.copy(), keys)
  else:
    init, integrate, get = integrator.obabo(pot, batch_fn, steps=inner, friction=1.0)
    ost = osgmc.obabo_init(theta.copy(), keys)
  state = init(_trees(theta, sizes, layout), key=keys)
  dk = prng.PRNGKey(0)
  eps = 2e-3
  for k in range(steps):
    state = integrate(state, scheduler.schedule(np.float32(eps), np.float32(1.0), 1.0, True))
    batches = []
    for _ in range(inner * (2 if which == "obabo" else 1)):
      dk, ix = odata.device_draw(dk, n, N)
      batches.append((X[ix], y[ix]))
    fns = [(lambda th, b=b: opot(th, b, N)) for b in batches]
    if which == "sghmc":
      ost = osgmc.friction_leapfrog_integrate(ost, fns, leaf_sizes, eps, 1.0)
    else:
      ost = osgmc.obabo_integrate(ost, list(zip(fns[0::2], fns[1::2])), leaf_sizes, eps, 1.0, 1.0)
    got = get(state)["variables"].flat.numpy()
    scale = np.abs(ost.theta).max(axis=1, keepdims=True)
    assert (np.abs(got - ost.theta) / scale).max() < 1e-5, (k, (np.abs(got - ost.theta) / scale).max())
    np.testing.assert_allclose(get(state)["energy"].numpy(), ost.potential, rtol=1e-5)
    assert np.array_equal(state.key.numpy(), ost.key)          # noise stream: bit-exact


def test_mlp_alias_sghmc_runs_c3_model(gpu):
  """alias.sghmc and alias.obabo run the 784-512-512-10 classifier end to end (8 chains,
  batch 256): samples come back with the pytree structure, finite, and moved."""
  from jax_sgmc_b200 import alias, data, glm, nn, potential
  sizes, C, N = (784, 512, 512, 10), 8, 4096
  rng = np.random.default_rng(0)
  X = rng.random((N, 784)).astype(np.float32)
  y = rng.integers(0, 10, N).astype(np.float32)
  loader = data.DeviceNumpyDataLoader(x=X, y=y)
  pot = potential.minibatch_potential(glm.GaussianPrior(10.0), nn.MLPClassifier())
  inits = [nn.init_params(prng.PRNGKey(c), sizes) for c in range(C)]
  for make in (alias.sghmc, alias.obabo):
    run = make(pot, loader, cache_size=1, batch_size=256, first_step_size=1e-5,
               last_step_size=1e-6, burn_in=2, accepted_samples=3, integration_steps=5,
               friction=1.0, progress_bar=False)
    res = run(*inits, iterations=6)
    assert len(res) == C and res[0]["sample_count"] == 3
    w0 = res[3]["samples"]["variables"]["layer_0"]["w"]
    assert w0.shape == (3, 784, 512) and np.all(np.isfinite(w0))
    assert np.abs(w0[-1] - inits[3]["layer_0"]["w"]).max() > 0
