"""BASELINE.json configs[4]: Bayesian CNN classifier potential + gradient on the GPU
(csrc/mlp.cu: im2col + chain-batched GEMMs + col2im) against the oracle's restatement
(oracle/sgmc.py::CNNClassifier, itself pinned by finite differences and a direct convolution
in tests/test_oracle_cnn.py), then the MH-corrected samplers of that config (SGGMC, AMAGOLD)
driven through the operator API on it, compared per iteration with the oracle's solvers.
Tolerance: rtol 1e-5 on potentials, 1e-5 of the gradient's scale per chain."""
import numpy as np
import pytest

from oracle import data as odata
from oracle import prng
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu


def _problem(image, channels, strides, classes, C, N, seed):
  rng = np.random.default_rng(seed)
  w_off, b_off, P = osgmc.cnn_layout(image, channels, strides, classes)
  model = osgmc.CNNClassifier(image, channels, strides, classes, w_off, b_off)
  geo, F = model.geometry()
  theta = np.zeros((C, P), np.float32)
  for l, (H, W, ci, Ho, Wo, co, st) in enumerate(geo):
    theta[:, w_off[l]:w_off[l] + 9 * ci * co] = rng.standard_normal((C, 9 * ci * co)) * np.sqrt(2.0 / (9 * ci))
    theta[:, b_off[l]:b_off[l] + co] = rng.standard_normal((C, co)) * 0.1
  L = len(geo)
  theta[:, w_off[L]:w_off[L] + F * classes] = rng.standard_normal((C, F * classes)) * np.sqrt(2.0 / F)
  theta[:, b_off[L]:b_off[L] + classes] = rng.standard_normal((C, classes)) * 0.1
  X = rng.random((N, int(np.prod(image)))).astype(np.float32)
  y = rng.integers(0, classes, N).astype(np.float32)
  return model, theta, X, y, P


def _trees(theta, model):
  """Host pytrees {"conv_l": {"b", "w"}, "head": {...}} whose ravel is theta's rows."""
  geo, F = model.geometry()
  out = []
  for row in theta:
    t = {}
    for l, (H, W, ci, Ho, Wo, co, st) in enumerate(geo):
      t[f"conv_{l}"] = {"w": row[model.w_off[l]:model.w_off[l] + 9 * ci * co].reshape(3, 3, ci, co),
                        "b": row[model.b_off[l]:model.b_off[l] + co]}
    L = len(geo)
    t["head"] = {"w": row[model.w_off[L]:model.w_off[L] + F * model.n_classes].reshape(F, -1),
                 "b": row[model.b_off[L]:model.b_off[L] + model.n_classes]}
    out.append(t)
  return out


def _close_grad(got, want, tol=1e-5):
  scale = np.abs(want).max(axis=1, keepdims=True)
  assert (np.abs(got - want) / scale).max() < tol, (np.abs(got - want) / scale).max()


@pytest.mark.parametrize("image,channels,strides,classes,C,n,masked", [
    ((6, 5, 2), (3, 4), (2, 1), 3, 3, 7, False),
    ((9, 8, 3), (5, 6), (2, 2), 10, 4, 19, True),
    ((8, 8, 1), (4,), (1,), 2, 2, 33, False),          # one convolution, stride 1
    ((16, 16, 3), (8, 16), (2, 2), 10, 2, 40, True),   # rows beyond one GEMM tile
    ((32, 32, 3), (4, 16), (2, 2), 10, 2, 9, True),    # 1024 head features: split-K head
    ((8, 8, 2), (72,), (1,), 3, 2, 5, False),          # > 64 channels: 128-wide tiles, 9 head splits
    ((7, 6, 4), (8, 4), (1, 2), 3, 2, 6, True),        # channels % 4 == 0 from the input on: vector im2col / col2im
])
def test_cnn_potential_and_gradient_match_oracle(gpu, image, channels, strides, classes, C, n, masked):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  N = 300
  model, theta, X, y, P = _problem(image, channels, strides, classes, C, N, seed=sum(image) + n)
  rng = np.random.default_rng(1)
  idx = rng.integers(0, N, n).astype(np.int32)
  mask = (rng.random(n) < 0.7).astype(np.float32) if masked else None
  for prior, T in ((("gaussian", 0, P, 3.0), 1.0), (("flat", 0, 0, 1.0), 2.5)):
    spec = ops.cnn_spec(image[0], image[1], (image[2],) + tuple(channels), strides, classes,
                        model.w_off, model.b_off, prior[0], prior[1], prior[2], prior[3], T)
    U, var = DA((C,), np.float32), DA((C,), np.float32)
    g, ell = DA((C, P), np.float32), DA((C, n), np.float32)
    ops.cnn_potential_grad(spec, DA.from_numpy(theta), DA.from_numpy(X), DA.from_numpy(y),
                           DA.from_numpy(idx), N, U, var, g, ell,
                           mask=None if mask is None else DA.from_numpy(mask))
    pot = osgmc.minibatch_potential(model, osgmc.Prior(*prior), T)
    wU, well, wg = pot(theta, (X[idx], y[idx]), N, mask=mask)
    np.testing.assert_allclose(ell.numpy(), well, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(U.numpy(), wU, rtol=1e-5)
    np.testing.assert_allclose(var.numpy(), well.astype(np.float64).var(axis=1), rtol=1e-4)
    _close_grad(g.numpy(), wg)
    U2 = DA((C,), np.float32)
    ops.cnn_potential_grad(spec, DA.from_numpy(theta), DA.from_numpy(X), DA.from_numpy(y),
                           DA.from_numpy(idx), N, U2,
                           mask=None if mask is None else DA.from_numpy(mask))   # potential only
    assert np.array_equal(U2.numpy(), U.numpy())


@pytest.mark.parametrize("kind", ["sggmc", "amagold"])
def test_mh_samplers_on_the_cnn_match_oracle(gpu, kind):
  """configs[4] in miniature: solver.sggmc / solver.amagold on the CNN potential through the
  operator API (nn.CNNClassifier behind minibatch_potential / full_potential) against the
  oracle: same accept / reject pattern, samples within 1e-5."""
  from jax_sgmc_b200 import data, glm, integrator, nn, ops, potential, scheduler, solver
  ops.set_option(ops.OPT_EXACT_UPDATE_MATH, 1)
  try:
    image, channels, strides, classes = (6, 5, 2), (3, 4), (2, 1), 3
    C, N, n, steps, iters = 4, 48, 8, 2, 4
    model, theta, X, y, P = _problem(image, channels, strides, classes, C, N, seed=5)
    theta *= 0.5
    loader = data.DeviceNumpyDataLoader(x=X.reshape((N,) + image), y=y)
    prior, lik = glm.GaussianPrior(2.0), nn.CNNClassifier(strides=strides)
    pot = potential.minibatch_potential(prior, lik, strategy="vmap")
    full = potential.full_potential(prior, lik, strategy="vmap")
    random_data = data.random_reference_data(loader, 1, n)
    full_map = data.full_reference_data(loader, 16, 16)
    eps, fr = 2e-3, (1.0 if kind == "sggmc" else 0.25)
    if kind == "sggmc":
      init, update, get = solver.sggmc(integrator.obabo(pot, random_data, steps, fr), full, full_map)
    else:
      init, update, get = solver.amagold(
          integrator.reversible_leapfrog(pot, random_data, steps, fr), full, full_map)
    keys = np.stack([prng.PRNGKey(30 + c) for c in range(C)])
    state = init(_trees(theta, model), key=keys)
    sched = scheduler.schedule(step_size=np.float32(eps), temperature=np.float32(1.0),
                               burn_in=np.float32(1.0), accept=True)
    o_prior = osgmc.Prior("gaussian", 0, P, 2.0)
    o_pot = osgmc.minibatch_potential(model, o_prior)
    o_full_fn = osgmc.full_potential(model, o_prior)
    ids = np.arange(int(np.ceil(N / 16)) * 16).reshape(-1, 16)
    batches = [(X[i % N], y[i % N], (i < N).astype(np.float32)) for i in ids]
    o_full = lambda th: o_full_fn(th, batches, N)
    dkey = prng.PRNGKey(0)

    def next_grad_fn():
      nonlocal dkey
      dkey, idx = odata.device_draw(dkey, n, N)
      return lambda th, idx=idx: o_pot(th, (X[idx], y[idx]), N)

    # the noise follows the pytree's leaves (integrator.random_tree): b, w per layer dict
    geo, F = model.geometry()
    sizes = []
    for (H, W, ci, Ho, Wo, co, st) in geo:
      sizes += [co, 9 * ci * co]
    sizes += [classes, F * classes]
    assert sum(sizes) == P
    o_state = osgmc.sggmc_init(theta, o_full, keys) if kind == "sggmc" else \
        osgmc.amagold_init(theta, o_full, keys, sizes=sizes)
    np.testing.assert_allclose(state.potential.numpy(), o_state.potential, rtol=1e-5)
    for it in range(iters):
      state, _ = update(state, sched)
      if kind == "sggmc":
        pairs = [(next_grad_fn(), next_grad_fn()) for _ in range(steps)]
        o_state, acc = osgmc.sggmc_update(o_state, pairs, o_full, sizes, eps, 1.0, fr)
      else:
        fns = [next_grad_fn() for _ in range(steps)]
        o_state, acc = osgmc.amagold_update(o_state, fns, o_full, sizes, eps, fr)
      assert np.array_equal(state.reject.numpy() == 0, acc), f"iteration {it}"
      got = get(state)["variables"].flat.numpy()
      scale = np.abs(o_state.integrator_state.theta).max(axis=1, keepdims=True)
      assert (np.abs(got - o_state.integrator_state.theta) / scale).max() < 1e-5
      np.testing.assert_allclose(state.potential.numpy(), o_state.potential, rtol=1e-5)
  finally:
    ops.set_option(ops.OPT_EXACT_UPDATE_MATH, 0)


def test_alias_sggmc_runs_the_cifar_shaped_cnn(gpu):
  """alias.sggmc / alias.amagold end to end on 32x32x3 images with the configs[4] network
  (conv3x3-32 stride 2, conv3x3-64 stride 2, dense-10; 60 362 parameters per chain)."""
  from jax_sgmc_b200 import alias, data, glm, nn, ops, potential
  rng = np.random.default_rng(0)
  N = 256
  X = rng.random((N, 32, 32, 3)).astype(np.float32)
  y = rng.integers(0, 10, N).astype(np.float32)
  lik = nn.CNNClassifier(strides=(2, 2))
  prior = glm.GaussianPrior(10.0)
  pot = potential.minibatch_potential(prior, lik, strategy="vmap")
  full = potential.full_potential(prior, lik, strategy="vmap")
  init = [nn.init_cnn_params(ops.prng_key(c), (32, 32, 3), (32, 64), (2, 2), 10) for c in range(2)]
  for make in (alias.sggmc, alias.amagold):
    run = make(pot, full, data.DeviceNumpyDataLoader(x=X, y=y), cache_size=1, batch_size=64,
               first_step_size=1e-4, last_step_size=5e-5, burn_in=2, progress_bar=False)
    res = run(*init, iterations=6)
    assert len(res) == 2
    for r in res:
      assert r["sample_count"] == 4
      assert r["samples"]["variables"]["head"]["w"].shape == (4, 4096, 10)
      assert r["samples"]["variables"]["conv_1"]["w"].shape == (4, 3, 3, 32, 64)
      assert np.all(np.isfinite(r["samples"]["variables"]["conv_0"]["w"]))
