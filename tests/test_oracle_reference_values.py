"""Pin the oracle on the numeric values the reference's own docs/tests print
(tests/golden/reference_values.json, each with its reference file:line)."""
import json
import os

import numpy as np

from oracle import data, prng, scheduler, sgmc

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden",
                                   "reference_values.json")))


class _SumNormal:
  """Likelihood of docs/usage/potential.rst:77-82: sum of norm.logpdf over the
  5 features of one observation, loc = sample mean, scale = sample std."""

  def loglik(self, theta, X, y):
    mean, std = theta[:, :5], theta[:, 5:6]
    z = (X[None, :, :] - mean[:, None, :]) / std[:, None, :]
    lp = -0.5 * np.log(2 * np.pi) - np.log(std[:, None, :]) - 0.5 * z * z
    return lp.sum(-1).astype(np.float32), None

  def vjp(self, theta, X, y, aux, cot):
    return np.zeros_like(theta)


def test_potential_doctest_values():
  g = GOLD["potential_doctest"]
  x = prng.normal(prng.PRNGKey(0), (100, 5))
  theta = np.concatenate([np.zeros(5), np.ones(1)]).astype(np.float32)[None]
  pot = sgmc.minibatch_potential(_SumNormal(), sgmc.Prior("flat"))
  idx = data.HostDraws(100, 5, seed=0).draw()
  U, ell, _ = pot(theta, (x[idx], None), 100)
  assert round(float(U[0])) == g["stochastic_potential_round"]
  assert round(float(np.var(ell[0]))) == g["likelihood_variance_round"]
  # full potential over batches of 3 with wrap-around masking
  full = sgmc.full_potential(_SumNormal(), sgmc.Prior("flat"))
  batches = []
  for s in range(0, 100, 3):
    ids = np.arange(s, s + 3)
    batches.append((x[ids % 100], None, (ids < 100).astype(np.float32)))
  Uf = full(theta, batches, 100)
  assert round(float(Uf[0])) == g["full_potential_round"]


def test_host_loader_draws():
  g = GOLD["host_loader_draws"]
  assert data.HostDraws(g["N"], 2, seed=0).draw().tolist() == g["seed0_mb2_first"]
  h = data.HostDraws(g["N"], 3, seed=0)
  draws = [h.draw().tolist() for _ in range(4)]
  assert draws[3] == g["seed0_mb3_fourth_random"]


class _Linear:
  """tests/test_potential.py:20-34: likelihood = sum(sample * observation),
  prior = sum(sample)."""

  def loglik(self, theta, X, y):
    return (theta @ X.T).astype(np.float32), None

  def vjp(self, theta, X, y, aux, cot):
    return (cot @ X).astype(np.float32)


class _SumPrior:
  def value(self, theta):
    return theta.sum(1).astype(np.float32)

  def grad(self, theta):
    return np.ones_like(theta)


def test_linear_potential_closed_form():
  """tests/test_potential.py:66-92: sample = ones(4), every observation is
  arange(4), observation_count = obs, batch_size = dim:
  U = -sum(0..3) * obs - 4 exactly."""
  theta = np.ones((1, 4), np.float32)
  for obs in (7, 11):
    for dim in (3, 5):
      pot = sgmc.minibatch_potential(_Linear(), _SumPrior())
      X = np.tile(np.arange(4, dtype=np.float32), (dim, 1))
      U, _, g = pot(theta, (X, None), obs)
      assert float(U[0]) == -6.0 * obs - 4.0
      # gradient of the closed form: -obs * arange(4) - 1
      assert np.allclose(g[0], -obs * np.arange(4) - 1.0, rtol=1e-6)


def test_scheduler_doc_values():
  g = GOLD["scheduler"]
  s = scheduler.polynomial_step_size(5, g["a"], g["b"], 0.1)
  assert round(float(s[0]), 2) == g["gamma_0.1"]["step0"]
  assert round(float(s[1]), 2) == g["gamma_0.1"]["step1"]
  s = scheduler.polynomial_step_size(5, g["a"], g["b"], 1.0)
  assert round(float(s[0]), 2) == g["gamma_1.0"]["step0"]
  assert round(float(s[1]), 2) == g["gamma_1.0"]["step1"]
  # first/last endpoints (tests/test_scheduler.py:81-102)
  s = scheduler.polynomial_step_size_first_last(1000, 0.05, 0.001)
  assert abs(float(s[0]) - 0.05) < 1e-6 and abs(float(s[-1]) - 0.001) < 1e-6
  b = scheduler.initial_burn_in(10, 3)
  assert b.tolist() == [0, 0, 0, 1, 1, 1, 1, 1, 1, 1]


def test_quickstart_dataset_matches_notebook_posterior():
  g = GOLD["quickstart_dataset"]
  x, y, w = data.quickstart_dataset()
  w_ls = np.linalg.lstsq(x.astype(np.float64), y.astype(np.float64),
                         rcond=None)[0].ravel()
  assert np.allclose(w_ls, g["numpyro_posterior_w"], atol=0.012)
  resid = y.ravel() - x @ w_ls
  assert abs(resid.std() - g["numpyro_posterior_sigma"]) < 0.01


def test_gradients_match_finite_differences():
  rng = np.random.default_rng(0)
  X = rng.standard_normal((16, 6)).astype(np.float32)
  for model, prior, theta, y in [
      (sgmc.GaussianLinear(6, 1, 0), sgmc.Prior("inv_sigma", 0),
       np.concatenate([[0.3], rng.standard_normal(6) * 0.1])[None],
       rng.standard_normal(16)),
      (sgmc.Logistic(6, 0), sgmc.Prior("gaussian", 0, 6, 10.0),
       (rng.standard_normal(6) * 0.5)[None], (rng.random(16) < 0.5) * 1.0)]:
    theta = theta.astype(np.float32)
    y = y.astype(np.float32)
    pot = sgmc.minibatch_potential(model, prior, temperature=1.0)
    U, _, g = pot(theta, (X, y), 1000)
    for p in range(theta.shape[1]):
      h = 1e-3
      tp, tm = theta.copy(), theta.copy()
      tp[0, p] += h
      tm[0, p] -= h
      fd = (pot(tp, (X, y), 1000)[0][0].astype(np.float64)
            - pot(tm, (X, y), 1000)[0][0]) / (2 * h)
      assert abs(fd - g[0, p]) <= 2e-2 * max(1.0, abs(fd)), (p, fd, g[0, p])
