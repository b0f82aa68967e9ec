"""CPU checks of the oracle's adaption restatements (the reference has no numeric test for
them: tests/test_adaption.py only tests the decorator): adaption.mass_matrix (Welford,
adaption.py:296-369) against NumPy's batch mean / variance, and adaption.fisher_information
(adaption.py:372-457) against an fp64 evaluation of its formulas with per-observation
gradients taken by finite differences of the log-likelihood."""
import numpy as np

from oracle import sgmc as osgmc


def test_mass_matrix_is_the_running_variance_and_changes_once():
  rng = np.random.default_rng(0)
  C, P, burn_in = 3, 6, 9
  xs = rng.standard_normal((14, C, P)).astype(np.float32) * 2 + 1
  st = osgmc.mass_matrix_init(xs[0])
  assert np.all(st.m_inv == 1) and np.all(st.m_sqrt == 1)
  for it, x in enumerate(xs, start=1):
    st = osgmc.mass_matrix_update(st, x, burn_in)
    np.testing.assert_allclose(st.mean, xs[:it].astype(np.float64).mean(0), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.ssq, it * xs[:it].astype(np.float64).var(0), rtol=1e-4, atol=1e-5)
    if it < burn_in:
      assert np.all(st.m_inv == 1)
    else:      # set in iteration burn_in from the statistics up to then, never again
      var9 = xs[:burn_in].astype(np.float64).var(0)
      np.testing.assert_allclose(st.m_inv, var9, rtol=1e-4)
      np.testing.assert_allclose(st.m_sqrt, 1 / np.sqrt(var9), rtol=1e-4)
  st = osgmc.mass_matrix_init(xs[0], init_cov=np.full(P, 4.0, np.float32))
  assert np.all(st.m_inv == 4) and np.all(st.m_sqrt == 0.5)


def _loglik64(kind, theta, X, y, d, w_off, aux_off):
  th = theta.astype(np.float64)
  z = th[:, w_off:w_off + d] @ X.astype(np.float64).T
  if kind == "logistic":
    if aux_off >= 0:
      z = z + th[:, aux_off][:, None]
    return y[None] * z - np.logaddexp(0.0, z)
  s = np.exp(th[:, aux_off])[:, None]
  r = y[None] - z
  return -0.5 * (r / s) ** 2 - np.log(s) - 0.5 * np.log(2 * np.pi)


def test_fisher_information_follows_the_reference_formulas():
  rng = np.random.default_rng(1)
  C, d, n, N, eps = 4, 5, 13, 200, 0.5
  X = rng.standard_normal((n, d)).astype(np.float32)
  for kind in ("logistic", "gaussian"):
    P, w_off, aux_off = d + 1, 1, 0
    y = (rng.random(n) < 0.5).astype(np.float32) if kind == "logistic" else \
        rng.standard_normal(n).astype(np.float32)
    theta = (rng.standard_normal((C, P)) * 0.4).astype(np.float32)
    model = osgmc.Logistic(d, w_off, aux_off) if kind == "logistic" else \
        osgmc.GaussianLinear(d, w_off, aux_off)
    _, _, grad = osgmc.minibatch_potential(model, osgmc.Prior("gaussian", 0, P, 3.0))(
        theta, (X, y), N)
    friction = (rng.random(P) * 0.3 + 0.02).astype(np.float32)
    ns, sc = osgmc.fisher_information_get(model, theta, (X, y), N, grad, friction, eps)
    # fp64: per-observation likelihood gradients by central differences
    per_obs = np.zeros((C, n, P))
    for j in range(P):
      e = np.zeros(P)
      e[j] = 1e-5
      per_obs[:, :, j] = (_loglik64(kind, theta + e, X, y, d, w_off, aux_off) -
                          _loglik64(kind, theta - e, X, y, d, w_off, aux_off)) / 2e-5
    m = grad.astype(np.float64) / N                                   # adaption.py:404
    ssq = ((per_obs - m[:, None, :]) ** 2).sum(1)                    # :414-420
    b = 0.5 * eps * ssq / (n - 1)                                     # :423-424
    corr = friction[None].astype(np.float64) - b                      # :427
    smallest = np.min(np.where(corr <= 0, np.inf, corr), axis=1, keepdims=True)
    pos = np.where(corr <= 0, smallest, corr)                         # :428-429
    assert (corr <= 0).any() and (corr > 0).any()                     # both branches exercised
    np.testing.assert_allclose(ns, np.sqrt(pos), rtol=2e-4)
    ok = (friction[None] - pos) > 1e-6
    want_sc = np.sqrt(np.where(ok, friction[None] - pos, 1.0))
    np.testing.assert_allclose(sc[ok], want_sc[ok], rtol=2e-3)
