"""CPU checks of the oracle's MLP classifier restatement (oracle/sgmc.py::MLPClassifier):
the reverse pass against central finite differences of an fp64 evaluation of the same
potential, the masked form, and the pytree layout helper.  (The reference has no numeric
test for a network potential -- its users rely on jax.grad -- so finite differences are
what pins the hand-derived gradient.)"""
import numpy as np

from oracle import sgmc as osgmc
from oracle import tree as otree


def _u64(theta, X, y, sizes, w_off, b_off, N, scale, mask=None, T=1.0):
  th = theta.astype(np.float64)
  C, n = th.shape[0], X.shape[0]
  h = np.broadcast_to(X.astype(np.float64)[None], (C,) + X.shape)
  L = len(sizes) - 1
  for l in range(L):
    i, o = sizes[l], sizes[l + 1]
    W = th[:, w_off[l]:w_off[l] + i * o].reshape(C, i, o)
    b = th[:, b_off[l]:b_off[l] + o]
    a = h @ W + b[:, None, :]
    h = np.tanh(a) if l < L - 1 else a
  m = h.max(2)
  lse = np.log(np.exp(h - m[..., None]).sum(2)) + m
  ell = np.take_along_axis(h, np.broadcast_to(y.astype(int)[None, :, None], (C, n, 1)), 2)[..., 0] - lse
  Lk = -N * ell.mean(1) if mask is None else -N / n * (ell @ mask.astype(np.float64))
  return (Lk + 0.5 * (th ** 2).sum(1) / scale ** 2) / T, ell


def test_mlp_oracle_gradient_matches_finite_differences():
  sizes = (9, 7, 6, 4)
  w_off, b_off, P = osgmc.mlp_layout(sizes)
  rng = np.random.default_rng(0)
  C, n, N = 3, 11, 200
  theta = (rng.standard_normal((C, P)) * 0.4).astype(np.float32)
  X = rng.random((n, sizes[0])).astype(np.float32)
  y = rng.integers(0, sizes[-1], n).astype(np.float32)
  for mask, T in ((None, 1.0), ((rng.random(n) < 0.6).astype(np.float32), 3.0)):
    pot = osgmc.minibatch_potential(osgmc.MLPClassifier(sizes, w_off, b_off),
                                    osgmc.Prior("gaussian", 0, P, 2.0), T)
    U, ell, g = pot(theta, (X, y), N, mask=mask)
    U0, ell0 = _u64(theta, X, y, sizes, w_off, b_off, N, 2.0, mask, T)
    np.testing.assert_allclose(U, U0, rtol=2e-6)
    np.testing.assert_allclose(ell, ell0, rtol=1e-5, atol=1e-6)
    for j in rng.integers(0, P, 40):
      e = np.zeros(P)
      e[j] = 1e-4
      fd = (_u64(theta + e, X, y, sizes, w_off, b_off, N, 2.0, mask, T)[0] -
            _u64(theta - e, X, y, sizes, w_off, b_off, N, 2.0, mask, T)[0]) / 2e-4
      assert np.abs(fd - g[:, j]).max() < 2e-5 * np.abs(g).max()


def test_mlp_layout_is_the_ravel_of_the_layer_dicts():
  sizes = (5, 4, 3)
  w_off, b_off, P = osgmc.mlp_layout(sizes)
  tree = {f"layer_{l}": {"w": np.full((sizes[l], sizes[l + 1]), 10 * l + 1, np.float32),
                         "b": np.full(sizes[l + 1], 10 * l + 2, np.float32)} for l in range(2)}
  flat, _ = otree.ravel_pytree(tree)
  assert flat.size == P
  for l in range(2):
    assert np.all(flat[w_off[l]:w_off[l] + sizes[l] * sizes[l + 1]] == 10 * l + 1)
    assert np.all(flat[b_off[l]:b_off[l] + sizes[l + 1]] == 10 * l + 2)
