"""CPU checks of the oracle's CNN classifier restatement (oracle/sgmc.py::CNNClassifier):
im2col against a direct evaluation of the convolution sum, col2im as the adjoint of im2col,
the reverse pass against central finite differences of an fp64 evaluation of the same
potential (the reference's users rely on jax.grad for a network likelihood, so finite
differences are what pins the hand-derived gradient), and the pytree layout helper."""
import numpy as np

from oracle import sgmc as osgmc
from oracle import tree as otree

IMAGE, CHANNELS, STRIDES, CLASSES = (6, 5, 2), (3, 4), (2, 1), 3


def _conv_direct(x, w, b, stride):
  """x [n, H, W, Cin], w [3, 3, Cin, Cout] -> [n, Ho, Wo, Cout], zero padding 1 (fp64)."""
  n, H, W, ci = x.shape
  co = w.shape[3]
  Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
  out = np.zeros((n, Ho, Wo, co))
  for ho in range(Ho):
    for wo in range(Wo):
      for kh in range(3):
        for kw in range(3):
          hi, wi = ho * stride + kh - 1, wo * stride + kw - 1
          if 0 <= hi < H and 0 <= wi < W:
            out[:, ho, wo, :] += x[:, hi, wi, :] @ w[kh, kw]
  return out + b


def _u64(theta, X, y, model, N, scale, mask=None, T=1.0):
  th = theta.astype(np.float64)
  C, n = th.shape[0], X.shape[0]
  geo, F = model.geometry()
  Us, ells = [], []
  for c in range(C):
    h = X.astype(np.float64).reshape((n,) + tuple(model.image))
    for l, (H, W, ci, Ho, Wo, co, st) in enumerate(geo):
      w = th[c, model.w_off[l]:model.w_off[l] + 9 * ci * co].reshape(3, 3, ci, co)
      b = th[c, model.b_off[l]:model.b_off[l] + co]
      h = np.tanh(_conv_direct(h, w, b, st))
    L = len(geo)
    Wh = th[c, model.w_off[L]:model.w_off[L] + F * model.n_classes].reshape(F, -1)
    logits = h.reshape(n, F) @ Wh + th[c, model.b_off[L]:model.b_off[L] + model.n_classes]
    m = logits.max(1)
    lse = np.log(np.exp(logits - m[:, None]).sum(1)) + m
    ell = logits[np.arange(n), y.astype(int)] - lse
    Lk = -N * ell.mean() if mask is None else -N / n * (ell @ mask.astype(np.float64))
    Us.append((Lk + 0.5 * (th[c] ** 2).sum() / scale ** 2) / T)
    ells.append(ell)
  return np.array(Us), np.array(ells)


def _model():
  w_off, b_off, P = osgmc.cnn_layout(IMAGE, CHANNELS, STRIDES, CLASSES)
  return osgmc.CNNClassifier(IMAGE, CHANNELS, STRIDES, CLASSES, w_off, b_off), P


def test_im2col_matches_the_direct_convolution_and_col2im_is_its_adjoint():
  rng = np.random.default_rng(0)
  for stride in (1, 2):
    x = rng.standard_normal((2, 7, 6, 3)).astype(np.float32)
    w = rng.standard_normal((3, 3, 3, 4))
    P = osgmc._im2col(x, stride)
    got = P.reshape(-1, 27).astype(np.float64) @ w.reshape(27, 4)
    want = _conv_direct(x.astype(np.float64), w, 0.0, stride)
    np.testing.assert_allclose(got.reshape(want.shape), want, rtol=1e-12, atol=1e-12)
    dP = rng.standard_normal(P.shape).astype(np.float32)
    back = osgmc._col2im(dP, 7, 6, 3, stride)
    # <im2col(x), dP> == <x, col2im(dP)>
    np.testing.assert_allclose(np.sum(P.astype(np.float64) * dP),
                               np.sum(x.astype(np.float64) * back), rtol=1e-5)


def test_cnn_oracle_gradient_matches_finite_differences():
  model, P = _model()
  rng = np.random.default_rng(1)
  C, n, N = 2, 5, 120
  theta = (rng.standard_normal((C, P)) * 0.3).astype(np.float32)
  X = rng.random((n, int(np.prod(IMAGE)))).astype(np.float32)
  y = rng.integers(0, CLASSES, n).astype(np.float32)
  for mask, T in ((None, 1.0), ((rng.random(n) < 0.6).astype(np.float32), 2.5)):
    pot = osgmc.minibatch_potential(model, osgmc.Prior("gaussian", 0, P, 2.0), T)
    U, ell, g = pot(theta, (X, y), N, mask=mask)
    U0, ell0 = _u64(theta, X, y, model, N, 2.0, mask, T)
    np.testing.assert_allclose(U, U0, rtol=3e-6)
    np.testing.assert_allclose(ell, ell0, rtol=2e-5, atol=2e-6)
    for j in rng.integers(0, P, 40):
      e = np.zeros(P)
      e[j] = 1e-4
      fd = (_u64(theta + e, X, y, model, N, 2.0, mask, T)[0] -
            _u64(theta - e, X, y, model, N, 2.0, mask, T)[0]) / 2e-4
      assert np.abs(fd - g[:, j]).max() < 3e-5 * np.abs(g).max(), j


def test_cnn_layout_is_the_ravel_of_the_layer_dicts():
  model, P = _model()
  geo, F = model.geometry()
  tree = {"head": {"w": np.full((F, CLASSES), 91, np.float32), "b": np.full(CLASSES, 92, np.float32)}}
  for l, (H, W, ci, Ho, Wo, co, st) in enumerate(geo):
    tree[f"conv_{l}"] = {"w": np.full((3, 3, ci, co), 10 * l + 1, np.float32),
                         "b": np.full(co, 10 * l + 2, np.float32)}
  flat, _ = otree.ravel_pytree(tree)
  assert flat.size == P
  for l, (H, W, ci, Ho, Wo, co, st) in enumerate(geo):
    assert np.all(flat[model.w_off[l]:model.w_off[l] + 9 * ci * co] == 10 * l + 1)
    assert np.all(flat[model.b_off[l]:model.b_off[l] + co] == 10 * l + 2)
  L = len(geo)
  assert np.all(flat[model.w_off[L]:model.w_off[L] + F * CLASSES] == 91)
  assert np.all(flat[model.b_off[L]:model.b_off[L] + CLASSES] == 92)
