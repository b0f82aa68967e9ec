"""NumPy f32 restatement of the jax-sgmc sampling hot path (oracle).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): never imported by the
product package.  Every function cites the reference lines it follows.
Arrays are chain-batched and flat: ``theta`` is f32[C, P] where P is the
raveled sample (``jax.flatten_util.ravel_pytree`` order, oracle/tree.py) and C
the number of independent chains (the reference's leading ``list_vmap`` axis,
util/list_map.py:53-56).  Evaluation order of every update follows SURVEY.md
Appendix A; all arithmetic is f32 with one rounding per operation (no FMA) so
the CUDA update kernels, which use explicit ``__fmul_rn``/``__fadd_rn``, can be
compared bit for bit when they are fed the same gradient.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, NamedTuple, Optional, Sequence

import math

import numpy as np

from . import prng

F32 = np.float32


# ----------------------------------------------------------------------------
# noise:  integrator.random_tree  (integrator.py:119-135)
# ----------------------------------------------------------------------------

def random_tree_flat(keys, sizes: Sequence[int], layout="original"):
  """Noise shaped like the raveled sample for every chain.

  ``splits = random.split(key, n_leaves)`` (integrator.py:131) and leaf l gets
  ``random.normal(splits[l], leaf.shape)`` (integrator.py:132); leaves are in
  tree_flatten order, so the flat noise is their concatenation.
  keys: uint32[C, 2] -> f32[C, P].
  """
  keys = np.asarray(keys, dtype=np.uint32)
  leaf_keys = prng.split(keys, len(sizes), layout)           # [C, L, 2]
  parts = [prng.normal(leaf_keys[..., l, :], (sz,), layout)
           for l, sz in enumerate(sizes)]
  return np.concatenate(parts, axis=-1).astype(F32)


# ----------------------------------------------------------------------------
# GLM likelihoods and priors (the "recognised" families of the north star)
# ----------------------------------------------------------------------------

@dataclass
class GaussianLinear:
  """Quickstart model (examples/quickstart.md:158-176).

  ``ell_i = norm.logpdf(y_i - x_i.w, scale=exp(log_sigma))`` with
  jax.scipy.stats.norm.logpdf's operation order:
  ``(log(2 pi s^2) + r^2 / s^2) / -2``.
  """
  d: int
  w_off: int
  log_sigma_off: int

  def loglik(self, theta, X, y):
    w = theta[:, self.w_off:self.w_off + self.d]
    sigma = np.exp(theta[:, self.log_sigma_off]).astype(F32)[:, None]
    r = (y[None, :] - (w @ X.T).astype(F32)).astype(F32)
    s2 = (sigma * sigma).astype(F32)
    ln = np.log((F32(2 * np.pi) * s2).astype(F32)).astype(F32)
    q = ((r * r).astype(F32) / s2).astype(F32)
    ell = ((ln + q).astype(F32) / F32(-2.0)).astype(F32)
    return ell, (r, s2)

  def vjp(self, theta, X, y, aux, cot):
    """sum_i cot[c,i] * d ell_i / d theta  -> f32[C, P]."""
    r, s2 = aux
    g = np.zeros_like(theta)
    dz = ((r / s2).astype(F32) * cot).astype(F32)              # d ell/d(x.w)
    g[:, self.w_off:self.w_off + self.d] = (dz @ X).astype(F32)
    dls = (((r * r).astype(F32) / s2).astype(F32) - F32(1.0)).astype(F32)
    g[:, self.log_sigma_off] = np.sum((dls * cot).astype(F32), axis=1,
                                      dtype=F32)
    return g


@dataclass
class Logistic:
  """Bayesian logistic regression (BASELINE.json configs[1]).

  ``ell_i = y_i z_i - softplus(z_i)``, ``z = x.w (+ b)``, which equals
  ``y log sigmoid(z) + (1-y) log(1-sigmoid(z))`` for y in {0,1};
  ``softplus(z) = max(z,0) + log1p(exp(-|z|))``.
  """
  d: int
  w_off: int
  b_off: int = -1

  def loglik(self, theta, X, y):
    w = theta[:, self.w_off:self.w_off + self.d]
    z = (w @ X.T).astype(F32)
    if self.b_off >= 0:
      z = (z + theta[:, self.b_off][:, None]).astype(F32)
    e = np.exp(-np.abs(z)).astype(F32)
    sp = (np.maximum(z, F32(0)) + np.log1p(e).astype(F32)).astype(F32)
    ell = ((y[None, :] * z).astype(F32) - sp).astype(F32)
    # sigmoid(z) = where(z>=0, 1/(1+e), e/(1+e))
    den = (F32(1.0) + e).astype(F32)
    sig = np.where(z >= 0, F32(1.0) / den, e / den).astype(F32)
    return ell, (sig,)

  def vjp(self, theta, X, y, aux, cot):
    (sig,) = aux
    g = np.zeros_like(theta)
    dz = ((y[None, :] - sig).astype(F32) * cot).astype(F32)
    g[:, self.w_off:self.w_off + self.d] = (dz @ X).astype(F32)
    if self.b_off >= 0:
      g[:, self.b_off] = np.sum(dz, axis=1, dtype=F32)
    return g


@dataclass
class MLPClassifier:
  """Bayesian MLP classifier (BASELINE.json configs[2]: 784-512-512-10, tanh).

  The likelihood a jax-sgmc user writes for it (examples/cifar.md:196-204 with a
  dense network instead of MobileNet): ``logits = apply(sample, x)``,
  ``ell = -softmax_cross_entropy_with_integer_labels(logits, label)``; the
  reference evaluates it per observation under ``vmap`` (potential.py:141-156)
  and differentiates it with ``jax.value_and_grad`` (integrator.py:593, :166).
  Restated: ``h_0 = x``, ``a_l = h_{l-1} W_l + b_l``, ``h_l = tanh(a_l)`` for all
  but the last layer, ``logits = a_L``;
  ``ell = logits[label] - logsumexp(logits)`` with the max-shifted logsumexp of
  ``jax.nn.log_softmax``.  The gradient is the hand-derived reverse pass (what
  reverse-mode AD emits): ``d logits = cot * (onehot - softmax)``,
  ``dW_l = h_{l-1}^T dA_l``, ``db_l = sum_i dA_l``, ``dH_{l-1} = dA_l W_l^T``,
  ``dA_{l-1} = dH_{l-1} * (1 - h_{l-1}^2)``.

  ``sizes``: layer widths (input, hidden..., classes); ``w_off`` / ``b_off``:
  offsets of ``W_l`` (row-major ``[in, out]``) and ``b_l`` in the raveled sample.
  Labels are stored as f32 class indices (the data loaders hold f32 arrays).
  """
  sizes: Sequence[int]
  w_off: Sequence[int]
  b_off: Sequence[int]

  def _params(self, theta, l):
    i, o = self.sizes[l], self.sizes[l + 1]
    W = theta[:, self.w_off[l]:self.w_off[l] + i * o].reshape(-1, i, o)
    b = theta[:, self.b_off[l]:self.b_off[l] + o]
    return W, b

  def loglik(self, theta, X, y):
    C, L = theta.shape[0], len(self.sizes) - 1
    h = np.broadcast_to(np.asarray(X, F32)[None], (C,) + X.shape)
    hs = [h]
    for l in range(L):
      W, b = self._params(theta, l)
      a = (np.matmul(h, W).astype(F32) + b[:, None, :]).astype(F32)
      h = np.tanh(a).astype(F32) if l < L - 1 else a
      hs.append(h)
    logits = hs[-1]
    m = logits.max(axis=2, keepdims=True)
    sh = (logits - m).astype(F32)
    ex = np.exp(sh).astype(F32)
    se = np.sum(ex, axis=2, keepdims=True, dtype=F32)
    lsm = (sh - np.log(se).astype(F32)).astype(F32)            # log_softmax
    lab = np.asarray(y).astype(np.int64)
    ell = np.take_along_axis(lsm, np.broadcast_to(lab[None, :, None], (C, len(lab), 1)),
                             axis=2)[..., 0].astype(F32)
    soft = (ex / se).astype(F32)
    return ell, (hs, soft, lab)

  def vjp(self, theta, X, y, aux, cot):
    hs, soft, lab = aux
    C, L = theta.shape[0], len(self.sizes) - 1
    g = np.zeros_like(theta)
    onehot = np.zeros(soft.shape[1:], F32)
    onehot[np.arange(len(lab)), lab] = 1
    dA = ((onehot[None] - soft).astype(F32) * cot[..., None]).astype(F32)
    for l in range(L - 1, -1, -1):
      i, o = self.sizes[l], self.sizes[l + 1]
      W, _ = self._params(theta, l)
      dW = np.matmul(np.swapaxes(hs[l], 1, 2), dA).astype(F32)          # [C, in, out]
      g[:, self.w_off[l]:self.w_off[l] + i * o] = dW.reshape(C, -1)
      g[:, self.b_off[l]:self.b_off[l] + o] = np.sum(dA, axis=1, dtype=F32)
      if l > 0:
        dH = np.matmul(dA, np.swapaxes(W, 1, 2)).astype(F32)
        dA = (dH * (F32(1.0) - (hs[l] * hs[l]).astype(F32)).astype(F32)).astype(F32)
    return g


def mlp_layout(sizes, order="haiku"):
  """Offsets of (W_l, b_l) in the raveled sample for the pytree
  ``{"layer_0": {"b": [out], "w": [in, out]}, ...}`` (tree_flatten visits dict keys
  sorted: b before w inside a layer; layers by name)."""
  w_off, b_off, off = [], [], 0
  for l in range(len(sizes) - 1):
    i, o = sizes[l], sizes[l + 1]
    b_off.append(off)
    off += o
    w_off.append(off)
    off += i * o
  return w_off, b_off, off


def _im2col(h, stride):
  """[..., n, H, W, C] -> patches [..., n, Ho, Wo, 9 C] of the 3x3 / zero-padding-1
  convolution, patch index (kh*3 + kw)*C + c (the row-major order of an HWIO filter)."""
  H, W = h.shape[-3], h.shape[-2]
  Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
  pad = [(0, 0)] * (h.ndim - 3) + [(1, 1), (1, 1), (0, 0)]
  hp = np.pad(h, pad)
  cols = []
  for kh in range(3):
    for kw in range(3):
      cols.append(hp[..., kh:kh + (Ho - 1) * stride + 1:stride,
                     kw:kw + (Wo - 1) * stride + 1:stride, :])
  return np.concatenate(cols, axis=-1)


def _col2im(dP, H, W, C, stride):
  """Adjoint of :func:`_im2col`: patch gradients [..., n, Ho, Wo, 9 C] -> [..., n, H, W, C]."""
  Ho, Wo = dP.shape[-3], dP.shape[-2]
  out = np.zeros(dP.shape[:-3] + (H + 2, W + 2, C), F32)
  for kh in range(3):
    for kw in range(3):
      k = (kh * 3 + kw) * C
      out[..., kh:kh + (Ho - 1) * stride + 1:stride,
          kw:kw + (Wo - 1) * stride + 1:stride, :] += dP[..., k:k + C]
  return out[..., 1:H + 1, 1:W + 1, :].astype(F32)


@dataclass
class CNNClassifier:
  """Bayesian CNN classifier (BASELINE.json configs[4]: SGGMC / AMAGOLD on CIFAR-10-shape
  data): the likelihood of examples/cifar.md:196-204 with a small convolutional network,
  ``ell = -softmax_cross_entropy_with_integer_labels(apply(sample, x), label)``, evaluated
  per observation under vmap (potential.py:141-156) and differentiated with
  ``jax.value_and_grad``.  Restated: 3x3 convolutions (zero padding 1, ``strides[l]``,
  filters HWIO ``[3, 3, cin, cout]``) as im2col + matmul, tanh, a dense head on the
  flattened NHWC feature map; the gradient is the hand-derived reverse pass
  (``dW = patches^T dZ``, ``dPatches = dZ W^T``, col2im, ``* (1 - h^2)``).

  ``image``: (H, W, Cin); ``channels``: conv output channels; ``w_off`` / ``b_off``: offsets
  of the conv layers' then the head's weights / biases in the raveled sample.  ``X`` rows
  are flattened NHWC images."""
  image: Sequence[int]
  channels: Sequence[int]
  strides: Sequence[int]
  n_classes: int
  w_off: Sequence[int]
  b_off: Sequence[int]

  def geometry(self):
    H, W, c = self.image
    out = []
    for co, st in zip(self.channels, self.strides):
      Ho, Wo = (H - 1) // st + 1, (W - 1) // st + 1
      out.append((H, W, c, Ho, Wo, co, st))
      H, W, c = Ho, Wo, co
    return out, H * W * c

  def loglik(self, theta, X, y):
    C, n = theta.shape[0], X.shape[0]
    geo, Fdim = self.geometry()
    h = np.broadcast_to(np.asarray(X, F32).reshape((1, n) + tuple(self.image)),
                        (C, n) + tuple(self.image))
    hs, Ps = [h], []
    for l, (H, W, ci, Ho, Wo, co, st) in enumerate(geo):
      Wl = theta[:, self.w_off[l]:self.w_off[l] + 9 * ci * co].reshape(C, 9 * ci, co)
      bl = theta[:, self.b_off[l]:self.b_off[l] + co]
      P = _im2col(h, st).reshape(C, n * Ho * Wo, 9 * ci)
      a = (np.matmul(P, Wl).astype(F32) + bl[:, None, :]).astype(F32)
      h = np.tanh(a).astype(F32).reshape(C, n, Ho, Wo, co)
      Ps.append(P)
      hs.append(h)
    L = len(geo)
    Wh = theta[:, self.w_off[L]:self.w_off[L] + Fdim * self.n_classes].reshape(C, Fdim, -1)
    bh = theta[:, self.b_off[L]:self.b_off[L] + self.n_classes]
    flat = h.reshape(C, n, Fdim)
    logits = (np.matmul(flat, Wh).astype(F32) + bh[:, None, :]).astype(F32)
    m = logits.max(axis=2, keepdims=True)
    sh = (logits - m).astype(F32)
    ex = np.exp(sh).astype(F32)
    se = np.sum(ex, axis=2, keepdims=True, dtype=F32)
    lsm = (sh - np.log(se).astype(F32)).astype(F32)
    lab = np.asarray(y).astype(np.int64)
    ell = np.take_along_axis(lsm, np.broadcast_to(lab[None, :, None], (C, n, 1)),
                             axis=2)[..., 0].astype(F32)
    return ell, (hs, Ps, (ex / se).astype(F32), lab)

  def vjp(self, theta, X, y, aux, cot):
    hs, Ps, soft, lab = aux
    C, n = theta.shape[0], len(lab)
    geo, Fdim = self.geometry()
    L = len(geo)
    g = np.zeros_like(theta)
    onehot = np.zeros(soft.shape[1:], F32)
    onehot[np.arange(n), lab] = 1
    dlog = ((onehot[None] - soft).astype(F32) * cot[..., None]).astype(F32)
    Wh = theta[:, self.w_off[L]:self.w_off[L] + Fdim * self.n_classes].reshape(C, Fdim, -1)
    flat = hs[L].reshape(C, n, Fdim)
    g[:, self.w_off[L]:self.w_off[L] + Fdim * self.n_classes] = \
        np.matmul(np.swapaxes(flat, 1, 2), dlog).astype(F32).reshape(C, -1)
    g[:, self.b_off[L]:self.b_off[L] + self.n_classes] = np.sum(dlog, axis=1, dtype=F32)
    dH = np.matmul(dlog, np.swapaxes(Wh, 1, 2)).astype(F32)
    dZ = (dH * (F32(1.0) - (flat * flat).astype(F32)).astype(F32)).astype(F32)
    for l in range(L - 1, -1, -1):
      H, W, ci, Ho, Wo, co, st = geo[l]
      dZm = dZ.reshape(C, n * Ho * Wo, co)
      Wl = theta[:, self.w_off[l]:self.w_off[l] + 9 * ci * co].reshape(C, 9 * ci, co)
      g[:, self.w_off[l]:self.w_off[l] + 9 * ci * co] = \
          np.matmul(np.swapaxes(Ps[l], 1, 2), dZm).astype(F32).reshape(C, -1)
      g[:, self.b_off[l]:self.b_off[l] + co] = np.sum(dZm, axis=1, dtype=F32)
      if l > 0:
        dP = np.matmul(dZm, np.swapaxes(Wl, 1, 2)).astype(F32).reshape(C, n, Ho, Wo, 9 * ci)
        dHl = _col2im(dP, H, W, ci, st)
        dZ = (dHl * (F32(1.0) - (hs[l] * hs[l]).astype(F32)).astype(F32)).astype(F32)
    return g


def cnn_layout(image, channels, strides, n_classes):
  """Offsets in the raveled sample of the pytree ``{"conv_0": {"b", "w"}, ..., "head":
  {"b", "w"}}`` (tree_flatten: keys sorted -- conv_* before head, b before w)."""
  H, W, c = image
  w_off, b_off, off = [], [], 0
  for co, st in zip(channels, strides):
    b_off.append(off)
    off += co
    w_off.append(off)
    off += 9 * c * co
    H, W, c = (H - 1) // st + 1, (W - 1) // st + 1, co
  b_off.append(off)
  off += n_classes
  w_off.append(off)
  off += H * W * c * n_classes
  return w_off, b_off, off


@dataclass
class Prior:
  """Log-priors used by the reference examples.

  kind ``"gaussian"``: sum over [off, off+size) of ``-0.5 (theta/scale)^2``
  (normalising constant dropped -- it has no gradient; the reference's
  examples/cifar.md:214-219 uses ``norm.logpdf`` sums, whose constant only
  shifts U).  kind ``"inv_sigma"``: ``1/exp(theta[off])`` -- the quickstart's
  (unusual) log-prior, examples/quickstart.md:172-173.  kind ``"flat"``: 0.
  """
  kind: str = "flat"
  off: int = 0
  size: int = 0
  scale: float = 1.0

  def value(self, theta):
    C = theta.shape[0]
    if self.kind == "flat":
      return np.zeros(C, F32)
    if self.kind == "gaussian":
      t = theta[:, self.off:self.off + self.size]
      inv = F32(1.0) / F32(self.scale * self.scale)
      return (F32(-0.5) * inv * np.sum((t * t).astype(F32), axis=1,
                                       dtype=F32)).astype(F32)
    if self.kind == "inv_sigma":
      return (F32(1.0) / np.exp(theta[:, self.off]).astype(F32)).astype(F32)
    raise ValueError(self.kind)

  def grad(self, theta):
    g = np.zeros_like(theta)
    if self.kind == "gaussian":
      inv = F32(1.0) / F32(self.scale * self.scale)
      g[:, self.off:self.off + self.size] = (
          -theta[:, self.off:self.off + self.size] * inv).astype(F32)
    elif self.kind == "inv_sigma":
      g[:, self.off] = (-(F32(1.0) / np.exp(theta[:, self.off]).astype(F32))
                        ).astype(F32)
    return g


# ----------------------------------------------------------------------------
# potential.minibatch_potential / full_potential   (potential.py:94-293)
# ----------------------------------------------------------------------------

def minibatch_potential(model, prior: Prior, temperature: float = 1.0):
  """``potential_fn(theta, (X_b, y_b), N, mask=None) -> (U, ell, grad)``.

  potential.py:159-214: ``L = -N * mean(ell)`` (:183) or
  ``-N/n * dot(ell, mask)`` (:185); ``U = (L - prior) / T`` (:210).  The
  gradient is the reverse-mode derivative of exactly that expression
  (integrator.py:166,593,792 call ``value_and_grad``): cotangent
  ``(1/T) * (-N) / n`` on every ell_i, ``-(1/T)`` on the prior.
  """
  T = F32(temperature)

  def potential_fn(theta, batch, N, mask=None):
    X, y = batch
    n = X.shape[0]
    theta = np.asarray(theta, F32)
    ell, aux = model.loglik(theta, X, y)
    if mask is None:
      L = (F32(-N) * (np.sum(ell, axis=1, dtype=F32) / F32(n))).astype(F32)
      cot = np.full(ell.shape, (F32(-N) / F32(n)) / T, dtype=F32)
    else:
      m = np.asarray(mask, F32)
      L = ((F32(-N) / F32(n)) * (ell @ m).astype(F32)).astype(F32)
      cot = (((F32(-N) / F32(n)) / T) * m)[None, :].astype(F32)
      cot = np.broadcast_to(cot, ell.shape)
    pv = prior.value(theta)
    U = ((L - pv).astype(F32) / T).astype(F32)
    g = model.vjp(theta, X, y, aux, cot)
    g = (g - (prior.grad(theta) / T).astype(F32)).astype(F32)
    return U, ell, g

  return potential_fn


def full_potential(model, prior: Prior, temperature: float = 1.0):
  """potential.py:219-293: ``U = (sum_b -dot(ell_b, mask_b) - prior) / T``.

  ``batches`` is an iterable of ``(X_b, y_b, mask_b)`` as produced by
  ``full_reference_data`` (data/core.py:586-627): the last batch wraps indices
  modulo N and masks the overhang (core.py:571-572).  The inner potential has
  zero prior and T=1 and is un-scaled by n/N (potential.py:258-271).
  """
  inner = minibatch_potential(model, Prior("flat"), 1.0)
  T = F32(temperature)

  def full_fn(theta, batches, N):
    theta = np.asarray(theta, F32)
    total = np.zeros(theta.shape[0], F32)
    for X, y, mask in batches:
      n = X.shape[0]
      u, _, _ = inner(theta, (X, y), N, mask=mask)
      total = (total + (u * F32(n) / F32(N)).astype(F32)).astype(F32)
    return ((total - prior.value(theta)).astype(F32) / T).astype(F32)

  return full_fn


# ----------------------------------------------------------------------------
# adaption.rms_prop   (adaption.py:225-293)
# ----------------------------------------------------------------------------

def rms_prop_init(theta, alpha=0.9, lmbd=1e-5):
  """adaption.py:238-252: v = ones_like(sample)."""
  return np.ones_like(theta, dtype=F32), F32(alpha), F32(lmbd)


def rms_prop_update(state, grad):
  """adaption.py:270-272: v' = alpha v + (1 - alpha) g^2 (raw gradient)."""
  v, alpha, lmbd = state
  one_m = (F32(1.0) - alpha).astype(F32)
  new_v = ((alpha * v).astype(F32)
           + (one_m * (grad * grad).astype(F32)).astype(F32)).astype(F32)
  return new_v, alpha, lmbd


def rms_prop_get(state):
  """adaption.py:289-291: G = (lmbd + sqrt(v))^-1, sqrt(G), Gamma = 0."""
  v, _, lmbd = state
  g = (F32(1.0) / (lmbd + np.sqrt(v).astype(F32)).astype(F32)).astype(F32)
  return g, np.sqrt(g).astype(F32), np.zeros_like(g)


# ----------------------------------------------------------------------------
# integrator.langevin_diffusion   (integrator.py:767-924)
# ----------------------------------------------------------------------------

class LangevinState(NamedTuple):
  theta: np.ndarray            # f32[C, P]
  key: np.ndarray              # u32[C, 2]
  v: Optional[np.ndarray]      # rms_prop second moment or None
  potential: np.ndarray        # f32[C]
  variance: np.ndarray         # f32[C]


def langevin_init(theta, keys=None, rms=False):
  """integrator.py:803-846.  Default key PRNGKey(0) for every chain (:804)."""
  theta = np.array(theta, dtype=F32)
  C = theta.shape[0]
  if keys is None:
    keys = np.tile(prng.PRNGKey(0), (C, 1))
  return LangevinState(theta, np.array(keys, dtype=np.uint32),
                       np.ones_like(theta) if rms else None,
                       np.zeros(C, F32), np.ones(C, F32))


def sgld_scales(step_size, temperature):
  """integrator.py:882-884: (-eps), sqrt(2*T*eps) as f32 scalars."""
  eps = F32(step_size)
  neg_eps = F32(-eps)
  noise_scale = np.sqrt(F32(F32(F32(2.0) * F32(temperature)) * eps)).astype(F32)
  return neg_eps, F32(noise_scale)


def sgld_apply(theta, grad, xi, step_size, temperature, v=None,
               alpha=0.9, lmbd=1e-5):
  """Elementwise part of one SGLD / pSGLD step.

  Returns (theta', v').  integrator.py:882-912, adaption.py:254-291.
  """
  neg_eps, ns = sgld_scales(step_size, temperature)
  sg = (neg_eps * grad).astype(F32)
  sn = (ns * xi).astype(F32)
  if v is None:
    delta = (sg + sn).astype(F32)
    new_v = None
  else:
    new_v, _, _ = rms_prop_update((v, F32(alpha), F32(lmbd)), grad)
    G, S, _ = rms_prop_get((new_v, F32(alpha), F32(lmbd)))
    # ((eps*Gamma) + (G*sg)) + (S*sn), Gamma == 0     (integrator.py:903-909)
    delta = ((F32(0.0) + (G * sg).astype(F32)).astype(F32)
             + (S * sn).astype(F32)).astype(F32)
  return (theta + delta).astype(F32), new_v


def langevin_update(state: LangevinState, grad_fn: Callable, sizes,
                    step_size, temperature, alpha=0.9, lmbd=1e-5,
                    layout="original"):
  """One ``langevin_diffusion.update_fn`` (integrator.py:860-922).

  ``grad_fn(theta) -> (U[C], ell[C,n], grad[C,P])`` is the value_and_grad of
  the stochastic potential on this step's minibatch (:875-879).
  """
  ks = prng.split(state.key, 2, layout)                     # :871
  new_key, sub = ks[..., 0, :], ks[..., 1, :]
  xi = random_tree_flat(sub, sizes, layout)                 # :874
  U, ell, g = grad_fn(state.theta)
  mean = (np.sum(ell, axis=1, dtype=F32) / F32(ell.shape[1])).astype(F32)
  dev = (ell - mean[:, None]).astype(F32)
  var = (np.sum((dev * dev).astype(F32), axis=1, dtype=F32)
         / F32(ell.shape[1])).astype(F32)                   # jnp.var  :880
  theta, v = sgld_apply(state.theta, g, xi, step_size, temperature,
                        state.v, alpha, lmbd)
  return LangevinState(theta, new_key, v, U.astype(F32), var)


# ----------------------------------------------------------------------------
# integrator.friction_leapfrog (SGHMC)   (integrator.py:563-765)
# ----------------------------------------------------------------------------

class LeapfrogState(NamedTuple):
  theta: np.ndarray
  momentum: np.ndarray
  key: np.ndarray
  potential: np.ndarray


def leapfrog_init(theta, keys=None):
  """integrator.py:668-713: momentum placeholder = the sample (:702)."""
  theta = np.array(theta, dtype=F32)
  C = theta.shape[0]
  if keys is None:
    keys = np.tile(prng.PRNGKey(0), (C, 1))
  return LeapfrogState(theta, theta.copy(), np.array(keys, np.uint32),
                       np.zeros(C, F32))


def sghmc_inner_apply(theta_new, p, m, grad, xi, step_size, friction, cb_diff_sqrt=None):
  """Momentum part of ``_body_fun`` after the gradient (integrator.py:616-655).

  m = M^-1 p (computed by the caller from the *old* momentum, :610).
  p1 = p + ((-eps*C) * m); p2 = p1 + ((-eps)*g); p3 = p2 + (C*(sqrt(2 eps)*xi)), or with a
  noise model (:632-650) p3 = p2 + sqrt(2 eps)*(cb_diff_sqrt*xi).
  """
  eps = F32(step_size)
  C = np.asarray(friction, F32)
  p1 = (p + ((F32(-eps) * C).astype(F32) * m).astype(F32)).astype(F32)
  p2 = (p1 + (F32(-eps) * grad).astype(F32)).astype(F32)
  ns = np.sqrt(F32(F32(2.0) * eps)).astype(F32)
  if cb_diff_sqrt is not None:
    return (p2 + (ns * (np.asarray(cb_diff_sqrt, F32) * xi).astype(F32)).astype(F32)).astype(F32)
  p3 = (p2 + (C * (ns * xi).astype(F32)).astype(F32)).astype(F32)
  return p3


def friction_leapfrog_integrate(state: LeapfrogState, grad_fns, sizes,
                                step_size, friction=0.25, mass=None,
                                layout="original", noise_model_fn=None):
  """``friction_leapfrog.integrate`` (integrator.py:716-757).

  ``noise_model_fn(theta_new, grad, step_index) -> cb_diff_sqrt`` stands for
  ``get_noise_model`` evaluated on the step's minibatch (:632-647).

  ``grad_fns`` is a sequence of ``steps`` callables, one per inner step (each
  inner step draws a fresh minibatch, :621).  ``mass`` is the flat diagonal
  mass (f32[P]) or None for unit mass (:722-726).
  """
  P = state.theta.shape[1]
  mass = np.ones(P, F32) if mass is None else np.asarray(mass, F32)
  inv_m = (F32(1.0) / mass).astype(F32)                      # :108-110
  sqrt_m = np.sqrt(mass).astype(F32)                         # :111-113
  fr = np.asarray(friction, F32)
  fr = np.full(P, fr, F32) if fr.ndim == 0 else fr           # :729-733
  ks = prng.split(state.key, 2, layout)                      # :736
  key, sub = ks[..., 0, :], ks[..., 1, :]
  p = (sqrt_m * random_tree_flat(sub, sizes, layout)).astype(F32)  # :737-738
  theta = state.theta
  U = np.zeros(theta.shape[0], F32)
  eps = F32(step_size)
  step_i = 0
  for grad_fn in grad_fns:                                   # :749-755
    m = (inv_m * p).astype(F32)                              # :610
    theta = (theta + (eps * m).astype(F32)).astype(F32)      # :611-612
    U, _, g = grad_fn(theta)                                 # :621-625
    ks = prng.split(key, 2, layout)                          # :630
    key, sub = ks[..., 0, :], ks[..., 1, :]
    xi = random_tree_flat(sub, sizes, layout)                # :631
    cb = None if noise_model_fn is None else noise_model_fn(theta, g, step_i)
    p = sghmc_inner_apply(theta, p, m, g, xi, eps, fr, cb)
    step_i += 1
  return LeapfrogState(theta, p, key, U.astype(F32))


# ----------------------------------------------------------------------------
# integrator.obabo   (integrator.py:138-346)
# ----------------------------------------------------------------------------

class ObaboState(NamedTuple):
  theta: np.ndarray
  momentum: np.ndarray
  key: np.ndarray
  potential: np.ndarray
  kinetic_energy_start: np.ndarray
  kinetic_energy_end: np.ndarray


def obabo_init(theta, keys=None):
  """integrator.py:275-315: zero momentum (:303), zero KE accumulators."""
  theta = np.array(theta, dtype=F32)
  C = theta.shape[0]
  if keys is None:
    keys = np.tile(prng.PRNGKey(0), (C, 1))
  z = np.zeros(C, F32)
  return ObaboState(theta, np.zeros_like(theta), np.array(keys, np.uint32),
                    z, z.copy(), z.copy())


def _tree_dot(a, b, sizes):
  """util/tree_util.py:131-133: Python sum over per-leaf jnp.sum(a*b)."""
  total = None
  off = 0
  for sz in sizes:
    s = np.sum((a[:, off:off + sz] * b[:, off:off + sz]).astype(F32), axis=1,
               dtype=F32)
    total = s if total is None else (total + s).astype(F32)
    off += sz
  return total


def obabo_o_step(p, xi, step_size, temperature, friction, sqrt_m):
  """``_momentum_resampling`` (integrator.py:192-200)."""
  # exp evaluated in f64 libm and rounded once (same as the C host code)
  a = F32(math.exp(float(F32(-F32(friction) * F32(step_size)))))
  ns = np.sqrt(F32(F32(F32(1.0) - a) * F32(temperature))).astype(F32)
  sa = np.sqrt(a).astype(F32)
  return ((sa * p).astype(F32)
          + (ns * (sqrt_m * xi).astype(F32)).astype(F32)).astype(F32)


def obabo_integrate(state: ObaboState, grad_fn_pairs, sizes, step_size,
                    temperature=1.0, friction=1.0, mass=None,
                    layout="original", mass_matrix=None):
  """``obabo.integrate`` (integrator.py:318-338; step :203-273).

  ``grad_fn_pairs``: one ``(grad_fn_1, grad_fn_2)`` per step -- two minibatch
  draws and two gradient evaluations per step (:225, :243).  ``mass_matrix``: an
  adapted ``MassMatrix(inv, sqrt)`` (adaption.py:296-369; arrays broadcastable to the
  positions) used as handed over instead of ``init_mass(mass)`` (:320-321).
  """
  P = state.theta.shape[1]
  mass = np.ones(P, F32) if mass is None else np.asarray(mass, F32)
  inv_m = (F32(1.0) / mass).astype(F32)
  sqrt_m = np.sqrt(mass).astype(F32)
  if mass_matrix is not None:
    inv_m, sqrt_m = (np.asarray(a, F32) for a in mass_matrix)
  eps = F32(step_size)
  theta, p, key = state.theta, state.momentum, state.key
  ke_s, ke_e = state.kinetic_energy_start, state.kinetic_energy_end
  U = state.potential
  for g1_fn, g2_fn in grad_fn_pairs:
    ks = prng.split(key, 3, layout)                          # :208
    key, k1, k2 = ks[..., 0, :], ks[..., 1, :], ks[..., 2, :]
    p1 = obabo_o_step(p, random_tree_flat(k1, sizes, layout), eps,
                      temperature, friction, sqrt_m)         # :210-214
    ke_s = (ke_s + (F32(0.5) * _tree_dot(p1, (inv_m * p1).astype(F32), sizes)
                    ).astype(F32)).astype(F32)               # :222
    U1, _, g1 = g1_fn(theta)                                 # :225-229
    half = F32(F32(-1.0) * F32(F32(0.5) * eps))              # :188
    p2 = ((half * g1).astype(F32) + p1).astype(F32)          # :230-233
    theta = (theta + (eps * (inv_m * p2).astype(F32)).astype(F32)
             ).astype(F32)                                   # :236-240
    U2, _, g2 = g2_fn(theta)                                 # :243-247
    p3 = ((half * g2).astype(F32) + p2).astype(F32)          # :248-251
    p4 = obabo_o_step(p3, random_tree_flat(k2, sizes, layout), eps,
                      temperature, friction, sqrt_m)         # :253-257
    ke_e = (ke_e + (F32(0.5) * _tree_dot(p3, (inv_m * p3).astype(F32), sizes)
                    ).astype(F32)).astype(F32)               # :261
    U = (F32(0.5) * (U1 + U2).astype(F32)).astype(F32)       # :264
    p = p4
  return ObaboState(theta, p, key, U, ke_s, ke_e)


# ----------------------------------------------------------------------------
# solver.parallel_tempering (reSGLD)   (solver.py:220-299)
# ----------------------------------------------------------------------------

class TemperingState(NamedTuple):
  normal: LangevinState
  hot: LangevinState
  ssq: np.ndarray      # f32[S]   one entry per reSGLD system
  F: np.ndarray        # f32[S]
  step: int
  key: np.ndarray      # u32[S, 2]


def parallel_tempering_init(normal_theta, hot_theta, ssq_init=0.0, keys=None,
                            F=1.0, rms=False, layout="original"):
  """solver.py:248-259: ``key, split1, split2 = split(key, 3)``."""
  normal_theta = np.asarray(normal_theta, F32)
  S = normal_theta.shape[0]
  if keys is None:
    keys = np.tile(prng.PRNGKey(0), (S, 1))
  ks = prng.split(np.asarray(keys, np.uint32), 3, layout)
  key, s1, s2 = ks[..., 0, :], ks[..., 1, :], ks[..., 2, :]
  return TemperingState(langevin_init(normal_theta, s1, rms),
                        langevin_init(hot_theta, s2, rms),
                        np.full(S, ssq_init, F32), np.full(S, F, F32), 0, key)


def resgld_swap_decision(U_n, U_h, var_n, ssq, F, step, T_n, T_h, keys,
                         layout="original", sa_schedule=None):
  """Swap arithmetic of solver.py:273-291.

  Returns (exchange[S] bool, ssq'[S], key'[S,2], log_s, log_u).  The reference
  exchanges the chains iff ``not (log_u < log_s)`` (:287-291) -- inverted
  with respect to the paper; reproduced as is.
  """
  # sa_schedule(step) (:274), default 1 / n (:221)
  eta = F32(1.0) / F32(step) if sa_schedule is None else F32(sa_schedule(step))
  ssq = (((F32(1.0) - eta).astype(F32) * ssq).astype(F32)
         + (eta * var_n).astype(F32)).astype(F32)            # :275-276
  temps = (F32(1.0) / F32(T_n) - F32(1.0) / F32(T_h)).astype(F32)   # :279
  corr = ((temps * ssq).astype(F32) / F).astype(F32)         # :280
  log_s = (temps * ((U_n - U_h).astype(F32) - corr).astype(F32)).astype(F32)
  ks = prng.split(keys, 2, layout)                           # :283
  key, sub = ks[..., 0, :], ks[..., 1, :]
  u = prng.uniform(sub, (), layout=layout)                   # :284
  log_u = prng.log_libdevice(u)
  exchange = ~(log_u < log_s)
  return exchange, ssq, key, log_s, log_u


def parallel_tempering_update(state: TemperingState, grad_fn_normal,
                              grad_fn_hot, sizes, step_size, T_normal, T_hot,
                              step_size_hot=None, layout="original", sa_schedule=None):
  """solver.py:264-293: update both chains, then maybe exchange whole states."""
  step = state.step + 1                                      # :267
  eps_h = step_size if step_size_hot is None else step_size_hot
  normal = langevin_update(state.normal, grad_fn_normal, sizes, step_size,
                           T_normal, layout=layout)          # :270
  hot = langevin_update(state.hot, grad_fn_hot, sizes, eps_h, T_hot,
                        layout=layout)                       # :271
  exchange, ssq, key, _, _ = resgld_swap_decision(
      normal.potential, hot.potential, normal.variance, state.ssq, state.F,
      step, T_normal, T_hot, state.key, layout, sa_schedule)

  def pick(a, b):
    if a is None:
      return None, None
    m = exchange.reshape((-1,) + (1,) * (a.ndim - 1))
    return np.where(m, b, a), np.where(m, a, b)

  fields = [pick(a, b) for a, b in zip(normal, hot)]
  normal = LangevinState(*[f[0] for f in fields])
  hot = LangevinState(*[f[1] for f in fields])
  return TemperingState(normal, hot, ssq, state.F, step, key), exchange


# ----------------------------------------------------------------------------
# integrator.reversible_leapfrog (AMAGOLD)   (integrator.py:349-560)
# ----------------------------------------------------------------------------

def reversible_leapfrog_init(theta, keys=None, mass=None, sizes=None,
                             layout="original", mass_matrix=None):
  """``reversible_leapfrog.init_fn`` (integrator.py:472-512): ``key, split =
  split(key)``; momentum = sqrt(m) * random_tree(split, sample); potential 0."""
  theta = np.asarray(theta, F32)
  C, P = theta.shape
  if keys is None:
    keys = np.tile(prng.PRNGKey(0), (C, 1))
  sizes = [P] if sizes is None else sizes
  sqrt_m = np.ones(P, F32) if mass is None else np.sqrt(np.asarray(mass, F32)).astype(F32)
  if mass_matrix is not None:
    sqrt_m = np.asarray(mass_matrix[1], F32)
  ks = prng.split(np.asarray(keys, np.uint32), 2, layout)
  key, sub = ks[..., 0, :], ks[..., 1, :]
  p = (sqrt_m * random_tree_flat(sub, sizes, layout)).astype(F32)
  return LeapfrogState(theta, p, key, np.zeros(C, F32))


def reversible_leapfrog_integrate(state: LeapfrogState, grad_fns, sizes, step_size,
                                  friction=0.25, mass=None, layout="original",
                                  mass_matrix=None):
  """``reversible_leapfrog.integrate`` (integrator.py:514-553, body :403-466).

  ``grad_fns``: one ``theta -> (U, ell, grad)`` per inner step (only the
  gradient is used, :375).  Half position step, ``steps`` bodies (the first one
  skips the position update, :411-418), half position step; the accumulated
  energy starts at 0 (:531) and is returned in ``potential``.
  """
  P = state.theta.shape[1]
  mass = np.ones(P, F32) if mass is None else np.asarray(mass, F32)
  inv_m = (F32(1.0) / mass).astype(F32)
  sqrt_m = np.sqrt(mass).astype(F32)
  if mass_matrix is not None:                                # adapted MassMatrix(inv, sqrt)
    inv_m, sqrt_m = (np.asarray(a, F32) for a in mass_matrix)
  eps = F32(step_size)
  f = F32(friction)
  half_eps = F32(F32(0.5) * eps)

  def position_update(scale, pos, mom):                      # :378-381
    return (pos + (scale * (inv_m * mom).astype(F32)).astype(F32)).astype(F32)

  theta = position_update(half_eps, state.theta, state.momentum)   # :523-524
  p, key = state.momentum, state.key
  energy = np.zeros(theta.shape[0], F32)                     # :531
  noise_scale = np.sqrt(F32(F32(4.0 * friction) * eps)).astype(F32)   # :424
  decay = F32(F32(1.0) - F32(eps * f))                       # :436
  norm = F32(F32(1.0) / F32(F32(1.0) + F32(eps * f)))        # :446
  neg_eps = F32(F32(-1.0) * eps)                             # :439
  for step, g_fn in enumerate(grad_fns):
    if step > 0:
      theta = position_update(eps, theta, p)                 # :411-418
    ks = prng.split(key, 2, layout)                          # :420
    key, sub = ks[..., 0, :], ks[..., 1, :]
    noise = (sqrt_m * random_tree_flat(sub, sizes, layout)).astype(F32)   # :383-386
    scaled_noise = (noise_scale * noise).astype(F32)         # :423-425
    _, _, g = g_fn(theta)                                    # :427-431
    un = (((decay * p).astype(F32) + (neg_eps * g).astype(F32)).astype(F32)
          + scaled_noise).astype(F32)                        # :435-444
    pn = (norm * un).astype(F32)                             # :445-447
    dot = _tree_dot((p + pn).astype(F32), (inv_m * g).astype(F32), sizes)   # :395-399
    energy = ((half_eps * dot).astype(F32) + energy).astype(F32)   # :400, :455
    p = pn
  theta = position_update(half_eps, theta, p)                # :545-546
  return LeapfrogState(theta, p, key, energy)


# ----------------------------------------------------------------------------
# adaption.mass_matrix, diagonal   (adaption.py:296-369)
# ----------------------------------------------------------------------------

class MassState(NamedTuple):
  iteration: int
  mean: np.ndarray     # f32[C, P]
  ssq: np.ndarray      # f32[C, P]
  m_inv: np.ndarray    # f32[C, P]
  m_sqrt: np.ndarray   # f32[C, P]


def mass_matrix_init(theta, init_cov=None) -> MassState:
  """adaption.py:319-341: zero statistics; ``m_inv = init_cov`` (ones by default),
  ``m_sqrt = 1 / sqrt(init_cov)``."""
  theta = np.asarray(theta, F32)
  cov = np.ones_like(theta) if init_cov is None else \
      np.broadcast_to(np.asarray(init_cov, F32), theta.shape).astype(F32)
  return MassState(0, np.zeros_like(theta), np.zeros_like(theta), cov,
                   (F32(1.0) / np.sqrt(cov)).astype(F32))


def mass_matrix_update(state: MassState, sample, burn_in: int) -> MassState:
  """adaption.py:343-363 (Welford); the matrix is replaced once, in iteration
  ``burn_in`` (:310-313: ``inv = ssq / n``, ``sqrt = sqrt(n / ssq)``).  The iteration
  counter is an int32: ``(n - 1) / n`` and ``1 / n`` are f32 true divisions."""
  x = np.asarray(sample, F32)
  it = state.iteration + 1
  fi = F32(it)
  w_old, w_new = F32(F32(it - 1) / fi), F32(F32(1.0) / fi)
  new_mean = ((w_old * state.mean).astype(F32) + (w_new * x).astype(F32)).astype(F32)
  ssq = (state.ssq + ((x - state.mean).astype(F32) * (x - new_mean).astype(F32)).astype(F32)
         ).astype(F32)
  m_inv, m_sqrt = state.m_inv, state.m_sqrt
  if it == burn_in:
    with np.errstate(divide="ignore", invalid="ignore"):
      m_inv = (ssq / fi).astype(F32)
      m_sqrt = np.sqrt((fi / ssq).astype(F32)).astype(F32)
  return MassState(it, new_mean, ssq, m_inv, m_sqrt)


# ----------------------------------------------------------------------------
# adaption.fisher_information, diagonal   (adaption.py:372-457)
# ----------------------------------------------------------------------------

def fisher_information_get(model, theta, batch, N, sample_grad, friction, step_size):
  """``fisher_information(diagonal=True).get`` (adaption.py:391-438) for C chains.

  ``sample_grad``: gradient of the stochastic potential at ``theta`` on ``batch``;
  ``friction``: scalar or f32[P].  Per-observation gradients ``grad(likelihoods[i])`` come
  from the model's vjp with a one-hot cotangent.  Returns ``(noise_scale, scale)`` =
  ``(cb_diff_sqrt, b_sqrt)``, f32[C, P] each.  The reference subtracts the un-scaled
  POTENTIAL gradient from the per-observation LIKELIHOOD gradient (:404, :416); kept."""
  X, y = batch
  n = X.shape[0]
  theta = np.asarray(theta, F32)
  C, P = theta.shape
  m = (np.asarray(sample_grad, F32) / F32(N)).astype(F32)                     # :404
  ell, aux = model.loglik(theta, X, y)
  ssq = np.zeros((C, P), F32)
  for i in range(n):
    cot = np.zeros_like(ell)
    cot[:, i] = F32(1.0)
    gi = model.vjp(theta, X, y, aux, cot)                                     # :406-412
    d = (gi - m).astype(F32)
    ssq = (ssq + (d * d).astype(F32)).astype(F32)                             # :414-420
  v = (F32(1.0 / (n - 1)) * ssq).astype(F32)                                  # :423
  b = (F32(F32(0.5) * F32(step_size)) * v).astype(F32)                        # :424
  fr = np.broadcast_to(np.asarray(friction, F32), (C, P))
  corr = (fr - b).astype(F32)                                                 # :427
  smallest = np.min(np.where(corr <= 0, np.inf, corr), axis=1, keepdims=True)  # :428
  pos = np.where(corr <= 0, smallest, corr).astype(F32)                       # :429
  b_corr = (fr - pos).astype(F32)                                             # :432
  with np.errstate(invalid="ignore"):
    return np.sqrt(pos).astype(F32), np.sqrt(b_corr).astype(F32)              # :434-435


# ----------------------------------------------------------------------------
# solver.sggmc / solver.amagold: MH correction   (solver.py:301-577)
# ----------------------------------------------------------------------------

def mh_decision(mode, U_state, U_new, e0, e1, temperature, keys, layout="original"):
  """Accept/reject arithmetic of solver.sggmc (:524-539) / solver.amagold
  (:381-395).  Returns (accept[C] bool, key'[C,2], log_alpha, ratio)."""
  if mode == "sggmc":
    s = (((U_new - U_state).astype(F32) + e1).astype(F32) - e0).astype(F32)
    la = (F32(F32(-1.0) / F32(temperature)) * s).astype(F32)
    la = np.where(la <= 0, la, F32(0.0)).astype(F32)
  else:
    la = ((U_state - U_new).astype(F32) + e1).astype(F32)
    la = np.where(la > 0, F32(0.0), la).astype(F32)
  ks = prng.split(keys, 2, layout)
  key, sub = ks[..., 0, :], ks[..., 1, :]
  u = prng.uniform(sub, (), layout=layout)
  accept = prng.log_libdevice(u) < la
  return accept, key, la, np.exp(la.astype(np.float64)).astype(F32)


class MHState(NamedTuple):
  integrator_state: tuple   # ObaboState (sggmc) or LeapfrogState (amagold)
  potential: np.ndarray     # f32[C] full potential of the current sample
  key: np.ndarray           # u32[C, 2]
  acceptance_ratio: np.ndarray


def sggmc_init(theta, full_potential_fn, keys=None, layout="original"):
  """solver.sggmc.init (:462-500): ``key, split = split(key)``; the integrator
  gets ``key``, the solver keeps ``split``."""
  theta = np.asarray(theta, F32)
  C = theta.shape[0]
  if keys is None:
    keys = np.tile(prng.PRNGKey(0), (C, 1))
  ks = prng.split(np.asarray(keys, np.uint32), 2, layout)
  return MHState(obabo_init(theta, ks[..., 0, :]), full_potential_fn(theta),
                 ks[..., 1, :], np.zeros(C, F32))


def sggmc_update(state: MHState, grad_fn_pairs, full_potential_fn, sizes, step_size,
                 temperature=1.0, friction=1.0, mass=None, layout="original",
                 mass_matrix=None):
  """solver.sggmc.update (:502-566)."""
  old = state.integrator_state
  prop = obabo_integrate(old, grad_fn_pairs, sizes, step_size, temperature, friction,
                         mass, layout, mass_matrix)
  U_new = full_potential_fn(prop.theta)
  accept, key, _, ratio = mh_decision("sggmc", state.potential, U_new,
                                      prop.kinetic_energy_start, prop.kinetic_energy_end,
                                      temperature, state.key, layout)
  m = accept[:, None]
  zeros = np.zeros_like(prop.kinetic_energy_start)
  new_int = ObaboState(np.where(m, prop.theta, old.theta),
                       np.where(m, prop.momentum, old.momentum), prop.key,
                       np.where(accept, prop.potential, old.potential), zeros, zeros)
  return MHState(new_int, np.where(accept, U_new, state.potential).astype(F32), key,
                 ratio), accept


def amagold_init(theta, full_potential_fn, keys=None, mass=None, sizes=None,
                 layout="original", mass_matrix=None):
  """solver.amagold.init (:323-362)."""
  theta = np.asarray(theta, F32)
  C = theta.shape[0]
  if keys is None:
    keys = np.tile(prng.PRNGKey(0), (C, 1))
  ks = prng.split(np.asarray(keys, np.uint32), 2, layout)
  return MHState(reversible_leapfrog_init(theta, ks[..., 0, :], mass, sizes, layout,
                                          mass_matrix),
                 full_potential_fn(theta), ks[..., 1, :], np.zeros(C, F32))


def amagold_update(state: MHState, grad_fns, full_potential_fn, sizes, step_size,
                   friction=0.25, mass=None, layout="original", mass_matrix=None):
  """solver.amagold.update (:364-424): on rejection the old state is kept with
  the momentum negated (direction = -1, :392, :399)."""
  old = state.integrator_state
  prop = reversible_leapfrog_integrate(old, grad_fns, sizes, step_size, friction, mass,
                                       layout, mass_matrix)
  U_new = full_potential_fn(prop.theta)
  accept, key, _, ratio = mh_decision("amagold", state.potential, U_new, None,
                                      prop.potential, 1.0, state.key, layout)
  m = accept[:, None]
  new_int = LeapfrogState(np.where(m, prop.theta, old.theta),
                          np.where(m, prop.momentum, (F32(-1.0) * old.momentum).astype(F32)),
                          prop.key, np.where(accept, prop.potential, old.potential))
  return MHState(new_int, np.where(accept, U_new, state.potential).astype(F32), key,
                 ratio), accept
