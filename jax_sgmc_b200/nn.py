"""Recognised neural-network likelihoods (BASELINE.json configs[2] and configs[4]).

Like :mod:`jax_sgmc_b200.glm`, the objects here are *specifications*: the
likelihood a jax-sgmc user writes as ``-softmax_cross_entropy(apply(sample, x),
label)`` (reference examples/cifar.md:196-204) and lets ``jax.value_and_grad``
differentiate is evaluated, together with its hand-derived reverse pass, by the
chain-batched kernels of ``csrc/mlp.cu``.  ``potential.minibatch_potential``
accepts them in place of the Python callable.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import glm, ops


class MLPClassifier(glm._Spec):
  """Dense classifier: ``h_l = tanh(h_{l-1} W_l + b_l)``, ``logits = h_{L-1} W_L + b_L``,
  ``ell = log_softmax(logits)[label]``.

  ``layers``: sequence of ``(weight_path, bias_path)`` leaf paths into the sample
  pytree, first layer first; ``None`` = every dict of the sample that holds a 2-D
  ``"w"`` and a 1-D ``"b"`` leaf, in ``tree_flatten`` order (the layout of
  :func:`init_params` and of Haiku's ``hk.nets.MLP`` parameters).
  ``x`` / ``y``: names of the feature and the class-index leaves of the data."""
  family = "mlp_classifier"

  def __init__(self, x="x", y="y", layers: Optional[Sequence] = None, activation="tanh"):
    if activation != "tanh":
      raise NotImplementedError("activation: 'tanh'")
    self.x, self.y, self.layers, self.activation = x, y, layers, activation


class CNNClassifier(glm._Spec):
  """Convolutional classifier (BASELINE.json configs[4]): ``strides[l]``-strided 3x3
  convolutions with zero padding 1 and tanh, then a dense layer on the flattened NHWC
  feature map, ``ell = log_softmax(logits)[label]``.

  The sample holds one ``{"w": [3, 3, cin, cout], "b": [cout]}`` dict per convolution and
  one ``{"w": [features, classes], "b": [classes]}`` dict for the head, found in
  ``tree_flatten`` order (the layout of :func:`init_cnn_params`).  Observations ``x`` are
  ``[height, width, channels]`` images, ``y`` the class index."""
  family = "cnn_classifier"

  def __init__(self, x="x", y="y", strides: Sequence[int] = (2, 2)):
    self.x, self.y, self.strides = x, y, tuple(int(s) for s in strides)


def _find_layers(treedef, prefix=()):
  """Paths of the dicts holding {"w", "b"} leaves, in tree_flatten order."""
  out = []
  if treedef[0] == "dict":
    keys, subs = treedef[1], treedef[2]
    if "w" in keys and "b" in keys and all(
        subs[keys.index(k)][0] == "leaf" for k in ("w", "b")):
      out.append(prefix)
    else:
      for k, s in zip(keys, subs):
        out.extend(_find_layers(s, prefix + (k,)))
  elif treedef[0] in ("tuple", "list"):
    for i, s in enumerate(treedef[1]):
      out.extend(_find_layers(s, prefix + (i,)))
  return out


_SPEC_CACHE = {}


def resolve(likelihood: MLPClassifier, prior, sample, temperature: float):
  """``sgmc_mlp_spec`` for a ChainTree layout (cached per layout)."""
  key = (id(likelihood), id(prior), id(sample.treedef), tuple(sample.sizes), float(temperature))
  hit = _SPEC_CACHE.get(key)
  if hit is not None and hit[0] is likelihood and hit[1] is prior and hit[3] is sample.treedef:
    return hit[2]
  layers = likelihood.layers
  if layers is None:
    layers = [(p + ("w",), p + ("b",)) for p in _find_layers(sample.treedef)]
  if not layers:
    raise ValueError("the sample holds no {'w', 'b'} layer dicts")
  offs, shapes = sample.offsets(), sample.shapes
  sizes, w_off, b_off, used = [], [], [], 0
  for wp, bp in layers:
    wi, bi = sample.leaf_index(wp), sample.leaf_index(bp)
    ws, bs = shapes[wi], shapes[bi]
    if len(ws) != 2 or len(bs) != 1 or bs[0] != ws[1]:
      raise ValueError(f"layer {wp}: expected w [in, out] and b [out], got {ws} and {bs}")
    if sizes and sizes[-1] != ws[0]:
      raise ValueError(f"layer {wp}: input width {ws[0]} does not follow {sizes[-1]}")
    if not sizes:
      sizes.append(int(ws[0]))
    sizes.append(int(ws[1]))
    w_off.append(offs[wi])
    b_off.append(offs[bi])
    used += ws[0] * ws[1] + bs[0]
  if used != sample.n_params:
    raise ValueError("the sample has leaves the MLP does not use")
  kind, p_off, p_size, p_scale = "flat", 0, 0, 1.0
  if isinstance(prior, glm.GaussianPrior):
    kind, p_scale = "gaussian", prior.scale
    if prior.leaves is None:
      p_off, p_size = 0, sample.n_params
    else:
      idxs = sorted(sample.leaf_index(l) for l in prior.leaves)
      p_off = offs[idxs[0]]
      p_size = sum(sample.sizes[i] for i in idxs)
      assert offs[idxs[-1]] + sample.sizes[idxs[-1]] - p_off == p_size, \
          "prior leaves must be contiguous in the flat sample"
  elif not isinstance(prior, glm.FlatPrior):
    raise TypeError("the MLP potential takes a FlatPrior or a GaussianPrior")
  spec = ops.mlp_spec(sizes, w_off, b_off, kind, p_off, p_size, p_scale, temperature)
  if len(_SPEC_CACHE) > 64:
    _SPEC_CACHE.clear()
  _SPEC_CACHE[key] = (likelihood, prior, spec, sample.treedef)
  return spec


def init_params(key, sizes: Sequence[int], scale: str = "he_normal"):
  """Host pytree ``{"layer_0": {"b": [out], "w": [in, out]}, ...}``: weights
  ``normal(split(key, L)[l]) * sqrt(2 / in)`` from the package's jax.random-compatible
  generator, zero biases ("He-normal init from PRNGKey(c)")."""
  assert scale == "he_normal"
  key = np.asarray(key, np.uint32).reshape(1, 2)
  L = len(sizes) - 1
  from .device import DeviceArray
  ks = ops.split(DeviceArray.from_numpy(key), L).numpy().reshape(L, 2)
  out = {}
  for l in range(L):
    i, o = int(sizes[l]), int(sizes[l + 1])
    z = ops.normal(DeviceArray.from_numpy(ks[l:l + 1]), i * o).numpy().reshape(i, o)
    out[f"layer_{l}"] = {"w": (z * np.float32(np.sqrt(2.0 / i))).astype(np.float32),
                         "b": np.zeros(o, np.float32)}
  return out


def resolve_cnn(likelihood: CNNClassifier, prior, sample, temperature: float, image_shape):
  """``sgmc_cnn_spec`` for a ChainTree layout and an observation shape (cached per layout)."""
  key = (id(likelihood), id(prior), id(sample.treedef), tuple(sample.sizes), float(temperature),
         tuple(image_shape))
  hit = _SPEC_CACHE.get(key)
  if hit is not None and hit[0] is likelihood and hit[1] is prior and hit[3] is sample.treedef:
    return hit[2]
  if len(image_shape) != 3:
    raise ValueError(f"observations must be [height, width, channels] images, got {image_shape}")
  paths = _find_layers(sample.treedef)
  n_conv = len(likelihood.strides)
  if len(paths) != n_conv + 1:
    raise ValueError(f"expected {n_conv} convolution dicts and one head dict, found {len(paths)}")
  offs, shapes = sample.offsets(), sample.shapes
  H, W, cin = (int(v) for v in image_shape)
  channels, w_off, b_off, used = [cin], [], [], 0
  for l, p in enumerate(paths[:n_conv]):
    wi, bi = sample.leaf_index(p + ("w",)), sample.leaf_index(p + ("b",))
    ws, bs = shapes[wi], shapes[bi]
    if len(ws) != 4 or ws[0] != 3 or ws[1] != 3 or ws[2] != channels[-1] or tuple(bs) != (ws[3],):
      raise ValueError(f"convolution {p}: expected w [3, 3, {channels[-1]}, cout] and b [cout], "
                       f"got {ws} and {bs}")
    channels.append(int(ws[3]))
    w_off.append(offs[wi]); b_off.append(offs[bi])
    used += int(np.prod(ws)) + bs[0]
    st = likelihood.strides[l]
    H, W = (H - 1) // st + 1, (W - 1) // st + 1
  p = paths[n_conv]
  wi, bi = sample.leaf_index(p + ("w",)), sample.leaf_index(p + ("b",))
  ws, bs = shapes[wi], shapes[bi]
  if len(ws) != 2 or ws[0] != H * W * channels[-1] or tuple(bs) != (ws[1],):
    raise ValueError(f"head {p}: expected w [{H * W * channels[-1]}, classes] and b [classes], "
                     f"got {ws} and {bs}")
  w_off.append(offs[wi]); b_off.append(offs[bi])
  used += ws[0] * ws[1] + bs[0]
  if used != sample.n_params:
    raise ValueError("the sample has leaves the CNN does not use")
  kind, p_off, p_size, p_scale = "flat", 0, 0, 1.0
  if isinstance(prior, glm.GaussianPrior):
    if prior.leaves is not None:
      raise ValueError("the CNN potential takes a GaussianPrior on the whole sample")
    kind, p_scale, p_off, p_size = "gaussian", prior.scale, 0, sample.n_params
  elif not isinstance(prior, glm.FlatPrior):
    raise TypeError("the CNN potential takes a FlatPrior or a GaussianPrior")
  spec = ops.cnn_spec(image_shape[0], image_shape[1], channels, likelihood.strides, int(ws[1]),
                      w_off, b_off, kind, p_off, p_size, p_scale, temperature)
  if len(_SPEC_CACHE) > 64:
    _SPEC_CACHE.clear()
  _SPEC_CACHE[key] = (likelihood, prior, spec, sample.treedef)
  return spec


def init_cnn_params(key, image_shape, conv_channels: Sequence[int], strides: Sequence[int],
                    n_classes: int):
  """Host pytree ``{"conv_0": {"b", "w": [3, 3, cin, cout]}, ..., "head": {"b", "w"}}`` with
  He-normal weights from the package's jax.random-compatible generator and zero biases."""
  from .device import DeviceArray
  key = np.asarray(key, np.uint32).reshape(1, 2)
  L = len(conv_channels) + 1
  ks = ops.split(DeviceArray.from_numpy(key), L).numpy().reshape(L, 2)
  H, W, cin = (int(v) for v in image_shape)
  out = {}
  for l, (cout, st) in enumerate(zip(conv_channels, strides)):
    fan_in = 9 * cin
    z = ops.normal(DeviceArray.from_numpy(ks[l:l + 1]), fan_in * cout).numpy()
    out[f"conv_{l}"] = {"w": (z.reshape(3, 3, cin, cout) * np.float32(np.sqrt(2.0 / fan_in))
                              ).astype(np.float32), "b": np.zeros(cout, np.float32)}
    cin, H, W = int(cout), (H - 1) // st + 1, (W - 1) // st + 1
  F = H * W * cin
  z = ops.normal(DeviceArray.from_numpy(ks[L - 1:L]), F * n_classes).numpy()
  out["head"] = {"w": (z.reshape(F, n_classes) * np.float32(np.sqrt(2.0 / F))).astype(np.float32),
                 "b": np.zeros(n_classes, np.float32)}
  return out
