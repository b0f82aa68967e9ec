"""Recognised neural-network likelihoods (BASELINE.json configs[2]).

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
