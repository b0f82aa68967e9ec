"""Recognised GLM likelihoods and priors.

jax-sgmc takes arbitrary Python ``likelihood(sample, observation)`` /
``prior(sample)`` callables and differentiates them with ``jax.grad``.  The
B200 path replaces that with fused kernels for the *recognised* families named
by the north star ("canonical GLM likelihoods"); a user states the family with
one of the objects below instead of writing the formula.  They are
specifications, not host implementations: calling them raises (there is no CPU
fallback); any other callable is rejected by ``potential.minibatch_potential``
with a pointer to the JAX route (``jax.value_and_grad`` feeding the fused
update kernels through the C ABI).
"""
from __future__ import annotations

from typing import Optional, Sequence

from . import ops


class _Spec:
  def __call__(self, *args, **kwargs):
    raise NotImplementedError(
        f"{type(self).__name__} is a specification evaluated by the fused CUDA "
        "kernels; it has no host implementation")


class GaussianRegression(_Spec):
  """``norm.logpdf(y - x.w, scale=exp(log_sigma))`` -- the quickstart model
  (reference examples/quickstart.md:158-169)."""
  family = "gaussian"

  def __init__(self, x="x", y="y", weights="w", log_sigma="log_sigma"):
    self.x, self.y, self.weights, self.aux = x, y, weights, log_sigma


class LogisticRegression(_Spec):
  """``y log sigmoid(z) + (1-y) log(1-sigmoid(z))``, ``z = x.w (+ bias)``
  (BASELINE.json configs[1])."""
  family = "logistic"

  def __init__(self, x="x", y="y", weights="w", bias: Optional[str] = None):
    self.x, self.y, self.weights, self.aux = x, y, weights, bias


class FlatPrior(_Spec):
  kind = "flat"


class GaussianPrior(_Spec):
  """``sum -0.5 (theta/scale)^2`` over the given leaves (default: all)."""
  kind = "gaussian"

  def __init__(self, scale: float = 1.0, leaves: Optional[Sequence[str]] = None):
    self.scale, self.leaves = float(scale), leaves


class InvSigmaPrior(_Spec):
  """``1 / exp(log_sigma)`` -- the quickstart's log-prior
  (reference examples/quickstart.md:172-173)."""
  kind = "inv_sigma"

  def __init__(self, log_sigma="log_sigma"):
    self.leaf = log_sigma


_SPEC_CACHE = {}


def resolve(likelihood, prior, sample, temperature: float, x_absmax: float = 0.0):
  """Build the C-ABI ``sgmc_glm_spec`` for a ChainTree layout (cached per layout:
  this runs once per sampling step)."""
  key = (id(likelihood), id(prior), id(sample.treedef), tuple(sample.sizes),
         float(temperature), float(x_absmax))
  hit = _SPEC_CACHE.get(key)
  if hit is not None and hit[0] is likelihood and hit[1] is prior and hit[3] is sample.treedef:
    return hit[2]
  spec = _resolve(likelihood, prior, sample, temperature, x_absmax)
  if len(_SPEC_CACHE) > 256:
    _SPEC_CACHE.clear()
  _SPEC_CACHE[key] = (likelihood, prior, spec, sample.treedef)
  return spec


def _resolve(likelihood, prior, sample, temperature: float, x_absmax: float = 0.0):
  offs, sizes = sample.offsets(), sample.sizes
  wl = sample.leaf_index(likelihood.weights)
  d, w_off = sizes[wl], offs[wl]
  aux_off = -1
  if likelihood.aux is not None:
    al = sample.leaf_index(likelihood.aux)
    assert sizes[al] == 1, "auxiliary parameter must be a scalar leaf"
    aux_off = offs[al]
  elif likelihood.family == "gaussian":
    raise ValueError("GaussianRegression needs a log_sigma leaf")
  P = sample.n_params
  if P != d + (1 if aux_off >= 0 else 0):
    raise ValueError("the sample has leaves the GLM family does not use")
  kind, p_off, p_size, p_scale = "flat", 0, 0, 1.0
  if isinstance(prior, GaussianPrior):
    kind, p_scale = "gaussian", prior.scale
    if prior.leaves is None:
      p_off, p_size = 0, P
    else:
      idxs = sorted(sample.leaf_index(l) for l in prior.leaves)
      p_off = offs[idxs[0]]
      p_size = sum(sizes[i] for i in idxs)
      assert offs[idxs[-1]] + sizes[idxs[-1]] - p_off == p_size, \
          "prior leaves must be contiguous in the flat sample"
  elif isinstance(prior, InvSigmaPrior):
    kind, p_off, p_size = "inv_sigma", offs[sample.leaf_index(prior.leaf)], 1
  elif not isinstance(prior, FlatPrior):
    raise TypeError("unrecognised prior")
  return ops.glm_spec(likelihood.family, d, w_off, aux_off, kind, p_off, p_size,
                      p_scale, temperature, x_absmax)
