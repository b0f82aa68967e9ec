"""Recognised GLM likelihoods and priors.

jax-sgmc takes arbitrary Python ``likelihood(sample, observation)`` /
``prior(sample)`` callables and differentiates them with ``jax.grad``.  The
B200 path replaces that with fused kernels for the *recognised* families named
by the north star ("canonical GLM likelihoods"); a user states the family with
one of the objects below instead of writing the formula -- or hands over the
reference-style callables (``likelihood(sample, observation)``, ``prior(sample)``) and
``from_callable`` recognises the family by evaluating them on a few host points.  The
objects are specifications, not host implementations: calling them raises (there is no
CPU fallback); a callable that is none of the recognised closed forms is rejected with
a pointer to the JAX route (``jax.value_and_grad`` feeding the fused update kernels
through the C ABI).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

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


# ---------------------------------------------------------------------------------
# Recognising reference-style callables
# ---------------------------------------------------------------------------------

def _probe_points(sample_template: dict, observation_template: dict, rng, count=6):
  for _ in range(count):
    smp = {k: rng.standard_normal(np.shape(v)) * 0.7 for k, v in sample_template.items()}
    obs = {k: rng.standard_normal(np.shape(v)) for k, v in observation_template.items()}
    yield smp, obs


def _close(a, b) -> bool:
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  return a.shape == b.shape and bool(np.all(np.isfinite(a))) and \
      bool(np.allclose(a, b, rtol=1e-6, atol=1e-9))


def from_callable(likelihood, prior, sample_template: dict, observation_template: dict):
  """Recognise reference-style callables (potential.py:94-127: ``likelihood(sample,
  observation) -> log-likelihood`` of ONE observation, ``prior(sample) -> log-prior``)
  as one of the closed forms the fused kernels evaluate; returns ``(likelihood_spec,
  prior_spec)``.  The callables are only *evaluated* (NumPy inputs, float64) on a few
  random points and compared with every candidate; they are never differentiated and
  never run on the sampling path.  ``sample_template`` / ``observation_template``: flat
  dicts ``{leaf name: array}`` giving the shapes (one chain's sample, one observation).
  Raises ``TypeError`` when no candidate reproduces the callable."""
  if not (isinstance(sample_template, dict) and isinstance(observation_template, dict)):
    raise TypeError("from_callable needs flat dict samples and observations")
  rng = np.random.default_rng(1234)
  vec_s = [k for k, v in sample_template.items() if np.size(v) > 1 or np.ndim(v) >= 1]
  sca_s = [k for k, v in sample_template.items() if np.size(v) == 1]
  vec_o = [k for k, v in observation_template.items() if np.size(v) > 1 or np.ndim(v) >= 1]
  sca_o = [k for k, v in observation_template.items() if np.size(v) == 1]

  def z_of(smp, obs, w, x, b):
    z = float(np.dot(np.ravel(obs[x]), np.ravel(smp[w])))
    return z + (float(np.ravel(smp[b])[0]) if b else 0.0)

  cands = []
  for w in vec_s + [k for k in sca_s if k not in vec_s]:
    for x in vec_o + [k for k in sca_o if k not in vec_o]:
      if np.size(sample_template[w]) != np.size(observation_template[x]):
        continue
      for y in sca_o:
        if y == x:
          continue
        for aux in [None] + [k for k in sca_s if k != w]:
          cands.append(("logistic", w, x, y, aux))
          if aux is not None:
            cands.append(("gaussian", w, x, y, aux))
  lik_spec = None
  pts = list(_probe_points(sample_template, observation_template, rng))
  for family, w, x, y, aux in cands:
    ok = True
    for smp, obs in pts:
      if family == "logistic":
        obs = dict(obs)
        obs[y] = np.asarray(float(rng.random() < 0.5)).reshape(np.shape(obs[y]))
        z = z_of(smp, obs, w, x, aux)
        yv = float(np.ravel(obs[y])[0])
        want = -yv * np.logaddexp(0.0, -z) - (1.0 - yv) * np.logaddexp(0.0, z)
      else:
        z = z_of(smp, obs, w, x, None)
        ls = float(np.ravel(smp[aux])[0])
        r = float(np.ravel(obs[y])[0]) - z
        want = -0.5 * (r / np.exp(ls)) ** 2 - ls - 0.5 * np.log(2.0 * np.pi)
      try:
        got = np.squeeze(likelihood(smp, obs))
      except Exception:
        ok = False
        break
      if not _close(got, want):
        ok = False
        break
    if ok:
      used = {w} | ({aux} if aux else set())
      if set(sample_template) - used:
        continue                  # the sample has leaves this family would ignore
      lik_spec = LogisticRegression(x, y, w, aux) if family == "logistic" else \
          GaussianRegression(x, y, w, aux)
      break
  if lik_spec is None:
    raise TypeError(
        "the likelihood callable is none of the recognised closed forms (logistic "
        "regression with optional bias, gaussian linear regression with a log_sigma leaf); "
        "pass a jax_sgmc_b200.glm / nn specification, or use the JAX route (INTEGRATION.md)")

  # ---- prior: flat, gaussian on all / some leaves, 1 / exp(log_sigma) ----------------
  zero = {k: np.zeros(np.shape(v)) for k, v in sample_template.items()}
  smps = [smp for smp, _ in pts]
  try:
    p0 = float(np.squeeze(prior(zero)))
    vals = [float(np.squeeze(prior(s_))) for s_ in smps]
  except Exception as e:
    raise TypeError(f"the prior callable cannot be evaluated on NumPy samples: {e}") from e
  if all(_close(v, p0) for v in vals):
    return lik_spec, FlatPrior()
  names = list(sample_template)
  subsets = [names] + [[k] for k in names] + [[k for k in names if k != q] for q in names]
  for leaves in subsets:
    ss = [sum(float(np.sum(np.square(s_[k]))) for k in leaves) for s_ in smps]
    if ss[0] <= 0:
      continue
    inv_var = -2.0 * (vals[0] - p0) / ss[0]
    if inv_var > 0 and all(_close(v - p0, -0.5 * inv_var * q) for v, q in zip(vals, ss)):
      scale = float(1.0 / np.sqrt(inv_var))
      return lik_spec, GaussianPrior(scale, None if leaves == names else leaves)
  for k in sca_s:
    if all(_close(v, 1.0 / np.exp(float(np.ravel(s_[k])[0]))) for v, s_ in zip(vals, smps)):
      return lik_spec, InvSigmaPrior(k)
  raise TypeError(
      "the prior callable is none of the recognised closed forms (flat, isotropic gaussian "
      "on all or some leaves, 1 / exp(log_sigma)); pass a jax_sgmc_b200.glm prior")
