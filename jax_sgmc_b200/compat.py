"""Array namespace for user-written likelihood / prior callables.

jax-sgmc models are plain Python functions written against ``jax.numpy`` and
``jax.scipy.stats`` (reference examples/quickstart.md:158-173).  JAX is not part of
this stack: ``potential.minibatch_potential`` never differentiates such a function, it
only *evaluates* it on a few host points to recognise which closed form it is
(``glm.from_callable``).  For that evaluation the function needs an array namespace:

    from jax_sgmc_b200.compat import jnp, norm          # explicit

or, for model code that must stay untouched (``import jax.numpy as jnp`` /
``from jax.scipy.stats import norm``), ``install_jax_shim()`` registers the NumPy-backed
stand-ins below as ``jax.numpy`` / ``jax.scipy.stats`` / ``jax.nn`` -- only when no real
``jax`` is importable.  Nothing here runs on the sampling path.
"""
from __future__ import annotations

import sys
import types

import numpy as jnp  # noqa: F401  (the NumPy namespace covers what closed-form models use)
import numpy as np


class norm:      # jax.scipy.stats.norm
  @staticmethod
  def logpdf(x, loc=0.0, scale=1.0):
    z = (np.asarray(x) - loc) / scale
    return -0.5 * z * z - np.log(scale) - 0.5 * np.log(2.0 * np.pi)

  @staticmethod
  def pdf(x, loc=0.0, scale=1.0):
    return np.exp(norm.logpdf(x, loc, scale))


def sigmoid(x):
  return 1.0 / (1.0 + np.exp(-np.asarray(x)))


def log_sigmoid(x):
  x = np.asarray(x)
  return -np.logaddexp(0.0, -x)


def softplus(x):
  return np.logaddexp(0.0, np.asarray(x))


def install_jax_shim() -> bool:
  """Register NumPy-backed ``jax`` / ``jax.numpy`` / ``jax.scipy.stats`` / ``jax.nn``
  modules so that unmodified reference model code imports.  Does nothing (returns
  False) when a real ``jax`` can be imported."""
  if "jax" in sys.modules and not getattr(sys.modules["jax"], "_sgmc_b200_shim", False):
    return False
  try:
    import importlib.util
    if "jax" not in sys.modules and importlib.util.find_spec("jax") is not None:
      return False
  except (ImportError, ValueError):
    pass
  jax = types.ModuleType("jax")
  jax._sgmc_b200_shim = True
  jax.numpy = np
  scipy = types.ModuleType("jax.scipy")
  stats = types.ModuleType("jax.scipy.stats")
  stats.norm = norm
  special = types.ModuleType("jax.scipy.special")
  special.expit = sigmoid
  special.logsumexp = lambda a, axis=None: np.logaddexp.reduce(np.asarray(a), axis=axis)
  scipy.stats, scipy.special = stats, special
  nn = types.ModuleType("jax.nn")
  nn.sigmoid, nn.log_sigmoid, nn.softplus = sigmoid, log_sigmoid, softplus
  jax.scipy, jax.nn = scipy, nn
  sys.modules.update({"jax": jax, "jax.numpy": np, "jax.scipy": scipy, "jax.scipy.stats": stats,
                      "jax.scipy.special": special, "jax.nn": nn})
  return True
