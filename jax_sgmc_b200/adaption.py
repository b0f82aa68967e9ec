"""Adaption strategies, mirroring ``jax_sgmc.adaption`` for the hot path.

``rms_prop()`` (reference adaption.py:225-293) returns the usual
``(init, update, get)`` triplet wrapped the way the ``@adaption`` decorator
does (adaption.py:107-222): positional arguments are raveled pytrees and
``get`` returns ``Manifold(g_inv, sqrt_g_inv, gamma)`` of ``Tensor(ndim=1)``.
``integrator.langevin_diffusion`` recognises the strategy and runs the fused
pSGLD kernel (update + get + step in one pass); the triplet itself is backed
by the stand-alone kernels ``sgmc_rms_prop_update`` / ``sgmc_rms_prop_get``.

``mass_matrix`` and ``fisher_information`` are outside the scope of this path
(burn-in only, dense eigh / SVD) and raise.
"""
from __future__ import annotations

from typing import Any, NamedTuple

import numpy as np

from . import ops
from .device import DeviceArray
from .tree_util import ChainTree, Tensor


class Manifold(NamedTuple):
  """adaption.py:53-64."""
  g_inv: Any
  sqrt_g_inv: Any
  gamma: Any


class MassMatrix(NamedTuple):
  """adaption.py:66-76."""
  inv: Any
  sqrt: Any


class RmsPropState(NamedTuple):
  """(v, alpha, lmbd) of adaption.py:238-252, v as ``f32[C, P]`` on device."""
  v: ChainTree
  alpha: float
  lmbd: float


class _RmsProp(tuple):
  """The (init, update, get) triplet; the marker attribute lets the Langevin
  integrator pick the fused kernel."""
  fused_kind = "rms_prop"


def rms_prop():
  """adaption.py:225-293."""

  def init(sample: ChainTree, alpha: float = 0.9, lmbd: float = 1e-5) -> RmsPropState:
    ones = DeviceArray.full(sample.flat.shape, 1.0)           # v = ones_like (:251)
    return RmsPropState(ChainTree.like(sample, ones), float(alpha), float(lmbd))

  def update(state: RmsPropState, sample: ChainTree, sample_grad: ChainTree,
             *args, **kwargs) -> RmsPropState:
    del sample, args, kwargs
    ops.rms_prop_update(state.v.flat, sample_grad.flat, state.alpha)   # in place
    return state

  def get(state: RmsPropState, sample: ChainTree = None, sample_grad: ChainTree = None,
          *args, **kwargs) -> Manifold:
    del sample, sample_grad, args, kwargs
    g = DeviceArray(state.v.flat.shape, np.float32)
    s = DeviceArray(state.v.flat.shape, np.float32)
    ops.rms_prop_get(state.v.flat, g, s, state.lmbd)
    zeros = DeviceArray.zeros(state.v.flat.shape)
    return Manifold(Tensor(1, ChainTree.like(state.v, g)),
                    Tensor(1, ChainTree.like(state.v, s)),
                    Tensor(1, ChainTree.like(state.v, zeros)))

  return _RmsProp((init, update, get))


def mass_matrix(*args, **kwargs):
  raise NotImplementedError("adaption.mass_matrix is outside the accelerated "
                            "sampling path (SURVEY.md section 8)")


def fisher_information(*args, **kwargs):
  raise NotImplementedError("adaption.fisher_information is outside the "
                            "accelerated sampling path (SURVEY.md section 8)")
