"""Adaption strategies, mirroring ``jax_sgmc.adaption`` for the hot path.

``rms_prop()`` (reference adaption.py:225-293) returns the usual
``(init, update, get)`` triplet wrapped the way the ``@adaption`` decorator
does (adaption.py:107-222): positional arguments are raveled pytrees and
``get`` returns ``Manifold(g_inv, sqrt_g_inv, gamma)`` of ``Tensor(ndim=1)``.
``integrator.langevin_diffusion`` recognises the strategy and runs the fused
pSGLD kernel (update + get + step in one pass); the triplet itself is backed
by the stand-alone kernels ``sgmc_rms_prop_update`` / ``sgmc_rms_prop_get``.

``mass_matrix(diagonal=True)`` (adaption.py:296-369) keeps the Welford statistics of
all chains on the device (``sgmc_mass_matrix_update``) and hands the integrators a
per-chain ``MassMatrix(inv, sqrt)`` that the OBABO / reversible-leapfrog kernels read
directly (``sgmc_*_adapted``).  The dense variants (``diagonal=False``: eigh / SVD of a
P x P matrix per chain) stay outside this path and raise.

``fisher_information(minibatch_potential, diagonal=True)`` (adaption.py:372-457): the
empirical Fisher noise model of SGHMC for a recognised GLM potential, evaluated by
``sgmc_glm_fisher_diag`` on the minibatch of the leapfrog step; ``get`` returns
``NoiseModel(cb_diff_sqrt, b_sqrt)`` of per-chain ``Tensor(ndim=1)``.
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


class MassState:
  """(iteration, mean, ssq, m_inv, m_sqrt) of adaption.py:319-341 for C chains; the
  arrays are ``f32[C, P]`` on the device and updated in place.  ``matrix`` is the
  ``MassMatrix`` handed to the integrators (the same object on every ``get``)."""

  def __init__(self, sample: ChainTree, init_cov, burn_in: int):
    shape = sample.flat.shape
    self.iteration, self.burn_in = 0, int(burn_in)
    self.mean, self.ssq = DeviceArray.zeros(shape), DeviceArray.zeros(shape)
    if init_cov is None:                                   # ones_like(sample), :330-331
      cov = np.ones(shape, np.float32)
    elif isinstance(init_cov, ChainTree):
      cov = init_cov.flat.numpy()
    else:
      from .tree_util import tree_flatten
      leaves, _ = tree_flatten(init_cov)
      flat = np.concatenate([np.asarray(l, np.float32).ravel() for l in leaves])
      assert flat.size == shape[1], "init_cov does not match the sample"
      cov = np.broadcast_to(flat[None, :], shape)
    cov = np.ascontiguousarray(cov, np.float32)
    self.m_inv = DeviceArray.from_numpy(cov)               # m_inv = init_cov, :334
    self.m_sqrt = DeviceArray.from_numpy(                  # m_sqrt = 1 / sqrt(init_cov), :335
        (np.float32(1.0) / np.sqrt(cov)).astype(np.float32))
    self.matrix = MassMatrix(Tensor(1, ChainTree.like(sample, self.m_inv)),
                             Tensor(1, ChainTree.like(sample, self.m_sqrt)))


def mass_matrix(diagonal: bool = True, burn_in: int = 1000):
  """adaption.py:296-369: running mean / sum of squares of the accepted samples; in
  iteration ``burn_in`` the matrix is set once (``M^-1 = ssq / n``, ``M^1/2 = sqrt(n / ssq)``).
  Every chain adapts its own matrix."""
  if not diagonal:
    raise NotImplementedError("the dense mass matrix (eigh of a P x P matrix per chain) is "
                              "outside the accelerated sampling path (SURVEY.md section 8)")

  def init(sample, init_cov=None) -> MassState:
    return MassState(sample, init_cov, burn_in)

  def update(state: MassState, sample: ChainTree, *args, **kwargs) -> MassState:
    del args, kwargs
    state.iteration += 1                                                   # :350
    ops.mass_matrix_update(state.mean, state.ssq, state.m_inv, state.m_sqrt, sample.flat,
                           state.iteration, state.burn_in)                 # :351-361
    return state

  def get(state: MassState) -> MassMatrix:
    return state.matrix                                                    # :365-367

  return init, update, get


class NoiseModel(NamedTuple):
  """adaption.py:78-88."""
  cb_diff_sqrt: Any
  b_sqrt: Any


def fisher_information(minibatch_potential=None, diagonal: bool = True):
  """adaption.py:372-457: ``init`` / ``update`` do nothing; ``get(state, sample,
  sample_grad, friction, mini_batch=..., step_size=..., model_state=...)`` estimates the
  gradient noise from the per-observation likelihood gradients of the minibatch and returns
  the corrected noise scales.  ``friction``: a scalar or a per-parameter ``f32[P]`` device
  vector (what ``friction_leapfrog`` holds)."""
  assert minibatch_potential, "Fisher information requires potential function."
  if not diagonal:
    raise NotImplementedError("the dense Fisher noise model (an SVD of a P x P matrix per "
                              "chain) is outside the accelerated sampling path")
  from . import glm
  buffers = {}

  def init(*args):
    del args

  def update(*args, **kwargs):
    del args, kwargs

  def get(state, sample: ChainTree, sample_grad: ChainTree, friction, *args, mini_batch,
          flat_potential=None, step_size=1.0, model_state=None, **kwargs) -> NoiseModel:
    del state, args, flat_potential, model_state, kwargs
    pot = minibatch_potential
    if hasattr(pot, "_resolve"):                      # reference-style callables, probed lazily
      pot = pot._resolve(sample, mini_batch[0].loader)
    if not hasattr(pot, "likelihood") or not isinstance(
        pot.likelihood, (glm.LogisticRegression, glm.GaussianRegression)):
      raise NotImplementedError("the Fisher noise model needs a GLM potential")
    batch, info = mini_batch
    if batch.per_chain or batch.mask is not None:
      raise NotImplementedError("the Fisher noise model needs one unmasked minibatch shared "
                                "by the chains")
    spec = glm.resolve(pot.likelihood, pot.prior, sample, pot.temperature,
                       batch.loader.absmax(pot.likelihood.x))
    key = sample.flat.shape
    buf = buffers.get(key)
    if buf is None:
      buf = {"ns": DeviceArray(key, np.float32), "sc": DeviceArray(key, np.float32),
             "scratch": None}
      buffers[key] = buf
    vec = friction if isinstance(friction, DeviceArray) else None
    buf["scratch"] = ops.glm_fisher_diag(
        spec, sample.flat, batch.leaf(pot.likelihood.x), batch.leaf(pot.likelihood.y),
        batch.idx, batch.n, int(info.observation_count), sample_grad.flat, vec,
        0.0 if vec is not None else float(friction), float(step_size), buf["ns"], buf["sc"],
        buf["scratch"])
    return NoiseModel(Tensor(1, ChainTree.like(sample, buf["ns"])),
                      Tensor(1, ChainTree.like(sample, buf["sc"])))

  return init, update, get
