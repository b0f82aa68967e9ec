"""Integrators of the sampling hot path, mirroring ``jax_sgmc.integrator``.

``langevin_diffusion`` (reference integrator.py:767-924), ``friction_leapfrog``
(:563-765) and ``obabo`` (:138-346) keep their signatures and return the same
``(init_fn, update_fn | integrate, get_fn)`` triplets over the same state
NamedTuples.  States hold chain-batched device buffers (``ChainTree``): all
chains advance in one kernel launch (the reference's ``list_vmap`` batching,
util/list_map.py:93-127) and buffers are updated in place -- the returned state
aliases the donated input, as ``jit(..., donate_argnums)`` would.

Each step is: minibatch draw -> fused GLM potential + gradient -> ONE fused
elementwise pass that also draws the Gaussian noise in-kernel (threefry2x32
reproducing ``jax.random`` bit for bit).
"""
from __future__ import annotations

import weakref
from typing import Any, Callable, Dict, NamedTuple, Tuple

import numpy as np

from . import data as _data
from . import ops
from . import potential as _potential
from .adaption import MassMatrix
from .device import DeviceArray
from .io import Negated
from .tree_util import ChainTree, Tensor, tree_flatten

PyTree = Any


class KeyState:
  """``uint32[C, 2]`` PRNG keys with the ping-pong partner the kernels need
  (``keys_out`` must not alias ``keys_in``)."""

  def __init__(self, keys: np.ndarray):
    keys = np.asarray(keys, np.uint32).reshape(-1, 2)
    self._bufs = [DeviceArray.from_numpy(keys), DeviceArray(keys.shape, np.uint32)]
    self._cur = 0

  @property
  def current(self) -> DeviceArray:
    return self._bufs[self._cur]

  @property
  def next(self) -> DeviceArray:
    return self._bufs[1 - self._cur]

  def flip(self):
    self._cur = 1 - self._cur

  def numpy(self) -> np.ndarray:
    return self.current.numpy()


class LeapfrogState(NamedTuple):
  """integrator.py:35-52."""
  positions: ChainTree
  momentum: ChainTree
  potential: DeviceArray
  model_state: PyTree
  data_state: Any
  key: KeyState
  extra_fields: PyTree = None


class ObaboState(NamedTuple):
  """integrator.py:55-75."""
  positions: ChainTree
  momentum: ChainTree
  potential: DeviceArray
  model_state: PyTree
  data_state: Any
  key: KeyState
  kinetic_energy_start: DeviceArray
  kinetic_energy_end: DeviceArray


class LangevinState(NamedTuple):
  """integrator.py:78-96."""
  latent_variables: ChainTree
  model_state: PyTree
  key: KeyState
  adapt_state: Any
  data_state: Any
  potential: DeviceArray
  variance: DeviceArray


def _as_chain_tree(sample) -> ChainTree:
  """Accept a ChainTree, one host pytree (a single chain) or a list of them."""
  if isinstance(sample, ChainTree):
    return sample
  if isinstance(sample, (list, tuple)) and not hasattr(sample, "_fields") and \
      len(sample) > 0 and isinstance(sample[0], dict):
    return ChainTree.from_trees(list(sample))
  return ChainTree.from_trees([sample])


def _keys_for(key, n_chains: int) -> KeyState:
  """Default ``PRNGKey(0)`` for every chain (integrator.py:804, :300-301,
  :699-700); a single key is shared by all chains, ``[C, 2]`` keys are per
  chain."""
  if isinstance(key, KeyState):
    return key
  if key is None:
    key = ops.prng_key(0)
  key = np.asarray(key, np.uint32)
  if key.ndim == 1:
    key = np.tile(key, (n_chains, 1))
  assert key.shape == (n_chains, 2), key.shape
  return KeyState(key)


def _init_data_state(batch_init, batch_kwargs, n_chains: int):
  """One random-data state per chain, merged (the reference initialises every
  chain separately, alias.py:110-119, so a host loader hands out one stream per
  chain, seeded by the chain id, numpy_loader.py:263)."""
  states = [batch_init(**batch_kwargs) for _ in range(n_chains)]
  return states[0] if n_chains == 1 else _data.merge_cache_states(states)


def _flat_vector(tree, like: ChainTree) -> DeviceArray:
  """A per-parameter pytree (mass, friction) raveled to ``f32[P]``."""
  leaves, _ = tree_flatten(tree)
  flat = np.concatenate([np.asarray(l, np.float32).ravel() for l in leaves])
  assert flat.size == like.n_params, "pytree does not match the sample"
  return DeviceArray.from_numpy(flat)


def _mass_operand(m, like: ChainTree):
  """What the kernels take for a mass: None (unit mass), ``f32[P]`` (a constant diagonal
  mass pytree, the kernels derive M^-1 and M^1/2), or ``(inv, sqrt)`` of ``f32[C, P]`` for
  an adapted ``MassMatrix`` (adaption.mass_matrix, one matrix per chain)."""
  if m is None:
    return None
  if isinstance(m, MassMatrix) and isinstance(getattr(m.inv, "tensor", None), ChainTree):
    return (m.inv.tensor.flat, m.sqrt.tensor.flat)
  return _flat_vector(m, like)


def init_mass(mass) -> MassMatrix:
  """integrator.py:99-116 (diagonal mass as ``Tensor(ndim=1)`` of the inverse
  and the square root).  Evaluated inside the fused kernels; this helper keeps
  the API and returns host pytrees."""
  from .tree_util import tree_map
  inv = tree_map(lambda x: np.power(np.asarray(x, np.float32), np.float32(-1.0)), mass)
  sqrt = tree_map(lambda x: np.sqrt(np.asarray(x, np.float32)), mass)
  return MassMatrix(inv=Tensor(1, inv), sqrt=Tensor(1, sqrt))


def random_tree(key, a: ChainTree) -> ChainTree:
  """integrator.py:119-135 for every chain: ``splits = split(key, n_leaves)``,
  leaf l ~ ``normal(splits[l], leaf.shape)``.  ``key``: KeyState or uint32[C,2]."""
  ks = _keys_for(key, a.n_chains)
  noise = ops.normal_like(ks.current, a.sizes)
  return ChainTree.like(a, noise)


# -----------------------------------------------------------------------------

class _Scratch:
  """Scratch buffers (gradient, mass / friction vectors) of one integrator,
  looked up by the state's flat sample array.

  An entry is only valid for the very array object it was built for (a weak
  reference is kept and compared -- ``id()`` alone can be reused by a later
  array of a different shape) and for the same ``token`` objects (e.g. the
  ``mass`` argument of ``integrate``); entries of dead arrays are dropped.
  """

  def __init__(self):
    self._entries: Dict[int, Any] = {}

  def get(self, owner: DeviceArray, build: Callable[[], Any], *token):
    ent = self._entries.get(id(owner))
    if ent is not None:
      ref, tok, val = ent
      if ref() is owner and len(tok) == len(token) and all(
          a is b for a, b in zip(tok, token)):
        return val
    for k in [k for k, (r, _, _) in self._entries.items() if r() is None]:
      del self._entries[k]
    val = build()
    self._entries[id(owner)] = (weakref.ref(owner), token, val)
    return val


def langevin_diffusion(potential_fn, batch_fn, adaption=None
                       ) -> Tuple[Callable, Callable, Callable]:
  """integrator.py:767-924."""
  if adaption is not None:
    if getattr(adaption, "fused_kind", None) != "rms_prop":
      raise NotImplementedError("only adaption.rms_prop() is fused into the "
                                "Langevin update")
    adapt_init, _, _ = adaption
  batch_init, batch_get, _ = batch_fn
  stochastic_gradient = _potential.value_and_grad(potential_fn)
  scratch = _Scratch()

  def init_fn(init_sample, key=None, adaption_kwargs: Dict = None,
              batch_kwargs: Dict = None, init_model_state: PyTree = None
              ) -> LangevinState:
    adaption_kwargs = adaption_kwargs or {}
    batch_kwargs = batch_kwargs or {}
    sample = _as_chain_tree(init_sample)
    C = sample.n_chains
    adaption_state = None if adaption is None else adapt_init(sample, **adaption_kwargs)
    return LangevinState(
        key=_keys_for(key, C), latent_variables=sample, adapt_state=adaption_state,
        data_state=_init_data_state(batch_init, batch_kwargs, C),
        model_state=init_model_state,
        potential=DeviceArray.zeros((C,)),                    # :842
        variance=DeviceArray.full((C,), 1.0))                 # :843

  def get_fn(state: LangevinState) -> Dict[str, PyTree]:
    """integrator.py:851-855 (``likelihood`` is ``-potential``)."""
    return {"variables": state.latent_variables,
            "likelihood": Negated(state.potential),
            "model_state": state.model_state}

  def update_fn(state: LangevinState, parameters, temp_per_chain=None,
                pre_update_hook: Callable = None, carry_ok: bool = False) -> LangevinState:
    """integrator.py:860-922.  ``pre_update_hook`` (a callable with an ``event``
    attribute) orders the update launch after that event -- the sharded reSGLD
    puts its label exchange there.  ``carry_ok``: the caller never writes the
    sample between updates (see ``_PotentialFn.sgld_step``)."""
    theta = state.latent_variables
    data_state, mini_batch = batch_get(state.data_state, information=True)   # :872
    grad_buf = scratch.get(theta.flat, lambda: DeviceArray(theta.flat.shape, np.float32))
    assert grad_buf.shape == theta.flat.shape
    v = alpha = lmbd = None
    if adaption is not None:
      v, alpha, lmbd = state.adapt_state.v.flat, state.adapt_state.alpha, \
          state.adapt_state.lmbd
    new_model_state = state.model_state
    # one C call for value_and_grad + update when all chains share the minibatch
    if not potential_fn.sgld_step(
        theta, mini_batch, state.key.current, state.key.next,
        float(parameters.step_size), float(parameters.temperature), v=v,
        alpha=alpha if alpha is not None else 0.9, lmbd=lmbd if lmbd is not None else 1e-5,
        temp_per_chain=temp_per_chain,
        wait_event=getattr(pre_update_hook, "event", None), grad_out=grad_buf,
        U_out=state.potential, var_out=state.variance, carry_ok=carry_ok):
      (_, (_, new_model_state)), grad = stochastic_gradient(               # :875-880
          theta, mini_batch, state=state.model_state, likelihoods=True,
          grad_out=grad_buf, U_out=state.potential, var_out=state.variance)
      if pre_update_hook is not None:
        pre_update_hook()
      # key, split = split(key); noise; scaled gradient / noise; adaption; theta' (:871-912)
      ops.sgld_update(theta.flat, grad.flat, state.key.current, state.key.next,
                      theta.sizes, float(parameters.step_size),
                      float(parameters.temperature), temp_per_chain=temp_per_chain,
                      v=v, alpha=alpha if alpha is not None else 0.9,
                      lmbd=lmbd if lmbd is not None else 1e-5)
    state.key.flip()
    return LangevinState(key=state.key, latent_variables=theta,
                         adapt_state=state.adapt_state, data_state=data_state,
                         model_state=new_model_state, potential=state.potential,
                         variance=state.variance)

  def scan_fn(state: LangevinState, step_sizes, temperatures, keep, samples_out,
              scalars_out, kept: int):
    """len(step_sizes) update_fn steps + sample collection in one C call.
    Returns ``(state, kept)`` or None when this configuration needs the step
    loop (per-chain minibatches, streamed data, stateful models)."""
    source_fn = getattr(batch_get, "scan_source", None)
    if source_fn is None or state.model_state is not None:
      return None
    theta = state.latent_variables
    steps = len(step_sizes)
    source = source_fn(state.data_state, steps)
    if source is None:
      return None
    grad_buf = scratch.get(theta.flat, lambda: DeviceArray(theta.flat.shape, np.float32))
    assert grad_buf.shape == theta.flat.shape
    v, alpha, lmbd = None, 0.9, 1e-5
    if adaption is not None:
      v, alpha, lmbd = state.adapt_state.v.flat, state.adapt_state.alpha, \
          state.adapt_state.lmbd
    kept = potential_fn.sgld_scan(
        theta, source, state.key.current, state.key.next, step_sizes, temperatures, keep,
        samples_out, scalars_out, kept, v=v, alpha=alpha, lmbd=lmbd, grad_out=grad_buf,
        U_out=state.potential, var_out=state.variance)
    if kept is None:          # the native scan declined (after consuming nothing)
      return None
    if steps % 2:
      state.key.flip()
    return state, kept

  update_fn.scan = scan_fn
  return init_fn, update_fn, get_fn


# -----------------------------------------------------------------------------
def friction_leapfrog(potential_fn, batch_fn, steps: int = 10, friction=0.25,
                      const_mass: PyTree = None, noise_model=None
                      ) -> Tuple[Callable, Callable, Callable]:
  """integrator.py:563-765 (SGHMC).  ``noise_model``: an ``adaption.fisher_information``
  triplet; its ``get`` is evaluated at the new positions on the step's minibatch and the
  corrected scale multiplies the injected noise (:632-650)."""
  get_noise_model = noise_model[2] if noise_model else None
  init_data, get_data, _ = batch_fn
  stochastic_gradient = _potential.value_and_grad(potential_fn)
  scratch = _Scratch()

  def init_fn(init_sample, key=None, batch_kwargs: Dict = None,
              init_model_state: PyTree = None) -> LeapfrogState:
    batch_kwargs = batch_kwargs or {}
    sample = _as_chain_tree(init_sample)
    C = sample.n_chains
    return LeapfrogState(
        potential=DeviceArray.zeros((C,)), key=_keys_for(key, C), positions=sample,
        momentum=sample.copy(),                    # placeholder = the sample (:702)
        data_state=_init_data_state(init_data, batch_kwargs, C),
        model_state=init_model_state,
        extra_fields=None)

  def integrate(state: LeapfrogState, parameters, mass: PyTree = None) -> LeapfrogState:
    theta, p = state.positions, state.momentum
    eps = float(parameters.step_size)
    def build():
      m = mass if mass is not None else const_mass
      fr_leaves = tree_flatten(friction)[0]
      scalar_fr = len(fr_leaves) == 1 and np.ndim(fr_leaves[0]) == 0
      return {
          "grad": DeviceArray(theta.flat.shape, np.float32),
          "mass": None if m is None else _flat_vector(m, theta),
          "friction": None if scalar_fr else _flat_vector(friction, theta),
          "friction_scalar": float(fr_leaves[0]) if scalar_fr else 0.0}
    sc = scratch.get(theta.flat, build, mass)      # rebuilt when `mass` changes
    assert sc["grad"].shape == theta.flat.shape
    # resample momentum (:736-738) fused with the first position update (:610-612)
    ops.sghmc_begin(theta.flat, p.flat, state.key.current, state.key.next,
                    theta.sizes, eps, sc["mass"])
    state.key.flip()
    data_state, model_state = state.data_state, state.model_state
    for s in range(steps):                                               # :749-755
      data_state, mini_batch = get_data(data_state, information=True)    # :621
      (_, model_state), grad = stochastic_gradient(                      # :622-625
          theta, mini_batch, state=model_state, grad_out=sc["grad"],
          U_out=state.potential)
      noise_mul = None
      if get_noise_model is not None:                                    # :632-650
        fr = sc["friction"] if sc["friction"] is not None else sc["friction_scalar"]
        noise_mul = get_noise_model(None, theta, grad, fr, mini_batch=mini_batch,
                                    step_size=eps, model_state=model_state
                                    ).cb_diff_sqrt.tensor.flat
      ops.sghmc_step(theta.flat, p.flat, grad.flat, state.key.current,
                     state.key.next, theta.sizes, eps, sc["friction_scalar"],
                     sc["friction"], sc["mass"], last=(s == steps - 1), noise_mul=noise_mul)
      state.key.flip()
    return LeapfrogState(positions=theta, momentum=p, key=state.key,
                         potential=state.potential, model_state=model_state,
                         data_state=data_state, extra_fields=state.extra_fields)

  def get_fn(state: LeapfrogState) -> Dict[str, PyTree]:
    return {"variables": state.positions, "energy": state.potential,
            "model_state": state.model_state}

  return init_fn, integrate, get_fn


# -----------------------------------------------------------------------------
def obabo(potential_fn, batch_fn, steps: int = 10, friction: float = 1.0,
          const_mass: PyTree = None) -> Tuple[Callable, Callable, Callable]:
  """integrator.py:138-346."""
  init_data, get_data, _ = batch_fn
  stochastic_gradient = _potential.value_and_grad(potential_fn)
  scratch = _Scratch()

  def init_fn(init_sample, key=None, batch_kwargs: Dict = None,
              init_model_state: PyTree = None) -> ObaboState:
    batch_kwargs = batch_kwargs or {}
    sample = _as_chain_tree(init_sample)
    C = sample.n_chains
    zeros = ChainTree.like(sample, DeviceArray.zeros(sample.flat.shape))  # :303
    return ObaboState(
        kinetic_energy_start=DeviceArray.zeros((C,)),
        kinetic_energy_end=DeviceArray.zeros((C,)), potential=DeviceArray.zeros((C,)),
        key=_keys_for(key, C), positions=sample, momentum=zeros,
        data_state=_init_data_state(init_data, batch_kwargs, C),
        model_state=init_model_state)

  def integrate(state: ObaboState, parameters, mass: PyTree = None) -> ObaboState:
    theta, p = state.positions, state.momentum
    eps, T = float(parameters.step_size), float(parameters.temperature)
    C = theta.n_chains
    def build():
      m = mass if mass is not None else const_mass
      return {
          "grad": DeviceArray(theta.flat.shape, np.float32),
          "U1": DeviceArray((C,), np.float32), "U2": DeviceArray((C,), np.float32),
          "mass": _mass_operand(m, theta)}
    sc = scratch.get(theta.flat, build, mass)      # rebuilt when `mass` changes
    assert sc["grad"].shape == theta.flat.shape
    data_state, model_state = state.data_state, state.model_state
    for _ in range(steps):                                               # :333-336
      data_state, mb = get_data(data_state, information=True)            # :225
      (_, model_state), g1 = stochastic_gradient(theta, mb, state=model_state,
                                                 grad_out=sc["grad"], U_out=sc["U1"])
      ops.obabo_pass_a(theta.flat, p.flat, g1.flat, state.kinetic_energy_start,
                       state.key.current, state.key.next, theta.sizes, eps, T,
                       float(friction), sc["mass"])                      # :210-240
      data_state, mb = get_data(data_state, information=True)            # :243
      (_, model_state), g2 = stochastic_gradient(theta, mb, state=model_state,
                                                 grad_out=sc["grad"], U_out=sc["U2"])
      ops.obabo_pass_b(p.flat, g2.flat, state.kinetic_energy_end,
                       state.key.current, theta.sizes, eps, T, float(friction),
                       sc["mass"])                                       # :248-261
      state.key.flip()
      ops.axpby(state.potential, 0.5, sc["U1"], 0.5, sc["U2"])           # :264
    return ObaboState(positions=theta, momentum=p, key=state.key,
                      potential=state.potential, model_state=model_state,
                      data_state=data_state,
                      kinetic_energy_start=state.kinetic_energy_start,
                      kinetic_energy_end=state.kinetic_energy_end)

  def get_fn(state) -> Dict[str, PyTree]:
    return {"variables": state.positions, "energy": state.potential,
            "model_state": state.model_state}

  return init_fn, integrate, get_fn


def reversible_leapfrog(potential_fn, batch_fn, steps: int = 10, friction=0.25,
                        const_mass: PyTree = None) -> Tuple[Callable, Callable, Callable]:
  """integrator.py:349-560 (the AMAGOLD integrator): half position step,
  ``steps`` momentum updates with friction and accumulated energy difference,
  half position step."""
  init_data, get_data, _ = batch_fn
  stochastic_gradient = _potential.value_and_grad(potential_fn)
  scratch = _Scratch()

  def _mass_vector(mass, theta):
    return _mass_operand(mass if mass is not None else const_mass, theta)

  def init_fn(init_sample, key=None, batch_kwargs: Dict = None,
              init_model_state: PyTree = None, mass: PyTree = None) -> LeapfrogState:
    batch_kwargs = batch_kwargs or {}
    sample = _as_chain_tree(init_sample)
    C = sample.n_chains
    ks = ops.split(_keys_for(key, C).current, 2).numpy()            # :500
    momentum = random_tree(ks[:, 1], sample)                        # :501, :383-386
    m = _mass_vector(mass, sample)
    if isinstance(m, tuple):                       # adapted: sqrt(M) per chain on the device
      ops.tree_ewise(2, momentum.flat, 1.0, m[1], momentum.flat)
    elif m is not None:
      momentum.flat.copy_from_host(momentum.flat.numpy() * np.sqrt(m.numpy())[None, :])
    return LeapfrogState(
        potential=DeviceArray.zeros((C,)), key=KeyState(ks[:, 0]), positions=sample,
        momentum=momentum, data_state=_init_data_state(init_data, batch_kwargs, C),
        model_state=init_model_state, extra_fields=None)

  def integrate(state: LeapfrogState, parameters, mass: PyTree = None) -> LeapfrogState:
    theta, p = state.positions, state.momentum
    eps = np.float32(parameters.step_size)
    def build():
      m = _mass_vector(mass, theta)
      inv_full = None
      if isinstance(m, tuple):
        inv_full = m[0]      # adapted: M^-1 per chain, read in place
      elif m is not None:    # inv_m broadcast over the chains for the opening half step
        inv_full = DeviceArray.from_numpy(
            np.tile((np.float32(1.0) / m.numpy())[None, :], (theta.n_chains, 1)))
      return {
          "grad": DeviceArray(theta.flat.shape, np.float32), "mass": m,
          "inv_full": inv_full, "tmp": None if m is None else DeviceArray(theta.flat.shape,
                                                                         np.float32),
          "U": DeviceArray((theta.n_chains,), np.float32)}
    sc = scratch.get(theta.flat, build, mass)      # rebuilt when `mass` changes
    assert sc["grad"].shape == theta.flat.shape
    half = float(np.float32(0.5) * eps)
    if sc["mass"] is None:                                           # :523-524
      ops.axpby(theta.flat, 1.0, theta.flat, half, p.flat)
    else:
      ops.tree_ewise(2, sc["tmp"], 1.0, sc["inv_full"], p.flat)
      ops.axpby(theta.flat, 1.0, theta.flat, half, sc["tmp"])
    state.potential.zero_()                                          # :531
    data_state, model_state = state.data_state, state.model_state
    for s in range(steps):                                           # :538-541
      data_state, mini_batch = get_data(data_state, information=True)   # :427
      (_, model_state), grad = stochastic_gradient(
          theta, mini_batch, state=model_state, grad_out=sc["grad"], U_out=sc["U"])
      ops.revleapfrog_step(theta.flat, p.flat, grad.flat, state.potential,
                           state.key.current, state.key.next, theta.sizes, float(eps),
                           float(friction), sc["mass"], last=(s == steps - 1))
      state.key.flip()
    return LeapfrogState(positions=theta, momentum=p, key=state.key,
                         potential=state.potential, model_state=model_state,
                         data_state=data_state, extra_fields=state.extra_fields)

  def get_fn(state: LeapfrogState) -> Dict[str, PyTree]:
    return {"variables": state.positions, "energy": state.potential,
            "model_state": state.model_state}

  return init_fn, integrate, get_fn
