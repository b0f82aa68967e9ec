"""Solvers of the sampling hot path, mirroring ``jax_sgmc.solver``.

``mcmc`` (reference solver.py:65-185), ``sgmc`` (:188-217) and
``parallel_tempering`` (:220-299).  ``mcmc`` drives ALL chains of a call in
lock-step through the batched kernels (the reference's ``strategy='vmap'``,
solver.py:179-180); ``strategy='map'`` of the alias solvers (chains one after
another, :165-171) gives the same per-chain results because chains are
independent, so both strategy names are accepted and run batched.

Multi-GPU: ``ShardedTempering`` shards reSGLD replicas over ranks and
exchanges the per-replica energies with an all-gather (NCCL on the device,
``torch.distributed`` gloo in the CPU tests).
"""
from __future__ import annotations

from functools import partial
from typing import Any, Callable, Dict, List, Tuple

import numpy as np

from . import io, ops
from .device import DeviceArray
from .integrator import KeyState, LangevinState, _as_chain_tree


def mcmc(solver, scheduler, strategy="map", saving=None, loading=None):
  """solver.py:65-185."""
  if strategy not in ("map", "vmap", "pmap"):
    raise NotImplementedError(f"Strategy {strategy} is unknown. ")
  if loading is not None:
    raise NotImplementedError("Loading of checkpoints is currently not supported.")
  init_saving, save, postprocess_saving = io.no_save() if saving is None else saving
  scheduler_init, scheduler_next, scheduler_get = scheduler
  _, solver_update, solver_get = solver

  def run(*states, schedulers=None, iterations: int = int(1e5)):
    iterations = int(iterations)
    results = []
    for init_state in states:       # every state is already a batch of chains
      main_scheduler, static_information = scheduler_init(iterations)
      scheduler_states = [main_scheduler]
      if schedulers is not None:
        scheduler_states += [scheduler_init(iterations, **kw)[0] for kw in schedulers]
      saving_state = init_saving(solver_get(init_state),
                                 (init_state, scheduler_states), static_information)
      state = init_state
      for _ in range(iterations):                                   # solver.py:152-160
        schedules = [scheduler_get(s) for s in scheduler_states]    # :97-101
        state, stats = solver_update(state, *schedules)             # :102
        keep = bool(schedules[0].burn_in) and bool(schedules[0].accept)   # :104
        saving_state, _ = save(saving_state, keep, solver_get(state),
                               scheduler_state=scheduler_states, solver_state=state)
        kw = stats if stats is not None else {}                     # :113-122
        scheduler_states = [scheduler_next(s, **kw) for s in scheduler_states]
      results.extend(postprocess_saving(saving_state, None))
    return results

  return run


def sgmc(integrator) -> Tuple[Callable, Callable, Callable]:
  """solver.py:188-217."""
  init_integrator, update_integrator, get_integrator = integrator

  def init(*args, **kwargs):
    return init_integrator(*args, **kwargs)

  def update(state, schedule):
    return update_integrator(state, schedule), None

  def get(state) -> Dict[str, Any]:
    return get_integrator(state)

  return init, update, get


class TemperingState:
  """(normal_chain, hot_chain, ssq, F, step, key) of solver.py:259 for S
  independent reSGLD systems (one per chain row)."""

  def __init__(self, normal, hot, ssq, F, step, key):
    self.normal, self.hot, self.ssq, self.F, self.step, self.key = \
        normal, hot, ssq, F, step, key
    S = ssq.shape[0]
    self.exchange = DeviceArray((S,), np.int32)

  def __iter__(self):
    return iter((self.normal, self.hot, self.ssq, self.F, self.step, self.key))

  def __getitem__(self, i):
    return tuple(self)[i]


def _swap_langevin(a: LangevinState, b: LangevinState, exchange: DeviceArray):
  """lax.cond swap of the whole chain states (solver.py:287-291): positions,
  keys, RMSprop state, potential and variance rows are exchanged where
  ``exchange`` is set."""
  ops.swap_rows(a.latent_variables.flat, b.latent_variables.flat, exchange)
  ops.swap_rows(a.key.current, b.key.current, exchange)
  ops.swap_rows(a.potential, b.potential, exchange)
  ops.swap_rows(a.variance, b.variance, exchange)
  if a.adapt_state is not None:
    ops.swap_rows(a.adapt_state.v.flat, b.adapt_state.v.flat, exchange)


def parallel_tempering(integrator, sa_schedule: Callable = lambda n: 1 / n
                       ) -> Tuple[Callable, Callable, Callable]:
  """solver.py:220-299 (reSGLD, two temperatures per system)."""
  del sa_schedule          # 1/n is fused in the decision kernel (solver.py:221)
  init_integrator, update_integrator, get_integrator = integrator

  def init(normal_sample, tempered_sample, ssq_init=0.0, key=None, F=1.0, **kwargs):
    normal_sample = _as_chain_tree(normal_sample)
    tempered_sample = _as_chain_tree(tempered_sample)
    S = normal_sample.n_chains
    key = ops.prng_key(0) if key is None else np.asarray(key, np.uint32)
    if key.ndim == 1:
      key = np.tile(key, (S, 1))
    ks = ops.split(DeviceArray.from_numpy(key), 3).numpy()          # :254
    normal_chain = init_integrator(normal_sample, key=ks[:, 1], **kwargs)   # :256
    hot_chain = init_integrator(tempered_sample, key=ks[:, 2], **kwargs)    # :257
    return TemperingState(normal_chain, hot_chain,
                          DeviceArray.full((S,), float(ssq_init)),
                          DeviceArray.full((S,), float(F)), 0, KeyState(ks[:, 0]))

  def update(state: TemperingState, normal_schedule, hot_schedule):
    state.step += 1                                                  # :267
    state.normal = update_integrator(state.normal, normal_schedule)  # :270
    state.hot = update_integrator(state.hot, hot_schedule)           # :271
    # ssq / log_s / log_u / decision (:273-286), fused
    ops.resgld_decide(state.normal.potential, state.hot.potential,
                      state.normal.variance, state.ssq, state.F, state.step,
                      float(normal_schedule.temperature),
                      float(hot_schedule.temperature), state.key.current,
                      state.key.next, state.exchange)
    state.key.flip()
    _swap_langevin(state.normal, state.hot, state.exchange)          # :287-291
    return state, None

  def get(state) -> Dict[str, Any]:
    return get_integrator(state.normal)                              # :296-297

  return init, update, get


def amagold(*args, **kwargs):
  raise NotImplementedError("solver.amagold is the next tier (SURVEY.md 8f)")


def sggmc(*args, **kwargs):
  raise NotImplementedError("solver.sggmc is the next tier (SURVEY.md 8f)")
