"""Solvers of the sampling hot path, mirroring ``jax_sgmc.solver``.

``mcmc`` (reference solver.py:65-185), ``sgmc`` (:188-217) and
``parallel_tempering`` (:220-299).  ``mcmc`` drives ALL chains of a call in
lock-step through the batched kernels (the reference's ``strategy='vmap'``,
solver.py:179-180); ``strategy='map'`` of the alias solvers (chains one after
another, :165-171) gives the same per-chain results because chains are
independent, so both strategy names are accepted and run batched.

Multi-GPU: ``ShardedTempering`` shards reSGLD replicas over ranks and
exchanges the per-replica energies with an all-gather (NCCL on the device,
host arrays over the socket control plane / gloo in the CPU tests).
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
      if schedulers is None and _native_scan(solver_update, scheduler_get, main_scheduler,
                                            iterations, state, saving_state):
        results.extend(postprocess_saving(saving_state, None))
        continue
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


def _native_scan(solver_update, scheduler_get, scheduler_state, iterations, state,
                 saving_state) -> bool:
  """Run the whole scan of solver.py:152-160 in native code when the solver has a
  scan (sgmc over langevin_diffusion on a GLM potential), the schedules are
  static and the samples go to the in-HBM buffer.  SGMC_NATIVE_SCAN=0 disables."""
  import os
  scan = getattr(solver_update, "scan", None)
  precompute = getattr(scheduler_get, "precompute", None)
  if scan is None or precompute is None or os.environ.get("SGMC_NATIVE_SCAN", "1") == "0":
    return False
  if not isinstance(saving_state, io._SavingState) or saving_state.ring is not None \
      or saving_state.extra or saving_state.scalar_key != "likelihood":
    return False
  arrays = precompute(scheduler_state, iterations)
  if arrays is None:
    return False
  eps, tau, keep = arrays
  collect = saving_state.capacity > 0
  # the progress bar of scheduler.update_fn (one line every `every` iterations): the scan
  # is cut at those iterations, otherwise it is one native call
  pb = scheduler_state.progress_bar_state
  every = pb["every"] if pb is not None and pb["enabled"] else iterations
  it0 = scheduler_state.state[0]
  done = 0
  while done < iterations:
    k = min(iterations - done, every - (it0 + done) % every)
    if pb is not None and pb["enabled"] and (it0 + done) % every == 0:
      print(f"[Step {it0 + done}/{pb['iterations']}]"
            f"({100 * (it0 + done) // pb['iterations']:.0f}%)", flush=True)
    out = scan(state, eps[done:done + k], tau[done:done + k], keep[done:done + k],
               saving_state.variables if collect else None,
               saving_state.scalars if collect else None, saving_state.count)
    if out is None:
      assert done == 0, "the native scan refused a later chunk of a run it had accepted"
      return False
    _, saving_state.count = out
    done += k
  saving_state.negate = True          # langevin.get_fn reports likelihood = -potential
  return True


def sgmc(integrator) -> Tuple[Callable, Callable, Callable]:
  """solver.py:188-217."""
  init_integrator, update_integrator, get_integrator = integrator

  def init(*args, **kwargs):
    return init_integrator(*args, **kwargs)

  code = getattr(update_integrator, "__code__", None)
  takes_carry = code is not None and "carry_ok" in code.co_varnames[:code.co_argcount]

  def update(state, schedule):
    # sgmc never touches the sample between updates: the Langevin integrator may
    # carry its operand form from step to step
    if takes_carry:
      return update_integrator(state, schedule, carry_ok=True), None
    return update_integrator(state, schedule), None

  def get(state) -> Dict[str, Any]:
    return get_integrator(state)

  update.scan = getattr(update_integrator, "scan", None)    # native scan, if any
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


def _default_sa_schedule(n):
  return 1 / n


def parallel_tempering(integrator, sa_schedule: Callable = _default_sa_schedule
                       ) -> Tuple[Callable, Callable, Callable]:
  """solver.py:220-299 (reSGLD, two temperatures per system).  The default
  ``sa_schedule`` (1 / n, solver.py:221) is evaluated inside the decision kernel; any
  other schedule is evaluated on the host per step (solver.py:274-276) and handed to
  the kernel as ``eta``."""
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
                      state.key.next, state.exchange,
                      eta=None if sa_schedule is _default_sa_schedule
                      else float(np.float32(sa_schedule(state.step))))
    state.key.flip()
    _swap_langevin(state.normal, state.hot, state.exchange)          # :287-291
    return state, None

  def get(state) -> Dict[str, Any]:
    return get_integrator(state.normal)                              # :296-297

  return init, update, get


class MHState:
  """SGGMCState / AMAGOLDState (solver.py:45-62) for C chains advancing together.

  ``potential`` is the full-data potential of the current sample; ``saved`` holds
  the pre-proposal copies the rejected chains are restored from."""

  def __init__(self, integrator_state, potential, full_data_state, key, C, P):
    self.integrator_state = integrator_state
    self.potential = potential
    self.full_data_state = full_data_state
    self.key = key
    self.mass_state = None
    self.reject = DeviceArray((C,), np.int32)
    self.ratio = DeviceArray.zeros((C,))
    self.kinetic = DeviceArray.zeros((C,))
    self.step_size = 0.0
    self.saved = {"theta": DeviceArray((C, P), np.float32),
                  "momentum": DeviceArray((C, P), np.float32),
                  "potential": DeviceArray((C,), np.float32)}

  @property
  def acceptance_ratio(self):
    return self.ratio, self.step_size, self.kinetic


def _mh_solver(kind, integrator_fn, full_potential_fn, full_data_map, mass_adaption):
  init_mass = update_mass = get_mass = None
  if mass_adaption is not None:                                       # :322-323 / :459-460
    init_mass, update_mass, get_mass = mass_adaption
  init_integrator, update_integrator, get_integrator = integrator_fn
  init_full_data, full_data_map_fn, _ = full_data_map

  def init(init_sample, key=None, initial_mass=None, full_data_kwargs: dict = None,
           **kwargs) -> MHState:
    sample = _as_chain_tree(init_sample)
    C = sample.n_chains
    mass_state = init_mass(sample, initial_mass) if init_mass else None   # :330-335 / :486-489
    if mass_state is not None and kind == "amagold":
      kwargs["mass"] = get_mass(mass_state)        # the initial momentum draw uses it (:353)
    full_data_state = init_full_data(**(full_data_kwargs or {}))
    potential, (full_data_state, model_state) = full_potential_fn(    # :474-478 / :341-345
        sample, full_data_state, full_data_map_fn, state=kwargs.get("init_model_state"))
    kwargs["init_model_state"] = model_state
    key = ops.prng_key(0) if key is None else np.asarray(key, np.uint32)
    if key.ndim == 1:
      key = np.tile(key, (C, 1))
    ks = ops.split(DeviceArray.from_numpy(key), 2).numpy()            # :483 / :349
    integrator_state = init_integrator(sample, key=ks[:, 0], **kwargs)
    state = MHState(integrator_state, potential, full_data_state, KeyState(ks[:, 1]),
                    C, sample.n_params)
    state.mass_state = mass_state
    return state

  def update(state: MHState, schedule):
    old = state.integrator_state
    # the integrators update in place: keep what a rejected chain goes back to
    state.saved["theta"].copy_from(old.positions.flat)
    state.saved["momentum"].copy_from(old.momentum.flat)
    state.saved["potential"].copy_from(old.potential)
    mass = get_mass(state.mass_state) if get_mass else None           # :503-506 / :366-369
    proposal = update_integrator(old, schedule, mass=mass)            # :512-515 / :372-375
    new_potential, (full_data_state, _) = full_potential_fn(          # :518-522 / :378-382
        proposal.positions, state.full_data_state, full_data_map_fn,
        state=proposal.model_state)
    if kind == "sggmc":
      ops.mh_decide("sggmc", state.potential, new_potential,
                    proposal.kinetic_energy_start, proposal.kinetic_energy_end,
                    float(schedule.temperature), state.key.current, state.key.next,
                    state.reject, state.ratio)                        # :524-539
      ops.axpby(state.kinetic, 1.0, proposal.kinetic_energy_end, -1.0,
                proposal.kinetic_energy_start)                        # :563
    else:
      ops.mh_decide("amagold", state.potential, new_potential, None, proposal.potential,
                    1.0, state.key.current, state.key.next, state.reject,
                    state.ratio)                                      # :381-395
      # a rejected chain restarts from the old state with the momentum flipped (:392-399)
      ops.tree_ewise(0, state.saved["momentum"], -1.0, state.saved["momentum"])
    state.key.flip()
    ops.swap_rows(proposal.positions.flat, state.saved["theta"], state.reject)
    ops.swap_rows(proposal.momentum.flat, state.saved["momentum"], state.reject)
    ops.swap_rows(proposal.potential, state.saved["potential"], state.reject)
    if kind == "sggmc":                                               # :551-552
      proposal.kinetic_energy_start.zero_()
      proposal.kinetic_energy_end.zero_()
    if update_mass:          # adapt the mass on the accepted sample (:552-553 / :409-410)
      state.mass_state = update_mass(state.mass_state, proposal.positions)
    state.integrator_state = proposal       # data_state / key of the proposal (:547-550)
    state.full_data_state = full_data_state
    state.step_size = float(schedule.step_size)
    return state, {"acceptance_ratio": state.ratio}

  def get(state: MHState) -> Dict[str, Any]:
    out = get_integrator(state.integrator_state)
    out["acceptance_ratio"] = state.ratio
    out["step_size"] = np.full((state.ratio.shape[0],), state.step_size, np.float32)
    if kind == "sggmc":
      out["kinetic_energy"] = state.kinetic
      out["potential"] = state.potential
    return out

  return init, update, get


def amagold(integrator_fn, full_potential_fn, full_data_map, mass_adaption=None
            ) -> Tuple[Callable, Callable, Callable]:
  """solver.py:301-432: reversible leapfrog proposal + amortised MH correction."""
  return _mh_solver("amagold", integrator_fn, full_potential_fn, full_data_map,
                    mass_adaption)


def sggmc(integrator_fn, full_potential_fn, full_data_map, mass_adaption=None
          ) -> Tuple[Callable, Callable, Callable]:
  """solver.py:434-577: OBABO proposal + MH correction on the full potential."""
  return _mh_solver("sggmc", integrator_fn, full_potential_fn, full_data_map,
                    mass_adaption)
