"""Schedules of the solver parameters, mirroring ``jax_sgmc.scheduler``.

Host-side scalar glue: the hot path consumes
four scalars per iteration, ``schedule(step_size, temperature, burn_in,
accept)`` (reference scheduler.py:48-52).  Same triplet structure and
arithmetic as the reference (f32 NumPy instead of jnp).
"""
from __future__ import annotations

from collections import namedtuple
from typing import Callable, Tuple

import numpy as np

F32 = np.float32

specific_scheduler = namedtuple("specific_scheduler", ["init", "update", "get"])
schedule = namedtuple("schedule", ["step_size", "temperature", "burn_in", "accept"])
scheduler_state = namedtuple("scheduler_state",
                             ["state", "step_size_state", "temperature_state",
                              "burn_in_state", "thinning_state",
                              "progress_bar_state"])
static_information = namedtuple("static_information", ["samples_collected"])


def _keep_state(state, iteration, **kw):
  """update() of the schedulers whose state never changes.  Carrying THIS function is the
  opt-in tag that lets ``precompute`` evaluate a scheduler ahead of the run; a scheduler
  with any other ``update`` (user-defined, ``adaptive_step_size``) is driven step by
  step, as in the reference."""
  del iteration, kw
  return state


def is_static(sched: specific_scheduler) -> bool:
  return sched.update is _keep_state


def init_scheduler(step_size: specific_scheduler = None,
                   temperature: specific_scheduler = None,
                   burn_in: specific_scheduler = None,
                   thinning: specific_scheduler = None,
                   progress_bar: bool = True,
                   progress_bar_steps: int = 20) -> Tuple[Callable, Callable, Callable]:
  """scheduler.py:93-237.  Defaults: constant step size 1.0 (:118-119),
  temperature 1.0, no burn in, accept everything."""
  if step_size is None:
    step_size = polynomial_step_size(a=1, b=1, gamma=0.0)
  if temperature is None:
    temperature = constant_temperature(tau=1.0)
  if burn_in is None:
    burn_in = initial_burn_in(n=0)
  if thinning is None:
    thinning = specific_scheduler(lambda iterations: (None, iterations),
                                  _keep_state, lambda *a, **k: True)

  def init_fn(iterations: int, **kw):
    thinning_state, total = thinning.init(iterations, **kw.get("thinning", {}))
    burn_in_state, collected = burn_in.init(iterations, **kw.get("burn_in", {}))
    total = min(total, collected)
    pb = None
    if progress_bar:
      pb = {"every": max(1, iterations // kw.get("progress_bar_steps", progress_bar_steps)),
            "iterations": iterations, "enabled": kw.get("enabled", True)}
    st = scheduler_state(
        state=(0, iterations),
        step_size_state=step_size.init(iterations, **kw.get("step_size", {})),
        temperature_state=temperature.init(iterations, **kw.get("temperature", {})),
        burn_in_state=burn_in_state, thinning_state=thinning_state,
        progress_bar_state=pb)
    return st, static_information(samples_collected=total)

  def update_fn(state: scheduler_state, **kw) -> scheduler_state:
    it, total = state.state
    pb = state.progress_bar_state
    if pb is not None and pb["enabled"] and (it % pb["every"] == 0):
      print(f"[Step {it}/{pb['iterations']}]({100 * it // pb['iterations']:.0f}%)",
            flush=True)
    return scheduler_state(
        state=(it + 1, total),
        step_size_state=step_size.update(state.step_size_state, it, **kw),
        temperature_state=temperature.update(state.temperature_state, it, **kw),
        burn_in_state=burn_in.update(state.burn_in_state, it, **kw),
        thinning_state=thinning.update(state.thinning_state, it, **kw),
        progress_bar_state=pb)

  def get_fn(state: scheduler_state, **kw) -> schedule:
    it, _ = state.state
    return schedule(
        step_size=F32(step_size.get(state.step_size_state, it, **kw)),
        temperature=F32(temperature.get(state.temperature_state, it, **kw)),
        burn_in=F32(burn_in.get(state.burn_in_state, it, **kw)),
        accept=bool(thinning.get(state.thinning_state, it, **kw)))

  def precompute(state: scheduler_state, iterations: int):
    """(step sizes f32[K], temperatures f32[K], keep bool[K]) of the next K
    iterations when every specific scheduler is tagged static (the built-in polynomial /
    constant / burn-in / thinning schedules) -- what lets solver.mcmc hand the whole
    scan to native code -- else None: the step loop then calls update() / get() per
    iteration."""
    parts = (step_size, temperature, burn_in, thinning)
    if not all(is_static(p) for p in parts):
      return None
    it0 = state.state[0]
    its = range(it0, it0 + iterations)
    eps = np.array([step_size.get(state.step_size_state, i) for i in its], F32)
    tau = np.array([temperature.get(state.temperature_state, i) for i in its], F32)
    keep = np.array([bool(burn_in.get(state.burn_in_state, i))
                     and bool(thinning.get(state.thinning_state, i)) for i in its])
    return eps, tau, keep

  get_fn.precompute = precompute
  return init_fn, update_fn, get_fn


def constant_temperature(tau: float = 1.0) -> specific_scheduler:
  """scheduler.py:245-276."""
  return specific_scheduler(lambda iterations, tau=tau: tau,
                            _keep_state,
                            lambda state, iteration, **kw: state)


def polynomial_step_size(a: float = 1.0, b: float = 1.0, gamma: float = 0.33
                         ) -> specific_scheduler:
  """scheduler.py:447-491: ``a * (b + n) ** (-gamma)`` precomputed (f32)."""

  def init_fn(iterations: int, a=a, b=b, gamma=gamma):
    assert gamma >= 0, f"Gamma must be positive: gamma = {gamma}"
    assert a > 0, f"a must be positive: a = {a}"
    assert b > 0, f"b must be greater than zero: b = {b}"
    n = np.arange(iterations).astype(F32)
    unscaled = np.power((F32(b) + n).astype(F32), F32(-gamma)).astype(F32)
    return (F32(a) * unscaled).astype(F32)

  return specific_scheduler(init_fn, _keep_state,
                            lambda state, iteration, **kw: state[iteration])


def polynomial_step_size_first_last(first: float = 1.0, last: float = 1.0,
                                    gamma: float = 0.33) -> specific_scheduler:
  """scheduler.py:494-545."""

  def find_ab(its, gamma, first, last):
    gamma, first, last = F32(gamma), F32(first), F32(last)
    ginv = np.power(gamma, F32(-1.0)).astype(F32)
    fpow = np.power(first, -ginv).astype(F32)
    lpow = np.power(last, -ginv).astype(F32)
    apow = ((lpow - fpow).astype(F32) / F32(its - 1)).astype(F32)
    a = np.power(apow, -gamma).astype(F32)
    b = np.power((first / a).astype(F32), -ginv).astype(F32)
    return a, b

  def init_fn(iterations: int, first=first, last=last, gamma=gamma):
    assert gamma > 0, f"Gamma must be bigger than 0, is {gamma}"
    assert first >= last, (f"The first step size must be larger than the last:"
                           f" {first} !>= {last}")
    a, b = find_ab(iterations, gamma, first, last)
    return polynomial_step_size(a=a, b=b, gamma=gamma).init(iterations)

  return specific_scheduler(init_fn, _keep_state,
                            lambda state, iteration, **kw: state[iteration])


def adaptive_step_size(burn_in=0, initial_step_size=0.05, stabilization_constant=100,
                       decay_constant=0.75, speed_constant=0.05,
                       target_acceptance_rate=0.02) -> specific_scheduler:
  """scheduler.py:376-444: dual averaging of log(step size) on the MH acceptance
  ratio during burn in.  All chains of a call share one schedule here, so the
  statistic is the MEAN acceptance ratio of the chains (for a single chain this
  is the reference's update; the reference adapts every chain separately)."""

  def init_fn(iterations: int, burn_in=burn_in, initial_step_size=initial_step_size,
              stabilization_constant=stabilization_constant,
              decay_constant=decay_constant, speed_constant=speed_constant,
              target_acceptance_rate=target_acceptance_rate):
    del iterations
    x_bar = np.log(F32(initial_step_size)).astype(F32)
    return (burn_in, x_bar, F32(0.0), F32(target_acceptance_rate),
            F32(stabilization_constant), F32(decay_constant), F32(speed_constant),
            np.log(F32(10 * initial_step_size)).astype(F32))

  def update_fn(state, iteration: int, acceptance_ratio=0.0, **kw):
    del kw
    burn_in, x_bar, h_bar, alpha, t0, kappa, gamma, mu = state
    acc = F32(np.mean(np.asarray(acceptance_ratio, dtype=np.float32)))
    m = F32(iteration + 1)
    # the reference uses the closed-over target_acceptance_rate (:423), not `alpha`
    h_bar = F32(h_bar * F32(F32(1) - F32(1) / F32(m + t0)))
    h_bar = F32(h_bar + F32(F32(1) / F32(m + t0)) * F32(F32(target_acceptance_rate) - acc))
    x = F32(mu - F32(np.sqrt(m) / gamma) * h_bar)
    lr = F32(np.power(m, -kappa))
    x_new = F32(F32(x_bar * F32(F32(1) - lr)) + F32(lr * x))
    x_bar = x_new if iteration < burn_in else x_bar             # only during burn in
    return burn_in, x_bar, h_bar, alpha, t0, kappa, gamma, mu

  def get_fn(state, iteration: int, **kw):
    del iteration, kw
    return F32(np.exp(state[1]))

  return specific_scheduler(init_fn, update_fn, get_fn)


def initial_burn_in(n: int = 0) -> specific_scheduler:
  """scheduler.py:567-596: discard the first n steps (returns 0.0 / 1.0)."""
  return specific_scheduler(
      lambda iterations, n=n: (n, iterations - n),
      _keep_state,
      lambda state, iteration, **kw: F32(1.0) if state <= iteration else F32(0.0))


def random_thinning(step_size_schedule: specific_scheduler,
                    burn_in_schedule: specific_scheduler, selections: int,
                    key=None) -> specific_scheduler:
  """scheduler.py:599-666: ``selections`` iterations drawn without replacement
  with probability proportional to ``step_size * burn_in``
  (``jax.random.choice(key, arange(its), (selections,), replace=False, p)``,
  default key ``PRNGKey(0)``).

  jax implements this with the Gumbel top-k trick on
  ``gumbel(key, (n,)) + log(p)``; the uniforms come from the same threefry
  stream (device kernel), the logs are taken on the host [recall: not pinned by
  any reference test; thinning only selects which samples are kept].
  """

  def init_fn(iterations: int, step_size_schedule=step_size_schedule,
              burn_in_schedule=burn_in_schedule, selections=selections, key=key):
    from . import ops
    from .device import DeviceArray
    k = ops.prng_key(0) if key is None else np.asarray(key, np.uint32)
    ss = step_size_schedule.init(iterations)
    bs, _ = burn_in_schedule.init(iterations)
    probs = np.array([step_size_schedule.get(ss, i) * burn_in_schedule.get(bs, i)
                      for i in range(iterations)], dtype=F32)
    assert np.count_nonzero(probs) >= selections, "Cannot select enough values"
    p = probs / probs.sum(dtype=F32)
    tiny = np.finfo(F32).tiny
    u = ops.uniform(DeviceArray.from_numpy(k[None]), iterations, tiny, 1.0).numpy()[0]
    with np.errstate(divide="ignore"):
      g = -np.log(-np.log(u.astype(F32))) + np.log(p.astype(F32))
    accepted = np.argsort(-g, kind="stable")[:selections]
    lookup = np.zeros(iterations, dtype=bool)
    lookup[accepted] = True
    return lookup, selections

  return specific_scheduler(init_fn, _keep_state,
                            lambda state, iteration, **kw: bool(state[iteration]))
