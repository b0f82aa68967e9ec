"""Ready-to-use solvers, mirroring ``jax_sgmc.alias`` for the hot path.

``sgld`` (reference alias.py:30-120), ``re_sgld`` (:122-208), ``sghmc``
(:451-540) and ``obabo`` (:542-619) keep signatures and defaults; each returns
``run_fn(*init_samples, init_model_state=None, iterations=1000)`` producing one
result dict per chain.  All chains of a call advance together through the
batched kernels; with the reference's defaults (every chain on PRNGKey(0),
integrator.py:804) the per-chain results equal the reference's sequential
``strategy='map'`` runs.  ``keys=`` (not in the reference) seeds chains
individually.
"""
from __future__ import annotations

from functools import partial
from typing import Any, Union

import numpy as np

from . import adaption, data, integrator, io, scheduler, solver
from .tree_util import ChainTree

Pytree = Any


def _chains(init_samples) -> ChainTree:
  """One host pytree per chain (the reference's ``*init_samples``), or a single
  ChainTree that already holds every chain on the device."""
  init_samples = list(init_samples)
  if len(init_samples) == 1 and isinstance(init_samples[0], ChainTree):
    return init_samples[0]
  return ChainTree.from_trees(init_samples)


def _schedule(first_step_size, last_step_size, burn_in, accepted_samples,
              progress_bar, temperature=None):
  step_size_schedule = scheduler.polynomial_step_size_first_last(
      first=first_step_size, last=last_step_size)
  burn_in_schedule = scheduler.initial_burn_in(burn_in)
  thinning = scheduler.random_thinning(step_size_schedule, burn_in_schedule,
                                       selections=accepted_samples)
  kw = {} if temperature is None else {"temperature": temperature}
  return scheduler.init_scheduler(step_size=step_size_schedule,
                                  burn_in=burn_in_schedule, thinning=thinning,
                                  progress_bar=progress_bar, **kw)


def _saving(save_to_numpy):
  return io.save(io.MemoryCollector()) if save_to_numpy else None


def sgld(potential_fn, data_loader, cache_size: int = 512, batch_size: int = 32,
         first_step_size: float = 0.05, last_step_size: float = 0.001,
         burn_in: int = 0, accepted_samples: int = 1000, rms_prop: bool = False,
         alpha: float = 0.9, lmbd: float = 1e-5, save_to_numpy: bool = True,
         progress_bar: bool = True):
  """alias.py:30-120: SGLD with polynomial step size and optional RMSprop."""
  random_data = data.random_reference_data(data_loader, cache_size, batch_size)
  rms = adaption.rms_prop() if rms_prop else None
  rms_integrator = integrator.langevin_diffusion(potential_fn, random_data,
                                                 adaption=rms)
  schedule = _schedule(first_step_size, last_step_size, burn_in, accepted_samples,
                       progress_bar)
  sgld_solver = solver.sgmc(rms_integrator)
  mcmc = solver.mcmc(sgld_solver, schedule, strategy="map",
                     saving=_saving(save_to_numpy))

  def run_fn(*init_samples, init_model_state: Pytree = None, iterations=1000,
             keys=None):
    state = sgld_solver[0](_chains(init_samples), key=keys,
                           adaption_kwargs={"alpha": alpha, "lmbd": lmbd},
                           init_model_state=init_model_state)
    return mcmc(state, iterations=iterations)

  return run_fn


def re_sgld(potential_fn, data_loader, cache_size: int = 512, batch_size: int = 32,
            temperature: float = 1000.0, first_step_size: float = 0.05,
            last_step_size: float = 0.001, burn_in: int = 0,
            accepted_samples: int = 100, save_to_numpy: bool = True,
            progress_bar: bool = True):
  """alias.py:122-208: replica exchange SGLD; ``init_samples`` are
  ``(normal, tempered)`` tuples."""
  del progress_bar                                                  # :171
  random_data = data.random_reference_data(data_loader, cache_size, batch_size)
  resgld_integrator = integrator.langevin_diffusion(potential_fn, random_data)
  schedule = _schedule(first_step_size, last_step_size, burn_in, accepted_samples,
                       False, temperature=scheduler.constant_temperature(1.0))
  resgld_solver = solver.parallel_tempering(resgld_integrator)
  mcmc = solver.mcmc(resgld_solver, schedule, strategy="map",
                     saving=_saving(save_to_numpy))

  def run_fn(*init_samples, init_model_state: Pytree = None, iterations=1000,
             keys=None):
    normal, tempered = zip(*init_samples)
    state = resgld_solver[0](_chains(normal),
                             _chains(tempered), key=keys,
                             init_model_state=init_model_state)
    return mcmc(state, iterations=iterations,
                schedulers=[{"temperature": {"tau": temperature}}])   # :205-207

  return run_fn


def sghmc(potential_fn, data_loader, cache_size: int = 512, batch_size: int = 32,
          integration_steps: int = 10, friction: Union[float, Pytree] = 1.0,
          mass: Pytree = None, first_step_size: float = 0.05,
          last_step_size: float = 0.001, burn_in: int = 0,
          accepted_samples: int = 1000, adapt_noise_model: bool = False,
          diagonal_noise: bool = True, save_to_numpy: bool = True,
          progress_bar: bool = True):
  """alias.py:451-540."""
  random_data = data.random_reference_data(data_loader, cache_size, batch_size)
  noise_model = None
  if adapt_noise_model:                                                # alias.py:506-510
    noise_model = adaption.fisher_information(minibatch_potential=potential_fn,
                                              diagonal=diagonal_noise)
  leapfrog = integrator.friction_leapfrog(potential_fn, random_data,
                                          friction=friction, const_mass=mass,
                                          steps=integration_steps, noise_model=noise_model)
  schedule = _schedule(first_step_size, last_step_size, burn_in, accepted_samples,
                       progress_bar)
  sghmc_solver = solver.sgmc(leapfrog)
  mcmc = solver.mcmc(sghmc_solver, schedule, strategy="map",
                     saving=_saving(save_to_numpy))

  def run_fn(*init_samples, init_model_state: Pytree = None, iterations=1000,
             keys=None):
    state = sghmc_solver[0](_chains(init_samples), key=keys,
                            init_model_state=init_model_state)
    return mcmc(state, iterations=iterations)

  return run_fn


def obabo(potential_fn, data_loader, cache_size: int = 512, batch_size: int = 32,
          integration_steps: int = 10, friction: Union[float, Pytree] = 1.0,
          mass: Pytree = None, first_step_size: float = 0.05,
          last_step_size: float = 0.001, burn_in: int = 0,
          accepted_samples: int = 1000, save_to_numpy: bool = True,
          progress_bar: bool = True):
  """alias.py:542-619."""
  random_data = data.random_reference_data(data_loader, cache_size, batch_size)
  obabo_integrator = integrator.obabo(potential_fn=potential_fn, batch_fn=random_data,
                                      steps=integration_steps, friction=friction,
                                      const_mass=mass)
  schedule = _schedule(first_step_size, last_step_size, burn_in, accepted_samples,
                       progress_bar)
  obabo_solver = solver.sgmc(obabo_integrator)
  mcmc = solver.mcmc(obabo_solver, schedule, strategy="map",
                     saving=_saving(save_to_numpy))

  def run_fn(*init_samples, init_model_state: Pytree = None, iterations=1000,
             keys=None):
    state = obabo_solver[0](_chains(init_samples), key=keys,
                            init_model_state=init_model_state)
    return mcmc(state, iterations=iterations)

  return run_fn


def _mh_schedule(first_step_size, last_step_size, adaptive_step_size,
                 stabilization_constant, decay_constant, speed_constant,
                 target_acceptance_rate, burn_in, accepted_samples, progress_bar):
  """alias.py:285-311 / :405-431."""
  burn_in_schedule = scheduler.initial_burn_in(burn_in)
  if adaptive_step_size:
    step_size_schedule = scheduler.adaptive_step_size(
        burn_in=burn_in, initial_step_size=first_step_size,
        stabilization_constant=stabilization_constant, decay_constant=decay_constant,
        speed_constant=speed_constant, target_acceptance_rate=target_acceptance_rate)
    thinning = None
    assert accepted_samples is None, ("Thinning currently not supported for"
                                      " adaptive step size.")
  else:
    step_size_schedule = scheduler.polynomial_step_size_first_last(
        first=first_step_size, last=last_step_size)
    thinning = None if accepted_samples is None else scheduler.random_thinning(
        step_size_schedule, burn_in_schedule, selections=accepted_samples)
  return scheduler.init_scheduler(step_size=step_size_schedule, burn_in=burn_in_schedule,
                                  thinning=thinning, progress_bar=progress_bar)


def amagold(stochastic_potential_fn, full_potential_fn, data_loader,
            cache_size: int = 512, batch_size: int = 32, integration_steps: int = 10,
            friction: float = 0.25, first_step_size: float = 0.001,
            last_step_size: float = 0.001, adaptive_step_size: bool = False,
            stabilization_constant: int = 10, decay_constant: float = 0.75,
            speed_constant: float = 0.05, target_acceptance_rate: float = 0.25,
            burn_in: int = 0, accepted_samples: Union[int, None] = None,
            mass: Pytree = None, save_to_numpy: bool = True, progress_bar: bool = True):
  """alias.py:210-328."""
  random_data = data.random_reference_data(data_loader, cache_size, batch_size)
  full_data_map = data.full_reference_data(data_loader, cache_size, batch_size)
  reversible_leapfrog = integrator.reversible_leapfrog(
      stochastic_potential_fn, random_data, integration_steps, friction, mass)
  amagold_solver = solver.amagold(reversible_leapfrog, full_potential_fn, full_data_map)
  schedule = _mh_schedule(first_step_size, last_step_size, adaptive_step_size,
                          stabilization_constant, decay_constant, speed_constant,
                          target_acceptance_rate, burn_in, accepted_samples, progress_bar)
  mcmc = solver.mcmc(amagold_solver, schedule, strategy="map",
                     saving=_saving(save_to_numpy))

  def run_fn(*init_samples, init_model_state: Pytree = None, iterations=1000, keys=None):
    state = amagold_solver[0](_chains(init_samples), key=keys,
                              init_model_state=init_model_state)
    return mcmc(state, iterations=iterations)

  return run_fn


def sggmc(stochastic_potential_fn, full_potential_fn, data_loader,
          cache_size: int = 512, batch_size: int = 32, integration_steps: int = 10,
          friction_coefficient: float = 1.0, first_step_size: float = 0.001,
          last_step_size: float = 0.001, adaptive_step_size: bool = False,
          stabilization_constant: int = 10, decay_constant: float = 0.75,
          speed_constant: float = 0.05, target_acceptance_rate: float = 0.25,
          burn_in: int = 0, accepted_samples: Union[int, None] = None,
          mass: Pytree = None, save_to_numpy: bool = True, progress_bar: bool = True):
  """alias.py:330-449."""
  random_data = data.random_reference_data(data_loader, cache_size, batch_size)
  full_data_map = data.full_reference_data(data_loader, cache_size, batch_size)
  obabo_integrator = integrator.obabo(stochastic_potential_fn, random_data,
                                      integration_steps, friction_coefficient, mass)
  sggmc_solver = solver.sggmc(obabo_integrator, full_potential_fn, full_data_map)
  schedule = _mh_schedule(first_step_size, last_step_size, adaptive_step_size,
                          stabilization_constant, decay_constant, speed_constant,
                          target_acceptance_rate, burn_in, accepted_samples, progress_bar)
  mcmc = solver.mcmc(sggmc_solver, schedule, strategy="map",
                     saving=_saving(save_to_numpy))

  def run_fn(*init_samples, init_model_state: Pytree = None, iterations=1000, keys=None):
    state = sggmc_solver[0](_chains(init_samples), key=keys,
                            init_model_state=init_model_state)
    return mcmc(state, iterations=iterations)

  return run_fn
