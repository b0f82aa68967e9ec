"""Sample collection, mirroring the slice of ``jax_sgmc.io`` the solvers use.

Reference: ``io.save`` (io.py:626-772), ``io.no_save`` (:775-879),
``MemoryCollector`` (:502-619).  The reference ships every kept sample to the
host with ``host_callback.id_tap`` (io.py:703); here kept samples are copied
device-to-device into a preallocated ``[samples, C, P]`` buffer and downloaded
once at the end.  Output format per chain is the reference's:
``{"sample_count": n, "samples": {"variables": pytree[n, ...], "likelihood" |
"energy": array[n], "model_state": None}}``.
"""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np

from .device import DeviceArray
from .tree_util import ChainTree, unravel_rows


class Negated:
  """Lazy ``-array`` (``langevin.get_fn`` reports ``likelihood = -potential``,
  integrator.py:853-855); the sign is applied when the value reaches the host."""

  def __init__(self, array: DeviceArray):
    self.array = array

  def numpy(self):
    return -self.array.numpy()


class MemoryCollector:
  """io.py:502-619 (API marker; the storage is the device buffer below)."""

  def __init__(self, save_dir=None):
    self.save_dir = save_dir


class JSONCollector(MemoryCollector):
  def __init__(self, *a, **k):
    raise NotImplementedError("JSONCollector is host glue outside this path")


class HDF5Collector(MemoryCollector):
  def __init__(self, *a, **k):
    raise NotImplementedError("HDF5Collector is host glue outside this path")


class _SavingState:
  def __init__(self, template: Dict[str, Any], capacity: int):
    var: ChainTree = template["variables"]
    C, P = var.flat.shape
    self.template = var
    self.scalar_key = "likelihood" if "likelihood" in template else "energy"
    self.capacity = int(capacity)
    self.variables = DeviceArray((max(self.capacity, 1), C, P), np.float32)
    self.scalars = DeviceArray((max(self.capacity, 1), C), np.float32)
    # further per-chain scalars of the solver's get() (acceptance_ratio, step_size,
    # kinetic_energy, potential of the MH solvers; solver.py:426-431, :568-575)
    self.extra = {}
    for k, v in template.items():
      if k in ("variables", "model_state", self.scalar_key):
        continue
      if isinstance(v, DeviceArray) and v.shape == (C,):
        self.extra[k] = DeviceArray((max(self.capacity, 1), C), np.float32)
      elif isinstance(v, np.ndarray) and v.shape == (C,):
        self.extra[k] = np.zeros((max(self.capacity, 1), C), np.float32)
    self.count = 0


def _make(checkpoint_every: int = 0):
  if checkpoint_every != 0:
    raise NotImplementedError("Checkpointing is not supported")    # io.py:681-682

  def init_saving(init_sample, init_checkpoint, static_information):
    del init_checkpoint
    return _SavingState(init_sample, static_information.samples_collected)

  def save(state: _SavingState, keep, sample, **unused):
    del unused
    if keep and state.count < state.capacity:
      state.variables.row_slice(state.count, state.count + 1).copy_from(
          sample["variables"].flat)
      sc = sample[state.scalar_key]
      state.negate = isinstance(sc, Negated)     # "likelihood" = -U, kept lazy
      state.scalars.row_slice(state.count, state.count + 1).copy_from(
          sc.array if state.negate else sc)
      for k, buf in state.extra.items():
        if isinstance(buf, DeviceArray):
          buf.row_slice(state.count, state.count + 1).copy_from(sample[k])
        else:
          buf[state.count] = sample[k]
      state.count += 1
    return state, None

  def postprocess(state: _SavingState, unused_saved=None) -> List[Dict[str, Any]]:
    n = state.count
    var = state.variables.numpy()[:n]            # [n, C, P]
    sca = state.scalars.numpy()[:n]              # [n, C]
    if getattr(state, "negate", False):
      sca = -sca                                 # get_fn labels -U (integrator.py:853-855)
    extra = {k: (b.numpy() if isinstance(b, DeviceArray) else b)[:n]
             for k, b in state.extra.items()}
    out = []
    for c in range(var.shape[1]):
      tree = unravel_rows(var[:, c], state.template.treedef, state.template.shapes)
      samples = {"variables": tree, state.scalar_key: sca[:, c].copy(), "model_state": None}
      samples.update({k: b[:, c].copy() for k, b in extra.items()})
      out.append({"sample_count": n, "samples": samples})
    return out

  return init_saving, save, postprocess


def save(data_collector: MemoryCollector = None, checkpoint_every: int = 0):
  """io.py:626-772."""
  del data_collector
  return _make(checkpoint_every)


def no_save():
  """io.py:775-879."""
  return _make(0)


def load(*args, **kwargs):
  raise NotImplementedError("Loading of checkpoints is currently not supported.")
