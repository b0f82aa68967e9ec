"""Sample collection, mirroring the slice of ``jax_sgmc.io`` the solvers use.

Reference: ``io.save`` (io.py:626-772), ``io.no_save`` (:775-879),
``MemoryCollector`` (:502-619).  The reference ships every kept sample to the
host with ``host_callback.id_tap`` (io.py:703).  Here a kept sample costs the
sampling stream one device-to-device copy: small runs keep all samples in a
preallocated ``[samples, C, P]`` device buffer and download it once at the end;
large runs (or ``MemoryCollector(stream_to_host=True)``) copy the sample into
one of a few staging slots and a second stream moves it to pinned host memory
while the chains keep stepping, so the HBM footprint
does not grow with the number of kept samples.  Output format per chain is the
reference's:
``{"sample_count": n, "samples": {"variables": pytree[n, ...], "likelihood" |
"energy": array[n], "model_state": None}}``.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional

import numpy as np

from . import _lib
from .device import DeviceArray, Event, Stream, current_stream
from .tree_util import ChainTree, unravel_rows

# kept samples above this many bytes are streamed to the host instead of being
# held in HBM until the end of the run
DEVICE_BUFFER_LIMIT = 1 << 30
# ... and land directly in a page-locked result array up to this size
PINNED_RESULT_LIMIT = 24 << 30
RING_SLOTS = 4


class Negated:
  """Lazy ``-array`` (``langevin.get_fn`` reports ``likelihood = -potential``,
  integrator.py:853-855); the sign is applied when the value reaches the host."""

  def __init__(self, array: DeviceArray):
    self.array = array

  def numpy(self):
    return -self.array.numpy()


class MemoryCollector:
  """io.py:502-619.  ``stream_to_host``: True / False force the host ring /
  the device buffer; None picks by size (``DEVICE_BUFFER_LIMIT``)."""

  def __init__(self, save_dir=None, stream_to_host: Optional[bool] = None):
    self.save_dir = save_dir
    self.stream_to_host = stream_to_host


class _PinnedOwner:
  def __init__(self, ptr):
    self.ptr = ptr

  def __del__(self):
    try:
      _lib.call("sgmc_host_free", self.ptr)
    except Exception:  # interpreter shutdown
      pass


def _pinned_array(n_floats: int, write_combined: bool = False):
  """float32[n] backed by page-locked host memory; freed when the last numpy
  view of it dies.  Returns (array, address).  ``write_combined``: for staging the host
  only ever writes (reading it back from the CPU is very slow)."""
  h = C.c_void_p()
  _lib.call("sgmc_host_alloc_wc" if write_combined else "sgmc_host_alloc", C.byref(h),
            max(n_floats, 1) * 4)
  raw = (C.c_float * max(n_floats, 1)).from_address(h.value)
  raw._owner = _PinnedOwner(h)
  return np.ctypeslib.as_array(raw), h.value


class _HostRing:
  """Kept samples -> host memory while the chains keep stepping.

  push(): the sampling stream copies theta into one of ``RING_SLOTS`` device
  staging slots (d2d, the only cost on that stream) and records an event; the
  copy stream waits for it and moves the slot to the host (D2H), straight into
  the page-locked result array when it fits ``PINNED_RESULT_LIMIT``, otherwise
  into a pinned slot that is drained into a pageable result array.  The host
  only ever waits for a copy issued ``RING_SLOTS`` kept samples ago."""

  def __init__(self, capacity: int, C_: int, P: int):
    self.row = C_ * P
    self.shape = (max(capacity, 1), C_, P)
    self.direct = self.shape[0] * self.row * 4 <= PINNED_RESULT_LIMIT
    self.copy_stream = Stream.create()
    self.dev = DeviceArray((RING_SLOTS, C_, P), np.float32)
    if self.direct:
      flat, self.out_addr = _pinned_array(self.shape[0] * self.row)
      self.out = flat.reshape(self.shape)
      self.slots = None
    else:
      self.out = np.empty(self.shape, np.float32)
      self.slots = [_pinned_array(self.row) for _ in range(RING_SLOTS)]
    self.staged = [Event() for _ in range(RING_SLOTS)]
    self.landed = [Event() for _ in range(RING_SLOTS)]
    self.holds = [-1] * RING_SLOTS        # sample index in flight per slot

  def _drain(self, r: int):
    if self.holds[r] >= 0:
      self.landed[r].sync()
      if not self.direct:
        self.out[self.holds[r]].reshape(-1)[:] = self.slots[r][0]
      self.holds[r] = -1

  def push(self, index: int, flat: DeviceArray):
    r = index % RING_SLOTS
    self._drain(r)            # the slot's previous D2H has finished: safe to overwrite
    stream = current_stream()
    slot = self.dev.row_slice(r, r + 1)
    slot.copy_from(flat, stream)
    self.staged[r].record(stream)
    self.copy_stream.wait_event(self.staged[r])
    dst = self.out_addr + index * self.row * 4 if self.direct else self.slots[r][1]
    _lib.call("sgmc_memcpy_d2h", C.c_void_p(dst), C.c_void_p(slot.ptr), self.row * 4,
              self.copy_stream.handle)
    self.landed[r].record(self.copy_stream)
    self.holds[r] = index

  def finish(self, n: int) -> np.ndarray:
    for r in range(RING_SLOTS):
      self._drain(r)
    return self.out[:n]


class JSONCollector(MemoryCollector):
  def __init__(self, *a, **k):
    raise NotImplementedError("JSONCollector is host glue outside this path")


class HDF5Collector(MemoryCollector):
  def __init__(self, *a, **k):
    raise NotImplementedError("HDF5Collector is host glue outside this path")


class _SavingState:
  def __init__(self, template: Dict[str, Any], capacity: int,
               stream_to_host: Optional[bool] = None):
    var: ChainTree = template["variables"]
    C, P = var.flat.shape
    self.template = var
    self.scalar_key = "likelihood" if "likelihood" in template else "energy"
    self.capacity = int(capacity)
    if stream_to_host is None:
      stream_to_host = self.capacity * C * P * 4 > DEVICE_BUFFER_LIMIT
    self.ring = _HostRing(self.capacity, C, P) if stream_to_host else None
    self.variables = None if stream_to_host else \
        DeviceArray((max(self.capacity, 1), C, P), np.float32)
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


def _make(checkpoint_every: int = 0, stream_to_host: Optional[bool] = None):
  if checkpoint_every != 0:
    raise NotImplementedError("Checkpointing is not supported")    # io.py:681-682

  def init_saving(init_sample, init_checkpoint, static_information):
    del init_checkpoint
    return _SavingState(init_sample, static_information.samples_collected, stream_to_host)

  def save(state: _SavingState, keep, sample, **unused):
    del unused
    if keep and state.count < state.capacity:
      if state.ring is not None:
        state.ring.push(state.count, sample["variables"].flat)
      else:
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
    var = state.ring.finish(n) if state.ring is not None \
        else state.variables.numpy()[:n]         # [n, C, P]
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
  return _make(checkpoint_every, getattr(data_collector, "stream_to_host", None))


def no_save():
  """io.py:775-879."""
  return _make(0)


def load(*args, **kwargs):
  raise NotImplementedError("Loading of checkpoints is currently not supported.")
