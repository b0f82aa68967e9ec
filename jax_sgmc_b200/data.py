"""Reference-data access for the sampling hot path (thin host glue).

Mirrors the parts of ``jax_sgmc.data`` the hot path touches: the loader
classes, ``MiniBatchInformation`` (data/core.py:49-61),
``random_reference_data`` (data/core.py:477-520) and ``full_reference_data``
(:522-627).  B200-first redesign: the *whole data set lives in HBM* (180 GB)
and a minibatch is just an index vector -- the gather is fused into the
potential kernels -- instead of shipping ``cache_size`` batches of rows from
the host through ``io_callback`` (core.py:664-791).

Index-draw semantics are the reference's:
* ``DeviceNumpyDataLoader``: ``key, split = split(key)``;
  ``randint(split, (n,), 0, N)`` on the device (numpy_loader.py:128-141),
  default key ``PRNGKey(0)`` for every chain (:124).
* ``NumpyDataLoader``: NumPy PCG64, ``default_rng(SeedSequence(seed)
  .spawn(1)[0])`` with ``seed = chain_id`` (numpy_loader.py:263-265) and
  ``rng.choice`` / shuffle / in-epoch draws (:382-460), ``cache_size`` batches
  of indices drawn per refill.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, NamedTuple, Optional, Tuple

import numpy as np

from . import ops
from .device import DeviceArray


class MiniBatchInformation(NamedTuple):
  """data/core.py:49-61."""
  observation_count: int
  mask: Any
  batch_size: int


class BatchRef:
  """A minibatch as indices into the HBM-resident data set.

  ``idx`` is ``int32[n]`` (one minibatch shared by all chains) or
  ``int32[C, n]`` (one per chain); ``None`` means rows ``0..n-1``.
  ``mask`` is ``f32[n]`` or None.
  """

  def __init__(self, loader: "DataLoader", idx: Optional[DeviceArray], n: int,
               mask: Optional[DeviceArray] = None, host_idx=None,
               rows: Optional[Dict[str, DeviceArray]] = None):
    self.loader, self.idx, self.n, self.mask = loader, idx, int(n), mask
    self.host_idx = host_idx
    # streaming loader: the minibatch rows themselves ({leaf: f32[n, ...]} on
    # the device, already gathered on the host), idx is None
    self.rows = rows

  def leaf(self, name: str) -> DeviceArray:
    """The array the potential kernels read ``name`` from: the gathered rows of
    a streamed batch, else the HBM-resident data set (indexed by ``idx``)."""
    return self.rows[name] if self.rows is not None else self.loader.device_data[name]

  @property
  def per_chain(self) -> bool:
    return self.idx is not None and self.idx.ndim == 2

  def materialize(self) -> Dict[str, np.ndarray]:
    """Gathered rows on the host (debugging / tests): {name: array[n, ...]}."""
    if self.rows is not None:
      return {k: v.numpy() for k, v in self.rows.items()}
    idx = self.idx.numpy() if self.idx is not None else np.arange(self.n)
    return {k: v.numpy()[idx] for k, v in self.loader.device_data.items()}


MiniBatch = Tuple[BatchRef, MiniBatchInformation]


class DataLoader:
  """Reference data resident in HBM (one DeviceArray per named leaf)."""

  def __init__(self, copy=True, **reference_data):
    del copy
    assert len(reference_data) > 0, "Observations are required."
    counts = {int(np.shape(v)[0]) for v in reference_data.values()}
    assert len(counts) == 1, "All arrays must have the same leading dimension."
    self._observation_count = counts.pop()
    self.host_shapes = {k: np.shape(v) for k, v in reference_data.items()}
    self.device_data = {
        k: (v if isinstance(v, DeviceArray)
            else DeviceArray.from_numpy(np.asarray(v, dtype=np.float32)))
        for k, v in reference_data.items()}

  @property
  def static_information(self):
    return {"observation_count": self._observation_count}

  def absmax(self, name: str) -> float:
    """max |leaf| over the whole data set (device reduction, computed once):
    the bound the tensor-core potential uses to scale its fp16 operands."""
    cache = self.__dict__.setdefault("_absmax", {})
    if name not in cache:
      cache[name] = ops.absmax(self.device_data[name])
    return cache[name]

  @property
  def _format(self):
    return {k: ((), v.shape[1:]) for k, v in self.device_data.items()}

  def initializer_batch(self, mb_size: int = None):
    """All-zero observation / batch (numpy_loader.py:42-62)."""
    return {k: np.zeros(((mb_size,) if mb_size else ()) + tuple(s[1:]), np.float32)
            for k, s in self.host_shapes.items()}


class DeviceNumpyDataLoader(DataLoader):
  """numpy_loader.py:89-141."""

  def init_random_data(self, *args, **kwargs):
    del args
    key = kwargs.get("key", None)
    return ops.prng_key(0) if key is None else np.asarray(key, np.uint32)


class NumpyDataLoader(DataLoader):
  """numpy_loader.py:147-460 (index draws only; rows stay in HBM)."""

  def __init__(self, copy=True, **reference_data):
    super().__init__(copy=copy, **reference_data)
    self._chains: List[dict] = []

  def register_random_pipeline(self, cache_size: int = 1, mb_size: int = None,
                               in_epochs: bool = False, shuffle: bool = False,
                               **kwargs) -> int:
    if mb_size > self._observation_count:
      raise ValueError(f"The batch size cannot be bigger than the observation "
                       f"count. Provided {mb_size} and {self._observation_count}")
    if not shuffle and in_epochs:
      raise ValueError("in_epochs = True can only be used for shuffle = True.")
    chain_id = len(self._chains)
    seed = kwargs.get("seed", chain_id)
    rng = np.random.default_rng(np.random.SeedSequence(seed).spawn(1)[0])
    self._chains.append({
        "type": "random", "rng": rng, "idx_offset": None, "in_epochs": in_epochs,
        "shuffle": shuffle, "remaining_samples": 0,
        "draws": math.ceil(self._observation_count / mb_size),
        "random_indices": None, "mb_size": mb_size, "cache_size": cache_size})
    return chain_id

  def register_ordered_pipeline(self, cache_size: int = 1, mb_size: int = None,
                                **kwargs) -> int:
    del kwargs
    assert mb_size <= self._observation_count
    chain_id = len(self._chains)
    self._chains.append({"type": "ordered", "rng": None, "idx_offset": 0,
                         "mb_size": mb_size, "cache_size": cache_size,
                         "in_epochs": False, "shuffle": False})
    return chain_id

  # -- index draws, numpy_loader.py:342-460 -----------------------------------------
  def get_indices(self, chain_id: int):
    chain = self._chains[chain_id]
    if chain["type"] == "ordered":
      fn = self._ordered_indices
    elif chain["in_epochs"]:
      fn = self._shuffle_in_epochs
    elif chain["shuffle"]:
      fn = self._shuffle_indices
    else:
      fn = self._draw_indices
    pairs = [fn(chain) for _ in range(chain["cache_size"])]
    idx, masks = zip(*pairs)
    return np.array(idx), np.array(masks, dtype=np.bool_)

  def _ordered_indices(self, chain):
    N, mb = self._observation_count, chain["mb_size"]
    idcs = np.arange(mb) + chain["idx_offset"]
    mask = idcs < N
    if chain["idx_offset"] + mb >= N:
      chain["idx_offset"] = 0
    else:
      chain["idx_offset"] += mb
    return np.mod(idcs, N), mask

  def _draw_indices(self, chain):
    # == rng.choice(np.arange(0, N), size=mb, replace=True) of numpy_loader.py:382-389,
    # which draws exactly these integers (tests/test_host_logic.py pins the equality)
    sel = chain["rng"].integers(0, self._observation_count, size=chain["mb_size"])
    return sel, np.ones(chain["mb_size"], dtype=np.bool_)

  def _shuffle_indices(self, chain):
    N, mb = self._observation_count, chain["mb_size"]
    floor_draws = math.floor(N / mb)
    ceil_draws = floor_draws + 2
    if chain["remaining_samples"] < mb:
      new_indices = chain["rng"].choice(np.arange(0, N), size=N, replace=False)
      if chain["random_indices"] is None:
        chain["draws"] = 0
        chain["random_indices"] = np.zeros(ceil_draws * mb, dtype=np.int_)
      update_idxs = np.mod(np.arange(N) + chain["draws"] * mb
                           + chain["remaining_samples"], ceil_draws * mb)
      chain["random_indices"][update_idxs] = new_indices
      chain["remaining_samples"] += N
    mask = np.ones(mb, dtype=np.bool_)
    sel_idx = np.mod(np.arange(mb) + chain["draws"] * mb, mb * ceil_draws)
    sel = np.copy(chain["random_indices"][sel_idx])
    chain["draws"] = (chain["draws"] + 1) % ceil_draws
    chain["remaining_samples"] -= mb
    return sel, mask

  def _shuffle_in_epochs(self, chain):
    N, mb = self._observation_count, chain["mb_size"]
    ceil_draws = math.ceil(N / mb)
    if chain["draws"] == ceil_draws:
      new_indices = chain["rng"].choice(np.arange(0, N), size=N, replace=False)
      if chain["random_indices"] is None:
        chain["draws"] = 0
        chain["random_indices"] = np.zeros(ceil_draws * mb, dtype=np.int_)
      chain["random_indices"][0:N] = new_indices
      chain["draws"] = 0
    start, end = mb * chain["draws"], mb * (chain["draws"] + 1)
    mask = np.arange(start, end) < N
    sel = np.copy(chain["random_indices"][start:end])
    chain["draws"] += 1
    return sel, mask


class StreamingNumpyDataLoader(NumpyDataLoader):
  """A host-resident data set that does NOT fit (or should not live) in HBM:
  the reference's ``NumpyDataLoader`` cache refills (data/core.py:664-791,
  numpy_loader.py:342-389) with the ``io_callback`` replaced by pinned-memory
  double buffering -- while the chains consume the cached block, the next block
  of ``cache_size`` minibatches is gathered on the host into page-locked memory
  and copied to the device on a second stream.

  Index streams are ``NumpyDataLoader``'s (same PCG64 draws).  All chains of a
  solver call share ONE stream of minibatches (``shared`` = the first chain's
  pipeline): shipping a separate batch per chain would multiply the host link
  traffic by the number of chains."""

  def __init__(self, copy=True, **reference_data):
    del copy
    assert len(reference_data) > 0, "Observations are required."
    counts = {int(np.shape(v)[0]) for v in reference_data.values()}
    assert len(counts) == 1, "All arrays must have the same leading dimension."
    self._observation_count = counts.pop()
    self.host_data = {k: np.ascontiguousarray(v, dtype=np.float32)
                      for k, v in reference_data.items()}
    self.host_shapes = {k: v.shape for k, v in self.host_data.items()}
    self.device_data = {}            # nothing is resident
    self._chains: List[dict] = []
    self.upload_comm = None          # see shard_upload
    self.pull = False                # True: the GPU reads the rows over the host link itself
    self._mapped = {}

  def mapped(self, name: str) -> int:
    """Device-visible address of ``host_data[name]``: the array's pages are locked and
    mapped in place on first use (cudaHostRegister, no copy), so kernels can read
    minibatch rows straight out of host memory (sgmc_pull_rows)."""
    if name not in self._mapped:
      self._mapped[name] = ops.host_register(self.host_data[name])
    return self._mapped[name]

  def __del__(self):
    for name in list(getattr(self, "_mapped", {})):
      try:
        ops.host_unregister(self.host_data[name])
      except Exception:       # interpreter shutdown / context already gone
        pass
      self._mapped.pop(name, None)

  def shard_upload(self, comm):
    """Chain-sharded multi-GPU runs: every rank consumes the same minibatches, so
    each rank gathers / uploads only rows [rank n / R, (rank + 1) n / R) of a batch
    and the rows are all-gathered over NVLink on the device (``comm``: a
    ``dist.NcclCommunicator``).  The host link then carries 1 / R of the batch per
    rank instead of R identical copies through one host."""
    self.upload_comm = comm

  def absmax(self, name: str) -> float:
    cache = self.__dict__.setdefault("_absmax", {})
    if name not in cache:
      cache[name] = float(np.abs(self.host_data[name]).max())
    return cache[name]

  @property
  def _format(self):
    return {k: ((), v.shape[1:]) for k, v in self.host_data.items()}


class _StreamSlot:
  """One cached block: pinned staging + device copy of ``cache`` minibatches."""

  def __init__(self, loader: StreamingNumpyDataLoader, cache: int, n: int):
    from .device import Event
    from .io import _pinned_array
    self.host, self.dev, self.addr = {}, {}, {}
    for k, v in loader.host_data.items():
      row = int(np.prod(v.shape[1:], dtype=np.int64))
      arr, addr = _pinned_array(cache * n * row)
      self.host[k] = arr.reshape((cache * n,) + v.shape[1:])
      self.addr[k] = addr
      self.dev[k] = DeviceArray((cache, n) + v.shape[1:], np.float32)
    self.copied = Event()
    self.consumed = Event()
    self.host_idx = None


class CacheState:
  """State of the random-batch functional (data/core.py:240-263).

  device mode: ``keys`` ping-pong pair of ``uint32[2]`` device buffers.
  host mode  : per-chain ids, the cached index block ``int32[C, cache, n]`` on
  the device and the current line in the cache.
  """

  def __init__(self, mode, **fields):
    self.mode = mode
    self.__dict__.update(fields)


def random_reference_data(data_loader: DataLoader, cached_batches_count: int,
                          mb_size: int, verify_calls: bool = False):
  """data/core.py:477-520 -> ``(init_fn, get_fn, release_fn)``."""
  del verify_calls
  N = data_loader.static_information["observation_count"]
  if N < mb_size:
    raise ValueError(f"Batch size cannot be bigger than the number of total "
                     f"observations. Got {N} and {mb_size}.")
  if cached_batches_count <= 0 or mb_size <= 0:
    raise ValueError(f"Cache size and batch size must be positive, got"
                     f"{cached_batches_count} and {mb_size}.")
  info = MiniBatchInformation(observation_count=N, mask=None, batch_size=mb_size)

  if isinstance(data_loader, DeviceNumpyDataLoader):
    if cached_batches_count != 1:
      raise ValueError("No caching on device.")

    def init_fn(**kwargs) -> CacheState:
      key = data_loader.init_random_data(**kwargs)
      return CacheState("device", keys=[DeviceArray.from_numpy(key),
                                        DeviceArray((2,), np.uint32)],
                        flip=0, idx=DeviceArray((mb_size,), np.int32),
                        host_key=np.asarray(key, np.uint32))

    def get_fn(state: CacheState, information: bool = False):
      ops.minibatch_draw(state.keys[state.flip], state.keys[1 - state.flip],
                         state.idx, N)
      state.flip = 1 - state.flip
      batch = BatchRef(data_loader, state.idx, mb_size)
      return (state, (batch, info)) if information else (state, batch)

    def scan_source(state: CacheState, steps: int):
      """What a native scan needs to draw the next ``steps`` minibatches itself."""
      src = {"loader": data_loader, "n": mb_size, "N": N, "idx_all": None,
             "key_a": state.keys[state.flip], "key_b": state.keys[1 - state.flip],
             "idx_buf": state.idx}
      state.flip = (state.flip + steps) % 2          # the keys ping-pong once per draw
      return src

    get_fn.scan_source = scan_source
    return init_fn, get_fn, lambda: None

  if isinstance(data_loader, StreamingNumpyDataLoader):
    import ctypes as C
    from . import _lib
    from .device import Stream, current_stream

    def init_fn(**kwargs) -> CacheState:
      chain_id = data_loader.register_random_pipeline(
          cached_batches_count, mb_size, **kwargs)
      return CacheState("stream", chain_ids=[chain_id], line=cached_batches_count,
                        cache_size=cached_batches_count, slots=None, cur=1,
                        copy_stream=None, staged=False)

    def _stage(state: CacheState, slot: "_StreamSlot"):
      """Draw the next block of indices, gather its rows into the slot's pinned
      memory and enqueue the H2D copy on the copy stream."""
      idx, _ = data_loader.get_indices(state.chain_ids[0])        # [cache, n]
      slot.host_idx = idx
      flat = idx.reshape(-1)
      # host: the previous copy OUT of this pinned buffer has finished (the host
      # runs ahead of the device); device: the chains are done with the slot
      slot.copied.sync()
      state.copy_stream.wait_event(slot.consumed)
      for k, v in data_loader.host_data.items():
        np.take(v, flat, axis=0, out=slot.host[k])
        _lib.call("sgmc_memcpy_h2d", C.c_void_p(slot.dev[k].ptr), C.c_void_p(slot.addr[k]),
                  slot.dev[k].nbytes, state.copy_stream.handle)
      slot.copied.record(state.copy_stream)

    def get_fn(state: CacheState, information: bool = False):
      main = current_stream()
      if state.slots is None:
        state.copy_stream = Stream.create()
        state.slots = [_StreamSlot(data_loader, cached_batches_count, mb_size)
                       for _ in range(2)]
        for sl in state.slots:
          sl.consumed.record(main)
        _stage(state, state.slots[0])
      if state.line == state.cache_size:               # block exhausted: switch slots
        old = state.slots[state.cur]
        old.consumed.record(main)
        state.cur = 1 - state.cur
        main.wait_event(state.slots[state.cur].copied)
        state.line = 0
        _stage(state, old)                             # prefetch the block after this one
      slot = state.slots[state.cur]
      rows = {k: v.row_slice(state.line, state.line + 1).reshape((mb_size,) + v.shape[2:])
              for k, v in slot.dev.items()}
      batch = BatchRef(data_loader, None, mb_size, rows=rows,
                       host_idx=slot.host_idx[state.line][None])
      state.line += 1
      return (state, (batch, info)) if information else (state, batch)

    def scan_source(state: CacheState, steps: int):
      """What the native host scan needs: the loader, and a function handing out
      the index rows of the next k minibatches (the chain's PCG64 pipeline, block by
      block as get_fn would draw them).  Only before get_fn started streaming."""
      del steps
      if state.slots is not None or len(state.chain_ids) != 1:
        return None
      left = []

      def draw(k: int) -> np.ndarray:
        rows, need = [], k
        while need > 0:
          if not left:
            blk, _ = data_loader.get_indices(state.chain_ids[0])     # [cache, n]
            left.extend(np.asarray(blk, np.int32))
          take = min(need, len(left))
          rows.extend(left[:take])
          del left[:take]
          need -= take
        return np.stack(rows)

      return {"loader": data_loader, "n": mb_size, "N": N, "host_stream": True,
              "draw": draw, "chunk": cached_batches_count}

    get_fn.scan_source = scan_source
    return init_fn, get_fn, lambda: None

  if isinstance(data_loader, NumpyDataLoader):
    def init_fn(**kwargs) -> CacheState:
      chain_id = data_loader.register_random_pipeline(
          cached_batches_count, mb_size, **kwargs)
      return CacheState("host", chain_ids=[chain_id], cache=None, line=0,
                        cache_size=cached_batches_count)

    def _refill(state: CacheState):
      blocks = [data_loader.get_indices(c)[0] for c in state.chain_ids]
      state.host_cache = np.stack(blocks).astype(np.int32)        # [C, cache, n]
      state.cache = DeviceArray.from_numpy(
          np.ascontiguousarray(state.host_cache.transpose(1, 0, 2)))  # [cache, C, n]
      state.line = 0

    def get_fn(state: CacheState, information: bool = False):
      if state.cache is None or state.line == state.cache_size:
        _refill(state)
      rows = state.cache.row_slice(state.line, state.line + 1)
      C = len(state.chain_ids)
      idx = rows.reshape(mb_size) if C == 1 else rows.reshape(C, mb_size)
      batch = BatchRef(data_loader, idx, mb_size,
                       host_idx=state.host_cache[:, state.line])
      state.line += 1
      return (state, (batch, info)) if information else (state, batch)

    def scan_source(state: CacheState, steps: int):
      """The index rows of the next ``steps`` minibatches (one chain): the rest of
      the current cache block, then freshly drawn blocks -- the sequence get_fn
      would hand out -- as one device array int32[steps][n]."""
      if len(state.chain_ids) != 1:
        return None                        # one index stream per chain: step loop
      rows, need = [], steps
      while need > 0:
        if state.cache is None or state.line == state.cache_size:
          _refill(state)
        take = min(need, state.cache_size - state.line)
        rows.append(state.host_cache[0, state.line:state.line + take])
        state.line += take
        need -= take
      idx_all = DeviceArray.from_numpy(np.concatenate(rows).astype(np.int32))
      return {"loader": data_loader, "n": mb_size, "N": N, "idx_all": idx_all}

    get_fn.scan_source = scan_source
    return init_fn, get_fn, lambda: None

  raise TypeError("The DataLoader must inherit from HostDataLoader or "
                  "DeviceDataLoader")


def merge_cache_states(states: List[CacheState]) -> CacheState:
  """Stack the per-chain data states of C chains (list_vmap equivalent).

  Device mode requires identical keys (the reference default: every chain
  starts from PRNGKey(0), numpy_loader.py:124) so that one minibatch is shared.
  """
  first = states[0]
  if first.mode == "stream":      # one shared stream of minibatches for all chains
    return first
  if first.mode == "device":
    for s in states[1:]:
      if not np.array_equal(s.host_key, first.host_key):
        raise NotImplementedError(
            "chains with different device data keys are not batched; run them "
            "as separate solver calls")
    return first
  merged = CacheState("host", chain_ids=[c for s in states for c in s.chain_ids],
                      cache=None, line=0, cache_size=first.cache_size)
  return merged


def full_reference_data(data_loader: DataLoader, cached_batches_count: int = 100,
                        mb_size: int = None):
  """data/core.py:522-627: map a function over the whole data set in batches.

  Returns ``(init_fn, map_fn, release_fn)``; ``map_fn(fun, data_state, carry,
  masking=True, information=True)`` calls ``fun(batch, mask, carry)`` for every
  batch; the last batch wraps indices modulo N and masks the overhang
  (core.py:571-572, :944-945).  Quirk kept: a device loader uses
  ``cached_batches_count`` as the batch size (core.py:552-553).
  """
  N = data_loader.static_information["observation_count"]
  if isinstance(data_loader, StreamingNumpyDataLoader):
    raise NotImplementedError("full-data passes over a streamed (host-resident) data set "
                              "are not implemented; use NumpyDataLoader (HBM-resident)")
  if isinstance(data_loader, DeviceNumpyDataLoader):
    mb_size = cached_batches_count
  if mb_size is None or mb_size <= 0:
    raise ValueError("mb_size must be positive")
  n_batches = math.ceil(N / mb_size)
  info = MiniBatchInformation(observation_count=N, mask=None, batch_size=mb_size)
  ids = np.arange(n_batches * mb_size).reshape(n_batches, mb_size)
  d_idx = DeviceArray.from_numpy(np.mod(ids, N).astype(np.int32))
  d_mask = DeviceArray.from_numpy((ids < N).astype(np.float32))

  def init_fn(**kwargs):
    del kwargs
    return CacheState("full")

  def map_fn(fun, data_state, carry, masking: bool = False,
             information: bool = False):
    results = []
    for b in range(n_batches):
      idx = d_idx.row_slice(b, b + 1).reshape(mb_size)
      mask = d_mask.row_slice(b, b + 1).reshape(mb_size)
      batch = BatchRef(data_loader, idx, mb_size, mask=mask)
      arg = (batch, info) if information else batch
      if masking:
        res, carry = fun(arg, mask, carry)
      else:
        res, carry = fun(arg, carry)
      results.append(res)
    return data_state, (results, carry)

  # lets potential.full_potential run all batches inside one C call
  map_fn.loader, map_fn.mb_size = data_loader, mb_size
  return init_fn, map_fn, lambda: None
