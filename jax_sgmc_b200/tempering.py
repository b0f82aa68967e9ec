"""reSGLD with replicas sharded over GPUs (BASELINE.json configs[3]).

``sharded_tempering`` generalises ``solver.parallel_tempering`` (reference
solver.py:220-299) to a ladder of R temperatures, replica r living on GPU
``r * world / R``.  Per step every rank updates its local replicas with the
fused Langevin kernels, then ONE all-gather (NCCL over NVLink) shares the
per-replica ``(U, var)`` rows and every rank evaluates the identical swap
decisions on the device (``sgmc_resgld_ladder_step``) from identical keys.
Replicas exchange temperature labels, not parameters, so nothing else crosses
NVLink; the reference exchanges whole chain states (solver.py:287-291), which
is the same Markov chain up to relabelling.  For R = 2 the decisions, keys and
cold-chain samples reproduce the reference (tests/test_gpu_tempering.py).

``overlap_exchange=True`` moves the all-gather and the decision kernels to a
second stream: the swap of step k only has to be known when step k+1 *updates*
(the temperature enters the noise scale, not the potential), so it runs
concurrently with step k+1's minibatch draw and potential / gradient kernels.
Same results bit for bit.  Measured on 8 B200: 96.9 us per step
against 111.7 us with the exchange on the sampling stream; ``bench.py`` turns it on for
N > 1.  It stays an argument (default off) because it costs a second stream and an event
per step on a single GPU, where there is no latency to hide.
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import numpy as np

from . import ops
from .device import DeviceArray, Event, Stream, current_stream
from .dist import LocalCommunicator, shard_range
from .integrator import KeyState, _as_chain_tree


class _WaitFor:
  """pre_update_hook of langevin_diffusion: order the update after ``event``."""

  def __init__(self, stream, event):
    self.stream, self.event = stream, event

  def __call__(self):
    self.stream.wait_event(self.event)


class ShardedTemperingState:
  """``exchange`` / ``temp_index`` are written by the exchange stream: reading
  them through these properties first waits for the pending exchange."""

  def __init__(self, exchange, temp_index, **kw):
    self.__dict__.update(kw)
    self._exchange, self._temp_index = exchange, temp_index
    self.pending: Event = None        # end of the exchange in flight (overlap mode)

  def wait(self):
    if self.pending is not None:
      self.pending.sync()

  @property
  def exchange(self):
    self.wait()
    return self._exchange

  @property
  def temp_index(self):
    self.wait()
    return self._temp_index


def sharded_tempering(integrator, temperatures: Sequence[float], comm=None,
                      overlap_exchange: bool = False):
  """Returns ``(init, update, get)`` like the reference's solvers.

  ``init(samples, ssq_init=0.0, key=PRNGKey(0), F=1.0, **kw)``: ``samples`` is
  a list of R per-replica initial samples (each a ChainTree / list of B host
  pytrees), or one such sample used for every replica.
  ``update(state, schedule)``: one reSGLD step.  The ladder's temperatures are the
  ``temperatures`` argument; ``schedule.temperature`` must be 1.0 (as in ``alias.re_sgld``,
  alias.py:185) -- the fused update reads the per-chain temperature labels, not the
  schedule's scalar.  ``comm``: ``LocalCommunicator`` (one process), ``NcclCommunicator`` or
  ``PeerCommunicator``; host-side communicators are rejected.
  """
  comm = comm or LocalCommunicator()
  temps_host = np.asarray(temperatures, np.float32)
  R = int(temps_host.size)
  assert R >= 2 and R % comm.world == 0, "replicas must divide evenly over ranks"
  r0, r1 = shard_range(R, comm.rank, comm.world)
  n_local = r1 - r0
  init_integrator, update_integrator, get_integrator = integrator
  xstream = Stream.create() if overlap_exchange else None
  ready, done = Event(), Event()       # (U, var) snapshot taken / exchange finished

  def init(samples, ssq_init=0.0, key=None, F=1.0, **kwargs):
    if not (isinstance(samples, (list, tuple)) and len(samples) == R
            and not isinstance(samples[0], dict)):
      samples = [samples] * R
    trees = [_as_chain_tree(s) for s in samples]
    B = trees[0].n_chains
    key = ops.prng_key(0) if key is None else np.asarray(key, np.uint32)
    if key.ndim == 1:
      key = np.tile(key, (B, 1))
    # key, split_1..split_R = split(key, R + 1): for R = 2 this is the
    # reference's 3-way split (solver.py:254)
    ks = ops.split(DeviceArray.from_numpy(key), R + 1).numpy()          # [B, R+1, 2]
    uv = DeviceArray.zeros((n_local, 2, B))                              # (U, var) rows
    replicas = []
    for l, r in enumerate(range(r0, r1)):
      st = init_integrator(trees[r].copy() if r > r0 and trees[r] is trees[r0]
                           else trees[r], key=ks[:, 1 + r], **kwargs)
      row = uv.row_slice(l, l + 1).reshape(2, B)
      st = st._replace(potential=row.row_slice(0, 1).reshape(B),
                       variance=row.row_slice(1, 2).reshape(B))
      replicas.append(st)
    if R == 2:
      pair_keys = ks[:, 0][None]                                         # [1, B, 2]
    else:   # extension: one exchange stream per neighbouring pair
      pair_keys = np.ascontiguousarray(
          ops.split(DeviceArray.from_numpy(ks[:, 0]), R - 1).numpy().transpose(1, 0, 2))
    holder = np.tile(np.arange(R, dtype=np.int32)[:, None], (1, B))
    t_local = np.tile(temps_host[r0:r1, None], (1, B)).astype(np.float32)
    return ShardedTemperingState(
        replicas=replicas, uv=uv, gathered=DeviceArray.zeros((R, 2, B)),
        holder=DeviceArray.from_numpy(holder),
        ssq=DeviceArray.full((R - 1, B), float(ssq_init)),
        F=DeviceArray.full((B,), float(F)), temps=DeviceArray.from_numpy(temps_host),
        keys=KeyState(pair_keys.reshape(-1, 2)),
        exchange=DeviceArray.zeros((R - 1, B), np.int32),
        temp_per_chain=DeviceArray.from_numpy(t_local),
        temp_index=DeviceArray.from_numpy(
            np.tile(np.arange(r0, r1, dtype=np.int32)[:, None], (1, B))),
        uv_send=DeviceArray.zeros((n_local, 2, B)), step=0, B=B)

  comm_handle = getattr(comm, "_comm", None)
  native = isinstance(comm, LocalCommunicator) or comm_handle is not None
  if not native and not hasattr(comm, "timeouts"):
    raise TypeError("sharded_tempering needs a device-side communicator (LocalCommunicator, "
                    "NcclCommunicator or PeerCommunicator)")

  def update(state: ShardedTemperingState, schedule, *unused):
    if float(schedule.temperature) != 1.0:
      raise ValueError("sharded_tempering: the ladder's temperatures are fixed at build time; "
                       "schedule.temperature must be 1.0")
    state.step += 1
    B = state.B
    main = current_stream()
    hook = None
    if overlap_exchange and state.pending is not None:
      # the previous exchange must have written the labels before this step's
      # UPDATE kernel reads them; draw / potential / gradient do not depend on it
      hook = _WaitFor(main, done)
    for l in range(n_local):
      t_row = state.temp_per_chain.row_slice(l, l + 1).reshape(B)
      # labels are exchanged, never parameters: the replica's sample is only written by
      # its own updates (with one replica per rank the operand form is carried)
      state.replicas[l] = update_integrator(state.replicas[l], schedule,
                                            temp_per_chain=t_row,
                                            pre_update_hook=hook if l == 0 else None,
                                            carry_ok=n_local == 1)
    if native:
      # snapshot + all-gather + decision kernels: one C call
      ops.resgld_sharded_exchange(
          comm_handle, main, xstream, ready if overlap_exchange else None,
          done if overlap_exchange else None, state.uv, state.uv_send, state.gathered,
          state.holder, state.ssq, state.F, state.temps, state.keys.current,
          state.keys.next, state._exchange, R, B, state.step, r0, n_local,
          state.temp_per_chain, state._temp_index)
    else:
      # peer-memory all-gather (returns a view of the rows in its window) or a
      # host-side communicator, then the decision kernels
      assert not overlap_exchange, "overlap needs the NCCL / local communicator"
      gathered = comm.allgather(state.uv, state.gathered)
      if state.step % 256 == 0 and comm.timeouts():
        # a peer's rows did not arrive within the bounded spin: the decisions of that step
        # were taken on stale energies -- stop instead of sampling a different chain
        raise RuntimeError(f"peer-memory exchange timed out {comm.timeouts()} time(s)")
      ops.resgld_ladder_step(gathered, state.holder, state.ssq, state.F,
                             state.temps, state.keys.current, state.keys.next,
                             state._exchange, R, B, state.step, r0, n_local,
                             state.temp_per_chain, state._temp_index)
    state.keys.flip()
    if overlap_exchange:
      state.pending = done
    return state, None

  def get(state: ShardedTemperingState) -> Dict[str, Any]:
    """Local replicas' variables plus the temperature index of every system
    (index 0 marks the sample of the un-tempered chain, solver.py:296-297)."""
    if state.pending is not None:      # order later reads on the sampling stream
      current_stream().wait_event(state.pending)
    return {"variables": [get_integrator(s)["variables"] for s in state.replicas],
            "likelihood": [get_integrator(s)["likelihood"] for s in state.replicas],
            "temperature_index": state._temp_index, "model_state": None}

  return init, update, get


def cold_samples(sample: Dict[str, Any]) -> np.ndarray:
  """Host helper (world_size 1): rows of the replicas currently holding the
  lowest temperature, ``f32[B, P]``."""
  tidx = sample["temperature_index"].numpy()                       # [n_local, B]
  flats = np.stack([v.flat.numpy() for v in sample["variables"]])  # [n_local, B, P]
  sel = np.argmin(tidx, axis=0)
  return flats[sel, np.arange(flats.shape[1])]
