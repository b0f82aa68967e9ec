"""Multi-GPU plumbing: one process per GPU.

The data path needs exactly one collective -- the all-gather of per-replica
``(U, var)`` scalars in sharded reSGLD -- issued through NCCL by the C ABI
(``sgmc_nccl_allgather``) on the compute stream.  Everything else (chains
sharded over ranks) is communication free.  ``GlooCommunicator`` offers the
same interface on host arrays through ``torch.distributed`` (gloo) so the
host-side logic is testable on CPU boxes with world_size 2; the unique id of
the NCCL communicator is distributed over the same control plane.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .device import DeviceArray, current_stream, vp


def env_rank_world() -> Tuple[int, int, int]:
  return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
          int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
  """Contiguous, balanced [begin, end) slice of ``total`` items for ``rank``."""
  base, rem = divmod(total, world)
  begin = rank * base + min(rank, rem)
  return begin, begin + base + (1 if rank < rem else 0)


class LocalCommunicator:
  """world_size 1: the all-gather is the identity (no library involved)."""
  rank, world = 0, 1

  def allgather(self, send, recv, stream=None):
    if isinstance(send, DeviceArray):
      recv.copy_from(send, stream)
    else:
      recv[...] = np.asarray(send).reshape(recv.shape)
    return recv

  def barrier(self):
    pass


class GlooCommunicator:
  """Host arrays over torch.distributed (gloo): CPU tests / control plane."""

  def __init__(self, init: bool = True):
    import torch.distributed as dist
    self._dist = dist
    if init and not dist.is_initialized():
      os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
      dist.init_process_group("gloo")
    self.rank, self.world = dist.get_rank(), dist.get_world_size()

  def allgather(self, send: np.ndarray, recv: np.ndarray, stream=None) -> np.ndarray:
    import torch
    del stream
    t = torch.from_numpy(np.ascontiguousarray(send))
    outs = [torch.empty_like(t) for _ in range(self.world)]
    self._dist.all_gather(outs, t)
    recv[...] = np.stack([o.numpy() for o in outs]).reshape(recv.shape)
    return recv

  def broadcast_bytes(self, payload: Optional[bytes], nbytes: int, src: int = 0) -> bytes:
    import torch
    t = torch.zeros(nbytes, dtype=torch.uint8)
    if self.rank == src:
      t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    self._dist.broadcast(t, src)
    return bytes(t.numpy().tobytes())

  def barrier(self):
    self._dist.barrier()


class PeerCommunicator:
  """All-gather of a tiny payload through peer memory (NVLink / NVSwitch
  stores + flags, ``csrc/p2p_exchange.cu``) -- the latency path of the reSGLD
  exchange; no NCCL involved.  The windows' IPC handles travel over the gloo
  control plane once."""

  def __init__(self, ctl: GlooCommunicator, bytes_per_rank: int):
    assert bytes_per_rank % 16 == 0, "payload per rank must be a multiple of 16 bytes"
    self.rank, self.world, self.bytes = ctl.rank, ctl.world, int(bytes_per_rank)
    lib = _lib.load()
    nbytes = int(lib.sgmc_p2p_window_bytes(self.world, self.bytes))
    self.window = DeviceArray.zeros((nbytes,), np.uint8)
    current_stream().sync()
    handle = (C.c_char * 64)()
    _lib.call("sgmc_p2p_export", C.c_void_p(self.window.ptr), handle)
    handles = np.zeros((self.world, 64), np.uint8)
    ctl.allgather(np.frombuffer(handle.raw, np.uint8).copy(), handles)
    self._opened, ptrs = [], []
    for r in range(self.world):
      if r == self.rank:
        ptrs.append(self.window.ptr)
        continue
      p = C.c_void_p()
      buf = (C.c_char * 64).from_buffer_copy(handles[r].tobytes())
      _lib.call("sgmc_p2p_open", buf, C.byref(p))
      self._opened.append(p)
      ptrs.append(p.value)
    self.peer_table = DeviceArray.from_numpy(np.array(ptrs, np.uint64))
    self.seq = 0
    ctl.barrier()          # every window is mapped before anyone stores into it
    self._ctl = ctl

  def allgather(self, send: DeviceArray, recv: DeviceArray = None, stream=None) -> DeviceArray:
    """Returns a view of the gathered rows ([world] + send.shape) inside the
    local window; ``recv`` is ignored (nothing is copied)."""
    assert send.nbytes == self.bytes
    self.seq += 1
    s = (stream or current_stream()).handle
    _lib.call("sgmc_p2p_allgather", s, C.c_void_p(self.peer_table.ptr), self.rank,
              self.world, vp(send), self.bytes, self.seq)
    off = (self.seq & 1) * self.world * self.bytes
    return DeviceArray((self.world,) + tuple(send.shape), send.dtype,
                       ptr=self.window.ptr + off, owner=self.window)

  def timeouts(self) -> int:
    n = C.c_uint()
    _lib.call("sgmc_p2p_timeouts", C.byref(n))
    return int(n.value)

  def barrier(self):
    self._ctl.barrier()

  def __del__(self):
    for p in getattr(self, "_opened", []):
      try:
        _lib.call("sgmc_p2p_close", p)
      except Exception:
        pass


class NcclCommunicator:
  """Device buffers over NCCL / NVLink through the C ABI."""

  def __init__(self, rank: int, world: int, unique_id: bytes):
    assert len(unique_id) == 128
    self.rank, self.world = rank, world
    self._comm = C.c_void_p()
    buf = (C.c_char * 128).from_buffer_copy(unique_id)
    _lib.call("sgmc_nccl_init", C.byref(self._comm), buf, world, rank)

  @staticmethod
  def create_unique_id() -> bytes:
    buf = (C.c_char * 128)()
    _lib.call("sgmc_nccl_unique_id", buf)
    return bytes(buf.raw)

  @classmethod
  def from_control_plane(cls, ctl: GlooCommunicator) -> "NcclCommunicator":
    uid = cls.create_unique_id() if ctl.rank == 0 else None
    uid = ctl.broadcast_bytes(uid, 128, 0)
    return cls(ctl.rank, ctl.world, uid)

  def allgather(self, send: DeviceArray, recv: DeviceArray, stream=None) -> DeviceArray:
    s = (stream or current_stream()).handle
    _lib.call("sgmc_nccl_allgather", self._comm, s, vp(send), vp(recv), send.nbytes)
    return recv

  def allreduce_sum(self, send: DeviceArray, recv: DeviceArray, stream=None):
    s = (stream or current_stream()).handle
    _lib.call("sgmc_nccl_allreduce_sum_f32", self._comm, s, vp(send), vp(recv),
              send.size)
    return recv

  def barrier(self):
    pass

  def __del__(self):
    try:
      if self._comm.value:
        _lib.call("sgmc_nccl_destroy", self._comm)
    except Exception:
      pass
