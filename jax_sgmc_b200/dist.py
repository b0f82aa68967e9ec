"""Multi-GPU plumbing: one process per GPU.

The data path needs exactly one collective -- the all-gather of per-replica
``(U, var)`` scalars in sharded reSGLD -- issued through NCCL by the C ABI
(``sgmc_nccl_allgather``) on the compute stream.  Everything else (chains
sharded over ranks) is communication free.  ``SocketCommunicator`` is the
host-side control plane (plain TCP, no framework: the package imports neither
torch nor jax): it distributes the NCCL unique id and the peer-memory IPC
handles and offers the same ``allgather`` interface on host arrays, so the
host-side logic is testable on CPU boxes with world_size 2 (the tests also run
it over ``torch.distributed`` gloo through ``tests/_gloo_comm.py``).
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


class SocketCommunicator:
  """Host-side control plane without any framework: a star of TCP connections
  (rank 0 listens on ``MASTER_ADDR``:``SGMC_CONTROL_PORT``, default
  ``MASTER_PORT`` + 29; every other rank connects once).  Small host payloads
  only -- the NCCL unique id, peer-memory IPC handles, barriers, max-over-ranks
  of a timing -- never the data path.  Same interface as the communicators
  below (``allgather`` on NumPy arrays, ``broadcast_bytes``, ``barrier``)."""

  def __init__(self, rank: Optional[int] = None, world: Optional[int] = None,
               addr: Optional[str] = None, port: Optional[int] = None, timeout: float = 120.0):
    import socket
    import time
    er, ew, _ = env_rank_world()
    self.rank = er if rank is None else int(rank)
    self.world = ew if world is None else int(world)
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    if port is None:
      port = int(os.environ.get("SGMC_CONTROL_PORT",
                                int(os.environ.get("MASTER_PORT", "29500")) + 29))
    self._peers = {}
    if self.world == 1:
      return
    if self.rank == 0:
      srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
      srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
      srv.bind((addr, port))
      srv.listen(self.world)
      srv.settimeout(timeout)
      for _ in range(self.world - 1):
        conn, _ = srv.accept()
        conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        r = int.from_bytes(self._recv_exact(conn, 4), "little")
        self._peers[r] = conn
      srv.close()
    else:
      deadline = time.time() + timeout
      while True:
        try:
          conn = socket.create_connection((addr, port), timeout=timeout)
          break
        except OSError:
          if time.time() > deadline:
            raise
          time.sleep(0.05)
      conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
      conn.sendall(self.rank.to_bytes(4, "little"))
      self._peers[0] = conn

  @staticmethod
  def _recv_exact(conn, n: int) -> bytes:
    chunks, got = [], 0
    while got < n:
      c = conn.recv(n - got)
      if not c:
        raise ConnectionError("control plane peer closed the connection")
      chunks.append(c)
      got += len(c)
    return b"".join(chunks)

  def _send_msg(self, conn, payload: bytes):
    conn.sendall(len(payload).to_bytes(8, "little") + payload)

  def _recv_msg(self, conn) -> bytes:
    n = int.from_bytes(self._recv_exact(conn, 8), "little")
    return self._recv_exact(conn, n)

  def allgather_bytes(self, payload: bytes):
    """Every rank contributes ``payload``; returns the list ordered by rank."""
    if self.world == 1:
      return [payload]
    if self.rank == 0:
      parts = [payload] + [None] * (self.world - 1)
      for r, conn in self._peers.items():
        parts[r] = self._recv_msg(conn)
      blob = b"".join(len(p).to_bytes(8, "little") + p for p in parts)
      for conn in self._peers.values():
        self._send_msg(conn, blob)
      return parts
    conn = self._peers[0]
    self._send_msg(conn, payload)
    blob, parts, off = self._recv_msg(conn), [], 0
    for _ in range(self.world):
      n = int.from_bytes(blob[off:off + 8], "little")
      parts.append(blob[off + 8:off + 8 + n])
      off += 8 + n
    return parts

  def allgather(self, send: np.ndarray, recv: np.ndarray, stream=None) -> np.ndarray:
    del stream
    send = np.ascontiguousarray(send)
    parts = self.allgather_bytes(send.tobytes())
    recv[...] = np.stack([np.frombuffer(p, send.dtype).reshape(send.shape)
                          for p in parts]).reshape(recv.shape)
    return recv

  def broadcast_bytes(self, payload: Optional[bytes], nbytes: int, src: int = 0) -> bytes:
    parts = self.allgather_bytes(payload if self.rank == src else b"")
    assert len(parts[src]) == nbytes
    return parts[src]

  def barrier(self):
    self.allgather_bytes(b"")

  def max(self, x: float) -> float:
    parts = self.allgather_bytes(np.float64(x).tobytes())
    return float(max(np.frombuffer(p, np.float64)[0] for p in parts))

  def sum(self, x: float) -> float:
    parts = self.allgather_bytes(np.float64(x).tobytes())
    return float(sum(np.frombuffer(p, np.float64)[0] for p in parts))

  def close(self):
    for conn in self._peers.values():
      try:
        conn.close()
      except OSError:
        pass
    self._peers = {}


class PeerCommunicator:
  """All-gather of a tiny payload through peer memory (NVLink / NVSwitch
  stores + flags, ``csrc/p2p_exchange.cu``) -- the latency path of the reSGLD
  exchange; no NCCL involved.  The windows' IPC handles travel over the gloo
  control plane once."""

  def __init__(self, ctl, bytes_per_rank: int):
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
  def from_control_plane(cls, ctl) -> "NcclCommunicator":
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
