"""``torch.distributed`` (gloo) flavour of the host control plane, for the CPU tests
only: same interface as ``jax_sgmc_b200.dist.SocketCommunicator``.  The product
package itself never imports torch."""
import os
from typing import Optional

import numpy as np


class GlooCommunicator:
  """Host arrays over torch.distributed (gloo)."""

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
