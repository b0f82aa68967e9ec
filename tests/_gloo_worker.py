"""world_size-2 worker for tests/test_dist_gloo.py (CPU, gloo backend).

Exercises the N>1 host logic of the sharded paths without a GPU:
* chain sharding (`dist.shard_range`) covers all chains exactly once;
* the replica all-gather plumbing (`GlooCommunicator.allgather`) delivers the
  rank-major [R][2][B] layout the ladder kernel expects;
* the replicated-decision protocol: every rank, evaluating the swap rule on
  the gathered values with the same keys (here through the oracle's restatement
  of solver.py:273-291 standing in for the CUDA kernel), reaches identical
  decisions and label tables;
* the NCCL unique-id broadcast path (`broadcast_bytes`);
* bench.py's max-over-ranks reduction.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from jax_sgmc_b200 import dist  # noqa: E402
from oracle import prng  # noqa: E402
from oracle import sgmc as osgmc  # noqa: E402


def main():
  # SGMC_TEST_COMM=socket: the package's own framework-free control plane;
  # otherwise torch.distributed (gloo) through the test-side shim
  use_socket = os.environ.get("SGMC_TEST_COMM") == "socket"
  if use_socket:
    comm = dist.SocketCommunicator()
  else:
    from _gloo_comm import GlooCommunicator
    comm = GlooCommunicator()
  rank, world = comm.rank, comm.world
  assert world == 2

  # chain sharding
  lo, hi = dist.shard_range(4097, rank, world)
  sizes = np.zeros((world, 2), np.int64)
  comm.allgather(np.array([lo, hi], np.int64), sizes)
  assert sizes[0, 0] == 0 and sizes[-1, 1] == 4097 and sizes[0, 1] == sizes[1, 0]

  # replica all-gather: R = 4 replicas, 2 per rank, B systems
  R, B = 4, 5
  r0, r1 = dist.shard_range(R, rank, world)
  rng = np.random.default_rng(100 + rank)
  local = rng.standard_normal((r1 - r0, 2, B)).astype(np.float32)
  local[:, 1] = np.abs(local[:, 1])
  gathered = np.zeros((R, 2, B), np.float32)
  comm.allgather(local, gathered)
  assert np.array_equal(gathered[r0:r1], local)
  other = np.random.default_rng(100 + (1 - rank)).standard_normal((2, 2, B)).astype(np.float32)
  other[:, 1] = np.abs(other[:, 1])
  o0 = dist.shard_range(R, 1 - rank, world)[0]
  assert np.array_equal(gathered[o0:o0 + 2], other)

  # replicated decisions: same inputs + same keys on every rank -> same table
  temps = np.array([1.0, 3.0, 10.0, 30.0], np.float32)
  holder = np.tile(np.arange(R)[:, None], (1, B))
  keys = np.stack([prng.PRNGKey(7 + b) for b in range(B)])
  ssq = np.zeros(B, np.float32)
  for step in (1, 2, 3):
    for p in range(R - 1):
      if p % 2 != step % 2:
        continue
      lo_r, hi_r = holder[p], holder[p + 1]
      U_n = gathered[lo_r, 0, np.arange(B)]
      var_n = gathered[lo_r, 1, np.arange(B)]
      U_h = gathered[hi_r, 0, np.arange(B)]
      ex, ssq, keys, _, _ = osgmc.resgld_swap_decision(
          U_n, U_h, var_n, ssq, np.ones(B, np.float32), step, temps[p], temps[p + 1], keys)
      swapped = holder.copy()
      swapped[p, ex], swapped[p + 1, ex] = holder[p + 1, ex], holder[p, ex]
      holder = swapped
  tables = np.zeros((world,) + holder.shape, np.int64)
  comm.allgather(holder.astype(np.int64), tables)
  assert np.array_equal(tables[0], tables[1]), "ranks disagree on the label table"
  assert sorted(holder[:, 0].tolist()) == list(range(R))       # still a permutation

  # unique-id style broadcast
  payload = bytes(range(128)) if rank == 0 else None
  got = comm.broadcast_bytes(payload, 128, 0)
  assert got == bytes(range(128))

  # max over ranks as bench.py does it
  if use_socket:
    assert comm.max(float(rank + 1)) == 2.0 and comm.sum(float(rank + 1)) == 3.0
    comm.barrier()
    comm.close()
  else:
    import torch
    import torch.distributed as td
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    assert float(t[0]) == 2.0
    comm.barrier()
    td.destroy_process_group()
  print(f"rank {rank} ok")


if __name__ == "__main__":
  main()
