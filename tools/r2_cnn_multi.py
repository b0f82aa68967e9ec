"""C5 record of the bench under torchrun, alone (tool): CNN potential + gradient with the
minibatch rows sharded over the ranks."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jax_sgmc_b200 import _lib, device, dist  # noqa: E402
from jax_sgmc_b200.device import Stream  # noqa: E402


class A:
  pass


rank, world, local = bench.dist_env()
ctl = bench.Control(rank, world)
_lib.load()
device.set_device(local)
s = Stream.create()
device.set_current_stream(s)
nccl = dist.NcclCommunicator.from_control_plane(ctl) if world > 1 else None
for shard in ([False, True] if world > 1 else [False]):
  rec = bench.bench_cnn(A(), ctl, nccl if shard else None, s)
  if rank == 0:
    print("sharded" if shard else "replicated", json.dumps({k: rec[k] for k in ("n_gpus", "rows_per_rank", "us_per_evaluation")}), flush=True)
ctl.close()
