"""N>1 host logic on CPU: two processes over torch.distributed (gloo)."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def test_world_size_2_gloo():
  env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
         "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", str(_free_port()),
         os.path.join(ROOT, "tests", "_gloo_worker.py")]
  out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280)
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
  assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


def test_shard_range_balanced():
  sys.path.insert(0, ROOT)
  from jax_sgmc_b200 import dist
  for total in (0, 1, 7, 8, 4096, 4097):
    for world in (1, 2, 3, 8):
      ranges = [dist.shard_range(total, r, world) for r in range(world)]
      assert ranges[0][0] == 0 and ranges[-1][1] == total
      assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
      sizes = [b - a for a, b in ranges]
      assert max(sizes) - min(sizes) <= 1
