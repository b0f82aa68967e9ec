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


def test_world_size_2_socket_control_plane():
  """The same worker over the package's own control plane (plain TCP, no torch in
  the product package): two processes started directly, RANK / WORLD_SIZE /
  MASTER_* in the environment as torchrun would set them."""
  port = _free_port()
  procs = []
  for r in range(2):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(r),
               WORLD_SIZE="2", LOCAL_RANK=str(r), SGMC_TEST_COMM="socket",
               SGMC_CONTROL_PORT=str(port), OMP_NUM_THREADS="1")
    procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py")],
                                  env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                  text=True))
  outs = [p.communicate(timeout=280)[0] for p in procs]
  for r, (p, o) in enumerate(zip(procs, outs)):
    assert p.returncode == 0, o[-3000:]
    assert f"rank {r} ok" in o


def test_package_does_not_import_torch():
  """north star: no PyTorch in the product.  Importing every module of the package
  must not pull torch (or jax) in."""
  code = ("import sys, pkgutil, importlib, jax_sgmc_b200 as p\n"
          "for m in pkgutil.iter_modules(p.__path__):\n"
          "    if m.name != 'build': importlib.import_module('jax_sgmc_b200.' + m.name)\n"
          "assert 'torch' not in sys.modules and 'jax' not in sys.modules, "
          "[k for k in sys.modules if k.startswith(('torch', 'jax.'))][:5]\n")
  out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True,
                       timeout=120)
  assert out.returncode == 0, out.stdout + out.stderr


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
