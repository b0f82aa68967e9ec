"""In-tree build of libsgmc_b200.so (nvcc, sm_100a only).

Run as ``python -m jax_sgmc_b200.build`` or through ``__graft_entry__.build()``.
The shared object is written next to this file (``_C/libsgmc_b200.so``) so it
travels with the source tree; it is git-ignored.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libsgmc_b200.so")

SOURCES = ["runtime.cu", "prng_kernels.cu", "update_kernels.cu", "glm_simt.cu",
           "glm_tc.cu", "resgld.cu", "nccl_shim.cu", "adaption_kernels.cu",
           "resgld_ladder.cu", "misc_kernels.cu", "tree_kernels.cu", "p2p_exchange.cu", "mlp.cu", "host_pull.cu", "fisher.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-std=c++17", "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
] + ([f"-DSGMC_TC_BK={os.environ['SGMC_TC_BK']}"] if os.environ.get("SGMC_TC_BK") else []) \
  + (["-DSGMC_TC_DEBUG"] if os.environ.get("SGMC_TC_DEBUG") else [])


def _nvcc() -> str:
  for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or
                 not os.path.isabs(cand)):
      return cand
  raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
  h = hashlib.sha256()
  for p in sorted(paths):
    with open(p, "rb") as f:
      h.update(p.encode())
      h.update(f.read())
  h.update(" ".join(NVCC_FLAGS).encode())
  return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
  os.makedirs(OUT_DIR, exist_ok=True)
  srcs = [os.path.join(CSRC, s) for s in SOURCES
          if os.path.exists(os.path.join(CSRC, s))]
  deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
      os.path.join(HERE, "..", "include", "sgmc_b200.h")]
  stamp = os.path.join(OUT_DIR, "build.stamp")
  digest = _digest(deps)
  if (not force and os.path.exists(LIB) and os.path.exists(stamp)
      and open(stamp).read().strip() == digest):
    return LIB
  objs = []
  procs = []
  for s in srcs:
    o = os.path.join(OUT_DIR, os.path.basename(s) + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", s, "-o", o]
    if verbose:
      cmd.insert(1, "-Xptxas=-v")
      print(" ".join(cmd), flush=True)
    procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True)))
    objs.append(o)
  failed = False
  for s, p in procs:
    out, _ = p.communicate()
    if p.returncode != 0:
      failed = True
      sys.stderr.write(f"nvcc failed for {s}:\n{out}\n")
    elif verbose and out:
      print(out)
  if failed:
    raise RuntimeError("libsgmc_b200 build failed")
  link = [_nvcc(), "-shared", "-o", LIB, *objs, "-lcudart", "-ldl",
          "-gencode", "arch=compute_100a,code=sm_100a"]
  if verbose:
    print(" ".join(link), flush=True)
  subprocess.check_call(link)
  with open(stamp, "w") as f:
    f.write(digest)
  return LIB


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
