"""ctypes loader of the oracle's C restatement (oracle/c/sgmc_oracle.c).

TEST INFRASTRUCTURE: cross-checks the NumPy oracle (hardware fmaf / directed
rounding vs the NumPy emulation) and is the multi-threaded CPU baseline of
bench.py.  Built by ``make -C oracle/c`` (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libsgmc_oracle.so")
_lib = None


def load(build_if_missing: bool = True):
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(_LIB) and build_if_missing:
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "c")])
  lib = C.CDLL(_LIB)
  lib.oracle_normal_like.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
  lib.oracle_normal_like.restype = None
  lib.oracle_sgld_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float,
                                   C.c_float]
  lib.oracle_sgld_step.restype = None
  _lib = lib
  return lib


def _p(a):
  return None if a is None else a.ctypes.data_as(C.c_void_p)


def normal_like(keys: np.ndarray, sizes) -> np.ndarray:
  keys = np.ascontiguousarray(keys, np.uint32)
  sz = np.asarray(sizes, np.int64)
  out = np.empty((keys.shape[0], int(sz.sum())), np.float32)
  load().oracle_normal_like(_p(keys), keys.shape[0], _p(sz), len(sz), _p(out))
  return out


def sgld_step(theta, v, grad, keys, sizes, step_size, temperature, alpha=0.9, lmbd=1e-5):
  """In-place SGLD (v None) / pSGLD step on C-contiguous f32 / u32 arrays."""
  sz = np.asarray(sizes, np.int64)
  assert theta.flags.c_contiguous and grad.flags.c_contiguous and keys.flags.c_contiguous
  load().oracle_sgld_step(_p(theta), _p(v), _p(grad), _p(keys), theta.shape[0], _p(sz),
                          len(sz), step_size, temperature, alpha, lmbd)
