"""The C-ABI library loads and exports every symbol include/sgmc_b200.h
declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from jax_sgmc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  text = open(os.path.join(ROOT, "include", "sgmc_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(sgmc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
  lib = _lib.load()
  names = _declared()
  assert len(names) >= 40
  for n in names:
    assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_binding_covers_header():
  assert set(_declared()) == set(_lib.exported_names())


def test_no_torch_or_jax_dependency():
  """The shared object links only the CUDA runtime (no torch, no NCCL at link
  time: NCCL is dlopen'ed)."""
  import subprocess
  out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
  assert "torch" not in out and "libnccl" not in out and "jax" not in out


def test_errors_are_reported_not_swallowed():
  lib = _lib.load()
  assert lib.sgmc_version() >= 100
  # argument validation happens before any CUDA call
  ls = (ctypes.c_int64 * 1)(8)
  rc = lib.sgmc_sgld_update(None, None, None, None, None, 1, ls, 0, 0.1, 1.0,
                            None, 0)
  assert rc != 0
  assert b"alias" in lib.sgmc_last_error() or b"n_leaves" in lib.sgmc_last_error()


def test_binding_argument_counts_match_the_header():
  """Every ctypes prototype has as many arguments as the declaration in
  include/sgmc_b200.h (the binding is written by hand: guard against drift)."""
  text = open(os.path.join(ROOT, "include", "sgmc_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  decls = dict()
  for name, args in re.findall(r"\b(sgmc_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
    args = " ".join(args.split())
    decls[name] = 0 if args in ("", "void") else args.count(",") + 1
  protos = {**{k: v for k, v in _lib.PROTOTYPES.items()},
            **{k: v[0] for k, v in _lib.SPECIAL.items()}}
  assert set(protos) <= set(decls)
  for name, argtypes in protos.items():
    assert len(argtypes) == decls[name], (name, len(argtypes), decls[name])
