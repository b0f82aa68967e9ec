"""Pytrees of chain-batched device buffers.

Mirrors the parts of ``jax.tree_util`` / ``jax_sgmc.util`` the hot path relies
on (util/tree_util.py:27-133, util/list_map.py:53-56).  Leaves are visited in
``jax.tree_util.tree_flatten`` order (dict keys sorted, tuples / lists /
namedtuples positionally, ``None`` = empty subtree): this is the order
``integrator.random_tree`` hands out per-leaf noise keys (integrator.py:
130-134) and ``flatten_util.ravel_pytree`` concatenates leaves
(adaption.py:92-104), so it defines the flat layout the kernels see.

A :class:`ChainTree` is the native representation of "one pytree per chain":
all chains stacked on a leading axis and raveled to one ``f32[C, P]`` device
buffer (the reference's ``list_vmap`` stacking, util/list_map.py:53-56).
"""
from __future__ import annotations

from typing import Any, Callable, List, NamedTuple, Sequence, Tuple

import numpy as np

from .device import DeviceArray

PyTree = Any


class Tensor(NamedTuple):
  """util/tree_util.py:27-36: ndim 0 scalar, 1 diagonal (pytree), 2 dense."""
  ndim: int
  tensor: PyTree


def tree_flatten(tree) -> Tuple[List[Any], Any]:
  leaves: List[Any] = []

  def rec(t):
    if t is None:
      return ("none",)
    if isinstance(t, ChainTree):
      leaves.append(t)
      return ("leaf",)
    if isinstance(t, dict):
      keys = sorted(t.keys())
      return ("dict", keys, [rec(t[k]) for k in keys])
    if isinstance(t, tuple) and hasattr(t, "_fields"):
      return ("namedtuple", type(t), [rec(x) for x in t])
    if isinstance(t, (tuple, list)):
      return ("tuple" if isinstance(t, tuple) else "list", [rec(x) for x in t])
    leaves.append(t)
    return ("leaf",)

  return leaves, rec(tree)


def tree_unflatten(treedef, leaves):
  it = iter(leaves)

  def rec(d):
    kind = d[0]
    if kind == "none":
      return None
    if kind == "leaf":
      return next(it)
    if kind == "dict":
      return {k: rec(s) for k, s in zip(d[1], d[2])}
    if kind == "namedtuple":
      return d[1](*[rec(s) for s in d[2]])
    if kind == "tuple":
      return tuple(rec(s) for s in d[1])
    return [rec(s) for s in d[1]]

  return rec(treedef)


def tree_map(fn: Callable, tree, *rest):
  leaves, treedef = tree_flatten(tree)
  others = [tree_flatten(r)[0] for r in rest]
  return tree_unflatten(treedef, [fn(*xs) for xs in zip(leaves, *others)])


def tree_leaves(tree):
  return tree_flatten(tree)[0]


class ChainTree:
  """C pytrees with identical structure, stacked and raveled on the device.

  ``flat`` is ``f32[C, P]``; ``sizes`` / ``shapes`` describe the leaves in
  tree_flatten order, ``treedef`` rebuilds the pytree.
  """

  def __init__(self, flat: DeviceArray, treedef, shapes: Sequence[Tuple[int, ...]]):
    self.flat = flat
    self.treedef = treedef
    self.shapes = [tuple(s) for s in shapes]
    self.sizes = [int(np.prod(s, dtype=np.int64)) for s in self.shapes]
    assert sum(self.sizes) == flat.shape[1], (self.sizes, flat.shape)

  # -- construction --------------------------------------------------------------
  @classmethod
  def from_trees(cls, trees: Sequence[PyTree]) -> "ChainTree":
    """Stack host pytrees (one per chain) -> device ``f32[C, P]``."""
    rows, treedef, shapes = [], None, None
    for t in trees:
      leaves, td = tree_flatten(t)
      leaves = [np.asarray(l, dtype=np.float32) for l in leaves]
      if treedef is None:
        treedef, shapes = td, [l.shape for l in leaves]
      else:
        assert td == treedef and [l.shape for l in leaves] == shapes, \
            "all chains must share the pytree structure"
      rows.append(np.concatenate([l.ravel() for l in leaves]) if leaves
                  else np.zeros(0, np.float32))
    return cls(DeviceArray.from_numpy(np.stack(rows).astype(np.float32)),
               treedef, shapes)

  @classmethod
  def like(cls, other: "ChainTree", flat: DeviceArray) -> "ChainTree":
    return cls(flat, other.treedef, other.shapes)

  # -- views -----------------------------------------------------------------------
  @property
  def n_chains(self) -> int:
    return self.flat.shape[0]

  @property
  def n_params(self) -> int:
    return self.flat.shape[1]

  def offsets(self) -> List[int]:
    out, off = [], 0
    for s in self.sizes:
      out.append(off)
      off += s
    return out

  def leaf_index(self, path) -> int:
    """Index (tree_flatten order) of the leaf reached by ``path`` (a dict key
    or a sequence of keys / positions)."""
    if not isinstance(path, (tuple, list)):
      path = (path,)
    counter = [0]
    found = []

    def rec(d, p):
      kind = d[0]
      if kind == "none":
        return
      if kind == "leaf":
        if p == ():
          found.append(counter[0])
        counter[0] += 1
        return
      if kind == "dict":
        for k, s in zip(d[1], d[2]):
          rec(s, p[1:] if (p and p[0] == k) else (None,))
        return
      subs = d[2] if kind == "namedtuple" else d[1]
      for i, s in enumerate(subs):
        rec(s, p[1:] if (p and p[0] == i) else (None,))

    rec(self.treedef, tuple(path))
    if len(found) != 1:
      raise KeyError(f"no unique leaf at {path}")
    return found[0]

  def copy(self) -> "ChainTree":
    return ChainTree(self.flat.copy(), self.treedef, self.shapes)

  def to_host(self, flat: np.ndarray = None) -> PyTree:
    """Pytree of numpy arrays with a leading chain axis ``[C, *leaf_shape]``."""
    flat = self.flat.numpy() if flat is None else flat
    return unravel_rows(flat, self.treedef, self.shapes)

  def chain_to_host(self, chain: int) -> PyTree:
    host = self.to_host()
    return tree_map(lambda l: l[chain], host)

  def __repr__(self):
    return f"ChainTree(chains={self.n_chains}, params={self.n_params}, leaves={len(self.sizes)})"


# -- pytree vector-space operations (util/tree_util.py:38-133) --------------------
# Every ChainTree operand is a flat f32[C, P] device buffer, so the leaf-wise
# tree_map of the reference becomes one elementwise kernel.

def _ewise(op: int, alpha, a: "ChainTree", b: "ChainTree" = None) -> "ChainTree":
  from . import ops
  out = DeviceArray(a.flat.shape, np.float32)
  ops.tree_ewise(op, out, alpha, a.flat, None if b is None else b.flat)
  return ChainTree.like(a, out)


def tree_scale(alpha, tree: "ChainTree") -> "ChainTree":
  """util/tree_util.py:83-96: ``alpha * x`` on every leaf."""
  return _ewise(0, float(alpha), tree)


def tree_add(tree_a: "ChainTree", tree_b: "ChainTree") -> "ChainTree":
  """util/tree_util.py:99-110."""
  return _ewise(1, 0.0, tree_a, tree_b)


def tree_multiply(tree_a: "ChainTree", tree_b: "ChainTree") -> "ChainTree":
  """util/tree_util.py:58-80."""
  return _ewise(2, 0.0, tree_a, tree_b)


def tree_dot(tree_a: "ChainTree", tree_b: "ChainTree") -> DeviceArray:
  """util/tree_util.py:120-133, one scalar per chain: ``f32[C]``."""
  from . import ops
  out = DeviceArray((tree_a.n_chains,), np.float32)
  ops.tree_dot(out, tree_a.flat, tree_b.flat)
  return out


def tensor_matmul(matrix: Tensor, vector: "ChainTree") -> "ChainTree":
  """util/tree_util.py:38-56: scalar (ndim 0) and diagonal (ndim 1) tensors;
  dense matrices (ndim 2) are outside the fused path."""
  if matrix.ndim == 0:
    return tree_scale(matrix.tensor, vector)
  if matrix.ndim == 1:
    return tree_multiply(matrix.tensor, vector)
  raise NotImplementedError(f"Cannot multiply matrix with dimension {matrix.ndim}")


def unravel_rows(flat: np.ndarray, treedef, shapes) -> PyTree:
  """Inverse of the chain-wise ravel for a host array ``[..., P]``."""
  out, off = [], 0
  lead = flat.shape[:-1]
  for shp in shapes:
    sz = int(np.prod(shp, dtype=np.int64))
    out.append(flat[..., off:off + sz].reshape(lead + tuple(shp)))
    off += sz
  return tree_unflatten(treedef, out)
