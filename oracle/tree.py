"""Minimal pytree helpers with ``jax.tree_util`` ordering (oracle only).

TEST INFRASTRUCTURE.  Dict leaves are visited in sorted-key order, tuples,
lists and namedtuples positionally, ``None`` is an empty subtree -- the order
``jax.tree_util.tree_flatten`` uses and therefore the order
``integrator.random_tree`` (integrator.py:130-134) assigns per-leaf keys and
``flatten_util.ravel_pytree`` (adaption.py:92-104) concatenates leaves.
"""
from __future__ import annotations

import numpy as np


def tree_flatten(tree):
  leaves = []

  def rec(t):
    if t is None:
      return ("none",)
    if isinstance(t, dict):
      keys = sorted(t.keys())
      return ("dict", keys, [rec(t[k]) for k in keys])
    if isinstance(t, tuple) and hasattr(t, "_fields"):
      return ("namedtuple", type(t), [rec(x) for x in t])
    if isinstance(t, (tuple, list)):
      return ("tuple" if isinstance(t, tuple) else "list", [rec(x) for x in t])
    leaves.append(np.asarray(t))
    return ("leaf",)

  treedef = rec(tree)
  return leaves, treedef


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


def tree_map(fn, tree, *rest):
  leaves, treedef = tree_flatten(tree)
  others = [tree_flatten(r)[0] for r in rest]
  return tree_unflatten(treedef, [fn(*xs) for xs in zip(leaves, *others)])


def ravel_pytree(tree):
  """``jax.flatten_util.ravel_pytree``: concat of raveled leaves + inverse."""
  leaves, treedef = tree_flatten(tree)
  shapes = [l.shape for l in leaves]
  sizes = [int(np.prod(s, dtype=np.int64)) for s in shapes]
  flat = (np.concatenate([np.ravel(l) for l in leaves]).astype(np.float32)
          if leaves else np.zeros((0,), np.float32))

  def unravel(vec):
    out, off = [], 0
    for shp, sz in zip(shapes, sizes):
      out.append(np.asarray(vec[off:off + sz]).reshape(shp))
      off += sz
    return tree_unflatten(treedef, out)

  return flat, unravel


def leaf_sizes(tree):
  return [int(l.size) for l in tree_flatten(tree)[0]]
