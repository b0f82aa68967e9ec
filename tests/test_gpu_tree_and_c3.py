"""(1) util.tree_* operations == flat NumPy ops, bit-exact (the reference checks
them with assert_equal, tests/test_tree_operations.py:35-77).
(2) BASELINE.json configs[2] shapes: SGHMC / OBABO updates on the pytree of a
784-512-512-10 MLP (669 706 parameters per chain, six leaves, the last one
ragged) against the oracle, bit-exact."""
import numpy as np
import pytest

from oracle import prng
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu

MLP_SIZES = [784 * 512, 512, 512 * 512, 512, 512 * 10, 10]


def _bits(a):
  return np.ascontiguousarray(a).view(np.uint32)


def test_tree_operations_bit_exact(gpu):
  from jax_sgmc_b200 import tree_util as tu
  rng = np.random.default_rng(0)
  trees_a = [{"a": rng.standard_normal((3, 4)), "b": [rng.standard_normal(5), rng.standard_normal(())]}
             for _ in range(4)]
  trees_b = [{"a": rng.standard_normal((3, 4)), "b": [rng.standard_normal(5), rng.standard_normal(())]}
             for _ in range(4)]
  A, B = tu.ChainTree.from_trees(trees_a), tu.ChainTree.from_trees(trees_b)
  fa, fb = A.flat.numpy(), B.flat.numpy()
  assert fa.shape == (4, 18)
  assert np.array_equal(tu.tree_scale(0.37, A).flat.numpy(), np.float32(0.37) * fa)
  assert np.array_equal(tu.tree_add(A, B).flat.numpy(), fa + fb)
  assert np.array_equal(tu.tree_multiply(A, B).flat.numpy(), fa * fb)
  np.testing.assert_allclose(tu.tree_dot(A, B).numpy(), (fa * fb).sum(1), rtol=1e-6)
  assert np.array_equal(tu.tensor_matmul(tu.Tensor(1, B), A).flat.numpy(), fa * fb)
  assert np.array_equal(tu.tensor_matmul(tu.Tensor(0, 2.0), A).flat.numpy(), 2 * fa)
  with pytest.raises(NotImplementedError):
    tu.tensor_matmul(tu.Tensor(2, B), A)
  # round trip through the pytree view
  host = tu.tree_add(A, B).to_host()
  assert np.allclose(host["a"][2], np.float32(trees_a[2]["a"]) + np.float32(trees_b[2]["a"]))
  assert host["b"][1].shape == (4,)


def test_c3_mlp_pytree_sghmc_step_bit_exact(gpu):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C, P, eps = 2, sum(MLP_SIZES), 0.01
  rng = np.random.default_rng(1)
  theta = (rng.standard_normal((C, P)) * 0.05).astype(np.float32)
  grads = [(rng.standard_normal((C, P))).astype(np.float32) for _ in range(2)]
  keys = np.stack([prng.PRNGKey(c) for c in range(C)])
  st = osgmc.leapfrog_init(theta, keys)
  z = np.zeros(C, np.float32)
  want = osgmc.friction_leapfrog_integrate(
      st, [(lambda th, g=g: (z, None, g)) for g in grads], MLP_SIZES, eps, 1.0)
  d_theta, d_p = DA.from_numpy(theta), DA.from_numpy(theta)
  d_k = [DA.from_numpy(keys), DA((C, 2), np.uint32)]
  ops.sghmc_begin(d_theta, d_p, d_k[0], d_k[1], MLP_SIZES, eps)
  for s in range(2):
    ops.sghmc_step(d_theta, d_p, DA.from_numpy(grads[s]), d_k[(s + 1) % 2], d_k[s % 2],
                   MLP_SIZES, eps, friction=1.0, last=(s == 1))
  assert np.array_equal(d_k[1].numpy(), want.key)
  assert np.array_equal(_bits(d_theta.numpy()), _bits(want.theta))
  assert np.array_equal(_bits(d_p.numpy()), _bits(want.momentum))


def test_c3_mlp_pytree_obabo_step_bit_exact(gpu):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C, P, eps = 2, sum(MLP_SIZES), 0.01
  rng = np.random.default_rng(2)
  theta = (rng.standard_normal((C, P)) * 0.05).astype(np.float32)
  g1 = rng.standard_normal((C, P)).astype(np.float32)
  g2 = rng.standard_normal((C, P)).astype(np.float32)
  keys = np.stack([prng.PRNGKey(10 + c) for c in range(C)])
  z = np.zeros(C, np.float32)
  want = osgmc.obabo_integrate(osgmc.obabo_init(theta, keys),
                               [((lambda th: (z, None, g1)), (lambda th: (z, None, g2)))],
                               MLP_SIZES, eps, 1.0, 1.0)
  d_theta, d_p = DA.from_numpy(theta), DA.zeros(theta.shape)
  d_ks, d_ke = DA.zeros((C,)), DA.zeros((C,))
  d_kin, d_kout = DA.from_numpy(keys), DA((C, 2), np.uint32)
  ops.obabo_pass_a(d_theta, d_p, DA.from_numpy(g1), d_ks, d_kin, d_kout, MLP_SIZES, eps)
  ops.obabo_pass_b(d_p, DA.from_numpy(g2), d_ke, d_kin, MLP_SIZES, eps)
  assert np.array_equal(d_kout.numpy(), want.key)
  assert np.array_equal(_bits(d_theta.numpy()), _bits(want.theta))
  assert np.array_equal(_bits(d_p.numpy()), _bits(want.momentum))
  np.testing.assert_allclose(d_ks.numpy(), want.kinetic_energy_start, rtol=1e-5)
  np.testing.assert_allclose(d_ke.numpy(), want.kinetic_energy_end, rtol=1e-5)
