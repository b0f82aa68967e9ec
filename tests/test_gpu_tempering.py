"""reSGLD: solver.parallel_tempering (state swap) and the sharded label-swap
ladder must both reproduce the oracle's restatement of solver.py:264-293 for
R = 2 (same keys, same minibatches): identical exchange decisions, cold-chain
trajectories within rtol 1e-5."""
import numpy as np
import pytest

from oracle import data as odata
from oracle import prng
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu

D, S, NB, N, K = 8, 7, 16, 200, 120
EPS, T_HOT = 5e-3, 30.0


def _problem():
  from jax_sgmc_b200 import data, glm, integrator, potential
  X, y, _ = odata.logistic_dataset(N, D, seed=6)
  loader = data.DeviceNumpyDataLoader(x=X, y=y)
  pot = potential.minibatch_potential(glm.GaussianPrior(3.0), glm.LogisticRegression(),
                                      path="simt")
  batch_fn = data.random_reference_data(loader, 1, NB)
  integ = integrator.langevin_diffusion(pot, batch_fn)
  rng = np.random.default_rng(0)
  init_n = (rng.standard_normal((S, D)) * 0.1).astype(np.float32)
  init_h = (rng.standard_normal((S, D)) * 0.1).astype(np.float32)
  keys = np.stack([prng.PRNGKey(40 + s) for s in range(S)])
  return X, y, integ, init_n, init_h, keys


def _oracle_run(X, y, init_n, init_h, keys, sa_schedule=None):
  opot = osgmc.minibatch_potential(osgmc.Logistic(D, 0), osgmc.Prior("gaussian", 0, D, 3.0))
  st = osgmc.parallel_tempering_init(init_n, init_h, keys=keys)
  dk = prng.PRNGKey(0)
  cold, exch = [], []
  for _ in range(K):
    dk, idx = odata.device_draw(dk, NB, N)
    Xb, yb = X[idx], y[idx]
    fn = lambda th: opot(th, (Xb, yb), N)
    st, ex = osgmc.parallel_tempering_update(st, fn, fn, [D], EPS, 1.0, T_HOT,
                                             sa_schedule=sa_schedule)
    cold.append(st.normal.theta.copy())
    exch.append(ex.copy())
  return np.stack(cold), np.stack(exch), st


def test_parallel_tempering_matches_oracle(gpu):
  from jax_sgmc_b200 import scheduler, solver
  X, y, integ, init_n, init_h, keys = _problem()
  init, update, get = solver.parallel_tempering(integ)
  state = init([{"w": r} for r in init_n], [{"w": r} for r in init_h], key=keys)
  cold, exch = [], []
  for _ in range(K):
    state, _ = update(state, scheduler.schedule(EPS, 1.0, 1.0, True),
                      scheduler.schedule(EPS, T_HOT, 1.0, True))
    cold.append(get(state)["variables"].flat.numpy())
    exch.append(state.exchange.numpy().astype(bool))
  w_cold, w_exch, ost = _oracle_run(X, y, init_n, init_h, keys)
  assert np.array_equal(np.stack(exch), w_exch)
  assert 0 < w_exch.sum() < w_exch.size
  err = np.abs(np.stack(cold) - w_cold).max() / np.abs(w_cold).max()
  assert err < 1e-5, err
  assert np.array_equal(state.key.numpy(), ost.key)
  np.testing.assert_allclose(state.ssq.numpy(), ost.ssq, rtol=1e-4)


def test_parallel_tempering_custom_sa_schedule(gpu):
  """A user-supplied stochastic-approximation schedule (solver.py:221, :274-276) is
  applied, not replaced by 1 / n: decisions and ssq follow the oracle run with it."""
  from jax_sgmc_b200 import scheduler, solver
  X, y, integ, init_n, init_h, keys = _problem()
  sa = lambda n: 1.0 / (n + 10.0) ** 0.6
  init, update, _ = solver.parallel_tempering(integ, sa_schedule=sa)
  state = init([{"w": r} for r in init_n], [{"w": r} for r in init_h], key=keys)
  exch = []
  for _ in range(K):
    state, _ = update(state, scheduler.schedule(EPS, 1.0, 1.0, True),
                      scheduler.schedule(EPS, T_HOT, 1.0, True))
    exch.append(state.exchange.numpy().astype(bool))
  _, w_exch, ost = _oracle_run(X, y, init_n, init_h, keys, sa_schedule=sa)
  _, d_exch, dst = _oracle_run(X, y, init_n, init_h, keys)
  assert np.array_equal(np.stack(exch), w_exch)
  np.testing.assert_allclose(state.ssq.numpy(), ost.ssq, rtol=1e-4)
  assert not np.allclose(ost.ssq, dst.ssq, rtol=1e-3)       # the schedule matters


def test_swap_rows_beyond_one_grid_dimension(gpu):
  """More rows than gridDim.y holds (65 535): rows are exchanged in slices."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  rows, w = 70_001, 3
  rng = np.random.default_rng(2)
  a = rng.standard_normal((rows, w)).astype(np.float32)
  b = rng.standard_normal((rows, w)).astype(np.float32)
  ex = (rng.random(rows) < 0.5).astype(np.int32)
  da, db = DA.from_numpy(a), DA.from_numpy(b)
  ops.swap_rows(da, db, DA.from_numpy(ex))
  m = ex.astype(bool)[:, None]
  assert np.array_equal(da.numpy(), np.where(m, b, a))
  assert np.array_equal(db.numpy(), np.where(m, a, b))


@pytest.mark.parametrize("overlap", [False, True])
def test_sharded_label_swap_ladder_equals_state_swap(gpu, overlap):
  """R = 2 on one rank: exchanging temperature labels (what crosses NVLink in
  the multi-GPU layout) gives the reference's cold chain -- also with the
  exchange running on its own stream under the next step's potential."""
  from jax_sgmc_b200 import scheduler, tempering
  X, y, integ, init_n, init_h, keys = _problem()
  init, update, get = tempering.sharded_tempering(integ, [1.0, T_HOT],
                                                  overlap_exchange=overlap)
  state = init([[{"w": r} for r in init_n], [{"w": r} for r in init_h]], key=keys)
  cold, exch = [], []
  for _ in range(K):
    state, _ = update(state, scheduler.schedule(EPS, 1.0, 1.0, True))
    cold.append(tempering.cold_samples(get(state)))
    exch.append(state.exchange.numpy()[0].astype(bool))
  w_cold, w_exch, _ = _oracle_run(X, y, init_n, init_h, keys)
  assert np.array_equal(np.stack(exch), w_exch)
  err = np.abs(np.stack(cold) - w_cold).max() / np.abs(w_cold).max()
  assert err < 1e-5, err


@pytest.mark.parametrize("overlap", [False, True])
def test_ladder_four_replicas_keeps_a_permutation(gpu, overlap):
  """R = 4 (extension, no reference oracle): labels stay a permutation, every
  pair gets attempted, hot replicas spread wider than cold ones."""
  from jax_sgmc_b200 import scheduler, tempering
  X, y, integ, init_n, _, keys = _problem()
  temps = [1.0, 4.0, 16.0, 64.0]
  init, update, get = tempering.sharded_tempering(integ, temps, overlap_exchange=overlap)
  state = init([{"w": r} for r in init_n], key=keys)
  seen = np.zeros((3, S), bool)
  for _ in range(200):
    state, _ = update(state, scheduler.schedule(EPS, 1.0, 1.0, True))
    seen |= state.exchange.numpy().astype(bool)
  state.wait()
  holder = state.holder.numpy()
  assert all(sorted(holder[:, b].tolist()) == [0, 1, 2, 3] for b in range(S))
  assert seen.any(axis=1).all()
  tidx = state.temp_index.numpy()
  assert all(sorted(tidx[:, b].tolist()) == [0, 1, 2, 3] for b in range(S))
