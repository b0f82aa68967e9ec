"""The oracle's restatement of the MH-corrected samplers (solver.py:301-577,
integrator.py:349-560) must itself sample the right distribution: with exact
gradients and the exact potential of N(0, s^2 I) both SGGMC and AMAGOLD leave
it invariant (the reference checks its own implementation the same way,
tests/test_alias.py:165-201).  CPU only."""
import numpy as np
from scipy import stats as scpstats

from oracle import prng
from oracle import sgmc as osgmc

S = 0.5


def _target(P):
  full = lambda th: (0.5 * np.sum(th.astype(np.float64) ** 2, axis=1) / S ** 2).astype(np.float32)
  grad = lambda th: (full(th), None, (th / np.float32(S ** 2)).astype(np.float32))
  return full, grad


def _keys(C, base=0):
  return np.stack([prng.PRNGKey(base + c) for c in range(C)])


def test_oracle_sggmc_samples_the_gaussian():
  C, P, steps = 64, 2, 3
  full, grad = _target(P)
  st = osgmc.sggmc_init(np.zeros((C, P), np.float32), full, _keys(C))
  kept, acc = [], []
  for it in range(260):
    st, a = osgmc.sggmc_update(st, [(grad, grad)] * steps, full, [P], 0.25, 1.0, 1.0)
    acc.append(a.mean())
    if it >= 60 and it % 5 == 0:
      kept.append(st.integrator_state.theta.copy())
  x = np.concatenate(kept).ravel()
  assert 0.5 < np.mean(acc) <= 1.0
  assert abs(x.mean()) < 0.03
  assert abs(x.std() - S) < 0.03
  assert scpstats.kstest(x[::7] / S, "norm").pvalue > 0.01


def test_oracle_amagold_samples_the_gaussian():
  C, P, steps = 64, 2, 3
  full, grad = _target(P)
  st = osgmc.amagold_init(np.zeros((C, P), np.float32), full, _keys(C, 500), sizes=[P])
  kept, acc = [], []
  for it in range(260):
    st, a = osgmc.amagold_update(st, [grad] * steps, full, [P], 0.2, 0.25)
    acc.append(a.mean())
    if it >= 60 and it % 5 == 0:
      kept.append(st.integrator_state.theta.copy())
  x = np.concatenate(kept).ravel()
  assert 0.5 < np.mean(acc) <= 1.0
  assert abs(x.mean()) < 0.03
  assert abs(x.std() - S) < 0.03
  assert scpstats.kstest(x[::7] / S, "norm").pvalue > 0.01


def test_oracle_mh_rejection_restores_the_state():
  """A proposal with a huge potential is rejected: positions and (for AMAGOLD,
  negated) momentum of the old state come back, the key stream still advances."""
  C, P = 4, 3
  full_ok, grad = _target(P)
  calls = {"n": 0}

  def full(th):                     # first call (init) honest, then a wall
    calls["n"] += 1
    return full_ok(th) if calls["n"] == 1 else full_ok(th) + np.float32(1e6)

  theta = np.ones((C, P), np.float32)
  st = osgmc.amagold_init(theta, full, _keys(C, 9), sizes=[P])
  p0 = st.integrator_state.momentum.copy()
  new, acc = osgmc.amagold_update(st, [grad] * 2, full, [P], 0.1, 0.25)
  assert not acc.any()
  assert np.array_equal(new.integrator_state.theta, theta)
  assert np.array_equal(new.integrator_state.momentum, -p0)
  assert not np.array_equal(new.key, st.key)
  assert np.array_equal(new.potential, st.potential)
