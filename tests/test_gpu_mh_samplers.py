"""SGGMC / AMAGOLD: the reversible-leapfrog kernel,
the Metropolis-Hastings decision kernel and the two solvers against the oracle's
restatement of integrator.py:349-560 and solver.py:301-577, plus the reference's
own statistical acceptance tests (tests/test_alias.py:165-201)."""
import numpy as np
import pytest
from scipy import stats as scpstats

from oracle import data as odata
from oracle import prng
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu


def _keys(C, base=0):
  return np.stack([prng.PRNGKey(base + c) for c in range(C)])


@pytest.mark.parametrize("sizes,with_mass", [([64], False), ([1, 4, 19], True),
                                             ([1024], False), ([7, 2, 33], False)])
def test_revleapfrog_kernel_matches_oracle(gpu, sizes, with_mass):
  """Three inner steps with external gradients: theta, p, keys bit-exact (every
  op separately rounded, oracle order); accumulated energy to 1e-5 (reduction
  order)."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  rng = np.random.default_rng(3)
  C, P, steps = 5, sum(sizes), 3
  theta = rng.standard_normal((C, P)).astype(np.float32)
  mass = (np.abs(rng.standard_normal(P)) + 0.5).astype(np.float32) if with_mass else None
  grads = [(rng.standard_normal((C, P)) * 2).astype(np.float32) for _ in range(steps)]
  eps, fr = 0.013, 0.25
  st = osgmc.reversible_leapfrog_init(theta, _keys(C, 7), mass, sizes)
  want = osgmc.reversible_leapfrog_integrate(
      st, [lambda th, g=g: (None, None, g) for g in grads], sizes, eps, fr, mass)
  # device: opening half step, then the fused kernel per inner step
  inv_m = np.ones(P, np.float32) if mass is None else (np.float32(1) / mass)
  half = np.float32(0.5) * np.float32(eps)
  th0 = (st.theta + (half * (inv_m * st.momentum).astype(np.float32)).astype(np.float32)
         ).astype(np.float32)
  d_t, d_p = DA.from_numpy(th0), DA.from_numpy(st.momentum)
  d_e = DA.zeros((C,))
  kk = [DA.from_numpy(st.key), DA((C, 2), np.uint32)]
  d_m = None if mass is None else DA.from_numpy(mass)
  for s in range(steps):
    ops.revleapfrog_step(d_t, d_p, DA.from_numpy(grads[s]), d_e, kk[s % 2], kk[(s + 1) % 2],
                         sizes, eps, fr, d_m, last=(s == steps - 1))
  assert np.array_equal(kk[steps % 2].numpy(), want.key)
  assert np.array_equal(d_p.numpy().view(np.uint32), want.momentum.view(np.uint32))
  assert np.array_equal(d_t.numpy().view(np.uint32), want.theta.view(np.uint32))
  np.testing.assert_allclose(d_e.numpy(), want.potential, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("mode", ["sggmc", "amagold"])
def test_mh_decide_matches_oracle(gpu, mode):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  rng = np.random.default_rng(1)
  C = 1000
  U_old = (rng.standard_normal(C) * 5 + 100).astype(np.float32)
  U_new = (U_old + rng.standard_normal(C) * 1.5).astype(np.float32)
  e0 = np.abs(rng.standard_normal(C)).astype(np.float32)
  e1 = np.abs(rng.standard_normal(C)).astype(np.float32)
  U_new[:5] = U_old[:5]                 # log_alpha == +-(e1 - e0): both signs
  e1[:3] = e0[:3]                       # log_alpha == 0 exactly -> always accepted
  U_new[5] = np.nan                     # see the last assertion
  keys = _keys(C, 50)
  T = 1.7
  accept, key, la, ratio = osgmc.mh_decision(mode, U_old, U_new, e0, e1, T, keys)
  d_U = DA.from_numpy(U_old)
  rej, rat = DA((C,), np.int32), DA((C,), np.float32)
  k1 = DA((C, 2), np.uint32)
  ops.mh_decide(mode, d_U, DA.from_numpy(U_new), DA.from_numpy(e0), DA.from_numpy(e1), T,
                DA.from_numpy(keys), k1, rej, rat)
  assert np.array_equal(rej.numpy() == 0, accept)
  assert 0.2 < accept.mean() < 0.95
  assert np.array_equal(k1.numpy(), key)
  got_U = d_U.numpy()
  want_U = np.where(accept, U_new, U_old)
  assert np.array_equal(got_U.view(np.uint32), want_U.view(np.uint32))
  ok = ~np.isnan(ratio)
  np.testing.assert_allclose(rat.numpy()[ok], ratio[ok], rtol=2e-6)
  # a NaN potential: where(la <= 0, la, 0) turns it into log_alpha = 0 (accept) in
  # sggmc (solver.py:529), where(la > 0, 0, la) keeps the NaN (reject) in amagold (:383)
  assert accept[:3].all() and accept[5] == (mode == "sggmc")


def _logistic_problem(C, d, N, seed=0):
  X, y, _ = odata.logistic_dataset(N, d, seed=seed)
  rng = np.random.default_rng(seed + 1)
  theta = (rng.standard_normal((C, d)) * 0.1).astype(np.float32)
  return X, y, theta


def test_mass_matrix_update_matches_oracle(gpu):
  """adaption.mass_matrix(diagonal=True) (adaption.py:296-369): Welford statistics of
  every chain and the one-off matrix update in iteration burn_in, bit for bit against
  the oracle; get() hands out the same MassMatrix object every time."""
  from jax_sgmc_b200 import adaption
  from jax_sgmc_b200.tree_util import ChainTree
  rng = np.random.default_rng(11)
  C, burn_in = 5, 7
  tmpl = {"b": np.zeros((), np.float32), "w": np.zeros((3, 4), np.float32)}
  init, update, get = adaption.mass_matrix(burn_in=burn_in)
  first = rng.standard_normal((C, 13)).astype(np.float32)
  tree = ChainTree.from_trees([tmpl] * C)
  tree.flat.copy_from_host(first)
  cov = (np.abs(rng.standard_normal(13)) + 0.5).astype(np.float32)
  state = init(tree, {"b": cov[:1].reshape(()), "w": cov[1:].reshape(3, 4)})
  want = osgmc.mass_matrix_init(first, cov)
  m0 = get(state)
  assert np.array_equal(m0.inv.tensor.flat.numpy(), want.m_inv)
  assert np.array_equal(m0.sqrt.tensor.flat.numpy().view(np.uint32), want.m_sqrt.view(np.uint32))
  for it in range(1, 11):
    x = (rng.standard_normal((C, 13)) * (1 + it)).astype(np.float32)
    tree.flat.copy_from_host(x)
    state = update(state, tree)
    want = osgmc.mass_matrix_update(want, x, burn_in)
    assert get(state) is m0
    for got, ref in ((state.mean, want.mean), (state.ssq, want.ssq), (state.m_inv, want.m_inv),
                     (state.m_sqrt, want.m_sqrt)):
      assert np.array_equal(got.numpy().view(np.uint32), ref.view(np.uint32)), it
    assert (it < burn_in) == np.array_equal(want.m_inv, np.broadcast_to(cov, (C, 13)))
  with pytest.raises(NotImplementedError):
    adaption.mass_matrix(diagonal=False)


def test_adapted_mass_kernels_match_oracle(gpu):
  """The OBABO passes and the reversible-leapfrog step with a per-chain MassMatrix(inv,
  sqrt): theta, p, keys bit-exact against the oracle run with the same matrices."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  rng = np.random.default_rng(5)
  sizes = [3, 20, 41]
  C, P, steps = 4, sum(sizes), 2
  theta = rng.standard_normal((C, P)).astype(np.float32)
  inv = (np.abs(rng.standard_normal((C, P))) + 0.3).astype(np.float32)
  sqrt = (np.abs(rng.standard_normal((C, P))) + 0.3).astype(np.float32)   # independent arrays
  g = [(rng.standard_normal((C, P)) * 2).astype(np.float32) for _ in range(2 * steps)]
  eps, T, fr = 0.011, 1.3, 0.7
  d_inv, d_sqrt = DA.from_numpy(inv), DA.from_numpy(sqrt)
  # ---- OBABO
  st = osgmc.obabo_init(theta, _keys(C, 3))
  st = st._replace(momentum=rng.standard_normal((C, P)).astype(np.float32))
  pairs = [(lambda th, a=g[2 * s]: (np.zeros(C, np.float32), None, a),
            lambda th, b=g[2 * s + 1]: (np.zeros(C, np.float32), None, b)) for s in range(steps)]
  want = osgmc.obabo_integrate(st, pairs, sizes, eps, T, fr, mass_matrix=(inv, sqrt))
  d_t, d_p = DA.from_numpy(theta), DA.from_numpy(st.momentum)
  ke0, ke1 = DA.zeros((C,)), DA.zeros((C,))
  kk = [DA.from_numpy(st.key), DA((C, 2), np.uint32)]
  for s in range(steps):
    ops.obabo_pass_a(d_t, d_p, DA.from_numpy(g[2 * s]), ke0, kk[s % 2], kk[(s + 1) % 2], sizes,
                     eps, T, fr, (d_inv, d_sqrt))
    ops.obabo_pass_b(d_p, DA.from_numpy(g[2 * s + 1]), ke1, kk[s % 2], sizes, eps, T, fr,
                     (d_inv, d_sqrt))
  assert np.array_equal(kk[steps % 2].numpy(), want.key)
  assert np.array_equal(d_t.numpy().view(np.uint32), want.theta.view(np.uint32))
  assert np.array_equal(d_p.numpy().view(np.uint32), want.momentum.view(np.uint32))
  np.testing.assert_allclose(ke0.numpy(), want.kinetic_energy_start, rtol=1e-5)
  np.testing.assert_allclose(ke1.numpy(), want.kinetic_energy_end, rtol=1e-5)
  # ---- reversible leapfrog
  ls = osgmc.reversible_leapfrog_init(theta, _keys(C, 9), None, sizes, mass_matrix=(inv, sqrt))
  want = osgmc.reversible_leapfrog_integrate(
      ls, [lambda th, a=a: (None, None, a) for a in g[:steps]], sizes, eps, 0.25,
      mass_matrix=(inv, sqrt))
  half = np.float32(0.5) * np.float32(eps)
  th0 = (ls.theta + (half * (inv * ls.momentum).astype(np.float32)).astype(np.float32)
         ).astype(np.float32)
  d_t, d_p, d_e = DA.from_numpy(th0), DA.from_numpy(ls.momentum), DA.zeros((C,))
  kk = [DA.from_numpy(ls.key), DA((C, 2), np.uint32)]
  for s in range(steps):
    ops.revleapfrog_step(d_t, d_p, DA.from_numpy(g[s]), d_e, kk[s % 2], kk[(s + 1) % 2], sizes,
                         eps, 0.25, (d_inv, d_sqrt), last=(s == steps - 1))
  assert np.array_equal(kk[steps % 2].numpy(), want.key)
  assert np.array_equal(d_p.numpy().view(np.uint32), want.momentum.view(np.uint32))
  assert np.array_equal(d_t.numpy().view(np.uint32), want.theta.view(np.uint32))
  np.testing.assert_allclose(d_e.numpy(), want.potential, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("adapt_mass", [False, True])
@pytest.mark.parametrize("kind", ["sggmc", "amagold"])
def test_mh_solver_trajectory_matches_oracle(gpu, kind, adapt_mass):
  """Six MH iterations of solver.sggmc / solver.amagold driven through the
  operator API against the oracle: identical accept/reject pattern, identical
  keys, samples within 1e-5 (fp32 SIMT gradients, different reduction order).
  ``adapt_mass``: with ``mass_adaption=adaption.mass_matrix(burn_in=3)`` -- the matrix
  changes after the third iteration (solver.py:503-506, :552-553 / :366-369, :409-410)."""
  from jax_sgmc_b200 import adaption, data, glm, integrator, potential, scheduler, solver
  from jax_sgmc_b200 import ops
  ops.set_option(ops.OPT_EXACT_UPDATE_MATH, 1)
  try:
    C, d, N, n, steps, iters = 6, 8, 60, 12, 3, 6
    adapt = adaption.mass_matrix(burn_in=3) if adapt_mass else None
    X, y, theta = _logistic_problem(C, d, N)
    loader = data.DeviceNumpyDataLoader(x=X, y=y)
    prior, lik = glm.GaussianPrior(3.0), glm.LogisticRegression()
    pot = potential.minibatch_potential(prior, lik, strategy="vmap", path="simt")
    full = potential.full_potential(prior, lik, strategy="vmap", path="simt")
    random_data = data.random_reference_data(loader, 1, n)
    full_map = data.full_reference_data(loader, 16, 16)
    eps, fr = 0.02, (1.0 if kind == "sggmc" else 0.25)
    if kind == "sggmc":
      integ = integrator.obabo(pot, random_data, steps, fr)
      init, update, get = solver.sggmc(integ, full, full_map, mass_adaption=adapt)
    else:
      integ = integrator.reversible_leapfrog(pot, random_data, steps, fr)
      init, update, get = solver.amagold(integ, full, full_map, mass_adaption=adapt)
    keys = _keys(C, 20)
    state = init([{"w": t} for t in theta], key=keys)
    sched = scheduler.schedule(step_size=np.float32(eps), temperature=np.float32(1.0),
                               burn_in=np.float32(1.0), accept=True)

    # ---- oracle side: same data key stream (PRNGKey(0) shared by the chains) ----
    o_pot = osgmc.minibatch_potential(osgmc.Logistic(d, 0), osgmc.Prior("gaussian", 0, d, 3.0))
    o_full_fn = osgmc.full_potential(osgmc.Logistic(d, 0), osgmc.Prior("gaussian", 0, d, 3.0))
    ids = np.arange(int(np.ceil(N / 16)) * 16).reshape(-1, 16)
    batches = [(X[i % N], y[i % N], (i < N).astype(np.float32)) for i in ids]
    o_full = lambda th: o_full_fn(th, batches, N)
    dkey = prng.PRNGKey(0)

    def next_grad_fn():
      nonlocal dkey
      dkey, idx = odata.device_draw(dkey, n, N)
      return lambda th, idx=idx: o_pot(th, (X[idx], y[idx]), N)

    o_mass = osgmc.mass_matrix_init(theta) if adapt_mass else None
    mm = lambda: (o_mass.m_inv, o_mass.m_sqrt) if adapt_mass else None
    if kind == "sggmc":
      o_state = osgmc.sggmc_init(theta, o_full, keys)
    else:
      o_state = osgmc.amagold_init(theta, o_full, keys, sizes=[d], mass_matrix=mm())
    np.testing.assert_allclose(state.potential.numpy(), o_state.potential, rtol=1e-5)
    pattern = []
    for it in range(iters):
      state, stats = update(state, sched)
      if kind == "sggmc":
        pairs = [(next_grad_fn(), next_grad_fn()) for _ in range(steps)]
        o_state, acc = osgmc.sggmc_update(o_state, pairs, o_full, [d], eps, 1.0, fr,
                                          mass_matrix=mm())
      else:
        fns = [next_grad_fn() for _ in range(steps)]
        o_state, acc = osgmc.amagold_update(o_state, fns, o_full, [d], eps, fr,
                                            mass_matrix=mm())
      if adapt_mass:
        o_mass = osgmc.mass_matrix_update(o_mass, o_state.integrator_state.theta, 3)
      pattern.append(acc)
      assert np.array_equal(state.reject.numpy() == 0, acc), f"iteration {it}"
      assert np.array_equal(state.key.current.numpy(), o_state.key)
      got = get(state)["variables"].flat.numpy()
      np.testing.assert_allclose(got, o_state.integrator_state.theta, rtol=1e-5, atol=1e-6)
      np.testing.assert_allclose(state.integrator_state.momentum.flat.numpy(),
                                 o_state.integrator_state.momentum, rtol=1e-5, atol=1e-6)
      np.testing.assert_allclose(state.potential.numpy(), o_state.potential, rtol=1e-5)
      np.testing.assert_allclose(stats["acceptance_ratio"].numpy(),
                                 o_state.acceptance_ratio, rtol=1e-3, atol=1e-6)
    pattern = np.array(pattern)
    assert pattern.any()
    if not adapt_mass:
      assert not pattern.all(), "want both accepts and rejects in the test"
    if adapt_mass:
      assert state.mass_state.iteration == iters
      assert not np.allclose(state.mass_state.m_inv.numpy(), 1.0)
      np.testing.assert_allclose(state.mass_state.m_inv.numpy(), o_mass.m_inv, rtol=2e-4,
                                 atol=1e-9)
  finally:
    ops.set_option(ops.OPT_EXACT_UPDATE_MATH, 0)


def _ks_problem():
  """tests/test_alias.py:32-65 (see test_gpu_api._ks_problem)."""
  from jax_sgmc_b200 import data, glm, potential
  x = 0.5 * prng.normal(prng.PRNGKey(11), (100,))
  loader = data.DeviceNumpyDataLoader(x=np.zeros((2, 1), np.float32),
                                      y=np.zeros((2,), np.float32))
  prior, lik = glm.GaussianPrior(0.5), glm.LogisticRegression()
  pot = potential.minibatch_potential(prior, lik, strategy="vmap")
  full = potential.full_potential(prior, lik, strategy="vmap")
  init = {"w": np.array([x[0]], np.float32)}

  def check(samples):
    st = scpstats.kstest(np.ravel(samples), x)
    assert st.pvalue > 0.05, f"KS p-value {st.pvalue}"

  return loader, pot, full, init, check


def test_ks_amagold_and_sggmc(gpu):
  """tests/test_alias.py:165-201: both MH samplers reproduce N(0, 0.5^2)."""
  from jax_sgmc_b200 import alias
  loader, pot, full, init, check = _ks_problem()
  for make in (alias.amagold, alias.sggmc):
    run = make(pot, full, loader, cache_size=1, batch_size=1, first_step_size=0.5,
               last_step_size=0.1, burn_in=100, progress_bar=False)
    res = run(init, iterations=400)[0]
    assert res["sample_count"] == 300
    assert {"acceptance_ratio", "step_size"} <= set(res["samples"].keys())
    check(res["samples"]["variables"]["w"][::3])      # thin: successive MH samples correlate


def test_adaptive_step_size_through_sggmc(gpu):
  """alias.sggmc(adaptive_step_size=True): the dual-averaging schedule moves the
  step size during burn in and freezes it afterwards (scheduler.py:376-444)."""
  from jax_sgmc_b200 import alias
  loader, pot, full, init, check = _ks_problem()
  run = alias.sggmc(pot, full, loader, cache_size=1, batch_size=1, first_step_size=0.05,
                    adaptive_step_size=True, burn_in=60, target_acceptance_rate=0.6,
                    progress_bar=False)
  res = run(init, iterations=100)[0]
  eps = np.ravel(res["samples"]["step_size"])
  assert res["sample_count"] == 40
  assert np.all(eps == eps[0]) and eps[0] != np.float32(0.05)
  ratio = np.ravel(res["samples"]["acceptance_ratio"])
  assert np.all((ratio >= 0) & (ratio <= 1))


@pytest.mark.parametrize("C,d,N,mb,path", [(5, 8, 60, 16, "simt"), (64, 64, 1000, 128, "tc_parity"),
                                           (3, 8, 64, 16, "simt")])
def test_full_potential_one_call_equals_the_batch_loop(gpu, C, d, N, mb, path):
  """sgmc_glm_full_potential (all batches inside one C call) against the generic
  full_data_map loop of potential.full_potential: same bits; and the oracle."""
  from jax_sgmc_b200 import data, glm, potential
  from jax_sgmc_b200.tree_util import ChainTree
  X, y, theta = _logistic_problem(C, d, N, seed=3)
  loader = data.NumpyDataLoader(x=X, y=y)
  prior, lik = glm.GaussianPrior(3.0), glm.LogisticRegression()
  full = potential.full_potential(prior, lik, strategy="vmap", path=path, temperature=1.5)
  init, map_fn, _ = data.full_reference_data(loader, 4, mb)
  sample = ChainTree.from_trees([{"w": t} for t in theta])
  fast, _ = full(sample, init(), map_fn)
  loop_map = lambda *a, **k: map_fn(*a, **k)          # no .loader attribute: generic loop
  slow, _ = full(sample, init(), loop_map)
  assert np.array_equal(fast.numpy().view(np.uint32), slow.numpy().view(np.uint32))
  o_full = osgmc.full_potential(osgmc.Logistic(d, 0), osgmc.Prior("gaussian", 0, d, 3.0), 1.5)
  ids = np.arange(int(np.ceil(N / mb)) * mb).reshape(-1, mb)
  batches = [(X[i % N], y[i % N], (i < N).astype(np.float32)) for i in ids]
  np.testing.assert_allclose(fast.numpy(), o_full(theta, batches, N), rtol=2e-5)


def test_energies_are_reduced_in_a_fixed_order(gpu):
  """The per-chain energies of the OBABO passes and of the reversible leapfrog are sums over
  hundreds of warp-tiles that run on different CTAs: repeated runs on the same inputs must
  give the same bits (one partial per (chain, tile), added in tile order by k_energy_finish),
  the accumulation onto a non-zero start value included."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  rng = np.random.default_rng(4)
  C, sizes = 5, [70001, 3, 130000, 517]        # 785 tiles of 256 parameters per chain, ragged leaves
  P = sum(sizes)
  theta = rng.standard_normal((C, P)).astype(np.float32)
  mom = rng.standard_normal((C, P)).astype(np.float32)
  grad = (rng.standard_normal((C, P)) * 3).astype(np.float32)
  keys = _keys(C, 21)
  start = rng.standard_normal(C).astype(np.float32)

  def run(which):
    d_t, d_p, d_g = DA.from_numpy(theta), DA.from_numpy(mom), DA.from_numpy(grad)
    e = DA.from_numpy(start)
    kin, kout = DA.from_numpy(keys), DA((C, 2), np.uint32)
    if which == "a":
      ops.obabo_pass_a(d_t, d_p, d_g, e, kin, kout, sizes, 0.01, 1.2, 0.8)
    elif which == "b":
      ops.obabo_pass_b(d_p, d_g, e, kin, sizes, 0.01, 1.2, 0.8)
    else:
      ops.revleapfrog_step(d_t, d_p, d_g, e, kin, kout, sizes, 0.01, 0.3)
    return e.numpy().view(np.uint32), d_p.numpy()

  for which in ("a", "b", "leapfrog"):
    first, p_after = run(which)
    for _ in range(4):
      again, _p = run(which)
      assert np.array_equal(first, again), which
    got = first.view(np.float32)
    assert np.all(np.isfinite(got)) and not np.array_equal(got, start)
    if which == "b":   # ke_end += 0.5 <p3, p3>, p3 = the momentum before the O step: check by value
      p3 = (np.float32(-0.005) * grad + mom).astype(np.float32)
      want = start + 0.5 * (p3.astype(np.float64) ** 2).sum(axis=1)
      np.testing.assert_allclose(got, want, rtol=1e-5)
