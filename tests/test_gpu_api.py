"""GPU tests of the operator API (potential / integrator / adaption / solver /
alias mirrors): trajectories against the oracle driven through the API, and
the reference's own statistical acceptance tests (tests/test_alias.py:
Kolmogorov-Smirnov p > 0.05 for every solver; linear-regression sigma)."""
import numpy as np
import pytest
from scipy import stats as scpstats

from oracle import data as odata
from oracle import prng
from oracle import scheduler as osched
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu


def _ks_problem():
  """tests/test_alias.py:32-65: target N(0, 0.5^2); the potential does not
  depend on the data.  Expressed with the recognised families: logistic
  likelihood on all-zero features (constant) + GaussianPrior(0.5)."""
  from jax_sgmc_b200 import data, glm, potential
  x = 0.5 * prng.normal(prng.PRNGKey(11), (100,))
  loader = data.DeviceNumpyDataLoader(x=np.zeros((2, 1), np.float32),
                                      y=np.zeros((2,), np.float32))
  pot = potential.minibatch_potential(glm.GaussianPrior(0.5),
                                      glm.LogisticRegression(), strategy="vmap")
  init = {"w": np.array([x[0]], np.float32)}

  def check(samples):
    st = scpstats.kstest(np.ravel(samples), x)
    assert st.pvalue > 0.05, f"KS p-value {st.pvalue}"

  return loader, pot, init, check


def test_ks_sgld_rms(gpu):
  """tests/test_alias.py:121-142."""
  from jax_sgmc_b200 import alias
  loader, pot, init, check = _ks_problem()
  run = alias.sgld(pot, loader, cache_size=1, batch_size=1, first_step_size=0.5,
                   last_step_size=0.1, burn_in=100, accepted_samples=100,
                   rms_prop=True, progress_bar=False)
  res = run(init, iterations=2000)
  assert res[0]["sample_count"] == 100
  check(res[0]["samples"]["variables"]["w"])


def test_ks_re_sgld(gpu):
  """tests/test_alias.py:144-165."""
  from jax_sgmc_b200 import alias
  loader, pot, init, check = _ks_problem()
  run = alias.re_sgld(pot, loader, cache_size=1, batch_size=1, first_step_size=0.1,
                      last_step_size=0.05, burn_in=100, accepted_samples=100,
                      temperature=1.5, progress_bar=False)
  res = run((init, init), iterations=1500)
  check(res[0]["samples"]["variables"]["w"])


def test_ks_sghmc_and_obabo(gpu):
  """tests/test_alias.py:205-244."""
  from jax_sgmc_b200 import alias
  loader, pot, init, check = _ks_problem()
  run = alias.sghmc(pot, loader, cache_size=1, batch_size=1, first_step_size=0.1,
                    last_step_size=0.05, burn_in=100, accepted_samples=100,
                    integration_steps=5, friction=1.0, progress_bar=False)
  check(run(init, iterations=1500)[0]["samples"]["variables"]["w"])
  run = alias.obabo(pot, loader, cache_size=1, batch_size=1, first_step_size=0.2,
                    last_step_size=0.1, burn_in=100, accepted_samples=100,
                    integration_steps=5, friction=1.0, progress_bar=False)
  check(run(init, iterations=1500)[0]["samples"]["variables"]["w"])


def test_quickstart_sgld_posterior(gpu):
  """BASELINE.json configs[0]: examples/quickstart Bayesian linear regression
  with alias.sgld (rms_prop), single chain -- the setting the reference tests on
  this data (tests/test_alias.py:250-265), shortened; posterior means must sit
  on the least-squares solution / the notebook's NumPyro posterior."""
  from jax_sgmc_b200 import alias, data, glm, potential
  x, y, _ = odata.quickstart_dataset()
  loader = data.NumpyDataLoader(x=x, y=y[:, 0])
  pot = potential.minibatch_potential(prior=glm.InvSigmaPrior("log_sigma"),
                                      likelihood=glm.GaussianRegression(
                                          weights="w", log_sigma="log_sigma"),
                                      strategy="vmap")
  run = alias.sgld(pot, loader, cache_size=512, batch_size=10, first_step_size=0.05,
                   last_step_size=0.001, burn_in=6000, accepted_samples=2000,
                   rms_prop=True, progress_bar=False)
  init = {"w": np.zeros((4, 1), np.float32), "log_sigma": np.array(2.5, np.float32)}
  res = run(init, iterations=12000)[0]["samples"]["variables"]
  w = res["w"].mean(axis=0).ravel()
  sigma = np.exp(res["log_sigma"]).mean()
  w_ls = np.linalg.lstsq(x.astype(np.float64), y.astype(np.float64), rcond=None)[0].ravel()
  assert np.abs(w - w_ls).max() < 0.08, (w, w_ls)
  assert abs(sigma - 0.5) < 0.1, sigma


def test_reference_callables_run_the_quickstart_unmodified(gpu):
  """The quickstart's own model code (examples/quickstart.md:158-173: plain likelihood /
  prior callables written against jax.numpy and jax.scipy.stats) handed to
  minibatch_potential as the reference does: recognised on first use as the gaussian
  regression + 1/sigma prior, then the very same samples as the specification objects --
  also with has_state=True (potential.py:131-137) for a likelihood that keeps its state."""
  import sys
  from jax_sgmc_b200 import alias, compat, data, glm, potential
  shimmed = compat.install_jax_shim()
  import jax.numpy as jnp
  from jax.scipy.stats import norm

  def model(sample, observations):
    weights = sample["w"]
    predictors = observations["x"]
    return jnp.dot(predictors, weights)

  def likelihood(sample, observations):
    sigma = jnp.exp(sample["log_sigma"])
    y = observations["y"]
    y_pred = model(sample, observations)
    return norm.logpdf(y - y_pred, scale=sigma)

  def prior(sample):
    return 1 / jnp.exp(sample["log_sigma"])

  x, y, _ = odata.quickstart_dataset()
  init = {"w": np.zeros((4, 1), np.float32), "log_sigma": np.array(2.5, np.float32)}
  out = []
  for pot in (potential.minibatch_potential(prior=prior, likelihood=likelihood, strategy="vmap"),
              potential.minibatch_potential(prior=glm.InvSigmaPrior("log_sigma"),
                                            likelihood=glm.GaussianRegression(
                                                weights="w", log_sigma="log_sigma"),
                                            strategy="vmap", has_state=True)):
    loader = data.NumpyDataLoader(x=x, y=y)
    run = alias.sgld(pot, loader, cache_size=64, batch_size=10, first_step_size=0.05,
                     last_step_size=0.001, burn_in=100, accepted_samples=50,
                     rms_prop=True, progress_bar=False)
    out.append(run(init, iterations=300)[0]["samples"])
  assert np.array_equal(out[0]["variables"]["w"], out[1]["variables"]["w"])
  assert np.array_equal(out[0]["variables"]["log_sigma"], out[1]["variables"]["log_sigma"])
  assert np.array_equal(out[0]["likelihood"], out[1]["likelihood"])
  bad = potential.minibatch_potential(prior=prior, likelihood=lambda s, o: jnp.sum(o["x"]) * 0.0,
                                      strategy="vmap")
  run = alias.sgld(bad, data.NumpyDataLoader(x=x, y=y), cache_size=8, batch_size=10,
                   accepted_samples=5, progress_bar=False)
  with pytest.raises(TypeError):
    run(init, iterations=10)
  if shimmed:
    for name in [m for m in sys.modules if m == "jax" or m.startswith("jax.")]:
      del sys.modules[name]


def test_langevin_api_matches_oracle_trajectory(gpu):
  """integrator.langevin_diffusion + adaption.rms_prop + solver.sgmc driven
  through the API for 300 steps == the oracle, same keys, same minibatches."""
  from jax_sgmc_b200 import adaption, data, glm, integrator, potential, scheduler, solver
  d, C, n, N, K = 16, 6, 32, 400, 300
  X, y, _ = odata.logistic_dataset(N, d, seed=2)
  loader = data.DeviceNumpyDataLoader(x=X, y=y)
  pot = potential.minibatch_potential(glm.GaussianPrior(10.0), glm.LogisticRegression(),
                                      path="simt")
  batch_fn = data.random_reference_data(loader, 1, n)
  integ = integrator.langevin_diffusion(pot, batch_fn, adaption=adaption.rms_prop())
  init_fn, update_fn, get_fn = solver.sgmc(integ)
  keys = np.stack([prng.PRNGKey(c) for c in range(C)])
  state = init_fn([{"w": np.zeros(d, np.float32)} for _ in range(C)], key=keys,
                  adaption_kwargs={"alpha": 0.9, "lmbd": 1e-5})
  eps = osched.polynomial_step_size_first_last(K, 2e-2, 2e-3)
  for k in range(K):
    state, _ = update_fn(state, scheduler.schedule(eps[k], 1.0, 1.0, True))
  # oracle
  opot = osgmc.minibatch_potential(osgmc.Logistic(d, 0),
                                   osgmc.Prior("gaussian", 0, d, 10.0))
  st = osgmc.langevin_init(np.zeros((C, d), np.float32), keys, rms=True)
  dk = prng.PRNGKey(0)
  for k in range(K):
    dk, idx = odata.device_draw(dk, n, N)
    Xb, yb = X[idx], y[idx]
    st = osgmc.langevin_update(st, lambda th: opot(th, (Xb, yb), N), [d], eps[k], 1.0)
  got = get_fn(state)["variables"].to_host()["w"]
  assert np.abs(got - st.theta).max() / np.abs(st.theta).max() < 1e-5
  np.testing.assert_allclose(get_fn(state)["likelihood"].numpy(), -st.potential, rtol=1e-5)
  assert np.array_equal(state.key.numpy(), st.key)


def test_adaption_triplet_matches_oracle(gpu):
  """adaption.rms_prop (init, update, get) stand-alone == adaption.py:238-291."""
  from jax_sgmc_b200 import adaption
  from jax_sgmc_b200.device import DeviceArray
  from jax_sgmc_b200.tree_util import ChainTree
  rng = np.random.default_rng(0)
  sample = ChainTree.from_trees([{"a": rng.standard_normal(5), "b": rng.standard_normal((2, 3))}
                                 for _ in range(3)])
  g = rng.standard_normal((3, 11)).astype(np.float32)
  init, update, get = adaption.rms_prop()
  st = init(sample, alpha=0.8, lmbd=1e-3)
  st = update(st, sample, ChainTree.like(sample, DeviceArray.from_numpy(g)))
  man = get(st)
  ov = osgmc.rms_prop_update(osgmc.rms_prop_init(np.zeros((3, 11)), 0.8, 1e-3), g)
  oG, oS, _ = osgmc.rms_prop_get(ov)
  assert man.g_inv.ndim == 1
  assert np.array_equal(st.v.flat.numpy(), ov[0])
  assert np.array_equal(man.g_inv.tensor.flat.numpy(), oG)
  assert np.array_equal(man.sqrt_g_inv.tensor.flat.numpy(), oS)
  assert set(man.g_inv.tensor.to_host()) == {"a", "b"}


def test_host_loader_gives_each_chain_its_own_stream(gpu):
  """Reference semantics: NumpyDataLoader chains are seeded by the chain id
  (numpy_loader.py:263), so two chains of one run see different minibatches."""
  from jax_sgmc_b200 import data, glm, integrator, potential, scheduler
  X, y, _ = odata.logistic_dataset(50, 8, seed=4)
  loader = data.NumpyDataLoader(x=X, y=y)
  batch_fn = data.random_reference_data(loader, 4, 5)
  pot = potential.minibatch_potential(glm.FlatPrior(), glm.LogisticRegression())
  init_fn, update_fn, get_fn = integrator.langevin_diffusion(pot, batch_fn)
  state = init_fn([{"w": np.zeros(8, np.float32)}] * 2)
  state = update_fn(state, scheduler.schedule(0.01, 1.0, 1.0, True))
  U = state.potential.numpy()
  opot = osgmc.minibatch_potential(osgmc.Logistic(8, 0), osgmc.Prior("flat"))
  for c in range(2):
    idx = odata.HostDraws(50, 5, seed=c).draw()
    wU, _, _ = opot(np.zeros((1, 8), np.float32), (X[idx], y[idx]), 50)
    np.testing.assert_allclose(U[c], wU[0], rtol=1e-5)


@pytest.mark.parametrize("direct", [True, False])
def test_sample_ring_equals_device_buffer(gpu, direct, monkeypatch):
  """io.save through the host ring (kept samples stream to pinned host memory
  on a copy stream while the chains keep stepping) returns
  exactly what the all-in-HBM buffer returns."""
  from jax_sgmc_b200 import adaption, data, glm, integrator, io, potential, scheduler, solver
  from jax_sgmc_b200.tree_util import ChainTree
  monkeypatch.setattr(io, "PINNED_RESULT_LIMIT", (24 << 30) if direct else 0)
  X, y, _ = odata.logistic_dataset(500, 16, seed=2)
  loader = data.DeviceNumpyDataLoader(x=X, y=y)
  pot = potential.minibatch_potential(glm.GaussianPrior(5.0), glm.LogisticRegression(),
                                      strategy="vmap", path="simt")
  init = [{"w": np.full(16, 0.01 * c, np.float32)} for c in range(7)]

  def run(collector):
    integ = integrator.langevin_diffusion(pot, data.random_reference_data(loader, 1, 32),
                                          adaption.rms_prop())
    sched = scheduler.init_scheduler(
        step_size=scheduler.polynomial_step_size_first_last(first=1e-2, last=1e-3),
        burn_in=scheduler.initial_burn_in(10), progress_bar=False)
    slv = solver.sgmc(integ)
    mcmc = solver.mcmc(slv, sched, saving=io.save(collector))
    return mcmc(slv[0](ChainTree.from_trees(init)), iterations=60)

  a = run(io.MemoryCollector(stream_to_host=False))
  b = run(io.MemoryCollector(stream_to_host=True))
  assert len(a) == len(b) == 7 and a[0]["sample_count"] == b[0]["sample_count"] == 50
  for ra, rb in zip(a, b):
    assert np.array_equal(ra["samples"]["variables"]["w"], rb["samples"]["variables"]["w"])
    assert np.array_equal(ra["samples"]["likelihood"], rb["samples"]["likelihood"])
  assert not np.array_equal(a[0]["samples"]["variables"]["w"][0],
                            a[0]["samples"]["variables"]["w"][-1])


@pytest.mark.parametrize("link", ["pull", "staged", "hybrid"])
def test_streaming_loader_equals_resident_loader(gpu, monkeypatch, link):
  """StreamingNumpyDataLoader feeds the chains the same minibatches as
  the HBM-resident NumpyDataLoader: identical samples, several cache refills deep --
  with the rows gathered on the host into pinned memory and copied by the DMA engine
  (default) and with the rows pulled by the GPU out of the mapped host arrays."""
  from jax_sgmc_b200 import alias, data, glm, potential
  monkeypatch.setenv("SGMC_HOST_PULL", "1" if link == "pull" else "0")
  monkeypatch.setenv("SGMC_HOST_PULL_FRACTION", "0.34" if link == "hybrid" else "0")
  X, y, _ = odata.logistic_dataset(700, 16, seed=4)
  pot = potential.minibatch_potential(glm.GaussianPrior(5.0), glm.LogisticRegression(),
                                      strategy="vmap", path="simt")
  init = {"w": np.zeros(16, np.float32)}
  out = []
  for cls in (data.NumpyDataLoader, data.StreamingNumpyDataLoader):
    loader = cls(x=X, y=y)
    run = alias.sgld(pot, loader, cache_size=7, batch_size=24, first_step_size=1e-2,
                     last_step_size=1e-3, burn_in=5, accepted_samples=40, rms_prop=True,
                     progress_bar=False)
    out.append(run(init, iterations=60)[0])
  a, b = out
  assert a["sample_count"] == b["sample_count"] == 40
  assert np.array_equal(a["samples"]["variables"]["w"], b["samples"]["variables"]["w"])
  assert np.array_equal(a["samples"]["likelihood"], b["samples"]["likelihood"])
  assert pot.host_link_mode.startswith(link)
  # many chains share the one stream of minibatches
  loader = data.StreamingNumpyDataLoader(x=X, y=y)
  run = alias.sgld(pot, loader, cache_size=4, batch_size=24, first_step_size=1e-2,
                   last_step_size=1e-3, accepted_samples=10, progress_bar=False)
  res = run(init, init, init, iterations=20, keys=np.stack([prng.PRNGKey(i) for i in range(3)]))
  assert len(res) == 3 and res[2]["samples"]["variables"]["w"].shape == (10, 16)
  assert not np.array_equal(res[0]["samples"]["variables"]["w"],
                            res[1]["samples"]["variables"]["w"])


@pytest.mark.parametrize("loader_kind,chains", [("device", 3), ("host", 1)])
@pytest.mark.parametrize("rms", [True, False])
def test_native_scan_equals_the_python_step_loop(gpu, monkeypatch, loader_kind, chains, rms):
  """solver.mcmc hands the whole scan to sgmc_glm_sgld_scan_device when it can
  (static schedules, langevin on a GLM potential, shared minibatches); the
  samples must be those of the step-by-step loop, bit for bit -- burn in,
  thinning, odd iteration counts, cache refills of the host loader included."""
  from jax_sgmc_b200 import alias, data, glm, potential
  X, y, _ = odata.logistic_dataset(300, 8, seed=7)
  pot = potential.minibatch_potential(glm.GaussianPrior(4.0), glm.LogisticRegression(),
                                      strategy="vmap")
  init = [{"w": np.full(8, 0.05 * c, np.float32)} for c in range(chains)]
  keys = np.stack([prng.PRNGKey(40 + c) for c in range(chains)])
  out = []
  for native in ("1", "0"):
    monkeypatch.setenv("SGMC_NATIVE_SCAN", native)
    cls = data.DeviceNumpyDataLoader if loader_kind == "device" else data.NumpyDataLoader
    loader = cls(x=X, y=y)
    run = alias.sgld(pot, loader, cache_size=1 if loader_kind == "device" else 7,
                     batch_size=16, first_step_size=2e-2, last_step_size=2e-3, burn_in=11,
                     accepted_samples=23, rms_prop=rms, progress_bar=False)
    out.append(run(*init, iterations=77, keys=keys))
  for a, b in zip(*out):
    assert a["sample_count"] == b["sample_count"] == 23
    assert np.array_equal(a["samples"]["variables"]["w"], b["samples"]["variables"]["w"])
    assert np.array_equal(a["samples"]["likelihood"], b["samples"]["likelihood"])


@pytest.mark.parametrize("iterations", [100, 1000])
def test_random_thinning_respects_burn_in(gpu, iterations):
  """tests/test_scheduler.py:125-160: with an arbitrary burn-in mask, exactly
  `selections` iterations are accepted and none of them is subject to burn in;
  iterations with larger step sizes are preferred (probability ~ step size)."""
  from jax_sgmc_b200 import scheduler
  accepted = (prng.uniform(prng.PRNGKey(0), (iterations,)) < 0.3)
  nonzero = np.nonzero(accepted)[0]
  burn_in = scheduler.specific_scheduler(lambda *a: (None, int(accepted.sum())),
                                         lambda *a, **k: None,
                                         lambda _, it, **k: bool(accepted[it]))
  step_size = scheduler.polynomial_step_size(a=1.0, b=1.0, gamma=1.0)
  selections = int(0.5 * nonzero.size)
  thinning = scheduler.random_thinning(step_size_schedule=step_size, burn_in_schedule=burn_in,
                                       selections=selections)
  state, total = thinning.init(iterations)
  chosen = [i for i in range(iterations) if thinning.get(state, i)]
  assert total == selections and len(chosen) == selections
  assert set(chosen) <= set(nonzero.tolist())
  # step size ~ 1/(1+i): early eligible iterations are (much) more likely to be kept
  assert np.mean(chosen) < np.mean(nonzero)
  # deterministic (default key PRNGKey(0))
  state2, _ = thinning.init(iterations)
  assert [i for i in range(iterations) if thinning.get(state2, i)] == chosen


def test_pull_rows_reads_the_mapped_host_array(gpu):
  """sgmc_host_register + sgmc_pull_rows: rows (and labels) of a host array the GPU
  reads over the host link by index -- whole batches, a rank's row slice, row lengths
  that are not a multiple of 4 floats."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  rng = np.random.default_rng(5)
  for N, d, n in ((5000, 256, 64), (777, 10, 33)):
    X = rng.standard_normal((N, d)).astype(np.float32)
    y = rng.standard_normal(N).astype(np.float32)
    Xm, ym = ops.host_register(X), ops.host_register(y)
    assert ops.host_register(X) == Xm                      # idempotent
    idx = rng.integers(0, N, n).astype(np.int32)
    d_idx = DA.from_numpy(idx)
    dst, lab = DA.zeros((n, d)), DA.zeros((n,))
    ops.pull_rows(Xm, ym, d_idx, n, 0, n, d, dst, lab)
    assert np.array_equal(dst.numpy(), X[idx])
    assert np.array_equal(lab.numpy(), y[idx])
    dst2 = DA.zeros((n, d))
    r0, rows = n // 3, n // 2
    ops.pull_rows(Xm, None, d_idx, n, r0, rows, d, dst2, n_ctas=3)
    want = np.zeros((n, d), np.float32)
    want[r0:r0 + rows] = X[idx[r0:r0 + rows]]
    assert np.array_equal(dst2.numpy(), want)
    ops.host_unregister(X)
    ops.host_unregister(y)
