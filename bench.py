#!/usr/bin/env python
"""Benchmark of the jax-sgmc sampling hot path on B200 (BASELINE.json metric).

Workload (configs[1], "C2"): Bayesian logistic regression, N = 1M synthetic
observations x 1024 features resident in HBM, minibatch 1024 shared by all
chains (the reference default: every chain's data key is PRNGKey(0)), 4096
parallel pSGLD (SGLD + RMSprop) chains per GPU, chain c seeded PRNGKey(c).
One "step" = one pass of the hot path over all chains of a rank:
    minibatch draw (threefry randint)  ->  GLM potential + gradient
    ->  fused pSGLD update with in-kernel jax.random noise.
metric = chain-steps/sec (chains x steps / seconds), whole job over all GPUs.
N > 1: chains are sharded over ranks (independent chains, no collective:
"weak" scaling, 4096 chains per GPU).

JSON keys follow the driver contract; see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "chain-steps/sec (BLR, 4096 chains)"
UNIT = "chain-steps/s"
BYTES_PER_PARAM = {"sgld": 12, "sgld_rms": 20}


def parse():
  p = argparse.ArgumentParser()
  p.add_argument("--gpus", type=int, default=1)
  p.add_argument("--steps", type=int, default=2000)
  p.add_argument("--warmup", type=int, default=10)
  p.add_argument("--impl", default="b200", choices=["b200", "reference"])
  p.add_argument("--chains", type=int, default=4096)
  p.add_argument("--features", type=int, default=1024)
  p.add_argument("--batch", type=int, default=1024)
  p.add_argument("--observations", type=int, default=1_000_000)
  p.add_argument("--path", default="auto",
                 choices=["auto", "simt", "tc_parity", "tc_throughput"])
  p.add_argument("--no-cpu-baseline", action="store_true")
  p.add_argument("--serial-launch", action="store_true",
                 help="disable programmatic dependent launch (A/B of the launch overlap)")
  p.add_argument("--no-resgld", action="store_true")
  p.add_argument("--resgld-systems", type=int, default=4096)
  p.add_argument("--resgld-steps", type=int, default=200)
  p.add_argument("--cpu-seconds", type=float, default=15.0)
  p.add_argument("--resgld-overlap", type=int, default=int(os.environ.get("SGMC_RESGLD_OVERLAP", "-1")),
                 help="1: run the ladder's exchange on a second stream under the next step's "
                      "potential (default: on for N > 1)")
  p.add_argument("--only-resgld", action="store_true",
                 help="tool mode: print only the reSGLD record (not the contract line)")
  return p.parse_args()


# ---------------------------------------------------------------------------
class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled during the timed region."""

  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
       "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index: int):
    self.gpu, self.proc, self.lines = gpu_index, None, []

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
           "--format=csv,noheader,nounits", "-lms", "50"],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.15)
    self.proc.terminate()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in self.lines:
      f = [x.strip() for x in ln.split(",")]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); mx.append(float(f[2]))
      except ValueError:
        continue
      for name, val in zip(names, f[5:9]):
        if val.lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  return rank, world, local


# ---------------------------------------------------------------------------
def cpu_reference_run(args, seconds: float, max_chains: int = 256, max_steps: int = 10000,
                      warmup: int = 0):
  """The oracle port timed on a bounded sample of the same workload on all host
  cores: `max_chains` chains of the C2 model, same d / batch; one step = draw
  the minibatch, potential + gradient (NumPy f32 on the threaded BLAS), noise +
  pSGLD update (the oracle's C restatement, threaded over chain slices).  At
  most `max_steps` steps and at most `seconds` of CPU time."""
  from concurrent.futures import ThreadPoolExecutor
  from oracle import cnative
  from oracle import data as odata
  from oracle import prng
  from oracle import sgmc as osgmc
  d, n = args.features, args.batch
  Ns = 20000                                   # bounded data sample
  rng = np.random.default_rng(0)
  X = (rng.standard_normal((Ns, d)) / np.sqrt(d)).astype(np.float32)
  w = rng.standard_normal(d).astype(np.float32)
  y = (rng.random(Ns) < 1 / (1 + np.exp(-(X @ w)))).astype(np.float32)
  C = max_chains
  cores = os.cpu_count() or 1
  keys = np.ascontiguousarray(np.stack([prng.PRNGKey(c) for c in range(C)]))
  theta = np.zeros((C, d), np.float32)
  v = np.ones((C, d), np.float32)
  pot = osgmc.minibatch_potential(osgmc.Logistic(d, 0),
                                  osgmc.Prior("gaussian", 0, d, 10.0))
  cnative.load()
  bounds = np.linspace(0, C, min(cores, C) + 1).astype(int)
  slices = [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
  pool = ThreadPoolExecutor(max_workers=len(slices))
  state = {"dk": prng.PRNGKey(0)}

  def one_step():
    state["dk"], idx = odata.device_draw(state["dk"], n, Ns)
    _, _, g = pot(theta, (X[idx], y[idx]), args.observations)
    g = np.ascontiguousarray(g)
    list(pool.map(lambda ab: cnative.sgld_step(theta[ab[0]:ab[1]], v[ab[0]:ab[1]],
                                               g[ab[0]:ab[1]], keys[ab[0]:ab[1]], [d],
                                               1e-3, 1.0), slices))

  for _ in range(warmup):
    one_step()
  steps, t0 = 0, time.perf_counter()
  while True:
    one_step()
    steps += 1
    el = time.perf_counter() - t0
    if el >= seconds or steps >= max_steps:
      break
  pool.shutdown()
  return {"value": C * steps / el, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"{C} chains x {steps} pSGLD steps, d={d}, batch={n}: oracle port "
                    f"(NumPy f32 + threaded BLAS potential, C noise/update on {len(slices)} "
                    f"threads), {el:.1f} s"}, steps, el


WORKLOAD = ("C2: Bayesian logistic regression, pSGLD (alias.sgld rms_prop=True), "
            "chains sharded over GPUs")


def run_reference(args):
  rank, world, _ = dist_env()
  if rank != 0:
    return
  # K timed steps after W warm-up steps of the bounded sample, capped at 150 s
  base, steps, el = cpu_reference_run(args, 150.0, max_steps=args.steps,
                                      warmup=min(args.warmup, 20))
  line = {
      "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
      "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 20),
      "ms_per_step": 1e3 * el / steps, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      # the b200 arm's workload; each CPU step is the bounded sample named in
      # cpu_baseline.sample (256 of the 4096 chains, same d / batch / arithmetic)
      "config": {"workload": WORKLOAD, "chains_per_gpu": args.chains,
                 "features": args.features, "batch": args.batch,
                 "observations": args.observations, "cpu_sample_chains": 256,
                 "parallelism": "host threads"},
      "cpu_baseline": base,
      "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
      "note": "jax is not installable in this image: the reference arm times "
              "the NumPy restatement of the reference (oracle/), not jax-sgmc",
  }
  print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
class Control:
  """Barrier + max-over-ranks on the host control plane (torch.distributed, gloo), and
  the carrier of the NCCL unique id.  Plumbing only."""

  def __init__(self, rank, world):
    self.rank, self.world = rank, world
    if world > 1:
      import torch.distributed as dist
      os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
      dist.init_process_group("gloo")
      self.dist = dist

  def barrier(self):
    if self.world > 1:
      self.dist.barrier()

  def _reduce(self, x: float, op) -> float:
    if self.world == 1:
      return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    self.dist.all_reduce(t, op=op)
    return float(t[0])

  def max(self, x: float) -> float:
    return self._reduce(x, self.dist.ReduceOp.MAX) if self.world > 1 else x

  def sum(self, x: float) -> float:
    return self._reduce(x, self.dist.ReduceOp.SUM) if self.world > 1 else x

  def broadcast_bytes(self, payload, nbytes: int, src: int = 0) -> bytes:
    if self.world == 1:
      return payload
    box = [payload if self.rank == src else None]
    self.dist.broadcast_object_list(box, src=src)
    assert len(box[0]) == nbytes
    return box[0]

  def close(self):
    if self.world > 1:
      self.dist.destroy_process_group()


def bench_resgld(args, ctl, nccl, stream, path):
  """C4: a reSGLD ladder of 8 replicas (geometric temperatures 1 ... 1000, every replica
  `--resgld-systems` independent systems of the C2 model) sharded over the ranks: per step
  every rank advances its replicas, ONE NCCL all-gather shares the (U, var) rows, every
  rank takes the identical swap decisions.  Strong scaling: the ladder is the same at
  every N."""
  from jax_sgmc_b200 import data, dist, glm, integrator, ops, potential, scheduler, tempering
  from jax_sgmc_b200.device import Event
  world, rank = ctl.world, ctl.rank
  R, B, d, n = 8, args.resgld_systems, args.features, args.batch
  if R % world:
    return None
  Nr = 100_000
  X, y, _ = ops.synth_logistic_data(0, Nr, d)
  loader = data.DeviceNumpyDataLoader(x=X, y=y)
  pot = potential.minibatch_potential(glm.GaussianPrior(10.0), glm.LogisticRegression(), path=path)
  integ = integrator.langevin_diffusion(pot, data.random_reference_data(loader, 1, n))
  comm = nccl if world > 1 else dist.LocalCommunicator()
  temps = list(np.geomspace(1.0, 1000.0, R).astype(np.float32))
  overlap = world > 1 if args.resgld_overlap < 0 else bool(args.resgld_overlap)
  init, update, _ = tempering.sharded_tempering(integ, temps, comm, overlap_exchange=overlap)
  from jax_sgmc_b200.tree_util import ChainTree
  from jax_sgmc_b200.device import DeviceArray as DA
  template = ChainTree.from_trees([{"w": np.zeros(d, np.float32)}])
  samples = [ChainTree.like(template, DA.zeros((B, d))) for _ in range(R)]
  keys = np.stack([ops.prng_key(100 + b) for b in range(B)])
  state = init(samples, key=keys)
  sch = scheduler.schedule(np.float32(1e-3), np.float32(1.0), 1.0, True)
  for _ in range(10):
    state, _ = update(state, sch)
  stream.sync()
  ctl.barrier()
  K = args.resgld_steps
  e0, e1 = Event(), Event()
  e0.record(stream)
  for _ in range(K):
    state, _ = update(state, sch)
  state.wait()
  e1.record(stream)
  e1.sync()
  ms = ctl.max(e0.elapsed_ms(e1))
  swaps = int(state.exchange.numpy().sum())
  return {"workload": "C4: reSGLD ladder, 8 replicas sharded over the ranks "
                      "(temperature labels exchanged, one all-gather of (U, var) per step)",
          "replicas": R, "systems_per_replica": B, "features": d, "batch": n, "n_gpus": world,
          "replicas_per_gpu": R // world, "exchange": "nccl_allgather" if world > 1 else "local",
          "overlap_exchange": overlap,
          "steps": K, "us_per_step": ms * 1e3 / K,
          "replica_chain_steps_per_s": R * B * K / (ms * 1e-3), "scaling": "strong",
          "swaps_in_last_step": swaps}


def bench_row_sharded(args, ctl, nccl, stream, path):
  """C5's collective on the C2 model: the minibatch ROWS of the GLM potential sharded over
  the ranks -- every rank evaluates n / R rows for all chains, then ncclAllReduce of the
  gradient (16.8 MB) and three scalars per chain (sgmc_glm_potential_grad_row_sharded).
  Reports the time of the sharded evaluation and the all-reduce's bus bandwidth
  (2 (R-1) / R x bytes / time, the nccl-tests convention)."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA, Event
  world, rank = ctl.world, ctl.rank
  C, d, n, N = args.chains, args.features, args.batch, 100_000
  if world < 2 or n % world or N % world or (n // world) % 8:
    return None
  X, y, _ = ops.synth_logistic_data(0, N, d)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0, x_absmax=ops.absmax(X))
  theta = DA.from_numpy((np.random.default_rng(0).standard_normal((C, d)) * 0.1).astype(np.float32))
  idx = DA.from_numpy(np.random.default_rng(1).integers(0, N, n).astype(np.int32))
  U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
  comm = nccl._comm.value
  ws, scratch = ops.glm_potential_grad_row_sharded(spec, theta, X, y, idx, N, U, var, g, n, rank,
                                                   world, comm, path=path)
  e0, e1 = Event(), Event()
  K = 50

  def timed(fn):
    for _ in range(3):
      fn()
    stream.sync()
    ctl.barrier()
    e0.record(stream)
    for _ in range(K):
      fn()
    e1.record(stream)
    e1.sync()
    return ctl.max(e0.elapsed_ms(e1)) / K

  ms_all = timed(lambda: ops.glm_potential_grad_row_sharded(
      spec, theta, X, y, idx, N, U, var, g, n, rank, world, comm, workspace=ws, scratch=scratch,
      path=path))
  ms_ar = timed(lambda: nccl.allreduce_sum(g, g))
  nbytes = C * d * 4
  return {"workload": "C2 model, minibatch rows sharded over the ranks, gradient all-reduce "
                      "(the collective of configs[4])", "n_gpus": world, "rows_per_rank": n // world,
          "us_per_evaluation": ms_all * 1e3, "allreduce_bytes": nbytes,
          "allreduce_us": ms_ar * 1e3,
          "allreduce_busbw_gbs": 2 * (world - 1) / world * nbytes / (ms_ar * 1e-3) / 1e9,
          "potential_path": path}


def bench_cnn(args, ctl, nccl, stream):
  """C5 (configs[4]): the CIFAR-10-shape CNN potential + gradient (conv3x3-32 / 2, conv3x3-64
  / 2, dense-10; 60 362 parameters per chain) for 8 chains on a minibatch of 1024 images whose
  rows are sharded over the ranks, gradient and potential all-reduced (the evaluation
  alias.sggmc / alias.amagold make per leapfrog step).  Strong scaling: the minibatch is
  the same at every N."""
  from jax_sgmc_b200 import data, glm, nn, ops, potential
  from jax_sgmc_b200.device import Event
  from jax_sgmc_b200.tree_util import ChainTree
  world = ctl.world
  C, n, N = 8, 1024, 8192
  if n % world:
    return None
  rng = np.random.default_rng(0)
  loader = data.DeviceNumpyDataLoader(x=rng.random((N, 32, 32, 3), dtype=np.float32),
                                      y=rng.integers(0, 10, N).astype(np.float32))
  sample = ChainTree.from_trees(
      [nn.init_cnn_params(ops.prng_key(c), (32, 32, 3), (32, 64), (2, 2), 10) for c in range(C)])
  pot = potential.minibatch_potential(glm.GaussianPrior(10.0), nn.CNNClassifier(strides=(2, 2)))
  if nccl is not None:
    pot.shard_rows(nccl)
  init_fn, get_fn, _ = data.random_reference_data(loader, 1, n)
  state = init_fn()
  state, ref = get_fn(state, information=True)
  from jax_sgmc_b200.device import DeviceArray as DA
  # caller-owned outputs, as the integrators pass them (no allocation inside the loop)
  g_buf, U_buf = DA((C, sample.n_params), np.float32), DA((C,), np.float32)
  for _ in range(2):
    pot.value_and_grad(sample, ref, grad_out=g_buf, U_out=U_buf)
  stream.sync()
  ctl.barrier()
  K = 10
  e0, e1 = Event(), Event()
  e0.record(stream)
  for _ in range(K):
    state, ref = get_fn(state, information=True)
    pot.value_and_grad(sample, ref, grad_out=g_buf, U_out=U_buf)
  e1.record(stream)
  e1.sync()
  ms = ctl.max(e0.elapsed_ms(e1)) / K
  # forward + reverse pass: 3 x the forward FLOPs minus the first layer's input gradient
  f1, f2, f3 = 2.0 * 256 * 27 * 32, 2.0 * 64 * 288 * 64, 2.0 * 4096 * 10
  flops = C * n * (3 * (f1 + f2 + f3) - f1)
  return {"workload": "C5: CNN potential + gradient (conv3x3-32/2, conv3x3-64/2, dense-10 on "
                      "32x32x3), minibatch rows sharded over the ranks, gradient all-reduce",
          "chains": C, "batch": n, "rows_per_rank": n // world, "n_gpus": world,
          "parameters_per_chain": int(sample.n_params), "us_per_evaluation": ms * 1e3,
          "chain_evaluations_per_s": C / (ms * 1e-3), "tflops_fp32": flops / (ms * 1e-3) / 1e12,
          "scaling": "strong"}


def run_b200(args):
  from jax_sgmc_b200 import _lib, alias, data, device, dist, glm, ops, potential
  from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream
  rank, world, local = dist_env()
  ctl = Control(rank, world)
  _lib.load()
  device.set_device(local)
  stream = Stream.create()
  device.set_current_stream(stream)
  if args.serial_launch:
    ops.set_option(ops.OPT_SERIAL_LAUNCH, 1)
  nccl = dist.NcclCommunicator.from_control_plane(ctl) if world > 1 else None

  C, d, n, N = args.chains, args.features, args.batch, args.observations
  path = args.path
  if path == "auto":
    # tensor-core path at fp32-level parity (fp16 hi/lo split operands)
    path = os.environ.get("SGMC_BENCH_PATH", "tc_parity")
  if args.only_resgld:
    rec = bench_resgld(args, ctl, nccl, stream, path)
    if rank == 0:
      print(json.dumps(rec), flush=True)
    ctl.close()
    return
  pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
  peaks = json.load(open(pk)) if os.path.exists(pk) else {}

  # ---- resident state ------------------------------------------------------
  X, y, _ = ops.synth_logistic_data(0, N, d)
  theta = DA.zeros((C, d))
  v = DA.full((C, d), 1.0)
  grad = DA.zeros((C, d))
  keys = [ops.prng_keys(range(rank * C, (rank + 1) * C)), DA((C, 2), np.uint32)]
  dkey = [DA.from_numpy(ops.prng_key(0)), DA((2,), np.uint32)]
  idx = DA((n,), np.int32)
  U, var = DA((C,), np.float32), DA((C,), np.float32)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0,
                      prior_size=d, prior_scale=10.0,
                      x_absmax=ops.absmax(X))   # data-set statistic, computed once
  ws = ops.glm_workspace(C, n, d, path)
  eps = 1e-3

  def scan(k):
    # k steps of solver.mcmc's scan in ONE C call (what alias.sgld runs for a data set
    # resident in HBM): per step minibatch draw (threefry randint) + operand staging one
    # step ahead on a side stream, the tcgen05 potential kernel, the fused pSGLD update
    ops.glm_sgld_scan_device(spec, theta, X, y, N, n, U, var, grad, keys[0], keys[1], [d],
                             np.full(k, eps, np.float32), np.ones(k, np.float32),
                             np.zeros(k, np.uint8), None, None, 0, data_key_a=dkey[0],
                             data_key_b=dkey[1], idx_buf=idx, idx_all=None, v=v, alpha=0.9,
                             lmbd=1e-5, workspace=ws, path=path)
    if k % 2:                                  # the chain / data keys ping-pong once per step
      keys.reverse()
      dkey.reverse()

  # Clock sampler runs from the warm-up on, so that every nvidia-smi sample is
  # taken under the same load as the timed region that follows immediately; the
  # warm-up is W steps, extended to ~0.4 s so the sampler is up before timing.
  sampler = ClockSampler(local)
  sampler.start()
  W = max(3, args.warmup)
  scan(W)
  stream.sync()
  t_w = time.perf_counter()
  while time.perf_counter() - t_w < 0.4:
    scan(200)
    stream.sync()

  # ---- timed region: exactly K steps ----------------------------------------
  ctl.barrier()
  device.synchronize()
  e0, e1 = Event(), Event()
  l0 = ops.launch_count()
  e0.record(stream)
  scan(args.steps)
  e1.record(stream)
  e1.sync()
  device.synchronize()
  launches = ops.launch_count() - l0
  ms = e0.elapsed_ms(e1)
  ctl.barrier()
  clocks = sampler.stop()
  ms = ctl.max(ms)
  value = world * C * args.steps / (ms * 1e-3)

  # ---- where the step's time goes: CUDA events between the launches of the carried
  # step (synchronises every step -- measurement only, not the number above) -----------
  step_profile = None
  if path != "simt":
    st = {"k": 0}
    ops.set_option(ops.OPT_STEP_PROFILE, 1)
    for i in range(120):
      k = st["k"]
      ops.minibatch_draw(dkey[k % 2], dkey[(k + 1) % 2], idx, N)
      ops.glm_sgld_step(spec, theta, X, y, idx, N, U, var, grad, keys[k % 2],
                        keys[(k + 1) % 2], eps, 1.0, v=v, alpha=0.9, lmbd=1e-5,
                        workspace=ws, path=path, write_grad=False,
                        carry=ops.STEP_CARRY if k else ops.STEP_CARRY_INIT)
      st["k"] = k + 1
      if i == 19:
        ops.step_profile(reset=True)
    pr = ops.step_profile()
    ops.set_option(ops.OPT_STEP_PROFILE, 0)
    stream.sync()
    step_profile = {"us_prepare": pr[0], "us_potential": pr[1], "us_update": pr[2],
                    "steps": pr[3], "note": "every step synchronised; prepare = minibatch "
                    "operand staging (off the critical path in the scan)"}

  # ---- roofline of the dominant HBM kernel (fused pSGLD update) --------------
  # rotating state sets so the working set (R x 67 MB) exceeds the 126 MB L2
  R = 6
  sets = [(DA.zeros((C, d)), DA.full((C, d), 1.0), DA.full((C, d), 0.01))
          for _ in range(R)]
  kk = [ops.prng_keys(range(C)), DA((C, 2), np.uint32)]
  for r in range(R):
    ops.sgld_update(sets[r][0], sets[r][2], kk[0], kk[1], [d], eps, 1.0, v=sets[r][1])
  stream.sync()
  reps = 20
  e0.record(stream)
  for i in range(reps * R):
    t_, v_, g_ = sets[i % R]
    ops.sgld_update(t_, g_, kk[i % 2], kk[(i + 1) % 2], [d], eps, 1.0, v=v_)
  e1.record(stream)
  e1.sync()
  upd_ms = e0.elapsed_ms(e1) / (reps * R)
  del sets
  alg_bytes = C * d * BYTES_PER_PARAM["sgld_rms"]
  achieved = alg_bytes / (upd_ms * 1e-3) / 1e9
  peak = peaks.get("hbm_gbs", 6650.0)
  traffic = None
  tp = os.path.join(ROOT, "profiles", "r02_update_traffic.json")
  if os.path.exists(tp) and (C, d) == (4096, 1024):
    # dram__bytes_read.sum + dram__bytes_write.sum per launch, parsed from the committed
    # ncu capture (tools/summarize_ncu.py writes this file)
    traffic = json.load(open(tp)).get("k_noise_pass_rms_bytes_per_launch")
  roofline = {"bound": "hbm", "kernel": "k_noise_pass<SgldOp<rms>>",
              "achieved": achieved, "peak": peak, "unit": "GB/s",
              "frac": achieved / peak, "traffic": traffic,
              "us_per_launch": upd_ms * 1e3,
              "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
              "algorithmic_bytes_per_launch": alg_bytes}
  if step_profile is not None:
    # the update of the carried step (k_sgld_apply_split: noise pre-generated under the
    # GEMM mainloops, also emits theta's fp16 operand form): same algorithmic bytes
    roofline["in_step_update"] = {
        "kernel": "k_sgld_apply_split", "us_per_launch": step_profile["us_update"],
        "achieved": alg_bytes / (step_profile["us_update"] * 1e-6) / 1e9,
        "frac": alg_bytes / (step_profile["us_update"] * 1e-6) / 1e9 / peak}

  # ---- tensor-pipe roofline of the GLM potential kernel --------------------------
  roofline_tensor = None
  if step_profile is not None:
    flops = 4.0 * n * d * C                      # algorithmic: 2ndC forward + 2ndC backward
    passes = 3 if path == "tc_parity" else 1
    pot_us = step_profile["us_potential"]
    tf = flops / (pot_us * 1e-6) / 1e12
    tpeak = peaks.get("bf16_tflops", 1650.0)
    roofline_tensor = {
        "bound": "tensor", "kernel": "k_glm_tc_pair (both contractions, one launch)",
        "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
        "us_per_call": pot_us, "algorithmic_flops_per_call": flops,
        "tensor_passes": passes, "hardware_frac": passes * tf / tpeak,
        "note": "peak = measured burst bf16 GEMM; CUDA events around the kernel inside the "
                "step; the parity path spends 3 fp16 MMA passes per algorithmic FLOP "
                "(fp32-level accuracy), hardware_frac counts them"}

  # ---- e2e: the operator API over a HOST-resident data set -------------------------
  # alias.sgld(potential, StreamingNumpyDataLoader(...)) -> solver.mcmc -> native host
  # scan.  Inside the timed region, every step: the host gathers the minibatch's rows
  # from the 4.1 GB array (threads, into page-locked memory), the rows travel H2D on a
  # copy stream (with N ranks each rank uploads 1/N of the rows and NCCL all-gathers them
  # over NVLink), the step runs, (U, var) of every chain travel D2H; the kept sample is
  # downloaded at the end.
  hX, hy = X.numpy(), y.numpy()
  del X, y
  loader = data.StreamingNumpyDataLoader(x=hX, y=hy)
  if nccl is not None:
    loader.shard_upload(nccl)
  pot = potential.minibatch_potential(glm.GaussianPrior(10.0), glm.LogisticRegression(),
                                      path=path)
  e2e_steps = max(args.steps, 4000)
  sampler_fn = alias.sgld(pot, loader, cache_size=512, batch_size=n, first_step_size=eps,
                          last_step_size=eps / 10, burn_in=0, accepted_samples=1,
                          rms_prop=True, progress_bar=False)
  from jax_sgmc_b200.tree_util import ChainTree
  init = ChainTree.like(ChainTree.from_trees([{"w": np.zeros(d, np.float32)}]), DA.zeros((C, d)))
  chain_keys = np.stack([ops.prng_key(rank * C + c) for c in range(C)])
  sampler_fn(init, iterations=1024, keys=chain_keys)          # warm-up (buffers, page locking)
  init = ChainTree.like(init, DA.zeros((C, d)))
  # three runs of e2e_steps iterations each; the MEDIAN is reported (the host side of this
  # leg shares its cores and memory bandwidth with other tenants of the box), all three are
  # listed in e2e.runs
  e2e_runs = []
  for _ in range(3):
    init = ChainTree.like(init, DA.zeros((C, d)))
    ctl.barrier()
    t0 = time.perf_counter()
    res = sampler_fn(init, iterations=e2e_steps, keys=chain_keys)
    device.synchronize()
    e2e_runs.append(ctl.max(time.perf_counter() - t0))
  e2e_s = sorted(e2e_runs)[1]
  kept = np.asarray(res[0]["samples"]["variables"]["w"])
  assert np.all(np.isfinite(kept))
  h2d = int(getattr(pot, "h2d_bytes_per_step", 0))
  d2h = int(getattr(pot, "d2h_bytes_per_step", 0))
  assert h2d > 0 and d2h > 0, "the e2e run did not take the host-stream scan"
  link = getattr(pot, "host_link_mode", "staged")
  how = ("the GPU pulls every minibatch's rows over the host link out of the page-locked, "
         "mapped host data set by index (index rows H2D per chunk)" if link == "pull" else
         "host gather of every minibatch from the host-resident data set, H2D of the rows "
         "from pinned memory")
  e2e = {"value": world * C * e2e_steps / e2e_s, "unit": UNIT,
         "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
         "seconds": e2e_s, "host_link": link,
         "runs": [world * C * e2e_steps / t for t in e2e_runs],
         "what": "through alias.sgld(minibatch_potential, StreamingNumpyDataLoader): " + how +
                 (f" (1/{world} per rank + NCCL all-gather)" if world > 1 else "") +
                 ", (U, var) of every chain D2H every step, kept sample downloaded; wall "
                 "clock around the whole run_fn call, median of three calls"}
  del loader, hX, hy

  resgld = None
  if not args.no_resgld:
    resgld = bench_resgld(args, ctl, nccl, stream, path)
  row_sharded = bench_row_sharded(args, ctl, nccl, stream, path)
  cnn = None if args.no_resgld else bench_cnn(args, ctl, nccl, stream)

  cpu_base = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cpu_base, _, _ = cpu_reference_run(args, args.cpu_seconds)

  if rank == 0:
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": W,
        "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "chains_per_gpu": C, "features": d, "batch": n,
                   "observations": N, "potential_path": path,
                   "l2": "inputs larger than L2: one step touches theta, v, grad (50 MB), "
                         f"the operand workspace ({ws.nbytes / 1e6:.0f} MB) and fresh rows of the "
                         f"{N * d * 4 / 1e9:.1f} GB data set -- more than the 126 MB L2, no flush "
                         "needed; the roofline kernel is timed on "
                         f"{R} rotating state sets ({R * 3 * C * d * 4 / 1e6:.0f} MB)",
                   "parallelism": f"chains x{world}"},
        "clocks": clocks, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_tensor": roofline_tensor,
        "step_profile": step_profile, "e2e": e2e,
        "cpu_baseline": cpu_base, "resgld": resgld, "row_sharded_gradient": row_sharded,
        "cnn_sharded_gradient": cnn,
    }
    print(json.dumps(line), flush=True)
  ctl.close()


def main():
  args = parse()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_b200(args)


if __name__ == "__main__":
  main()
