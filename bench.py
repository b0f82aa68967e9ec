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
  p.add_argument("--fused-epilogue", action="store_true",
                 help="experimental: pSGLD update inside the gradient GEMM's epilogue "
                      "(SGMC_OPT_FUSED_STEP_EPILOGUE)")
  p.add_argument("--cpu-seconds", type=float, default=15.0)
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


class Control:
  """Barrier + max-over-ranks on the host control plane (torch.distributed,
  gloo).  Plumbing only: no data-path collective exists for sharded chains."""

  def __init__(self, world):
    self.world = world
    if world > 1:
      import torch.distributed as dist
      os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
      dist.init_process_group("gloo")
      self.dist = dist

  def barrier(self):
    if self.world > 1:
      self.dist.barrier()

  def max(self, x: float) -> float:
    if self.world == 1:
      return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
    return float(t[0])

  def sum(self, x: float) -> float:
    if self.world == 1:
      return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
    return float(t[0])

  def close(self):
    if self.world > 1:
      self.dist.destroy_process_group()


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
      "config": {"workload": "C2 Bayesian logistic regression pSGLD (bounded CPU sample)",
                 "chains": 256, "features": args.features, "batch": args.batch},
      "cpu_baseline": base,
      "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
      "note": "jax is not installable in this image: the reference arm times "
              "the NumPy restatement of the reference (oracle/), not jax-sgmc",
  }
  print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def run_b200(args):
  from jax_sgmc_b200 import _lib, device, ops
  from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream
  rank, world, local = dist_env()
  ctl = Control(world)
  _lib.load()
  device.set_device(local)
  stream = Stream.create()
  device.set_current_stream(stream)
  if args.serial_launch:
    ops.set_option(ops.OPT_SERIAL_LAUNCH, 1)
  if args.fused_epilogue:
    ops.set_option(ops.OPT_FUSED_STEP_EPILOGUE, 1)

  C, d, n, N = args.chains, args.features, args.batch, args.observations
  path = args.path
  if path == "auto":
    # tensor-core path at fp32-level parity (fp16 hi/lo split operands)
    path = os.environ.get("SGMC_BENCH_PATH", "tc_parity")
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
  state = {"k": 0}

  def step():
    k = state["k"]
    ops.minibatch_draw(dkey[k % 2], dkey[(k + 1) % 2], idx, N)
    # the whole langevin_diffusion step in one C call (operand prepare, two
    # tcgen05 GEMMs, fused noise + pSGLD update)
    ops.glm_sgld_step(spec, theta, X, y, idx, N, U, var, grad, keys[k % 2],
                      keys[(k + 1) % 2], eps, 1.0, v=v, alpha=0.9, lmbd=1e-5,
                      workspace=ws, path=path, write_grad=False)
    state["k"] = k + 1

  # Clock sampler runs from the warm-up on, so that every nvidia-smi sample is
  # taken under the same load as the timed region that follows immediately; the
  # warm-up is W steps, extended to ~0.4 s so the sampler is up before timing.
  sampler = ClockSampler(local)
  sampler.start()
  t_w = time.perf_counter()
  n_w = 0
  while n_w < max(3, args.warmup) or time.perf_counter() - t_w < 0.4:
    step()
    n_w += 1
    if n_w % 50 == 0:
      stream.sync()
  stream.sync()

  # ---- timed region: exactly K steps ----------------------------------------
  ctl.barrier()
  device.synchronize()
  e0, e1 = Event(), Event()
  l0 = ops.launch_count()
  e0.record(stream)
  for _ in range(args.steps):
    step()
  e1.record(stream)
  e1.sync()
  device.synchronize()
  launches = ops.launch_count() - l0
  ms = e0.elapsed_ms(e1)
  ctl.barrier()
  clocks = sampler.stop()
  ms = ctl.max(ms)
  value = world * C * args.steps / (ms * 1e-3)

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
  alg_bytes = C * d * BYTES_PER_PARAM["sgld_rms"]
  achieved = alg_bytes / (upd_ms * 1e-3) / 1e9
  peak = peaks.get("hbm_gbs", 6650.0)
  roofline = {"bound": "hbm", "kernel": "k_noise_pass<SgldOp<rms>>",
              "achieved": achieved, "peak": peak, "unit": "GB/s",
              "frac": achieved / peak, "traffic": None,
              "us_per_launch": upd_ms * 1e3,
              "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
              "algorithmic_bytes_per_launch": alg_bytes}

  # ---- tensor-pipe roofline of the GLM potential op (prepare + 2 tcgen05 GEMMs) -
  roofline_tensor = None
  if path != "simt":
    e0.record(stream)
    for _ in range(reps):
      ops.glm_potential_grad(spec, theta, X, y, idx, N, U, var, grad, workspace=ws,
                             path=path)
    e1.record(stream)
    e1.sync()
    pot_ms = e0.elapsed_ms(e1) / reps
    flops = 4.0 * n * d * C                      # algorithmic: 2ndC forward + 2ndC backward
    tf = flops / (pot_ms * 1e-3) / 1e12
    tpeak = peaks.get("bf16_tflops_sustained", 1400.0)
    roofline_tensor = {
        "bound": "tensor", "kernel": "k_glm_tc_gemm (x2) + k_prepare_all",
        "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
        "us_per_call": pot_ms * 1e3, "algorithmic_flops_per_call": flops,
        "tensor_passes": 3 if path == "tc_parity" else 1,
        "note": "peak = measured sustained bf16 GEMM; the parity path spends 3 "
                "fp16 MMA passes per algorithmic FLOP (fp32-level accuracy)"}

  # ---- e2e: host data loader path ---------------------------------------------
  # Every step's minibatch rows travel host -> device from pinned memory and the
  # step's result (U, var per chain) travels device -> host, all inside the timed
  # region.  The loader prefetches: the copy of batch k+1 runs on a copy stream
  # while step k computes (double-buffered), and the host reads step k's result
  # after launching step k+1 -- the minibatch sequence does not depend on the
  # chain state, so this is the natural pipelining of the reference's host cache
  # (data/core.py:664-791).
  import ctypes as Ct
  copy_stream = Stream.create()
  hX = [Ct.c_void_p(), Ct.c_void_p()]
  hy = [Ct.c_void_p(), Ct.c_void_p()]
  hU = [Ct.c_void_p(), Ct.c_void_p()]
  Xb = [ops.gather_rows(X, idx), DA((n, d), np.float32)]
  yb = [ops.gather_rows(y.reshape(N, 1), idx), DA((n, 1), np.float32)]
  for b in range(2):
    _lib.call("sgmc_host_alloc", Ct.byref(hX[b]), n * d * 4)
    _lib.call("sgmc_host_alloc", Ct.byref(hy[b]), n * 4)
    _lib.call("sgmc_host_alloc", Ct.byref(hU[b]), C * 4 * 2)
    _lib.call("sgmc_memcpy_d2h", hX[b], Ct.c_void_p(Xb[0].ptr), n * d * 4, stream.handle)
    _lib.call("sgmc_memcpy_d2h", hy[b], Ct.c_void_p(yb[0].ptr), n * 4, stream.handle)
  stream.sync()
  copied = [Event(), Event()]
  computed = [Event(), Event()]
  result = [Event(), Event()]

  def enqueue_copy(j):
    b = j % 2
    copy_stream.wait_event(computed[b])          # buffer b free again (step j-2 done)
    _lib.call("sgmc_memcpy_h2d", Ct.c_void_p(Xb[b].ptr), hX[b], n * d * 4, copy_stream.handle)
    _lib.call("sgmc_memcpy_h2d", Ct.c_void_p(yb[b].ptr), hy[b], n * 4, copy_stream.handle)
    copied[b].record(copy_stream)

  def enqueue_step(j):
    b, k = j % 2, state["k"]
    stream.wait_event(copied[b])
    ops.glm_potential_grad(spec, theta, Xb[b], yb[b], None, N, U, var, grad,
                           workspace=ws, path=path, batch_size=n)
    ops.sgld_update(theta, grad, keys[k % 2], keys[(k + 1) % 2], [d], eps, 1.0,
                    v=v, alpha=0.9, lmbd=1e-5)
    computed[b].record(stream)
    _lib.call("sgmc_memcpy_d2h", hU[b], Ct.c_void_p(U.ptr), C * 4, stream.handle)
    _lib.call("sgmc_memcpy_d2h", Ct.c_void_p(hU[b].value + C * 4), Ct.c_void_p(var.ptr),
              C * 4, stream.handle)
    result[b].record(stream)
    state["k"] = k + 1

  def run_e2e(steps):
    for b in range(2):
      computed[b].record(stream)
    enqueue_copy(0)
    for j in range(steps):
      if j + 1 < steps:
        enqueue_copy(j + 1)                      # prefetch the next batch
      enqueue_step(j)
      if j >= 1:
        result[(j - 1) % 2].sync()               # host consumes step j-1's result
    result[(steps - 1) % 2].sync()

  run_e2e(4)
  ctl.barrier()
  e2e_steps = max(10, args.steps // 2)
  t0 = time.perf_counter()
  run_e2e(e2e_steps)
  e2e_s = ctl.max(time.perf_counter() - t0)
  e2e = {"value": world * C * e2e_steps / e2e_s, "unit": UNIT,
         "h2d_bytes_per_step": n * d * 4 + n * 4, "d2h_bytes_per_step": C * 8,
         "steps": e2e_steps,
         "what": "host data loader: every step's minibatch rows H2D from pinned host "
                 "memory (prefetched one step ahead on a copy stream), potential + "
                 "variance of every chain D2H and read by the host every step"}

  cpu_base = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cpu_base, _, _ = cpu_reference_run(args, args.cpu_seconds)

  if rank == 0:
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C2: Bayesian logistic regression, pSGLD "
                               "(alias.sgld rms_prop=True), chains sharded over GPUs",
                   "chains_per_gpu": C, "features": d, "batch": n,
                   "observations": N, "potential_path": path,
                   "l2": "state (theta,v,grad = 50 MB/GPU) is reused every step by "
                         "the algorithm itself; the roofline kernel is timed on "
                         f"{R} rotating state sets ({R * 3 * C * d * 4 / 1e6:.0f} MB > L2)",
                   "parallelism": f"chains x{world}"},
        "clocks": clocks, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_tensor": roofline_tensor, "e2e": e2e,
        "cpu_baseline": cpu_base,
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
