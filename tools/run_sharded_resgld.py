#!/usr/bin/env python
"""Sharded reSGLD over N GPUs (BASELINE.json configs[3]): one replica per rank,
NCCL all-gather of the (U, var) rows, label exchange.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P tools/run_sharded_resgld.py [--check] [--systems B] [--steps K]

--check (N = 2): the cold-chain samples of the 2-GPU run equal a single-process
2-replica run (tempering.sharded_tempering on LocalCommunicator) bit for bit.
Without --check: timing, system-steps/s over all ranks (max over ranks).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from jax_sgmc_b200 import data, device, dist, glm, integrator, ops, potential, scheduler, tempering  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream  # noqa: E402


def build(d, n, N, path):
  X, y, _ = ops.synth_logistic_data(0, N, d)
  loader = data.DeviceNumpyDataLoader(x=X, y=y)
  pot = potential.minibatch_potential(glm.GaussianPrior(10.0), glm.LogisticRegression(),
                                      path=path)
  batch_fn = data.random_reference_data(loader, 1, n)
  return integrator.langevin_diffusion(pot, batch_fn)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--check", action="store_true")
  ap.add_argument("--systems", type=int, default=512)
  ap.add_argument("--features", type=int, default=1024)
  ap.add_argument("--batch", type=int, default=1024)
  ap.add_argument("--observations", type=int, default=100000)
  ap.add_argument("--steps", type=int, default=100)
  ap.add_argument("--path", default="auto")
  ap.add_argument("--p2p", action="store_true",
                  help="peer-memory all-gather (NVLink stores + flags) instead of NCCL")
  ap.add_argument("--overlap", action="store_true",
                  help="run the exchange on a second stream under the next step's potential")
  a = ap.parse_args()
  rank, world, local = dist.env_rank_world()
  device.set_device(local)
  stream = Stream.create()
  device.set_current_stream(stream)
  ctl = dist.SocketCommunicator() if world > 1 else None
  R = max(world, 2)
  temps = list(np.geomspace(1.0, 1000.0, R).astype(np.float32))
  if a.check:
    a.systems, a.features, a.batch, a.observations, a.steps, a.path = 10, 16, 32, 400, 60, "simt"
  B, d = a.systems, a.features
  if world == 1:
    comm = dist.LocalCommunicator()
  elif a.p2p:
    comm = dist.PeerCommunicator(ctl, (R // world) * 2 * B * 4)
  else:
    comm = dist.NcclCommunicator.from_control_plane(ctl)
  integ = build(d, a.batch, a.observations, a.path)
  init, update, get = tempering.sharded_tempering(integ, temps, comm,
                                                  overlap_exchange=a.overlap)
  rng = np.random.default_rng(0)
  samples = [[{"w": (rng.standard_normal(d) * 0.1).astype(np.float32)} for _ in range(B)]
             for _ in range(R)]
  keys = np.stack([ops.prng_key(100 + b) for b in range(B)])
  state = init(samples, key=keys)
  sch = scheduler.schedule(np.float32(1e-3), np.float32(1.0), 1.0, True)

  if a.check:
    assert world == 2
    hist = []
    for _ in range(a.steps):
      state, _ = update(state, sch)
      hist.append((get(state)["variables"][0].flat.numpy(), state.temp_index.numpy()[0],
                   state.exchange.numpy().copy()))
    # reference: same ladder in ONE process (LocalCommunicator), run on rank 0's GPU
    integ1 = build(d, a.batch, a.observations, a.path)
    init1, update1, get1 = tempering.sharded_tempering(integ1, temps)
    st1 = init1(samples, key=keys)
    ok = True
    for k in range(a.steps):
      st1, _ = update1(st1, sch)
      th = get1(st1)["variables"][rank].flat.numpy()
      ti = st1.temp_index.numpy()[rank]
      ok &= np.array_equal(th, hist[k][0]) and np.array_equal(ti, hist[k][1])
      ok &= np.array_equal(st1.exchange.numpy(), hist[k][2])
    n_ex = int(sum(h[2].sum() for h in hist))
    print(f"rank {rank}: sharded == single-process: {ok}; exchanges {n_ex}", flush=True)
    assert ok and n_ex > 0
    if a.p2p:
      assert comm.timeouts() == 0
    ctl.barrier()
    return

  for _ in range(5):
    state, _ = update(state, sch)
  stream.sync()
  if ctl:
    ctl.barrier()
  e0, e1 = Event(), Event()
  e0.record(stream)
  for _ in range(a.steps):
    state, _ = update(state, sch)
  state.wait()
  e1.record(stream)
  e1.sync()
  ms = e0.elapsed_ms(e1)
  if ctl:
    ms = ctl.max(ms)
  if rank == 0:
    print(json.dumps({"workload": "reSGLD ladder, one replica per GPU", "n_gpus": world,
                      "overlap_exchange": bool(a.overlap),
                      "exchange": "p2p" if a.p2p else ("nccl" if world > 1 else "local"),
                      "replicas": R, "systems": B, "features": d, "batch": a.batch,
                      "ms_per_step": ms / a.steps,
                      "replica_chain_steps_per_s": R * B * a.steps / (ms * 1e-3)}), flush=True)


if __name__ == "__main__":
  main()
