#!/usr/bin/env python
"""Where the host time of a small-problem run goes: C1 (examples/quickstart linear
regression, one chain, alias.sgld with RMSprop) under cProfile."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jax_sgmc_b200 import alias, data, glm, potential
from oracle import data as odata

x, y, _ = odata.quickstart_dataset()
loader = data.NumpyDataLoader(x=x, y=y)
pot = potential.minibatch_potential(prior=glm.InvSigmaPrior("log_sigma"),
                                    likelihood=glm.GaussianRegression(), strategy="vmap")
run = alias.sgld(pot, loader, cache_size=512, batch_size=10, first_step_size=0.05,
                 last_step_size=0.001, burn_in=100, accepted_samples=50, rms_prop=True,
                 progress_bar=False)
init = {"w": np.zeros((4, 1), np.float32), "log_sigma": np.array(2.5, np.float32)}
run(init, iterations=400)
its = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
t0 = time.perf_counter()
run(init, iterations=its)
dt = time.perf_counter() - t0
print(f"{its} iterations in {dt:.2f} s = {dt / its * 1e6:.1f} us / iteration", flush=True)
pr = cProfile.Profile()
pr.enable()
run(init, iterations=its)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
