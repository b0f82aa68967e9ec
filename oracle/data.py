"""Restatement of the minibatch index draws and the synthetic datasets.

TEST INFRASTRUCTURE.  Index semantics: data/numpy_loader.py:128-141 (device
loader, ``jax.random.randint``) and :263-265, :382-389 (host loader, NumPy
PCG64).  Datasets: examples/quickstart.md:102-118 (C1) and the synthetic
logistic-regression set of SURVEY.md section 8d (C2).
"""
from __future__ import annotations

import numpy as np

from . import prng

F32 = np.float32


def device_draw(key, batch_size: int, observation_count: int,
                layout="original"):
  """numpy_loader.py:132-134: ``key, split = split(state)``;
  ``randint(split, (n,), 0, N)``.  Returns (new_key, int32 idx[n])."""
  ks = prng.split(key, 2, layout)
  return ks[0], prng.randint(ks[1], (batch_size,), 0, observation_count,
                             layout)


class HostDraws:
  """numpy_loader.py:263-265 + :382-389: one PCG64 stream per chain seeded by
  ``SeedSequence(seed).spawn(1)[0]`` (seed defaults to the chain id); each
  batch is ``rng.choice(arange(N), size=mb, replace=True)``."""

  def __init__(self, observation_count: int, mb_size: int, seed: int = 0):
    self.N, self.mb = observation_count, mb_size
    self.rng = np.random.default_rng(np.random.SeedSequence(seed).spawn(1)[0])

  def draw(self):
    return self.rng.choice(np.arange(0, self.N), size=self.mb, replace=True)


def quickstart_dataset():
  """examples/quickstart.md:102-118, regenerated with the restated PRNG."""
  N, samples = 4, 1000
  key = prng.PRNGKey(0)
  split1, split2, split3 = prng.split(key, 3)
  sigma = F32(0.5)
  w = prng.uniform(split3, (N, 1), minval=-1, maxval=1)
  noise = (sigma * prng.normal(split2, (samples, 1))).astype(F32)
  x = prng.uniform(split1, (samples, N), minval=-10, maxval=10)
  x = np.stack([(x[:, 0] + x[:, 1]).astype(F32), x[:, 1],
                (F32(0.1) * x[:, 2] - F32(0.5) * x[:, 3]).astype(F32),
                x[:, 3]]).transpose().astype(F32)
  y = ((x @ w).astype(F32) + noise).astype(F32)
  return np.ascontiguousarray(x), y, w


def logistic_dataset(n_obs: int, d: int, seed: int = 0):
  """X ~ N(0,1)/sqrt(d), w* ~ N(0,1),
  y ~ Bernoulli(sigmoid(X w*)) stored as f32, all from PRNGKey(seed) splits."""
  kx, kw, ky = prng.split(prng.PRNGKey(seed), 3)
  X = (prng.normal(kx, (n_obs, d)) / F32(np.sqrt(d))).astype(F32)
  w = prng.normal(kw, (d,))
  z = (X @ w).astype(F32)
  p = (1.0 / (1.0 + np.exp(-z.astype(np.float64))))
  u = prng.uniform(ky, (n_obs,))
  y = (u < p).astype(F32)
  return X, y, w
