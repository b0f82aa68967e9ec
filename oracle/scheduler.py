"""Restatement of the scalar schedules feeding the hot path (oracle).

TEST INFRASTRUCTURE.  scheduler.py:445-587.  The schedules are host-side
scalar glue: the hot path only consumes the four
scalars ``schedule(step_size, temperature, burn_in, accept)`` per iteration.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def polynomial_step_size(iterations: int, a=1.0, b=1.0, gamma=0.33):
  """scheduler.py:469-480: ``a * (b + n) ** (-gamma)`` for n in arange."""
  assert gamma >= 0 and a > 0 and b > 0
  n = np.arange(iterations).astype(F32)
  unscaled = np.power((F32(b) + n).astype(F32), F32(-gamma)).astype(F32)
  return (F32(a) * unscaled).astype(F32)


def find_ab(its, gamma, first, last):
  """scheduler.py:512-519."""
  gamma, first, last = F32(gamma), F32(first), F32(last)
  ginv = np.power(gamma, F32(-1.0)).astype(F32)
  fpow = np.power(first, -ginv).astype(F32)
  lpow = np.power(last, -ginv).astype(F32)
  apow = ((lpow - fpow).astype(F32) / F32(its - 1)).astype(F32)
  a = np.power(apow, -gamma).astype(F32)
  b = np.power((first / a).astype(F32), -ginv).astype(F32)
  return a, b


def polynomial_step_size_first_last(iterations, first=1.0, last=1.0,
                                    gamma=0.33):
  """scheduler.py:521-532."""
  assert gamma > 0 and first >= last
  a, b = find_ab(iterations, gamma, first, last)
  return polynomial_step_size(iterations, a, b, gamma)


def initial_burn_in(iterations: int, n: int = 0):
  """scheduler.py:578-587: 1.0 once ``n <= iteration`` else 0.0."""
  it = np.arange(iterations)
  return np.where(n <= it, F32(1.0), F32(0.0)).astype(F32)
