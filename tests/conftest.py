import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def _has_gpu() -> bool:
  try:
    from jax_sgmc_b200 import device
    return device.device_count() > 0
  except Exception:
    return False


@pytest.fixture(scope="session")
def gpu():
  """Session fixture for GPU tests: the library must load and see a device.

  On a GPU box a missing library is an ERROR (no silent skip): the product
  has no CPU fallback."""
  from jax_sgmc_b200 import _lib, device
  _lib.load()
  n = device.device_count()
  assert n > 0, "no CUDA device visible"
  device.set_device(0)
  return device
