"""Device memory plumbing for the host layer (no torch, no jax).

``DeviceArray`` is a thin owner of a ``cudaMalloc`` allocation with a shape and
dtype; it exists so the Python mirror of the jax-sgmc operator API can hold
pytrees of device buffers.  All arithmetic happens in libsgmc_b200 kernels;
there is no CPU fallback for any of it.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np

from . import _lib


class Stream:
  """A CUDA stream; ``Stream.default()`` is the legacy default stream (0)."""

  def __init__(self, handle: Optional[int] = None, owned: bool = False):
    self.handle = C.c_void_p(handle)
    self._owned = owned

  @classmethod
  def create(cls, high_priority: bool = False) -> "Stream":
    h = C.c_void_p()
    _lib.call("sgmc_stream_create_high_priority" if high_priority else "sgmc_stream_create",
              C.byref(h))
    return cls(h.value, owned=True)

  @classmethod
  def default(cls) -> "Stream":
    return cls(None)

  def sync(self):
    _lib.call("sgmc_stream_sync", self.handle)

  def wait_event(self, event: "Event"):
    """Make later work on this stream wait for ``event`` (no host blocking)."""
    _lib.call("sgmc_stream_wait_event", self.handle, event.handle)

  def __del__(self):
    if getattr(self, "_owned", False) and self.handle.value:
      try:
        _lib.call("sgmc_stream_destroy", self.handle)
      except Exception:  # interpreter shutdown
        pass


class Event:
  def __init__(self):
    self.handle = C.c_void_p()
    _lib.call("sgmc_event_create", C.byref(self.handle))

  def record(self, stream: Stream):
    _lib.call("sgmc_event_record", self.handle, stream.handle)

  def sync(self):
    _lib.call("sgmc_event_sync", self.handle)

  def elapsed_ms(self, later: "Event") -> float:
    ms = C.c_float()
    _lib.call("sgmc_event_elapsed_ms", self.handle, later.handle, C.byref(ms))
    return float(ms.value)

  def __del__(self):
    try:
      if self.handle.value:
        _lib.call("sgmc_event_destroy", self.handle)
    except Exception:
      pass


_current_stream = Stream.default()


def current_stream() -> Stream:
  return _current_stream


def set_current_stream(stream: Stream):
  global _current_stream
  _current_stream = stream


def set_device(index: int):
  _lib.call("sgmc_set_device", int(index))


def device_count() -> int:
  n = C.c_int()
  _lib.call("sgmc_device_count", C.byref(n))
  return n.value


def synchronize():
  _lib.call("sgmc_device_sync")


class DeviceArray:
  """An n-d array in device memory (row-major, contiguous)."""

  __slots__ = ("ptr", "shape", "dtype", "_owner", "nbytes", "__weakref__")

  def __init__(self, shape: Sequence[int], dtype, ptr: Optional[int] = None,
               owner=None):
    self.shape = tuple(int(s) for s in shape)
    self.dtype = np.dtype(dtype)
    self.nbytes = math.prod(self.shape) * self.dtype.itemsize
    if ptr is None:
      p = C.c_void_p()
      _lib.call("sgmc_malloc", C.byref(p), self.nbytes)
      self.ptr = p.value
      self._owner = None
    else:
      self.ptr = int(ptr)
      self._owner = owner if owner is not None else False

  # -- construction -----------------------------------------------------------
  @classmethod
  def from_numpy(cls, arr, dtype=None, stream: Optional[Stream] = None):
    arr = np.ascontiguousarray(arr, dtype=dtype)
    out = cls(arr.shape, arr.dtype)
    out.copy_from_host(arr, stream)
    return out

  @classmethod
  def zeros(cls, shape, dtype=np.float32, stream: Optional[Stream] = None):
    out = cls(shape, dtype)
    s = stream or current_stream()
    _lib.call("sgmc_memset", C.c_void_p(out.ptr), 0, out.nbytes, s.handle)
    return out

  @classmethod
  def full(cls, shape, value, dtype=np.float32):
    return cls.from_numpy(np.full(shape, value, dtype=dtype))

  # -- transfers ----------------------------------------------------------------
  def copy_from_host(self, arr: np.ndarray, stream: Optional[Stream] = None):
    arr = np.ascontiguousarray(arr, dtype=self.dtype)
    assert arr.nbytes == self.nbytes, (arr.shape, self.shape)
    s = stream or current_stream()
    _lib.call("sgmc_memcpy_h2d", C.c_void_p(self.ptr),
              arr.ctypes.data_as(C.c_void_p), self.nbytes, s.handle)
    s.sync()   # pageable source: do not let the caller free it early

  def copy_from_pinned(self, host_addr: int, nbytes: int, stream: Optional[Stream] = None):
    """Asynchronous H2D copy of ``nbytes`` from page-locked host memory at ``host_addr``
    (the caller keeps the source alive and untouched until the stream has passed)."""
    assert 0 <= nbytes <= self.nbytes
    s = stream or current_stream()
    _lib.call("sgmc_memcpy_h2d", C.c_void_p(self.ptr), C.c_void_p(host_addr), int(nbytes),
              s.handle)

  def numpy(self, stream: Optional[Stream] = None) -> np.ndarray:
    out = np.empty(self.shape, dtype=self.dtype)
    s = stream or current_stream()
    if self.nbytes:
      _lib.call("sgmc_memcpy_d2h", out.ctypes.data_as(C.c_void_p),
                C.c_void_p(self.ptr), self.nbytes, s.handle)
      s.sync()
    return out

  def copy(self, stream: Optional[Stream] = None) -> "DeviceArray":
    out = DeviceArray(self.shape, self.dtype)
    s = stream or current_stream()
    _lib.call("sgmc_memcpy_d2d", C.c_void_p(out.ptr), C.c_void_p(self.ptr),
              self.nbytes, s.handle)
    return out

  def zero_(self, stream: Optional[Stream] = None) -> "DeviceArray":
    s = stream or current_stream()
    _lib.call("sgmc_memset", C.c_void_p(self.ptr), 0, self.nbytes, s.handle)
    return self

  def copy_from(self, other: "DeviceArray", stream: Optional[Stream] = None):
    assert other.nbytes == self.nbytes
    s = stream or current_stream()
    _lib.call("sgmc_memcpy_d2d", C.c_void_p(self.ptr), C.c_void_p(other.ptr),
              self.nbytes, s.handle)

  # -- views ----------------------------------------------------------------------
  def reshape(self, *shape) -> "DeviceArray":
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
      shape = tuple(shape[0])
    n = math.prod(self.shape)
    shape = list(shape)
    if -1 in shape:
      k = shape.index(-1)
      rest = math.prod(s for s in shape if s != -1)
      shape[k] = n // max(rest, 1)
    assert math.prod(shape) == n, (shape, self.shape)
    return DeviceArray(shape, self.dtype, ptr=self.ptr, owner=self)

  def row_slice(self, start: int, stop: int) -> "DeviceArray":
    """View of rows [start, stop) along the leading axis."""
    row = math.prod(self.shape[1:]) * self.dtype.itemsize
    return DeviceArray((stop - start,) + self.shape[1:], self.dtype,
                       ptr=self.ptr + start * row, owner=self)

  @property
  def size(self) -> int:
    return math.prod(self.shape)

  @property
  def ndim(self) -> int:
    return len(self.shape)

  def __array__(self, dtype=None, copy=None):
    a = self.numpy()
    return a.astype(dtype) if dtype is not None else a

  def __repr__(self):
    return f"DeviceArray(shape={self.shape}, dtype={self.dtype})"

  def __del__(self):
    try:
      if getattr(self, "_owner", False) is None and self.ptr:
        _lib.call("sgmc_free", C.c_void_p(self.ptr))
    except Exception:
      pass


def vp(x) -> C.c_void_p:
  """Device pointer of a DeviceArray (or NULL for None) as a ctypes void*."""
  if x is None:
    return C.c_void_p(None)
  if isinstance(x, DeviceArray):
    return C.c_void_p(x.ptr)
  return C.c_void_p(int(x))


def i64_array(values):
  arr = (C.c_int64 * len(values))(*[int(v) for v in values])
  return arr
