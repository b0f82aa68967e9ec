"""ctypes binding of libsgmc_b200.so (the C ABI declared in include/sgmc_b200.h).

The product path has no CPU fallback: if the shared library is missing or an
entry point fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SGMC_LIB_PATH",
                          os.path.join(_HERE, "_C", "libsgmc_b200.so"))


class SgmcError(RuntimeError):
  """Raised when a libsgmc_b200 entry point returns a nonzero status."""


class GlmSpec(C.Structure):
  """``sgmc_glm_spec`` (include/sgmc_b200.h)."""
  _fields_ = [("family", C.c_int32), ("d", C.c_int32), ("w_off", C.c_int32),
              ("aux_off", C.c_int32), ("prior", C.c_int32),
              ("prior_off", C.c_int32), ("prior_size", C.c_int32),
              ("prior_scale", C.c_float), ("temperature", C.c_float),
              ("x_absmax", C.c_float)]


MLP_MAX_LAYERS = 8


class MlpSpec(C.Structure):
  """``sgmc_mlp_spec`` (include/sgmc_b200.h)."""
  _fields_ = [("n_layers", C.c_int32), ("sizes", C.c_int32 * (MLP_MAX_LAYERS + 1)),
              ("w_off", C.c_int64 * MLP_MAX_LAYERS), ("b_off", C.c_int64 * MLP_MAX_LAYERS),
              ("activation", C.c_int32), ("prior", C.c_int32), ("prior_off", C.c_int64),
              ("prior_size", C.c_int64), ("prior_scale", C.c_float),
              ("temperature", C.c_float)]


CNN_MAX_CONV = 4


class CnnSpec(C.Structure):
  """``sgmc_cnn_spec`` (include/sgmc_b200.h)."""
  _fields_ = [("n_conv", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
              ("channels", C.c_int32 * (CNN_MAX_CONV + 1)), ("stride", C.c_int32 * CNN_MAX_CONV),
              ("n_classes", C.c_int32), ("w_off", C.c_int64 * (CNN_MAX_CONV + 1)),
              ("b_off", C.c_int64 * (CNN_MAX_CONV + 1)), ("prior", C.c_int32),
              ("prior_off", C.c_int64), ("prior_size", C.c_int64), ("prior_scale", C.c_float),
              ("temperature", C.c_float)]


_vp, _i64, _i32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
_int = C.c_int

# name -> argtypes   (every function returns int status unless listed below)
PROTOTYPES = {
    "sgmc_set_option": [_int, _int],
    "sgmc_device_count": [C.POINTER(_int)],
    "sgmc_set_device": [_int],
    "sgmc_device_info": [_int, C.POINTER(_int), C.POINTER(_int),
                         C.POINTER(_int), C.POINTER(_sz)],
    "sgmc_malloc": [C.POINTER(_vp), _sz],
    "sgmc_free": [_vp],
    "sgmc_host_alloc": [C.POINTER(_vp), _sz],
    "sgmc_host_alloc_wc": [C.POINTER(_vp), _sz],
    "sgmc_host_free": [_vp],
    "sgmc_memcpy_h2d": [_vp, _vp, _sz, _vp],
    "sgmc_memcpy_d2h": [_vp, _vp, _sz, _vp],
    "sgmc_memcpy_d2d": [_vp, _vp, _sz, _vp],
    "sgmc_memset": [_vp, _int, _sz, _vp],
    "sgmc_stream_create": [C.POINTER(_vp)],
    "sgmc_stream_create_high_priority": [C.POINTER(_vp)],
    "sgmc_stream_destroy": [_vp],
    "sgmc_stream_sync": [_vp],
    "sgmc_device_sync": [],
    "sgmc_event_create": [C.POINTER(_vp)],
    "sgmc_event_destroy": [_vp],
    "sgmc_event_record": [_vp, _vp],
    "sgmc_event_sync": [_vp],
    "sgmc_event_elapsed_ms": [_vp, _vp, C.POINTER(_f32)],
    "sgmc_stream_wait_event": [_vp, _vp],
    "sgmc_prng_split": [_vp, _vp, _vp, _i64, _int, _int],
    "sgmc_random_bits": [_vp, _vp, _vp, _i64, _i64, _int],
    "sgmc_uniform": [_vp, _vp, _vp, _i64, _i64, _f32, _f32, _int],
    "sgmc_normal": [_vp, _vp, _vp, _i64, _i64, _int],
    "sgmc_normal_like": [_vp, _vp, _vp, _i64, C.POINTER(_i64), _int, _int],
    "sgmc_randint": [_vp, _vp, _vp, _i64, _i32, _i32, _int],
    "sgmc_minibatch_draw": [_vp, _vp, _vp, _vp, _i64, _i64, _int],
    "sgmc_gather_rows": [_vp, _vp, _vp, _vp, _i64, _i64],
    "sgmc_synth_logistic_data": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _int],
    "sgmc_sgld_update": [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64), _int,
                         _f32, _f32, _vp, _int],
    "sgmc_sgld_rms_update": [_vp, _vp, _vp, _vp, _vp, _vp, _i64,
                             C.POINTER(_i64), _int, _f32, _f32, _vp, _f32,
                             _f32, _int],
    "sgmc_rms_prop_update": [_vp, _vp, _vp, _i64, _f32],
    "sgmc_rms_prop_get": [_vp, _vp, _vp, _vp, _i64, _f32],
    "sgmc_mass_matrix_update": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64],
    "sgmc_axpby": [_vp, _vp, _f32, _vp, _f32, _vp, _i64],
    "sgmc_absmax": [_vp, _vp, _i64, _vp],
    "sgmc_tree_ewise": [_vp, _int, _vp, _f32, _vp, _vp, _i64],
    "sgmc_tree_dot": [_vp, _vp, _vp, _vp, _i64, _i64],
    "sgmc_sghmc_begin": [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64), _int,
                         _f32, _vp, _int],
    "sgmc_sghmc_step": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64),
                        _int, _f32, _f32, _vp, _vp, _int, _int],
    "sgmc_sghmc_step_noise_model": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64),
                                    _int, _f32, _f32, _vp, _vp, _vp, _int, _int],
    "sgmc_glm_fisher_diag": [_vp, C.POINTER(GlmSpec), _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64,
                             _vp, _vp, _f32, _f32, _vp, _vp, _vp],
    "sgmc_obabo_pass_a": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                          C.POINTER(_i64), _int, _f32, _f32, _f32, _vp, _int],
    "sgmc_obabo_pass_b": [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64), _int,
                          _f32, _f32, _f32, _vp, _int],
    "sgmc_obabo_pass_a_adapted": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                  C.POINTER(_i64), _int, _f32, _f32, _f32, _vp, _vp, _int],
    "sgmc_obabo_pass_b_adapted": [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64), _int,
                                  _f32, _f32, _f32, _vp, _vp, _int],
    "sgmc_revleapfrog_step_adapted": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64),
                                      _int, _f32, _f32, _vp, _vp, _int, _int],
    "sgmc_glm_potential_grad": [_vp, C.POINTER(GlmSpec), _vp, _i64, _i64, _vp,
                                _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp,
                                _vp, _sz, _int],
    "sgmc_glm_potential_grad_per_chain": [_vp, C.POINTER(GlmSpec), _vp, _i64, _i64, _vp,
                                          _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp,
                                          _vp, _sz],
    "sgmc_glm_sgld_scan_host": [_vp, _vp, C.POINTER(GlmSpec), _vp, _vp, _i64, _i64, _vp, _i64,
                                _i64, _i64, _i64, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32,
                                _f32, _vp, _sz, _int, _int, _vp, _int, _int, _vp, _vp, _vp, _i64,
                                C.POINTER(_i64)],
    "sgmc_glm_sgld_scan_hybrid": [_vp, _vp, C.POINTER(GlmSpec), _vp, _vp, _i64, _i64, _vp, _i64,
                                  _i64, _vp, _vp, _int, _i64, _i64, _i64, _vp, _int, _vp, _vp, _vp,
                                  _vp, _vp, _vp, _f32, _f32, _f32, _vp, _sz, _int, _int, _vp, _int,
                                  _int, _vp, _vp, _vp, _i64, C.POINTER(_i64)],
    "sgmc_glm_sgld_scan_pull": [_vp, _vp, C.POINTER(GlmSpec), _vp, _vp, _i64, _i64, _vp, _vp,
                                _vp, _int, _i64, _i64, _i64, _vp, _int, _vp, _vp, _vp, _vp, _vp,
                                _vp, _f32, _f32, _f32, _vp, _sz, _int, _int, _vp, _int, _int,
                                _vp, _vp, _vp, _i64, C.POINTER(_i64)],
    "sgmc_cnn_potential_grad": [_vp, C.POINTER(CnnSpec), _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64,
                                _i64, _vp, _vp, _vp, _vp, _vp, _sz],
    "sgmc_glm_potential_grad_row_sharded": [_vp, C.POINTER(GlmSpec), _vp, _i64, _i64, _vp, _vp,
                                            _vp, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _int, _vp,
                                            _int, _int, _vp],
    "sgmc_glm_row_shard_finalize": [_vp, _vp, _i64, _vp, _vp, _i64],
    "sgmc_mlp_potential_grad": [_vp, C.POINTER(MlpSpec), _vp, _i64, _i64, _vp, _vp, _vp, _vp,
                                _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz],
    "sgmc_host_register": [_vp, _sz, C.POINTER(_vp)],
    "sgmc_host_unregister": [_vp],
    "sgmc_pull_rows": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _int],
    "sgmc_host_gather_batches": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _int],
    "sgmc_glm_sgld_scan_device": [_vp, C.POINTER(GlmSpec), _vp, _vp, _i64, _i64, _vp, _vp, _i64,
                                  _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  C.POINTER(_i64), _int, _vp, _vp, _vp, _i64, _vp, _vp, _i64,
                                  C.POINTER(_i64), _f32, _f32, _vp, _sz, _int, _int],
    "sgmc_glm_full_potential": [_vp, C.POINTER(GlmSpec), _vp, _i64, _i64, _vp, _vp, _i64, _i64,
                                _vp, _vp, _vp, _vp, _vp, _sz, _int],
    "sgmc_glm_sgld_step": [_vp, C.POINTER(GlmSpec), _vp, _vp, _i64, _i64, _vp, _vp, _vp,
                           _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32,
                           _f32, _vp, _sz, _int, _int, _int, _vp, _vp, C.POINTER(_i64), _int, _int],
    "sgmc_glm_prepare_minibatch": [_vp, C.POINTER(GlmSpec), _i64, _vp, _vp, _i64, _vp, _sz, _int,
                                   _int],
    "sgmc_revleapfrog_step": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64), _int,
                              _f32, _f32, _vp, _int, _int],
    "sgmc_mh_decide": [_vp, _int, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _i64, _int],
    "sgmc_resgld_decide": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _vp,
                           _vp, _vp, _i64, _int],
    "sgmc_resgld_decide_eta": [_vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _vp,
                               _vp, _vp, _i64, _int],
    "sgmc_swap_rows": [_vp, _vp, _vp, _vp, _i64, _i64],
    "sgmc_resgld_ladder_step": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int,
                                _i64, _i64, _int, _int, _vp, _vp, _int],
    "sgmc_resgld_sharded_exchange": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp,
                                     _vp, _vp, _vp, _vp, _vp, _int, _i64, _i64, _int, _int,
                                     _vp, _vp, _int],
    "sgmc_p2p_export": [_vp, _vp],
    "sgmc_p2p_open": [_vp, C.POINTER(_vp)],
    "sgmc_p2p_close": [_vp],
    "sgmc_p2p_allgather": [_vp, _vp, _int, _int, _vp, _sz, C.c_uint],
    "sgmc_p2p_timeouts": [C.POINTER(C.c_uint)],
    "sgmc_nccl_unique_id": [_vp],
    "sgmc_nccl_init": [C.POINTER(_vp), _vp, _int, _int],
    "sgmc_nccl_destroy": [_vp],
    "sgmc_nccl_allgather": [_vp, _vp, _vp, _vp, _sz],
    "sgmc_nccl_allreduce_sum_f32": [_vp, _vp, _vp, _vp, _sz],
}
# functions whose return value is NOT a status code
SPECIAL = {
    "sgmc_last_error": ([], C.c_char_p),
    "sgmc_version": ([], _int),
    "sgmc_get_option": ([_int], _int),
    "sgmc_launch_count": ([], C.c_ulonglong),
    "sgmc_nccl_available": ([], _int),
    "sgmc_glm_workspace_bytes": ([_i64, _i64, _i64, _int], _sz),
    "sgmc_p2p_window_bytes": ([_int, _sz], _sz),
    "sgmc_mlp_workspace_bytes": ([C.POINTER(MlpSpec), _i64, _i64], _sz),
    "sgmc_glm_fisher_scratch_floats": ([_i64, _i64, _i64], _sz),
    "sgmc_cnn_workspace_bytes": ([C.POINTER(CnnSpec), _i64, _i64], _sz),
}

_lib = None


def load() -> C.CDLL:
  """Load the shared library (once).  Raises if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise SgmcError(
        f"{LIB_PATH} is missing: build it with `python -m jax_sgmc_b200.build` "
        "(there is no CPU fallback)")
  lib = C.CDLL(LIB_PATH)
  for name, argtypes in PROTOTYPES.items():
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = _int
  for name, (argtypes, restype) in SPECIAL.items():
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
  _lib = lib
  # SGMC_OPTIONS="4=1,5=256": process-wide sgmc_set_option calls for A/B runs of
  # the tools and the benchmark (option numbers: include/sgmc_b200.h)
  for item in filter(None, os.environ.get("SGMC_OPTIONS", "").split(",")):
    k, v = item.split("=")
    if lib.sgmc_set_option(int(k), int(v)) != 0:
      raise SgmcError(f"SGMC_OPTIONS: bad option {item!r}")
  return lib


def call(name: str, *args):
  """Call a status-returning entry point; raise SgmcError on failure."""
  lib = load()
  status = getattr(lib, name)(*args)
  if status != 0:
    msg = lib.sgmc_last_error().decode(errors="replace")
    raise SgmcError(f"{name} failed (status {status}): {msg}")


def exported_names():
  return sorted(list(PROTOTYPES) + list(SPECIAL))
