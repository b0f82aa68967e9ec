"""Array-level wrappers of the libsgmc_b200 entry points.

One Python function per C-ABI op (include/sgmc_b200.h), taking
``DeviceArray``s.  The operator-API modules (``potential``, ``integrator``,
``adaption``, ``solver``) are built on these; tests and ``bench.py`` call them
directly as "the C ABI".  Nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .device import DeviceArray, Stream, current_stream, vp, i64_array

ORIGINAL, PARTITIONABLE = 0, 1
_LAYOUTS = {"original": 0, "partitionable": 1, 0: 0, 1: 1}


def _layout(layout) -> int:
  return _LAYOUTS[layout]


def _s(stream: Optional[Stream]):
  return (stream or current_stream()).handle


# ---- PRNG ---------------------------------------------------------------------

def prng_key(seed: int) -> np.ndarray:
  """``jax.random.PRNGKey`` (x64 off): uint32[2] = [hi32, lo32] of the seed.
  Pure bit packing of a Python int; host-side by nature."""
  seed = int(seed)
  return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], np.uint32)


def prng_keys(seeds: Sequence[int]) -> DeviceArray:
  return DeviceArray.from_numpy(np.stack([prng_key(s) for s in seeds]))


def split(keys: DeviceArray, num: int = 2, layout=0, stream=None) -> DeviceArray:
  """random.split for every key: uint32[..., 2] -> uint32[..., num, 2]."""
  n_keys = keys.size // 2
  out = DeviceArray(keys.shape[:-1] + (num, 2), np.uint32)
  _lib.call("sgmc_prng_split", _s(stream), vp(keys), vp(out), n_keys, num,
            _layout(layout))
  return out


def random_bits(keys: DeviceArray, n: int, layout=0, stream=None) -> DeviceArray:
  n_keys = keys.size // 2
  out = DeviceArray(keys.shape[:-1] + (n,), np.uint32)
  _lib.call("sgmc_random_bits", _s(stream), vp(keys), vp(out), n_keys, n,
            _layout(layout))
  return out


def uniform(keys: DeviceArray, n: int, minval=0.0, maxval=1.0, layout=0,
            stream=None) -> DeviceArray:
  n_keys = keys.size // 2
  out = DeviceArray(keys.shape[:-1] + (n,), np.float32)
  _lib.call("sgmc_uniform", _s(stream), vp(keys), vp(out), n_keys, n,
            float(minval), float(maxval), _layout(layout))
  return out


def normal(keys: DeviceArray, n: int, layout=0, stream=None) -> DeviceArray:
  n_keys = keys.size // 2
  out = DeviceArray(keys.shape[:-1] + (n,), np.float32)
  _lib.call("sgmc_normal", _s(stream), vp(keys), vp(out), n_keys, n,
            _layout(layout))
  return out


def normal_like(keys: DeviceArray, leaf_sizes: Sequence[int], layout=0,
                stream=None, out: Optional[DeviceArray] = None) -> DeviceArray:
  """integrator.random_tree for every chain: f32[C, P]."""
  C_ = keys.size // 2
  P = int(sum(leaf_sizes))
  if out is None:
    out = DeviceArray((C_, P), np.float32)
  _lib.call("sgmc_normal_like", _s(stream), vp(keys), vp(out), C_,
            i64_array(leaf_sizes), len(leaf_sizes), _layout(layout))
  return out


def randint(key: DeviceArray, n: int, minval: int, maxval: int, layout=0,
            stream=None) -> DeviceArray:
  out = DeviceArray((n,), np.int32)
  _lib.call("sgmc_randint", _s(stream), vp(key), vp(out), n, int(minval),
            int(maxval), _layout(layout))
  return out


def minibatch_draw(key_in: DeviceArray, key_out: DeviceArray, idx: DeviceArray,
                   observation_count: int, layout=0, stream=None):
  _lib.call("sgmc_minibatch_draw", _s(stream), vp(key_in), vp(key_out), vp(idx),
            idx.size, int(observation_count), _layout(layout))


def gather_rows(src: DeviceArray, idx: DeviceArray, stream=None,
                out: Optional[DeviceArray] = None) -> DeviceArray:
  row = int(np.prod(src.shape[1:], dtype=np.int64))
  if out is None:
    out = DeviceArray((idx.size,) + src.shape[1:], np.float32)
  _lib.call("sgmc_gather_rows", _s(stream), vp(src), vp(idx), vp(out), idx.size,
            row)
  return out


def synth_logistic_data(seed: int, N: int, d: int, layout=0, stream=None):
  """Generate the C2 synthetic data set in HBM: returns (X, y, w_true)."""
  X = DeviceArray((N, d), np.float32)
  y = DeviceArray((N,), np.float32)
  w = DeviceArray((d,), np.float32)
  key = prng_key(seed)
  _lib.call("sgmc_synth_logistic_data", _s(stream),
            key.ctypes.data_as(C.c_void_p), vp(X), vp(y), vp(w), N, d,
            _layout(layout))
  return X, y, w


# ---- fused updates ----------------------------------------------------------------

def sgld_update(theta, grad, keys_in, keys_out, leaf_sizes, step_size,
                temperature=1.0, temp_per_chain=None, v=None, alpha=0.9,
                lmbd=1e-5, layout=0, stream=None):
  """One fused SGLD (v is None) or pSGLD step; theta / v updated in place."""
  C_ = theta.shape[0]
  ls = i64_array(leaf_sizes)
  if v is None:
    _lib.call("sgmc_sgld_update", _s(stream), vp(theta), vp(grad), vp(keys_in),
              vp(keys_out), C_, ls, len(leaf_sizes), float(step_size),
              float(temperature), vp(temp_per_chain), _layout(layout))
  else:
    _lib.call("sgmc_sgld_rms_update", _s(stream), vp(theta), vp(v), vp(grad),
              vp(keys_in), vp(keys_out), C_, ls, len(leaf_sizes),
              float(step_size), float(temperature), vp(temp_per_chain),
              float(alpha), float(lmbd), _layout(layout))


def rms_prop_update(v, grad, alpha=0.9, stream=None):
  _lib.call("sgmc_rms_prop_update", _s(stream), vp(v), vp(grad), v.size,
            float(alpha))


def rms_prop_get(v, g_inv, sqrt_g_inv, lmbd=1e-5, stream=None):
  _lib.call("sgmc_rms_prop_get", _s(stream), vp(v), vp(g_inv), vp(sqrt_g_inv),
            v.size, float(lmbd))


def tree_ewise(op: int, out, alpha, x, y=None, stream=None):
  """op 0: alpha*x, 1: x+y, 2: x*y (flat chain-batched buffers)."""
  _lib.call("sgmc_tree_ewise", _s(stream), int(op), vp(out), float(alpha), vp(x),
            vp(y if y is not None else x), out.size)


def tree_dot(out, x, y, stream=None):
  C_, P = x.shape
  _lib.call("sgmc_tree_dot", _s(stream), vp(out), vp(x), vp(y), C_, P)


def axpby(out, a, x, b, y, stream=None):
  _lib.call("sgmc_axpby", _s(stream), vp(out), float(a), vp(x), float(b), vp(y),
            out.size)


def sghmc_begin(theta, momentum, keys_in, keys_out, leaf_sizes, step_size,
                mass=None, layout=0, stream=None):
  _lib.call("sgmc_sghmc_begin", _s(stream), vp(theta), vp(momentum),
            vp(keys_in), vp(keys_out), theta.shape[0], i64_array(leaf_sizes),
            len(leaf_sizes), float(step_size), vp(mass), _layout(layout))


def sghmc_step(theta, momentum, grad, keys_in, keys_out, leaf_sizes, step_size,
               friction=0.25, friction_vec=None, mass=None, last=False,
               layout=0, stream=None, noise_mul=None):
  """``noise_mul``: cb_diff_sqrt f32[C, P] of a Fisher noise model (glm_fisher_diag)."""
  if noise_mul is not None:
    assert noise_mul.shape == theta.shape
    _lib.call("sgmc_sghmc_step_noise_model", _s(stream), vp(theta), vp(momentum), vp(grad),
              vp(keys_in), vp(keys_out), theta.shape[0], i64_array(leaf_sizes),
              len(leaf_sizes), float(step_size), float(friction),
              vp(friction_vec), vp(mass), vp(noise_mul), int(bool(last)), _layout(layout))
    return
  _lib.call("sgmc_sghmc_step", _s(stream), vp(theta), vp(momentum), vp(grad),
            vp(keys_in), vp(keys_out), theta.shape[0], i64_array(leaf_sizes),
            len(leaf_sizes), float(step_size), float(friction),
            vp(friction_vec), vp(mass), int(bool(last)), _layout(layout))


def _is_adapted(mass) -> bool:
  """A per-chain adapted mass matrix: ``MassMatrix(inv, sqrt)`` of f32[C, P] DeviceArrays
  (adaption.mass_matrix); anything else is a constant mass vector f32[P] or None."""
  return isinstance(mass, tuple) and len(mass) == 2


def obabo_pass_a(theta, momentum, grad, ke_start, keys_in, keys_out,
                 leaf_sizes, step_size, temperature=1.0, friction=1.0,
                 mass=None, layout=0, stream=None):
  if _is_adapted(mass):
    assert mass[0].shape == theta.shape and mass[1].shape == theta.shape
    _lib.call("sgmc_obabo_pass_a_adapted", _s(stream), vp(theta), vp(momentum), vp(grad),
              vp(ke_start), vp(keys_in), vp(keys_out), theta.shape[0],
              i64_array(leaf_sizes), len(leaf_sizes), float(step_size),
              float(temperature), float(friction), vp(mass[0]), vp(mass[1]), _layout(layout))
    return
  _lib.call("sgmc_obabo_pass_a", _s(stream), vp(theta), vp(momentum), vp(grad),
            vp(ke_start), vp(keys_in), vp(keys_out), theta.shape[0],
            i64_array(leaf_sizes), len(leaf_sizes), float(step_size),
            float(temperature), float(friction), vp(mass), _layout(layout))


def obabo_pass_b(momentum, grad, ke_end, keys_in, leaf_sizes, step_size,
                 temperature=1.0, friction=1.0, mass=None, layout=0,
                 stream=None):
  if _is_adapted(mass):
    assert mass[0].shape == momentum.shape and mass[1].shape == momentum.shape
    _lib.call("sgmc_obabo_pass_b_adapted", _s(stream), vp(momentum), vp(grad), vp(ke_end),
              vp(keys_in), momentum.shape[0], i64_array(leaf_sizes),
              len(leaf_sizes), float(step_size), float(temperature),
              float(friction), vp(mass[0]), vp(mass[1]), _layout(layout))
    return
  _lib.call("sgmc_obabo_pass_b", _s(stream), vp(momentum), vp(grad), vp(ke_end),
            vp(keys_in), momentum.shape[0], i64_array(leaf_sizes),
            len(leaf_sizes), float(step_size), float(temperature),
            float(friction), vp(mass), _layout(layout))


# ---- GLM potential -----------------------------------------------------------------

FAMILY = {"gaussian": 0, "logistic": 1}
PRIOR = {"flat": 0, "gaussian": 1, "inv_sigma": 2}
PATH = {"simt": 0, "tc_parity": 1, "tc_throughput": 2, 0: 0, 1: 1, 2: 2}


def glm_spec(family, d, w_off, aux_off=-1, prior="flat", prior_off=0,
             prior_size=0, prior_scale=1.0, temperature=1.0,
             x_absmax=0.0) -> _lib.GlmSpec:
  return _lib.GlmSpec(FAMILY[family], int(d), int(w_off), int(aux_off),
                      PRIOR[prior], int(prior_off), int(prior_size),
                      float(prior_scale), float(temperature), float(x_absmax))


def absmax(x: DeviceArray, stream=None) -> float:
  """max |x| of a device array (one-time data-set statistic; synchronises)."""
  out = DeviceArray.zeros((1,), np.float32)
  _lib.call("sgmc_absmax", _s(stream), vp(x), x.size, vp(out))
  return float(out.numpy()[0])


def glm_workspace(n_chains: int, batch_size: int, d: int, path=0) -> DeviceArray:
  nbytes = _lib.load().sgmc_glm_workspace_bytes(n_chains, batch_size, d,
                                                PATH[path])
  return DeviceArray((int(nbytes),), np.uint8)


def glm_potential_grad(spec, theta, X, y, idx, observation_count, potential,
                       variance=None, grad=None, ell=None, mask=None,
                       workspace=None, path=0, batch_size=None, stream=None):
  """U, var(ell), dU/dtheta for all chains on one shared minibatch."""
  C_, P = theta.shape
  n = int(batch_size if batch_size is not None
          else (idx.size if idx is not None else X.shape[0]))
  if workspace is None:
    workspace = glm_workspace(C_, n, spec.d, path)
  _lib.call("sgmc_glm_potential_grad", _s(stream), C.byref(spec), vp(theta), C_,
            P, vp(X), vp(y), vp(idx), vp(mask), n, int(observation_count),
            vp(potential), vp(variance), vp(grad), vp(ell), vp(workspace),
            workspace.nbytes, PATH[path])
  return workspace


def glm_potential_grad_per_chain(spec, theta, X, y, idx, observation_count, potential,
                                 variance=None, grad=None, ell=None, mask=None,
                                 workspace=None, stream=None):
  """One minibatch per chain: idx int32[C, n]; fp32 SIMT kernels, one launch set."""
  C_, P = theta.shape
  n = int(idx.shape[1])
  if workspace is None:
    workspace = glm_workspace(C_, n, spec.d, "simt")
  _lib.call("sgmc_glm_potential_grad_per_chain", _s(stream), C.byref(spec), vp(theta), C_,
            P, vp(X), vp(y), vp(idx), vp(mask), n, int(observation_count), vp(potential),
            vp(variance), vp(grad), vp(ell), vp(workspace), workspace.nbytes)
  return workspace


def glm_potential_grad_row_sharded(spec, theta, X, y, idx, observation_count, potential,
                                   variance, grad, batch_size, rank, n_ranks, nccl_comm=None,
                                   workspace=None, scratch=None, path=0, stream=None):
  """The minibatch's rows sharded over ``n_ranks`` ranks + all-reduce of the gradient
  (see sgmc_glm_potential_grad_row_sharded).  ``nccl_comm``: the raw handle of
  ``dist.NcclCommunicator`` (None: partials only).  Returns ``(workspace, scratch)``."""
  C_, P = theta.shape
  n = int(batch_size)
  n_r = n // n_ranks
  if workspace is None:
    workspace = glm_workspace(C_, n_r, spec.d, path)
  if scratch is None:
    scratch = DeviceArray((C_ * (n_r + 4),), np.float32)
  _lib.call("sgmc_glm_potential_grad_row_sharded", _s(stream), C.byref(spec), vp(theta), C_, P,
            vp(X), vp(y), vp(idx), n, int(observation_count), vp(potential), vp(variance),
            vp(grad), vp(workspace), workspace.nbytes, PATH[path],
            None if nccl_comm is None else C.c_void_p(nccl_comm), int(rank), int(n_ranks),
            vp(scratch))
  return workspace, scratch


def glm_row_shard_finalize(extras, batch_size, potential, variance, stream=None):
  _lib.call("sgmc_glm_row_shard_finalize", _s(stream), vp(extras), int(batch_size),
            vp(potential), vp(variance), potential.size)


def mlp_spec(sizes, w_off, b_off, prior="flat", prior_off=0, prior_size=0, prior_scale=1.0,
             temperature=1.0) -> _lib.MlpSpec:
  """``sgmc_mlp_spec``: dense layers sizes[l] -> sizes[l+1], tanh between them."""
  L = len(sizes) - 1
  assert 1 <= L <= _lib.MLP_MAX_LAYERS and len(w_off) == L and len(b_off) == L
  spec = _lib.MlpSpec()
  spec.n_layers = L
  for i, v in enumerate(sizes):
    spec.sizes[i] = int(v)
  for l in range(L):
    spec.w_off[l], spec.b_off[l] = int(w_off[l]), int(b_off[l])
  spec.activation = 0
  spec.prior, spec.prior_off, spec.prior_size = PRIOR[prior], int(prior_off), int(prior_size)
  spec.prior_scale, spec.temperature = float(prior_scale), float(temperature)
  return spec


def mlp_workspace(spec, n_chains: int, batch_size: int) -> DeviceArray:
  nbytes = _lib.load().sgmc_mlp_workspace_bytes(C.byref(spec), n_chains, batch_size)
  return DeviceArray((int(nbytes),), np.uint8)


def mlp_potential_grad(spec, theta, X, y, idx, observation_count, potential, variance=None,
                       grad=None, ell=None, mask=None, workspace=None, batch_size=None,
                       stream=None):
  """U, var(ell), dU/dtheta of the MLP classifier for all chains on one shared minibatch
  (see sgmc_mlp_potential_grad)."""
  C_, P = theta.shape
  n = int(batch_size if batch_size is not None
          else (idx.size if idx is not None else X.shape[0]))
  if workspace is None:
    workspace = mlp_workspace(spec, C_, n)
  _lib.call("sgmc_mlp_potential_grad", _s(stream), C.byref(spec), vp(theta), C_, P, vp(X),
            vp(y), vp(idx), vp(mask), n, int(observation_count), vp(potential), vp(variance),
            vp(grad), vp(ell), vp(workspace), workspace.nbytes)
  return workspace


def cnn_spec(height, width, channels, strides, n_classes, w_off, b_off, prior="flat",
             prior_off=0, prior_size=0, prior_scale=1.0, temperature=1.0):
  """``sgmc_cnn_spec``: ``channels`` = [input, conv_0 out, ...], one stride per conv layer,
  ``w_off`` / ``b_off`` = the conv layers' then the dense head's offsets in the flat sample."""
  sp = _lib.CnnSpec()
  n_conv = len(strides)
  assert 1 <= n_conv <= _lib.CNN_MAX_CONV and len(channels) == n_conv + 1
  assert len(w_off) == n_conv + 1 and len(b_off) == n_conv + 1
  sp.n_conv, sp.height, sp.width, sp.n_classes = n_conv, int(height), int(width), int(n_classes)
  for i, c in enumerate(channels):
    sp.channels[i] = int(c)
  for i, st in enumerate(strides):
    sp.stride[i] = int(st)
  for i in range(n_conv + 1):
    sp.w_off[i], sp.b_off[i] = int(w_off[i]), int(b_off[i])
  sp.prior, sp.prior_off, sp.prior_size = PRIOR[prior], int(prior_off), int(prior_size)
  sp.prior_scale, sp.temperature = float(prior_scale), float(temperature)
  return sp


def cnn_workspace(spec, n_chains: int, batch_size: int) -> DeviceArray:
  nbytes = _lib.load().sgmc_cnn_workspace_bytes(C.byref(spec), n_chains, batch_size)
  return DeviceArray((int(nbytes),), np.uint8)


def cnn_potential_grad(spec, theta, X, y, idx, observation_count, potential, variance=None,
                       grad=None, ell=None, mask=None, workspace=None, batch_size=None,
                       stream=None):
  """U, var(ell), dU/dtheta of the CNN classifier for all chains on one shared minibatch
  (see sgmc_cnn_potential_grad)."""
  C_, P = theta.shape
  n = int(batch_size if batch_size is not None
          else (idx.size if idx is not None else X.shape[0]))
  if workspace is None:
    workspace = cnn_workspace(spec, C_, n)
  _lib.call("sgmc_cnn_potential_grad", _s(stream), C.byref(spec), vp(theta), C_, P, vp(X),
            vp(y), vp(idx), vp(mask), n, int(observation_count), vp(potential), vp(variance),
            vp(grad), vp(ell), vp(workspace), workspace.nbytes)
  return workspace


def glm_sgld_scan_host(spec, theta, host_batches_ptr, host_batch_count, n_steps, batch_size,
                       observation_count,
                       device_slots, n_slots, potential_variance, host_results_ptr, grad,
                       keys_a, keys_b, step_sizes, copy_stream, temperature=1.0, v=None,
                       alpha=0.9, lmbd=1e-5, workspace=None, path=0, layout=0, stream=None,
                       nccl_comm=None, rank=0, n_ranks=1, keep=None, samples_out=None,
                       scalars_out=None, kept: int = 0) -> int:
  """n_steps pSGLD / SGLD steps over minibatches in pinned host memory, one C call
  (see sgmc_glm_sgld_scan_host).  ``host_batches_ptr`` / ``host_results_ptr`` are
  addresses of page-locked buffers; ``step_sizes`` a float32 NumPy array.
  ``nccl_comm`` (the handle of sgmc_nccl_init) switches to the sharded upload: the
  host batches then hold this rank's n / n_ranks rows + all labels."""
  C_, P = theta.shape
  ss = np.ascontiguousarray(step_sizes, np.float32)
  assert ss.size >= n_steps
  kp = None if keep is None else np.ascontiguousarray(keep, np.uint8)
  cnt = C.c_int64(int(kept))
  cap = 0 if samples_out is None else samples_out.shape[0]
  _lib.call("sgmc_glm_sgld_scan_host", _s(stream), copy_stream.handle, C.byref(spec),
            vp(theta), vp(v), C_, P, C.c_void_p(host_batches_ptr), int(host_batch_count),
            int(n_steps), int(batch_size), int(observation_count), vp(device_slots), int(n_slots),
            vp(potential_variance), C.c_void_p(host_results_ptr), vp(grad), vp(keys_a),
            vp(keys_b), ss.ctypes.data_as(C.c_void_p), float(temperature), float(alpha),
            float(lmbd), vp(workspace), workspace.nbytes, PATH[path], _layout(layout),
            None if nccl_comm is None else C.c_void_p(nccl_comm), int(rank), int(n_ranks),
            None if kp is None else kp.ctypes.data_as(C.c_void_p), vp(samples_out),
            vp(scalars_out), int(cap), C.byref(cnt))
  return int(cnt.value)


def glm_sgld_scan_pull(spec, theta, X_mapped: int, y_mapped: int, idx_all, n_steps, batch_size,
                       observation_count, device_slots, n_slots, potential_variance,
                       host_results_ptr, grad, keys_a, keys_b, step_sizes, copy_stream,
                       temperature=1.0, v=None, alpha=0.9, lmbd=1e-5, workspace=None, path=0,
                       layout=0, stream=None, nccl_comm=None, rank=0, n_ranks=1, keep=None,
                       samples_out=None, scalars_out=None, kept: int = 0, pull_ctas=0) -> int:
  """n_steps pSGLD / SGLD steps over a registered HOST data set whose minibatch rows the
  GPU pulls over the host link by index (see sgmc_glm_sgld_scan_pull).  ``X_mapped`` /
  ``y_mapped`` come from ``host_register``; ``idx_all`` is a device int32[n_steps, n]."""
  C_, P = theta.shape
  ss = np.ascontiguousarray(step_sizes, np.float32)
  assert ss.size >= n_steps
  kp = None if keep is None else np.ascontiguousarray(keep, np.uint8)
  cnt = C.c_int64(int(kept))
  cap = 0 if samples_out is None else samples_out.shape[0]
  _lib.call("sgmc_glm_sgld_scan_pull", _s(stream), copy_stream.handle, C.byref(spec),
            vp(theta), vp(v), C_, P, C.c_void_p(X_mapped), C.c_void_p(y_mapped), vp(idx_all),
            int(pull_ctas), int(n_steps), int(batch_size), int(observation_count),
            vp(device_slots), int(n_slots), vp(potential_variance),
            C.c_void_p(host_results_ptr), vp(grad), vp(keys_a), vp(keys_b),
            ss.ctypes.data_as(C.c_void_p), float(temperature), float(alpha), float(lmbd),
            vp(workspace), workspace.nbytes, PATH[path], _layout(layout),
            None if nccl_comm is None else C.c_void_p(nccl_comm), int(rank), int(n_ranks),
            None if kp is None else kp.ctypes.data_as(C.c_void_p), vp(samples_out),
            vp(scalars_out), int(cap), C.byref(cnt))
  return int(cnt.value)


def glm_sgld_scan_hybrid(spec, theta, host_batches_ptr, host_batch_count, rows_dma, X_mapped: int,
                         idx_all, n_steps, batch_size, observation_count, device_slots, n_slots,
                         potential_variance, host_results_ptr, grad, keys_a, keys_b, step_sizes,
                         copy_stream, temperature=1.0, v=None, alpha=0.9, lmbd=1e-5,
                         workspace=None, path=0, layout=0, stream=None, nccl_comm=None, rank=0,
                         n_ranks=1, keep=None, samples_out=None, scalars_out=None, kept: int = 0,
                         pull_ctas=0) -> int:
  """The host scan with both transports (see sgmc_glm_sgld_scan_hybrid): the first
  ``rows_dma`` rows of every minibatch (slice) from the staged host batches, the rest pulled
  by the GPU out of the registered data set."""
  C_, P = theta.shape
  ss = np.ascontiguousarray(step_sizes, np.float32)
  assert ss.size >= n_steps
  kp = None if keep is None else np.ascontiguousarray(keep, np.uint8)
  cnt = C.c_int64(int(kept))
  cap = 0 if samples_out is None else samples_out.shape[0]
  _lib.call("sgmc_glm_sgld_scan_hybrid", _s(stream), copy_stream.handle, C.byref(spec),
            vp(theta), vp(v), C_, P, C.c_void_p(host_batches_ptr), int(host_batch_count),
            int(rows_dma), C.c_void_p(X_mapped), vp(idx_all), int(pull_ctas), int(n_steps),
            int(batch_size), int(observation_count), vp(device_slots), int(n_slots),
            vp(potential_variance), C.c_void_p(host_results_ptr), vp(grad), vp(keys_a),
            vp(keys_b), ss.ctypes.data_as(C.c_void_p), float(temperature), float(alpha),
            float(lmbd), vp(workspace), workspace.nbytes, PATH[path], _layout(layout),
            None if nccl_comm is None else C.c_void_p(nccl_comm), int(rank), int(n_ranks),
            None if kp is None else kp.ctypes.data_as(C.c_void_p), vp(samples_out),
            vp(scalars_out), int(cap), C.byref(cnt))
  return int(cnt.value)


def host_gather_batches(dst_ptr: int, X: np.ndarray, y: np.ndarray, idx: np.ndarray, row0: int,
                        rows: int, n_threads: int):
  """Threaded host gather into a page-locked staging buffer (see
  sgmc_host_gather_batches); releases the GIL while it runs."""
  idx = np.ascontiguousarray(idx, np.int32)
  nb, n = idx.shape
  _lib.call("sgmc_host_gather_batches", C.c_void_p(dst_ptr), X.ctypes.data_as(C.c_void_p),
            y.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), int(nb), int(n),
            int(X.shape[1]), int(row0), int(rows), int(n_threads))


def host_register(arr: np.ndarray) -> int:
  """Lock the pages of a host array in place and map them for the device (see
  sgmc_host_register); returns the address kernels read it at."""
  assert arr.flags["C_CONTIGUOUS"]
  out = C.c_void_p()
  _lib.call("sgmc_host_register", C.c_void_p(arr.ctypes.data), int(arr.nbytes), C.byref(out))
  return int(out.value)


def host_unregister(arr: np.ndarray):
  _lib.call("sgmc_host_unregister", C.c_void_p(arr.ctypes.data))


def pull_rows(X_mapped: int, y_mapped, idx, batch_size, row0, rows, d, dst_rows, dst_labels=None,
              n_ctas=0, stream=None):
  """Rows idx[row0 : row0 + rows] of a registered host data set -> device (see
  sgmc_pull_rows).  ``dst_rows`` / ``dst_labels`` are DeviceArrays or raw addresses."""
  as_ptr = lambda a: C.c_void_p(a) if isinstance(a, int) else vp(a)
  _lib.call("sgmc_pull_rows", _s(stream), C.c_void_p(X_mapped),
            None if y_mapped is None else C.c_void_p(y_mapped), as_ptr(idx), int(batch_size),
            int(row0), int(rows), int(d), as_ptr(dst_rows),
            None if dst_labels is None else as_ptr(dst_labels), int(n_ctas))


def glm_sgld_scan_device(spec, theta, X, y, observation_count, batch_size, potential,
                         variance, grad, keys_a, keys_b, leaf_sizes, step_sizes,
                         temperatures, keep, samples_out, scalars_out, kept: int,
                         data_key_a=None, data_key_b=None, idx_buf=None, idx_all=None, v=None,
                         alpha=0.9, lmbd=1e-5, workspace=None, path=0, layout=0,
                         stream=None) -> int:
  """len(step_sizes) Langevin steps in one C call (see sgmc_glm_sgld_scan_device);
  returns the new number of collected samples."""
  C_, P = theta.shape
  ss = np.ascontiguousarray(step_sizes, np.float32)
  tt = np.ascontiguousarray(temperatures, np.float32)
  kp = np.ascontiguousarray(keep, np.uint8)
  assert ss.size == tt.size == kp.size
  cnt = C.c_int64(int(kept))
  cap = 0 if samples_out is None else samples_out.shape[0]
  _lib.call("sgmc_glm_sgld_scan_device", _s(stream), C.byref(spec), vp(theta), vp(v), C_, P,
            vp(X), vp(y), int(observation_count), int(batch_size), vp(data_key_a),
            vp(data_key_b), vp(idx_buf), vp(idx_all), vp(potential), vp(variance), vp(grad),
            vp(keys_a), vp(keys_b), i64_array(leaf_sizes), len(leaf_sizes),
            ss.ctypes.data_as(C.c_void_p), tt.ctypes.data_as(C.c_void_p),
            kp.ctypes.data_as(C.c_void_p), int(ss.size), vp(samples_out), vp(scalars_out),
            int(cap), C.byref(cnt), float(alpha), float(lmbd), vp(workspace),
            workspace.nbytes, PATH[path], _layout(layout))
  return int(cnt.value)


def glm_full_potential(spec, theta, X, y, observation_count, batch_size, potential,
                       scratch, wrap_idx, wrap_mask, workspace, path=0, stream=None):
  """potential.full_potential over the HBM-resident data set in one C call."""
  C_, P = theta.shape
  _lib.call("sgmc_glm_full_potential", _s(stream), C.byref(spec), vp(theta), C_, P, vp(X),
            vp(y), int(observation_count), int(batch_size), vp(potential), vp(scratch),
            vp(wrap_idx), vp(wrap_mask), vp(workspace), workspace.nbytes, PATH[path])


def glm_sgld_step(spec, theta, X, y, idx, observation_count, potential, variance,
                  grad, keys_in, keys_out, step_size, temperature=1.0, v=None,
                  alpha=0.9, lmbd=1e-5, mask=None, workspace=None, path=0,
                  batch_size=None, layout=0, write_grad=True, temp_per_chain=None,
                  wait_event=None, leaf_sizes=None, stream=None, carry=0):
  """One langevin_diffusion step on the GLM potential: potential / variance /
  gradient at the current theta, then the SGLD (v None) or pSGLD update in
  place -- inside the gradient GEMM's epilogue when the shapes allow."""
  C_, P = theta.shape
  n = int(batch_size if batch_size is not None
          else (idx.size if idx is not None else X.shape[0]))
  if workspace is None:
    workspace = glm_workspace(C_, n, spec.d, path)
  _lib.call("sgmc_glm_sgld_step", _s(stream), C.byref(spec), vp(theta), vp(v), C_, P,
            vp(X), vp(y), vp(idx), vp(mask), n, int(observation_count),
            vp(potential), vp(variance), vp(grad), vp(keys_in), vp(keys_out),
            float(step_size), float(temperature), float(alpha), float(lmbd),
            vp(workspace), workspace.nbytes, PATH[path], _layout(layout),
            1 if write_grad else 0, vp(temp_per_chain),
            None if wait_event is None else wait_event.handle,
            None if leaf_sizes is None else i64_array(leaf_sizes),
            0 if leaf_sizes is None else len(leaf_sizes), int(carry))
  return workspace


# ---- MH-corrected samplers ------------------------------------------------------------

def revleapfrog_step(theta, momentum, grad, energy, keys_in, keys_out, leaf_sizes,
                     step_size, friction, mass=None, last=False, layout=0, stream=None):
  if _is_adapted(mass):
    assert mass[0].shape == theta.shape and mass[1].shape == theta.shape
    _lib.call("sgmc_revleapfrog_step_adapted", _s(stream), vp(theta), vp(momentum), vp(grad),
              vp(energy), vp(keys_in), vp(keys_out), theta.shape[0], i64_array(leaf_sizes),
              len(leaf_sizes), float(step_size), float(friction), vp(mass[0]), vp(mass[1]),
              1 if last else 0, _layout(layout))
    return
  _lib.call("sgmc_revleapfrog_step", _s(stream), vp(theta), vp(momentum), vp(grad),
            vp(energy), vp(keys_in), vp(keys_out), theta.shape[0], i64_array(leaf_sizes),
            len(leaf_sizes), float(step_size), float(friction), vp(mass),
            1 if last else 0, _layout(layout))


def glm_fisher_diag(spec, theta, X, y, idx, batch_size, observation_count, grad, friction_vec,
                    friction_scalar, step_size, noise_scale, scale, scratch=None, stream=None):
  """adaption.fisher_information(diagonal=True).get for a GLM potential (see
  sgmc_glm_fisher_diag); returns the scratch buffer for reuse."""
  C_, P = theta.shape
  need = int(_lib.load().sgmc_glm_fisher_scratch_floats(C_, P, int(batch_size)))
  if scratch is None or scratch.size < need:
    scratch = DeviceArray((need,), np.float32)
  _lib.call("sgmc_glm_fisher_diag", _s(stream), C.byref(spec), vp(theta), C_, P, vp(X), vp(y),
            vp(idx), int(batch_size), int(observation_count), vp(grad), vp(friction_vec),
            float(friction_scalar), float(step_size), vp(noise_scale), vp(scale), vp(scratch))
  return scratch


def mass_matrix_update(mean, ssq, m_inv, m_sqrt, sample, iteration: int, burn_in: int,
                       stream=None):
  """adaption.mass_matrix(diagonal=True).update for all chains (sgmc_mass_matrix_update);
  ``iteration`` is the already incremented count."""
  assert mean.shape == sample.shape == ssq.shape == m_inv.shape == m_sqrt.shape
  _lib.call("sgmc_mass_matrix_update", _s(stream), vp(mean), vp(ssq), vp(m_inv), vp(m_sqrt),
            vp(sample), sample.size, int(iteration), int(burn_in))


def mh_decide(mode, U_state, U_new, e0, e1, temperature, keys_in, keys_out, reject,
              ratio, layout=0, stream=None):
  """mode 'sggmc' | 'amagold'; see sgmc_mh_decide."""
  _lib.call("sgmc_mh_decide", _s(stream), {"sggmc": 0, "amagold": 1}[mode], vp(U_state),
            vp(U_new), vp(e0), vp(e1), float(temperature), vp(keys_in), vp(keys_out),
            vp(reject), vp(ratio), U_state.size, _layout(layout))


# ---- reSGLD ------------------------------------------------------------------------

def resgld_decide(U_n, U_h, var_n, ssq, F, step, T_normal, T_hot, keys_in,
                  keys_out, exchange, layout=0, stream=None, eta=None):
  """``eta`` = sa_schedule(step) when the caller supplies its own schedule
  (solver.py:274-276); None = the reference default 1 / step."""
  if eta is None:
    _lib.call("sgmc_resgld_decide", _s(stream), vp(U_n), vp(U_h), vp(var_n),
              vp(ssq), vp(F), int(step), float(T_normal), float(T_hot),
              vp(keys_in), vp(keys_out), vp(exchange), exchange.size,
              _layout(layout))
  else:
    _lib.call("sgmc_resgld_decide_eta", _s(stream), vp(U_n), vp(U_h), vp(var_n),
              vp(ssq), vp(F), float(eta), float(T_normal), float(T_hot),
              vp(keys_in), vp(keys_out), vp(exchange), exchange.size,
              _layout(layout))


def resgld_ladder_step(gathered, holder, ssq, F, temps, keys_in, keys_out, exchange,
                       n_replicas, n_systems, step, first_local, n_local,
                       temp_per_chain, temp_index, layout=0, stream=None):
  _lib.call("sgmc_resgld_ladder_step", _s(stream), vp(gathered), vp(holder), vp(ssq),
            vp(F), vp(temps), vp(keys_in), vp(keys_out), vp(exchange),
            int(n_replicas), int(n_systems), int(step), int(first_local),
            int(n_local), vp(temp_per_chain), vp(temp_index), _layout(layout))


def resgld_sharded_exchange(comm_handle, main_stream, x_stream, ready, done, uv, uv_send,
                            gathered, holder, ssq, F, temps, keys_in, keys_out, exchange,
                            n_replicas, n_systems, step, first_local, n_local,
                            temp_per_chain, temp_index, layout=0):
  """All-gather + ladder decision in one C call; see sgmc_resgld_sharded_exchange."""
  _lib.call("sgmc_resgld_sharded_exchange", comm_handle, main_stream.handle,
            None if x_stream is None else x_stream.handle,
            None if ready is None else ready.handle, None if done is None else done.handle,
            vp(uv), vp(uv_send), uv.nbytes, vp(gathered), vp(holder), vp(ssq), vp(F),
            vp(temps), vp(keys_in), vp(keys_out), vp(exchange), int(n_replicas),
            int(n_systems), int(step), int(first_local), int(n_local), vp(temp_per_chain),
            vp(temp_index), _layout(layout))


def swap_rows(a: DeviceArray, b: DeviceArray, exchange: DeviceArray, stream=None):
  n_rows = exchange.size
  row_bytes = a.nbytes // max(n_rows, 1)
  _lib.call("sgmc_swap_rows", _s(stream), vp(a), vp(b), vp(exchange), n_rows,
            row_bytes)


OPT_EXACT_UPDATE_MATH = 0
OPT_SERIAL_LAUNCH = 1      # 1: disable programmatic dependent launch
OPT_FUSED_STEP_EPILOGUE = 2  # 1: glm_sgld_step updates inside the gradient GEMM's epilogue
OPT_STEP_NOISE_IN_GEMM = 3   # 1: glm_sgld_step generates the noise in the GEMMs' idle warps
STEP_CARRY_INIT, STEP_CARRY = 1, 2   # glm_sgld_step(carry=...): see sgmc_glm_sgld_step
OPT_TC_LEGACY = 4            # 1: round-1 kernel sequence instead of the persistent fused potential kernel
OPT_TC_TIMELINE = 6          # 1: k_glm_tc_pair records its phase timeline
OPT_TC_TILE_N = 7            # accumulator columns per CTA and tile: 256 (default) or 128
OPT_NO_SHADOW_NOISE = 8      # 1: noise in the update kernel instead of under the GEMM mainloops
OPT_NO_PIPELINE = 9          # 1: scans stage minibatches on the sampling stream (no side stream)
OPT_STEP_PROFILE = 10        # 1: per-kernel CUDA-event timing of the carried step
OPT_TC_MAX_PAIRS = 11        # > 0: cap on the CTA pairs of k_glm_tc_pair (chain groups side by side)
OPT_FUSED_PAIR_UPDATE = 13   # 1: the carried step's update as the tail phase of the potential kernel
OPT_STREAM_UPDATE = 12       # 1: streaming (bulk-copy) update kernel of the carried step (A/B)
OPT_TC_CTA_GROUP = 5         # 2 (default): tcgen05 cta_group::2 on CTA pairs; 1: single CTAs


def set_option(option: int, value: int):
  _lib.call("sgmc_set_option", int(option), int(value))


def launch_count() -> int:
  return int(_lib.load().sgmc_launch_count())


def step_profile(reset: bool = True):
  """(us_prepare, us_potential, us_update, steps) averaged since the last reset
  (needs set_option(OPT_STEP_PROFILE, 1) while stepping)."""
  lib = _lib.load()
  us = (C.c_double * 3)()
  n = C.c_longlong()
  lib.sgmc_debug_step_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
  lib.sgmc_debug_step_profile(us, C.byref(n), 1 if reset else 0)
  return float(us[0]), float(us[1]), float(us[2]), int(n.value)
