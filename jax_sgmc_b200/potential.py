"""Stochastic / full potential, mirroring ``jax_sgmc.potential``.

``minibatch_potential`` (reference potential.py:94-216) and ``full_potential``
(:219-293) keep their signatures.  The returned callables evaluate
``U = (-N mean(ell) - prior) / T`` (or the masked form) for *all chains at
once* on the device; ``value_and_grad(potential_fn)`` is what the integrators
call in place of ``jax.value_and_grad(potential_fn, argnums=0, has_aux=True)``
(integrator.py:166, :593, :792) and returns the gradient produced by the fused
GLM kernels (tcgen05 tensor-core path when the shape allows, fp32 SIMT
otherwise).
"""
from __future__ import annotations

import os
from typing import Any, Callable, Optional

import numpy as np

from . import glm, nn, ops
from .data import BatchRef, MiniBatchInformation
from .device import DeviceArray
from .tree_util import ChainTree

# 'auto': tensor cores whenever the shape qualifies; override with
# SGMC_GLM_PATH=simt|tc_parity|tc_throughput
DEFAULT_PATH = os.environ.get("SGMC_GLM_PATH", "auto")


def _select_path(path: str, spec, n_chains: int, n: int) -> str:
  if path != "auto":
    return path
  # The TMEM accumulator truncates (~0.5 ulp bias per 16-wide k-step, DESIGN.md
  # section 4.2): the 1e-5 parity bound holds up to K ~ 2048 in GEMM1 (K = d) and
  # ~4096 in GEMM2 (K = n); larger contractions stay on the fp32 SIMT path unless
  # the caller asks for a tensor-core path explicitly.
  tc_ok = (spec.family == ops.FAMILY["logistic"] and spec.aux_off < 0
           and spec.d % 8 == 0 and n % 8 == 0 and 64 <= spec.d <= 2048
           and 64 <= n <= 4096 and n_chains >= 64
           and spec.prior != ops.PRIOR["inv_sigma"])
  return "tc_parity" if tc_ok else "simt"


class _PotentialFn:
  """Callable with the reference's StochasticPotential protocol
  (potential.py:42-66)."""

  def __init__(self, prior, likelihood, temperature, path):
    self.prior, self.likelihood = prior, likelihood
    self.temperature, self.path = float(temperature), path
    self._buffers = {}
    self._carried = None      # (buffer key, sample pointer) of a running carried step sequence

  # -- internal ------------------------------------------------------------------
  def _run(self, sample: ChainTree, reference_data, mask, want_grad, want_ell,
           grad_out: Optional[DeviceArray] = None,
           U_out: Optional[DeviceArray] = None,
           var_out: Optional[DeviceArray] = None):
    batch, info = reference_data
    assert isinstance(batch, BatchRef), "reference_data must come from jax_sgmc_b200.data"
    spec = glm.resolve(self.likelihood, self.prior, sample, self.temperature,
                       batch.loader.absmax(self.likelihood.x))
    C, P, n = sample.n_chains, sample.n_params, batch.n
    N = int(info.observation_count)
    X = batch.leaf(self.likelihood.x)
    y = batch.leaf(self.likelihood.y)
    if mask is None:
      mask = batch.mask
    path = _select_path(self.path, spec, C, n)
    key = (C, P, n, path)
    self._carried = None          # this call reuses the workspace: a carried split ends here
    buf = self._buffers.get(key)
    if buf is None:
      buf = {"U": DeviceArray((C,), np.float32), "var": DeviceArray((C,), np.float32),
             "ws": ops.glm_workspace(C, n, spec.d, "simt" if batch.per_chain else path),
             "ell": None}
      self._buffers[key] = buf
    grad = None
    if want_grad:
      grad = grad_out if grad_out is not None else DeviceArray((C, P), np.float32)
    ell = None
    if want_ell:
      if buf["ell"] is None:
        buf["ell"] = DeviceArray((C, n), np.float32)
      ell = buf["ell"]
    U_buf = U_out if U_out is not None else buf["U"]
    var_buf = var_out if var_out is not None else buf["var"]
    if not batch.per_chain:
      ops.glm_potential_grad(spec, sample.flat, X, y, batch.idx, N, U_buf,
                             var_buf, grad, ell, mask=mask, workspace=buf["ws"],
                             path=path, batch_size=n)
    else:
      # one minibatch per chain (host loader with per-chain seeds): no operand
      # sharing between chains -> fp32 SIMT kernels, all chains in one launch set
      ops.glm_potential_grad_per_chain(spec, sample.flat, X, y, batch.idx, N, U_buf,
                                       var_buf, grad, ell, mask=mask,
                                       workspace=buf["ws"])
    return U_buf, var_buf, grad, ell

  # -- public protocol --------------------------------------------------------------
  def __call__(self, sample: ChainTree, reference_data, state: Any = None,
               mask=None, likelihoods: bool = False):
    U, _, _, ell = self._run(sample, reference_data, mask, False, likelihoods)
    if likelihoods:
      return U, (ell, state)
    return U, state

  def sgld_step(self, sample: ChainTree, reference_data, keys_in, keys_out, step_size,
                temperature, v=None, alpha=0.9, lmbd=1e-5, temp_per_chain=None,
                wait_event=None, grad_out=None, U_out=None, var_out=None,
                carry_ok: bool = False) -> bool:
    """The whole langevin_diffusion.update_fn body (integrator.py:860-922) --
    value_and_grad of this potential on the minibatch, then the SGLD / pSGLD
    update of ``sample`` in place -- as ONE C call (sgmc_glm_sgld_step).
    Returns False (nothing done) when the chains do not share the minibatch.

    ``carry_ok``: the caller guarantees that nothing but these calls writes
    ``sample`` between steps (solver.sgmc, the sharded tempering).  The operand form
    of the sample then travels from one step's update to the next potential inside
    the workspace (SGMC_STEP_CARRY); any other use of the workspace (a
    ``value_and_grad`` call) or a different sample array ends the sequence."""
    batch, info = reference_data
    if batch.per_chain or batch.mask is not None:
      return False
    spec = glm.resolve(self.likelihood, self.prior, sample, self.temperature,
                       batch.loader.absmax(self.likelihood.x))
    C, P, n = sample.n_chains, sample.n_params, batch.n
    path = _select_path(self.path, spec, C, n)
    key = (C, P, n, path)
    buf = self._buffers.get(key)
    if buf is None:
      buf = {"U": DeviceArray((C,), np.float32), "var": DeviceArray((C,), np.float32),
             "ws": ops.glm_workspace(C, n, spec.d, path), "ell": None}
      self._buffers[key] = buf
    carry = 0
    if carry_ok and path != "simt":
      token = (key, sample.flat.ptr)
      carry = ops.STEP_CARRY if self._carried == token else ops.STEP_CARRY_INIT
      self._carried = token
    else:
      self._carried = None
    ops.glm_sgld_step(
        spec, sample.flat, batch.leaf(self.likelihood.x),
        batch.leaf(self.likelihood.y), batch.idx,
        int(info.observation_count), U_out if U_out is not None else buf["U"],
        var_out if var_out is not None else buf["var"], grad_out, keys_in, keys_out,
        step_size, temperature, v=v, alpha=alpha, lmbd=lmbd, workspace=buf["ws"],
        path=path, batch_size=n, temp_per_chain=temp_per_chain, wait_event=wait_event,
        leaf_sizes=sample.sizes, carry=carry)
    return True

  def sgld_scan(self, sample: ChainTree, source, keys_a, keys_b, step_sizes, temperatures,
                keep, samples_out, scalars_out, kept: int, v=None, alpha=0.9, lmbd=1e-5,
                grad_out=None, U_out=None, var_out=None) -> int:
    """len(step_sizes) whole Langevin steps (draw, value_and_grad, update,
    collection) in ONE C call (sgmc_glm_sgld_scan_device); ``source`` comes from
    the data functional's ``scan_source``.  Returns the new sample count."""
    loader, n, N = source["loader"], source["n"], source["N"]
    if source.get("host_stream"):
      return self._sgld_scan_host_stream(sample, source, keys_a, keys_b, step_sizes,
                                         temperatures, keep, samples_out, scalars_out, kept,
                                         v, alpha, lmbd, grad_out, U_out, var_out)
    spec = glm.resolve(self.likelihood, self.prior, sample, self.temperature,
                       loader.absmax(self.likelihood.x))
    C = sample.n_chains
    path = _select_path(self.path, spec, C, n)
    key = (C, sample.n_params, n, path)
    self._carried = None          # the scan starts its own carried sequence
    buf = self._buffers.get(key)
    if buf is None:
      buf = {"U": DeviceArray((C,), np.float32), "var": DeviceArray((C,), np.float32),
             "ws": ops.glm_workspace(C, n, spec.d, path), "ell": None}
      self._buffers[key] = buf
    return ops.glm_sgld_scan_device(
        spec, sample.flat, loader.device_data[self.likelihood.x],
        loader.device_data[self.likelihood.y], N, n,
        U_out if U_out is not None else buf["U"],
        var_out if var_out is not None else buf["var"], grad_out, keys_a, keys_b,
        sample.sizes, step_sizes, temperatures, keep, samples_out, scalars_out, kept,
        data_key_a=source.get("key_a"), data_key_b=source.get("key_b"),
        idx_buf=source.get("idx_buf"), idx_all=source.get("idx_all"), v=v, alpha=alpha,
        lmbd=lmbd, workspace=buf["ws"], path=path)

  def _sgld_scan_host_stream(self, sample, source, keys_a, keys_b, step_sizes, temperatures,
                             keep, samples_out, scalars_out, kept, v, alpha, lmbd, grad_out,
                             U_out, var_out):
    """The scan over a HOST-resident data set (StreamingNumpyDataLoader).

    staged (default): chunks of minibatches are gathered by host threads into page-locked
    memory (sgmc_host_gather_batches) and copied by the DMA engine
    (sgmc_glm_sgld_scan_host) while the device works through the previous chunk.
    pull (``SGMC_HOST_PULL=1`` or ``loader.pull = True``; no host thread touches the data,
    but SM-issued host reads and the HBM-bound kernels slow each other down, see
    csrc/host_pull.cu): the data set is page-locked and mapped in place once
    (``loader.mapped``); per chunk only the index rows of the chain's NumPy PCG64 pipeline
    go to the device, and the GPU reads every minibatch's rows over the host link itself
    (sgmc_glm_sgld_scan_pull: rows two batches ahead, operand staging one step ahead,
    (U, var) of every step read back).
    Both give identical samples.  Returns None when the configuration needs the step
    loop."""
    import os
    import threading
    from .device import Stream, current_stream
    from .io import _pinned_array
    loader, n, N = source["loader"], source["n"], source["N"]
    temps = np.asarray(temperatures, np.float32)
    if temps.size and not np.all(temps == temps[0]):
      return None
    X = loader.host_data[self.likelihood.x]
    y = loader.host_data[self.likelihood.y].reshape(-1)
    if X.ndim != 2:
      return None
    spec = glm.resolve(self.likelihood, self.prior, sample, self.temperature,
                       loader.absmax(self.likelihood.x))
    C, d = sample.n_chains, int(X.shape[1])
    path = _select_path(self.path, spec, C, n)
    comm = loader.upload_comm
    world = comm.world if comm is not None else 1
    rank = comm.rank if comm is not None else 0
    if n % world:
      return None
    rows = n // world
    pull = bool(getattr(loader, "pull", False)) or os.environ.get("SGMC_HOST_PULL", "0") == "1"
    # hybrid: a fraction of every minibatch (slice) is pulled by the GPU while the host
    # gathers and the DMA engine copies the rest (sgmc_glm_sgld_scan_hybrid)
    frac = float(os.environ.get("SGMC_HOST_PULL_FRACTION", getattr(loader, "pull_fraction", 0.0)))
    rows_pull = 0 if pull else min(rows, int(round(frac * rows / 8.0)) * 8)
    hybrid = rows_pull > 0 and rows_pull < rows
    if rows_pull >= rows:
      pull = True
    rows_dma = rows - rows_pull if hybrid else rows
    host_stride, dev_stride = rows_dma * d + n, n * d + n
    K = len(step_sizes)
    CH = max(2, min(int(source["chunk"]), 512 if pull else 64, K))
    CH -= CH % 2                       # the chain keys ping-pong once per step
    key = ("host_stream", C, sample.n_params, n, path, CH, world, pull, rows_dma)
    buf = self._buffers.get(key)
    if buf is None:
      res, res_addr = _pinned_array(CH * 2 * C)
      buf = {"res": res, "res_addr": res_addr,
             "slots": DeviceArray((3 * dev_stride,), np.float32),
             "uv": DeviceArray((2, 2, C), np.float32), "copy": Stream.create(),
             "ws": ops.glm_workspace(C, n, spec.d, path)}
      if pull or hybrid:
        hidx, hidx_addr = _pinned_array(2 * CH * n)
        buf.update(hidx=hidx.view(np.int32).reshape(2, CH, n), hidx_addr=hidx_addr,
                   didx=DeviceArray((2, CH, n), np.int32))
      if not pull:
        ring, ring_addr = _pinned_array(2 * CH * host_stride,
                                        write_combined=os.environ.get("SGMC_RING_WC", "0") == "1")
        buf.update(ring=ring, ring_addr=ring_addr)
      self._buffers[key] = buf
    self._carried = None
    # gather threads (the gather is DRAM-bound: 8 threads move as much as 16 on a 16-vCPU
    # host), shared between the ranks of a node
    threads = max(1, int(os.environ.get("SGMC_GATHER_THREADS", os.cpu_count() or 1)) //
                  max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
    main = current_stream()
    Xm = loader.mapped(self.likelihood.x) if (pull or hybrid) else 0
    ym = loader.mapped(self.likelihood.y) if pull else 0

    def produce(count, half):
      # index draw (the chain's NumPy PCG64 pipeline) of one chunk, then (staged mode) the
      # row gather; runs in a worker thread beside the device's work on the previous chunk
      idx_rows = source["draw"](count)
      if pull or hybrid:
        buf["hidx"][half, :count] = idx_rows
      if not pull:
        ops.host_gather_batches(buf["ring_addr"] + half * CH * host_stride * 4, X, y, idx_rows,
                                rank * rows, rows_dma, threads)

    ss = np.ascontiguousarray(step_sizes, np.float32)
    kp = None if keep is None else np.ascontiguousarray(keep, np.uint8)
    nccl = getattr(comm, "_comm", None) and comm._comm.value
    done, half, last_k = 0, 0, 0
    produce(min(CH, K), 0)
    trace = []
    prof = os.environ.get("SGMC_HOST_SCAN_TRACE") == "1"
    import time as _time
    t_enq = t_gpu = t_join = t_host = 0.0
    while done < K:
      t_a = _time.perf_counter()
      k = min(CH, K - done)
      worker = None
      if done + k < K:
        worker = threading.Thread(target=produce, args=(min(CH, K - done - k), 1 - half))
        worker.start()
      common = dict(temperature=float(temps[0]) if temps.size else 1.0, v=v, alpha=alpha,
                    lmbd=lmbd, workspace=buf["ws"], path=path, nccl_comm=nccl, rank=rank,
                    n_ranks=world, keep=None if kp is None else kp[done:done + k],
                    samples_out=samples_out, scalars_out=scalars_out, kept=kept)
      if pull or hybrid:
        didx = buf["didx"].row_slice(half, half + 1)
        didx.copy_from_pinned(buf["hidx_addr"] + half * CH * n * 4, k * n * 4, buf["copy"])
      if hybrid:
        kept = ops.glm_sgld_scan_hybrid(
            spec, sample.flat, buf["ring_addr"] + half * CH * host_stride * 4, k, rows_dma, Xm,
            didx, k, n, N, buf["slots"], 3, buf["uv"], buf["res_addr"], grad_out, keys_a, keys_b,
            ss[done:done + k], buf["copy"],
            pull_ctas=int(os.environ.get("SGMC_PULL_CTAS", "0")), **common)
      elif pull:
        kept = ops.glm_sgld_scan_pull(
            spec, sample.flat, Xm, ym, didx, k, n, N, buf["slots"], 3, buf["uv"],
            buf["res_addr"], grad_out, keys_a, keys_b, ss[done:done + k], buf["copy"],
            pull_ctas=int(os.environ.get("SGMC_PULL_CTAS", "0")), **common)
      else:
        kept = ops.glm_sgld_scan_host(
            spec, sample.flat, buf["ring_addr"] + half * CH * host_stride * 4, k, k, n, N,
            buf["slots"], 3, buf["uv"], buf["res_addr"], grad_out, keys_a, keys_b,
            ss[done:done + k], buf["copy"], **common)
      t_b = _time.perf_counter()
      main.sync()
      buf["copy"].sync()
      t_c = _time.perf_counter()
      trace.append(buf["res"][:k * 2 * C].reshape(k, 2, C)[:, 0].mean(axis=1))   # host reads U
      t_d = _time.perf_counter()
      if worker is not None:
        worker.join()
      t_e = _time.perf_counter()
      t_enq += t_b - t_a; t_gpu += t_c - t_b; t_host += t_d - t_c; t_join += t_e - t_d
      done, half, last_k = done + k, 1 - half, k
    if prof:
      print(f"host scan: {K} steps, chunks of {CH}: enqueue {t_enq * 1e6 / K:.1f} us/step, "
            f"wait for the device {t_gpu * 1e6 / K:.1f}, host reads {t_host * 1e6 / K:.1f}, "
            f"wait for the gather worker {t_join * 1e6 / K:.1f}", flush=True)
    self.last_potential_trace = np.concatenate(trace) if trace else np.zeros(0, np.float32)
    if last_k:                          # LangevinState.potential / .variance of the last step
      last = buf["uv"].row_slice((last_k - 1) & 1, ((last_k - 1) & 1) + 1).reshape(2, C)
      if U_out is not None:
        U_out.copy_from(last.row_slice(0, 1).reshape(C))
      if var_out is not None:
        var_out.copy_from(last.row_slice(1, 2).reshape(C))
    self.host_link_mode = "pull" if pull else (f"hybrid({rows_pull}/{rows} rows pulled)" if hybrid
                                               else "staged")
    self.h2d_bytes_per_step = (rows * d + n) * 4 + (n * 4 if (pull or hybrid) else 0)   # + the index row
    self.d2h_bytes_per_step = 2 * C * 4
    return kept

  def value_and_grad(self, sample: ChainTree, reference_data, state: Any = None,
                     mask=None, likelihoods: bool = False,
                     grad_out: Optional[DeviceArray] = None,
                     U_out: Optional[DeviceArray] = None,
                     var_out: Optional[DeviceArray] = None):
    """((U, aux), grad) with aux = state or (var(ell), state).  ``grad_out`` /
    ``U_out`` / ``var_out`` let the caller own the output buffers."""
    U, var, grad, ell = self._run(sample, reference_data, mask, True, False,
                                  grad_out, U_out, var_out)
    self.last_variance = var
    g = ChainTree.like(sample, grad)
    if likelihoods:
      return (U, (var, state)), g
    return (U, state), g


class _MlpPotentialFn:
  """StochasticPotential protocol (potential.py:42-66) for the recognised dense-network
  likelihood (``nn.MLPClassifier``): chain-batched forward + hand-derived reverse pass
  (csrc/mlp.cu).  All chains share the minibatch."""

  def __init__(self, prior, likelihood, temperature):
    self.prior, self.likelihood, self.temperature = prior, likelihood, float(temperature)
    self._buffers = {}
    self.shard_comm = None

  def shard_rows(self, comm):
    """Minibatch-gradient sharding (BASELINE.json configs[4]; the reference's ``pmap`` batch
    strategy, potential.py:150): every rank evaluates rows ``[rank n / R, (rank + 1) n / R)``
    of the shared minibatch for all chains, then the gradient and the potential are
    all-reduced over ``comm`` (a ``dist.NcclCommunicator``).  With equal slices the mean over
    the ranks of ``(-N mean_r(ell) - prior) / T`` and of its gradient IS the unsharded value,
    so one ``ncclAllReduce(sum)`` of ``f32[C][P]`` + ``f32[C]`` and a scale by ``1 / R`` per
    evaluation; every rank then applies the identical update.  ``var(ell)`` (only
    ``langevin_diffusion`` / reSGLD read it) stays the rank's own."""
    self.shard_comm = comm
    return self

  def sgld_step(self, *args, **kwargs) -> bool:
    return False          # no fused whole-step call: the integrator runs potential, then update

  def _run(self, sample: ChainTree, reference_data, mask, want_grad, want_ell,
           grad_out=None, U_out=None, var_out=None):
    batch, info = reference_data
    assert isinstance(batch, BatchRef), "reference_data must come from jax_sgmc_b200.data"
    if batch.per_chain:
      raise NotImplementedError(
          "the MLP potential evaluates one minibatch shared by all chains (device loader, "
          "or a host loader with shared streams)")
    cnn = isinstance(self.likelihood, nn.CNNClassifier)
    X, y = batch.leaf(self.likelihood.x), batch.leaf(self.likelihood.y)
    if cnn:
      spec = nn.resolve_cnn(self.likelihood, self.prior, sample, self.temperature, X.shape[1:])
    else:
      spec = nn.resolve(self.likelihood, self.prior, sample, self.temperature)
    C, P, n = sample.n_chains, sample.n_params, batch.n
    N = int(info.observation_count)
    if mask is None:
      mask = batch.mask
    key = (C, P, n)
    buf = self._buffers.get(key)
    if buf is None:
      buf = {"U": DeviceArray((C,), np.float32), "var": DeviceArray((C,), np.float32),
             "ws": ops.cnn_workspace(spec, C, n) if cnn else ops.mlp_workspace(spec, C, n),
             "ell": None}
      self._buffers[key] = buf
    grad = None
    if want_grad:
      grad = grad_out if grad_out is not None else DeviceArray((C, P), np.float32)
    ell = None
    if want_ell:
      if buf["ell"] is None:
        buf["ell"] = DeviceArray((C, n), np.float32)
      ell = buf["ell"]
    U_buf = U_out if U_out is not None else buf["U"]
    var_buf = var_out if var_out is not None else buf["var"]
    evaluate = ops.cnn_potential_grad if cnn else ops.mlp_potential_grad
    comm = self.shard_comm
    if comm is None or comm.world == 1:
      evaluate(spec, sample.flat, X, y, batch.idx, N, U_buf, var_buf, grad, ell,
               mask=mask, workspace=buf["ws"], batch_size=n)
      return U_buf, var_buf, grad, ell
    # ---- rows of the minibatch sharded over the ranks + all-reduce (see shard_rows) ----
    R, r = comm.world, comm.rank
    if n % R or batch.idx is None or want_ell:
      raise ValueError("row sharding needs an indexed minibatch whose size is a multiple of "
                       "the rank count (per-observation likelihoods are not gathered)")
    n_r = n // R
    idx_r = DeviceArray((n_r,), np.int32, ptr=batch.idx.ptr + r * n_r * 4, owner=batch.idx)
    mask_r = None if mask is None else DeviceArray((n_r,), np.float32, ptr=mask.ptr + r * n_r * 4,
                                                   owner=mask)
    evaluate(spec, sample.flat, X, y, idx_r, N, U_buf, var_buf, grad, None,
             mask=mask_r, workspace=buf["ws"], batch_size=n_r)
    comm.allreduce_sum(U_buf, U_buf)
    ops.tree_ewise(0, U_buf, 1.0 / R, U_buf)
    if grad is not None:
      comm.allreduce_sum(grad, grad)
      ops.tree_ewise(0, grad, 1.0 / R, grad)
    return U_buf, var_buf, grad, ell

  def __call__(self, sample: ChainTree, reference_data, state: Any = None, mask=None,
               likelihoods: bool = False):
    U, _, _, ell = self._run(sample, reference_data, mask, False, likelihoods)
    if likelihoods:
      return U, (ell, state)
    return U, state

  def value_and_grad(self, sample: ChainTree, reference_data, state: Any = None, mask=None,
                     likelihoods: bool = False, grad_out=None, U_out=None, var_out=None):
    U, var, grad, _ = self._run(sample, reference_data, mask, True, False, grad_out, U_out,
                                var_out)
    self.last_variance = var
    g = ChainTree.like(sample, grad)
    if likelihoods:
      return (U, (var, state)), g
    return (U, state), g


class _ProbedPotentialFn:
  """``minibatch_potential(prior, likelihood)`` handed the reference's plain callables
  (potential.py:94-127).  The first call that brings a sample and a minibatch lets
  ``glm.from_callable`` recognise the closed form (the callables are evaluated on a few
  host points, once); from then on every call goes to the fused-kernel potential of the
  recognised specification.  A callable that is no recognised closed form raises
  ``TypeError`` at that first call."""

  def __init__(self, prior, likelihood, temperature, path):
    self._callables = (prior, likelihood)
    self._args = (temperature, path)
    self._impl = None

  def _resolve(self, sample: ChainTree, loader) -> "_PotentialFn":
    if self._impl is None:
      from .tree_util import tree_unflatten
      prior, likelihood = self._callables
      tmpl = tree_unflatten(sample.treedef, [np.zeros(s, np.float64) for s in sample.shapes])
      lik_spec, prior_spec = glm.from_callable(likelihood, prior, tmpl,
                                               loader.initializer_batch())
      self._impl = _PotentialFn(prior_spec, lik_spec, *self._args)
    return self._impl

  def __call__(self, sample, reference_data, *args, **kwargs):
    return self._resolve(sample, reference_data[0].loader)(sample, reference_data, *args,
                                                           **kwargs)

  def value_and_grad(self, sample, reference_data, *args, **kwargs):
    return self._resolve(sample, reference_data[0].loader).value_and_grad(
        sample, reference_data, *args, **kwargs)

  def sgld_step(self, sample, reference_data, *args, **kwargs):
    return self._resolve(sample, reference_data[0].loader).sgld_step(
        sample, reference_data, *args, **kwargs)

  def sgld_scan(self, sample, source, *args, **kwargs):
    return self._resolve(sample, source["loader"]).sgld_scan(sample, source, *args, **kwargs)

  def __getattr__(self, name):          # last_variance, h2d_bytes_per_step, ... of the real one
    impl = self.__dict__.get("_impl")
    if impl is None:
      raise AttributeError(name)
    return getattr(impl, name)


def value_and_grad(potential_fn: _PotentialFn) -> Callable:
  """Stand-in for ``jax.value_and_grad(potential_fn, argnums=0, has_aux=True)``.

  With ``likelihoods=True`` the aux carries ``var(ell)`` per chain (what the
  only caller, ``langevin_diffusion``, reduces the likelihoods to,
  integrator.py:880) instead of the raw per-observation values.
  """
  return potential_fn.value_and_grad


def minibatch_potential(prior, likelihood, strategy: str = "map",
                        has_state: bool = False, is_batched: bool = False,
                        temperature: float = 1., path: str = None):
  """potential.py:94-216.  ``prior`` / ``likelihood`` are ``jax_sgmc_b200.glm``
  specifications; ``strategy`` is accepted for API compatibility (the fused
  kernels always evaluate the whole minibatch in parallel)."""
  del is_batched
  if strategy not in ("map", "vmap", "pmap"):
    raise NotImplementedError(f"Strategy {strategy} is unknown")
  # has_state (potential.py:131-137, :174-177): the reference threads a model state
  # through the likelihood and keeps the state returned for observation 0.  The
  # recognised likelihood specifications are pure functions of (sample, observation) --
  # they never change a state -- so the state handed in is the state handed back, which
  # is what the reference computes for a likelihood that returns its state unchanged.
  del has_state
  if isinstance(likelihood, (nn.MLPClassifier, nn.CNNClassifier)):
    if not isinstance(prior, (glm.FlatPrior, glm.GaussianPrior)):
      raise TypeError("the network potentials take a FlatPrior or a GaussianPrior")
    return _MlpPotentialFn(prior, likelihood, temperature)
  lik_is_spec = isinstance(likelihood, (glm.GaussianRegression, glm.LogisticRegression))
  prior_is_spec = isinstance(prior, (glm.FlatPrior, glm.GaussianPrior, glm.InvSigmaPrior))
  if not lik_is_spec and not prior_is_spec and callable(likelihood) and callable(prior) \
      and not isinstance(likelihood, glm._Spec) and not isinstance(prior, glm._Spec):
    # the reference's own calling convention: plain Python callables, recognised as one
    # of the closed forms on first use (glm.from_callable) or rejected there
    return _ProbedPotentialFn(prior, likelihood, temperature, path or DEFAULT_PATH)
  if not lik_is_spec:
    raise TypeError(
        "likelihood must be a jax_sgmc_b200.glm / jax_sgmc_b200.nn specification "
        "(GaussianRegression, LogisticRegression, MLPClassifier) or, together with the "
        "prior, a reference-style callable of a recognised closed form; arbitrary "
        "callables need the JAX route, see INTEGRATION.md")
  if not prior_is_spec:
    raise TypeError("prior must be a jax_sgmc_b200.glm prior specification")
  return _PotentialFn(prior, likelihood, temperature, path or DEFAULT_PATH)


def full_potential(prior, likelihood, strategy: str = "map", has_state: bool = False,
                   is_batched: bool = False, temperature: float = 1.,
                   path: str = None):
  """potential.py:219-293: ``U = (sum_b -dot(ell_b, mask_b) - prior) / T``."""
  assert strategy != "pmap", "Pmap is currently not supported"
  batch_potential = minibatch_potential(glm.FlatPrior(), likelihood, strategy,
                                        has_state, is_batched, 1.0, path)
  prior_only = minibatch_potential(prior, likelihood, strategy, has_state,
                                   is_batched, 1.0, path)

  scratch = {}

  def _full_in_one_call(sample: ChainTree, loader, mb_size: int) -> DeviceArray:
    spec = glm.resolve(likelihood, prior, sample, temperature, loader.absmax(likelihood.x))
    C, n = sample.n_chains, int(mb_size)
    run_path = _select_path(path or DEFAULT_PATH, spec, C, n)
    key = (C, n, run_path)
    if key not in scratch:
      scratch[key] = {"s": DeviceArray((2 * C,), np.float32), "idx": DeviceArray((n,), np.int32),
                      "mask": DeviceArray((2 * n,), np.float32),
                      "ws": ops.glm_workspace(C, n, spec.d, run_path)}
    b = scratch[key]
    out = DeviceArray((C,), np.float32)
    ops.glm_full_potential(spec, sample.flat, loader.device_data[likelihood.x],
                           loader.device_data[likelihood.y],
                           loader.static_information["observation_count"], n, out, b["s"],
                           b["idx"], b["mask"], b["ws"], path=run_path)
    return out

  def sum_batched_evaluations(sample: ChainTree, data_state, full_data_map_fn,
                              state: Any = None):
    """Returns ``(U f32[C] on the device, (data_state, state))``; everything is
    enqueued, nothing synchronises (the MH solvers consume U on the device)."""
    loader = getattr(full_data_map_fn, "loader", None)
    if loader is not None and state is None and not has_state and \
        not isinstance(likelihood, (nn.MLPClassifier, nn.CNNClassifier)):
      # the standard full_reference_data pass over an HBM-resident data set: all
      # batches inside one C call (same arithmetic as the loop below)
      return _full_in_one_call(sample, loader, full_data_map_fn.mb_size), (data_state, state)
    first = []
    total = DeviceArray.zeros((sample.n_chains,))

    def body(reference_data, mask, carry):
      if not first:
        first.append(reference_data)
      U, _ = batch_potential(sample, reference_data, carry, mask)
      _, info = reference_data
      # undo the N/n scaling and add up (potential.py:264-271, :290)
      ops.axpby(total, 1.0, total, info.batch_size / info.observation_count, U)
      return None, carry

    data_state, (_, new_state) = full_data_map_fn(
        body, data_state, state, masking=True, information=True)
    # prior value: the potential of an all-masked batch is -prior (L = 0)
    ref = first[0]
    zero_mask = DeviceArray.zeros((ref[0].n,))
    Up, _ = prior_only(sample, ref, None, zero_mask)
    out = DeviceArray((sample.n_chains,), np.float32)
    ops.axpby(out, 1.0 / temperature, total, 1.0 / temperature, Up)    # (sum - prior) / T
    return out, (data_state, new_state)

  return sum_batched_evaluations
