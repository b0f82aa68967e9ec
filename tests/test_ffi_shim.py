"""The jax.ffi binding: jax_sgmc_b200/csrc/ffi_shim.cc holds one
XLA-FFI handler per compute entry of include/sgmc_b200.h.  JAX and its headers are not
installable in this image, so the file cannot be linked into a running XLA here; what CAN
be checked on this box is checked: the file is the generator's output, it compiles with
g++ against a minimal stand-in for xla/ffi/api/ffi.h (tests/stubs/) that type-checks every
handler's parameters against its binding and every forwarded call against the prototypes
of sgmc_b200.h, every handler symbol is exported, and every enqueue-only compute entry of
the header has a handler."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "jax_sgmc_b200", "csrc", "ffi_shim.cc")
sys.path.insert(0, os.path.join(ROOT, "tools"))

# Entries of the header that are NOT XLA custom calls, and why.
NOT_CUSTOM_CALLS = {
    # runtime plumbing of the JAX-less host layer (XLA owns memory, streams and events)
    "sgmc_version", "sgmc_device_count", "sgmc_set_device", "sgmc_device_info", "sgmc_malloc",
    "sgmc_free", "sgmc_host_alloc", "sgmc_host_alloc_wc", "sgmc_host_free", "sgmc_memcpy_h2d",
    "sgmc_memcpy_d2h", "sgmc_memcpy_d2d", "sgmc_memset", "sgmc_stream_create",
    "sgmc_stream_create_high_priority", "sgmc_stream_destroy", "sgmc_stream_sync",
    "sgmc_device_sync", "sgmc_event_create", "sgmc_event_destroy", "sgmc_event_record",
    "sgmc_event_sync", "sgmc_stream_wait_event", "sgmc_event_elapsed_ms", "sgmc_set_option",
    "sgmc_get_option", "sgmc_host_register", "sgmc_host_unregister",
    # whole scans: host schedules, several streams, host memory -- they replace lax.scan
    # itself and are driven from Python, not from inside an XLA computation
    "sgmc_glm_sgld_scan_host", "sgmc_glm_sgld_scan_pull", "sgmc_glm_sgld_scan_hybrid",
    "sgmc_glm_sgld_scan_device",
    "sgmc_host_gather_batches", "sgmc_pull_rows", "sgmc_glm_prepare_minibatch",
    # communicator handles / peer memory (set up from Python, one process per GPU)
    "sgmc_nccl_available", "sgmc_nccl_unique_id", "sgmc_nccl_init", "sgmc_nccl_destroy",
    "sgmc_nccl_allgather", "sgmc_nccl_allreduce_sum_f32", "sgmc_resgld_sharded_exchange",
    "sgmc_p2p_export", "sgmc_p2p_open", "sgmc_p2p_close", "sgmc_p2p_allgather",
    "sgmc_p2p_timeouts", "sgmc_glm_potential_grad_row_sharded", "sgmc_glm_row_shard_finalize",
    # bench / test data generator
    "sgmc_synth_logistic_data",
}


def _header_entries():
  text = open(os.path.join(ROOT, "include", "sgmc_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return set(re.findall(r"\bint\s+(sgmc_\w+)\s*\(", text))


def test_shim_is_the_generators_output():
  import gen_ffi_shim
  assert open(SHIM).read() == gen_ffi_shim.render(), \
      "ffi_shim.cc is stale: run python tools/gen_ffi_shim.py"


def test_every_compute_entry_has_a_handler():
  import gen_ffi_shim
  handlers = {"sgmc_" + name for name, _ in gen_ffi_shim.HANDLERS}
  entries = _header_entries()
  assert handlers <= entries, handlers - entries
  missing = entries - handlers - NOT_CUSTOM_CALLS
  assert not missing, f"compute entries without an FFI handler: {sorted(missing)}"
  assert not (handlers & NOT_CUSTOM_CALLS)


def test_shim_compiles_against_the_ffi_stand_in(tmp_path):
  import gen_ffi_shim
  obj = tmp_path / "ffi_shim.o"
  cmd = ["g++", "-std=c++17", "-O0", "-fPIC", "-Wall", "-Werror", "-c", SHIM, "-o", str(obj),
         "-I" + os.path.join(ROOT, "tests", "stubs"), "-I/usr/local/cuda/include"]
  res = subprocess.run(cmd, capture_output=True, text=True)
  assert res.returncode == 0, res.stderr[-4000:]
  nm = subprocess.run(["nm", "--defined-only", str(obj)], capture_output=True, text=True).stdout
  exported = set(re.findall(r"\bT (sgmc_ffi_\w+)", nm))
  assert exported == {"sgmc_ffi_" + name for name, _ in gen_ffi_shim.HANDLERS}
  # the forwarded launchers are the only undefined sgmc_ symbols, all declared in the header
  und = subprocess.run(["nm", "--undefined-only", str(obj)], capture_output=True, text=True).stdout
  called = set(re.findall(r"\bU (sgmc_\w+)", und))
  assert called == {"sgmc_" + name for name, _ in gen_ffi_shim.HANDLERS} | {"sgmc_last_error"}


def test_stand_in_rejects_a_handler_that_does_not_match_its_binding(tmp_path):
  """The check above has teeth: change one attribute type in the binding and the
  stand-in's static_assert fails the compile."""
  bad = tmp_path / "bad_shim.cc"
  text = open(SHIM).read()
  assert '.Attr<float>("alpha")' in text
  bad.write_text(text.replace('.Attr<float>("alpha")', '.Attr<int32_t>("alpha")', 1)
                 .replace('"../../include/sgmc_b200.h"',
                          '"' + os.path.join(ROOT, "include", "sgmc_b200.h") + '"'))
  res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", str(bad),
                        "-I" + os.path.join(ROOT, "tests", "stubs"), "-I/usr/local/cuda/include"],
                       capture_output=True, text=True)
  assert res.returncode != 0 and "do not match its binding" in res.stderr
