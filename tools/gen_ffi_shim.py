#!/usr/bin/env python
"""Generate jax_sgmc_b200/csrc/ffi_shim.cc: one XLA-FFI handler per compute entry of
include/sgmc_b200.h (the jax.ffi binding the north star names).

Every handler is pure plumbing -- the stream from the execution context, device buffers
-> typed_data(), scalars as attributes, in-place buffers as (argument, aliased result)
pairs -- so the file is generated from the table below instead of written by hand.
`python tools/gen_ffi_shim.py` rewrites the file; tests/test_ffi_shim.py checks that the
committed file is what this script emits, compiles it against a minimal stand-in for
xla/ffi/api/ffi.h (tests/stubs/) -- which type-checks every handler's parameter list
against its binding and every forwarded call against the prototypes of sgmc_b200.h --
and compares the exported handler symbols with the table.

Parameter kinds (launcher argument order):
  S                 the stream (execution context)
  in:DT:name        input buffer            -> name.typed_data()
  io:DT:name        in-place buffer: argument `name` + aliased result `name_out`
                    (input_output_aliases on the Python side) -> name_out->typed_data()
  out:DT:name       result buffer           -> name->typed_data()
  opt:DT:name       optional input buffer: an empty buffer means NULL
  ioopt:DT:name     optional in-place buffer (empty argument: NULL)
  ws:name           U8 scratch buffer       -> untyped_data(), size_bytes()
  a:T:name          scalar attribute (T = float | int32 | int64), passed as (CT)name
  span:name         int64 array attribute   -> name.begin(), (int)name.size()
  x:expr            an expression over the buffers (dimensions)
  glm               the sgmc_glm_spec, rebuilt from scalar attributes -> &spec
  mlp               the sgmc_mlp_spec, rebuilt from attributes        -> &spec
  cnn               the sgmc_cnn_spec, rebuilt from attributes        -> &spec
  null              a NULL pointer argument (a feature the FFI route does not expose)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "jax_sgmc_b200", "csrc", "ffi_shim.cc")

C0 = "{0}.dimensions()[0]"          # leading dimension of a buffer

HANDLERS = [
    # ---- jax.random ---------------------------------------------------------------------
    ("prng_split", ["S", "in:U32:keys", "out:U32:keys_out", "x:keys.element_count() / 2",
                    "a:int32:num", "a:int32:prng_layout"]),
    ("random_bits", ["S", "in:U32:keys", "out:U32:bits", "x:keys.element_count() / 2",
                     "a:int64:n", "a:int32:prng_layout"]),
    ("uniform", ["S", "in:U32:keys", "out:F32:values", "x:keys.element_count() / 2", "a:int64:n",
                 "a:float:minval", "a:float:maxval", "a:int32:prng_layout"]),
    ("normal", ["S", "in:U32:keys", "out:F32:values", "x:keys.element_count() / 2", "a:int64:n",
                "a:int32:prng_layout"]),
    ("normal_like", ["S", "in:U32:keys", "out:F32:noise", "x:keys.dimensions()[0]",
                     "span:leaf_sizes", "a:int32:prng_layout"]),
    ("randint", ["S", "in:U32:key", "out:S32:values", "x:values->element_count()",
                 "a:int32:minval", "a:int32:maxval", "a:int32:prng_layout"]),
    ("minibatch_draw", ["S", "in:U32:key", "out:U32:key_out", "out:S32:idx",
                        "x:idx->element_count()", "a:int64:observation_count",
                        "a:int32:prng_layout"]),
    ("gather_rows", ["S", "in:F32:src", "in:S32:idx", "out:F32:rows", "x:idx.element_count()",
                     "x:src.element_count() / src.dimensions()[0]"]),
    # ---- integrators / adaption ---------------------------------------------------------
    ("sgld_update", ["S", "io:F32:theta", "in:F32:grad", "in:U32:keys", "out:U32:keys_out",
                     "x:theta.dimensions()[0]", "span:leaf_sizes", "a:float:step_size",
                     "a:float:temperature", "opt:F32:temp_per_chain", "a:int32:prng_layout"]),
    ("sgld_rms_update", ["S", "io:F32:theta", "io:F32:v", "in:F32:grad", "in:U32:keys",
                         "out:U32:keys_out", "x:theta.dimensions()[0]", "span:leaf_sizes",
                         "a:float:step_size", "a:float:temperature", "opt:F32:temp_per_chain",
                         "a:float:alpha", "a:float:lmbd", "a:int32:prng_layout"]),
    ("rms_prop_update", ["S", "io:F32:v", "in:F32:grad", "x:v.element_count()", "a:float:alpha"]),
    ("rms_prop_get", ["S", "in:F32:v", "out:F32:g_inv", "out:F32:sqrt_g_inv",
                      "x:v.element_count()", "a:float:lmbd"]),
    ("mass_matrix_update", ["S", "io:F32:mean", "io:F32:ssq", "io:F32:m_inv", "io:F32:m_sqrt",
                            "in:F32:sample", "x:sample.element_count()", "a:int64:iteration",
                            "a:int64:burn_in"]),
    ("axpby", ["S", "out:F32:result", "a:float:a", "in:F32:x", "a:float:b", "in:F32:y",
               "x:x.element_count()"]),
    ("tree_ewise", ["S", "a:int32:op", "out:F32:result", "a:float:alpha", "in:F32:x", "in:F32:y",
                    "x:x.element_count()"]),
    ("tree_dot", ["S", "out:F32:result", "in:F32:x", "in:F32:y", "x:x.dimensions()[0]",
                  "x:x.element_count() / x.dimensions()[0]"]),
    ("sghmc_begin", ["S", "io:F32:theta", "io:F32:momentum", "in:U32:keys", "out:U32:keys_out",
                     "x:theta.dimensions()[0]", "span:leaf_sizes", "a:float:step_size",
                     "opt:F32:mass", "a:int32:prng_layout"]),
    ("sghmc_step", ["S", "io:F32:theta", "io:F32:momentum", "in:F32:grad", "in:U32:keys",
                    "out:U32:keys_out", "x:theta.dimensions()[0]", "span:leaf_sizes",
                    "a:float:step_size", "a:float:friction_scalar", "opt:F32:friction",
                    "opt:F32:mass", "a:int32:last", "a:int32:prng_layout"]),
    ("sghmc_step_noise_model", ["S", "io:F32:theta", "io:F32:momentum", "in:F32:grad",
                                "in:U32:keys", "out:U32:keys_out", "x:theta.dimensions()[0]",
                                "span:leaf_sizes", "a:float:step_size", "a:float:friction_scalar",
                                "opt:F32:friction", "opt:F32:mass", "in:F32:cb_diff_sqrt",
                                "a:int32:last", "a:int32:prng_layout"]),
    ("obabo_pass_a", ["S", "io:F32:theta", "io:F32:momentum", "in:F32:grad", "io:F32:ke_start",
                      "in:U32:keys", "out:U32:keys_out", "x:theta.dimensions()[0]",
                      "span:leaf_sizes", "a:float:step_size", "a:float:temperature",
                      "a:float:friction", "opt:F32:mass", "a:int32:prng_layout"]),
    ("obabo_pass_b", ["S", "io:F32:momentum", "in:F32:grad", "io:F32:ke_end", "in:U32:keys",
                      "x:momentum.dimensions()[0]", "span:leaf_sizes", "a:float:step_size",
                      "a:float:temperature", "a:float:friction", "opt:F32:mass",
                      "a:int32:prng_layout"]),
    ("obabo_pass_a_adapted", ["S", "io:F32:theta", "io:F32:momentum", "in:F32:grad",
                              "io:F32:ke_start", "in:U32:keys", "out:U32:keys_out",
                              "x:theta.dimensions()[0]", "span:leaf_sizes", "a:float:step_size",
                              "a:float:temperature", "a:float:friction", "in:F32:mass_inv",
                              "in:F32:mass_sqrt", "a:int32:prng_layout"]),
    ("obabo_pass_b_adapted", ["S", "io:F32:momentum", "in:F32:grad", "io:F32:ke_end",
                              "in:U32:keys", "x:momentum.dimensions()[0]", "span:leaf_sizes",
                              "a:float:step_size", "a:float:temperature", "a:float:friction",
                              "in:F32:mass_inv", "in:F32:mass_sqrt", "a:int32:prng_layout"]),
    ("revleapfrog_step", ["S", "io:F32:theta", "io:F32:momentum", "in:F32:grad", "io:F32:energy",
                          "in:U32:keys", "out:U32:keys_out", "x:theta.dimensions()[0]",
                          "span:leaf_sizes", "a:float:step_size", "a:float:friction",
                          "opt:F32:mass", "a:int32:last", "a:int32:prng_layout"]),
    ("revleapfrog_step_adapted", ["S", "io:F32:theta", "io:F32:momentum", "in:F32:grad",
                                  "io:F32:energy", "in:U32:keys", "out:U32:keys_out",
                                  "x:theta.dimensions()[0]", "span:leaf_sizes",
                                  "a:float:step_size", "a:float:friction", "in:F32:mass_inv",
                                  "in:F32:mass_sqrt", "a:int32:last", "a:int32:prng_layout"]),
    # ---- potentials -----------------------------------------------------------------------
    ("absmax", ["S", "in:F32:x", "x:x.element_count()", "out:F32:result"]),
    ("glm_potential_grad", ["S", "glm", "in:F32:theta", "x:theta.dimensions()[0]",
                            "x:theta.dimensions()[1]", "in:F32:X", "in:F32:y", "opt:S32:idx",
                            "opt:F32:mask", "a:int64:batch_size", "a:int64:observation_count",
                            "out:F32:potential", "out:F32:variance", "out:F32:grad", "null",
                            "ws:workspace", "a:int32:path"]),
    ("glm_potential_grad_per_chain", ["S", "glm", "in:F32:theta", "x:theta.dimensions()[0]",
                                      "x:theta.dimensions()[1]", "in:F32:X", "in:F32:y",
                                      "in:S32:idx", "opt:F32:mask", "a:int64:batch_size",
                                      "a:int64:observation_count", "out:F32:potential",
                                      "out:F32:variance", "out:F32:grad", "null", "ws:workspace"]),
    ("mlp_potential_grad", ["S", "mlp", "in:F32:theta", "x:theta.dimensions()[0]",
                            "x:theta.dimensions()[1]", "in:F32:X", "in:F32:y", "opt:S32:idx",
                            "opt:F32:mask", "a:int64:batch_size", "a:int64:observation_count",
                            "out:F32:potential", "out:F32:variance", "out:F32:grad", "null",
                            "ws:workspace"]),
    ("cnn_potential_grad", ["S", "cnn", "in:F32:theta", "x:theta.dimensions()[0]",
                            "x:theta.dimensions()[1]", "in:F32:X", "in:F32:y", "opt:S32:idx",
                            "opt:F32:mask", "a:int64:batch_size", "a:int64:observation_count",
                            "out:F32:potential", "out:F32:variance", "out:F32:grad", "null",
                            "ws:workspace"]),
    ("glm_full_potential", ["S", "glm", "in:F32:theta", "x:theta.dimensions()[0]",
                            "x:theta.dimensions()[1]", "in:F32:X", "in:F32:y",
                            "a:int64:observation_count", "a:int64:batch_size", "out:F32:potential",
                            "out:F32:scratch", "out:S32:wrap_idx", "out:F32:wrap_mask",
                            "ws:workspace", "a:int32:path"]),
    ("glm_fisher_diag", ["S", "glm", "in:F32:theta", "x:theta.dimensions()[0]",
                         "x:theta.dimensions()[1]", "in:F32:X", "in:F32:y", "opt:S32:idx",
                         "a:int64:batch_size", "a:int64:observation_count", "in:F32:grad",
                         "opt:F32:friction", "a:float:friction_scalar", "a:float:step_size",
                         "out:F32:noise_scale", "out:F32:scale", "out:F32:scratch"]),
    ("glm_sgld_step", ["S", "glm", "io:F32:theta", "ioopt:F32:v", "x:theta.dimensions()[0]",
                       "x:theta.dimensions()[1]", "in:F32:X", "in:F32:y", "opt:S32:idx",
                       "opt:F32:mask", "a:int64:batch_size", "a:int64:observation_count",
                       "out:F32:potential", "out:F32:variance", "out:F32:grad", "in:U32:keys",
                       "out:U32:keys_out", "a:float:step_size", "a:float:temperature",
                       "a:float:alpha", "a:float:lmbd", "ws:workspace", "a:int32:path",
                       "a:int32:prng_layout", "x:1", "opt:F32:temp_per_chain", "null", "x:nullptr",
                       "x:0", "x:0"]),
    # ---- solvers ------------------------------------------------------------------------------
    ("mh_decide", ["S", "a:int32:mode", "io:F32:U_state", "in:F32:U_new", "opt:F32:e0",
                   "in:F32:e1", "a:float:temperature", "in:U32:keys", "out:U32:keys_out",
                   "out:S32:reject", "out:F32:ratio", "x:U_state.element_count()",
                   "a:int32:prng_layout"]),
    ("resgld_decide", ["S", "in:F32:U_normal", "in:F32:U_hot", "in:F32:var_normal", "io:F32:ssq",
                       "in:F32:F", "a:int64:step", "a:float:T_normal", "a:float:T_hot",
                       "in:U32:keys", "out:U32:keys_out", "out:S32:exchange",
                       "x:ssq.element_count()", "a:int32:prng_layout"]),
    ("resgld_decide_eta", ["S", "in:F32:U_normal", "in:F32:U_hot", "in:F32:var_normal",
                           "io:F32:ssq", "in:F32:F", "a:float:eta", "a:float:T_normal",
                           "a:float:T_hot", "in:U32:keys", "out:U32:keys_out", "out:S32:exchange",
                           "x:ssq.element_count()", "a:int32:prng_layout"]),
    ("swap_rows", ["S", "io:F32:a", "io:F32:b", "in:S32:exchange", "x:exchange.element_count()",
                   "x:(int64_t)(a.size_bytes() / exchange.element_count())"]),
    ("resgld_ladder_step", ["S", "in:F32:gathered", "io:S32:holder", "io:F32:ssq", "in:F32:F",
                            "in:F32:temps", "in:U32:keys", "out:U32:keys_out", "out:S32:exchange",
                            "a:int32:n_replicas", "x:ssq.dimensions()[ssq.dimensions().size() - 1]",
                            "a:int64:step", "a:int32:first_local_replica",
                            "a:int32:n_local_replicas", "out:F32:temp_per_chain",
                            "out:S32:temp_index", "a:int32:prng_layout"]),
]

CT = {"float": "float", "int32": "int32_t", "int64": "int64_t"}
CAST = {"float": "", "int32": "(int)", "int64": ""}
GLM_ATTRS = [("int32", "family"), ("int32", "d"), ("int32", "w_off"), ("int32", "aux_off"),
             ("int32", "prior"), ("int32", "prior_off"), ("int32", "prior_size"),
             ("float", "prior_scale"), ("float", "potential_temperature"), ("float", "x_absmax")]
MLP_ATTRS = [("int32", "activation"), ("int32", "prior"), ("int64", "prior_off"),
             ("int64", "prior_size"), ("float", "prior_scale"), ("float", "potential_temperature")]

HEADER = '''// XLA-FFI adapters for libsgmc_b200 (jax.ffi custom calls).  GENERATED by
// tools/gen_ffi_shim.py -- edit the table there, not this file.
//
// NOT part of the default build: jax / the XLA FFI headers are not installable in the
// build image, so this file is compiled for real only where
//   python -c "import jax; print(jax.ffi.include_dir())"
// works:
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax;print(jax.ffi.include_dir())")
//       -I/usr/local/cuda/include ffi_shim.cc -o libsgmc_b200_ffi.so
//       -L_C -lsgmc_b200 -Wl,-rpath,'$ORIGIN/_C'
// Here it is compiled against a minimal stand-in for xla/ffi/api/ffi.h
// (tests/stubs/xla/ffi/api/ffi.h, tests/test_ffi_shim.py), which type-checks every
// handler against its binding and every forwarded call against include/sgmc_b200.h.
// The file contains no logic: each handler unpacks an XLA call frame (stream from the
// execution context, device buffers, scalar attributes) and forwards to the C-ABI
// launcher of the same name.  The launchers only enqueue work on the given stream and
// never allocate, so the calls are legal inside jit / lax.scan and XLA command buffers.
// In-place buffers are (argument, result) pairs the Python side aliases with
// input_output_aliases; optional inputs are empty buffers.  INTEGRATION.md shows the
// Python side (jax.ffi.register_ffi_target / jax.ffi.ffi_call).
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include <cstdint>

#include "../../include/sgmc_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Status(int rc) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, sgmc_last_error());
}

template <class B>
static auto OrNull(const B& b) -> decltype(b.typed_data()) {
  return b.element_count() ? b.typed_data() : nullptr;
}
'''

FOOTER = '''#endif  // __has_include("xla/ffi/api/ffi.h")
'''


def emit(name, params):
  args, attrs, rets, call, pre = [], [], [], [], []
  for p in params:
    f = p.split(":")
    k = f[0]
    if k == "S":
      call.append("stream")
    elif k == "in":
      args.append((f[1], f[2]))
      call.append(f"{f[2]}.typed_data()")
    elif k == "opt":
      args.append((f[1], f[2]))
      call.append(f"OrNull({f[2]})")
    elif k == "io":
      args.append((f[1], f[2]))
      rets.append((f[1], f[2] + "_out"))
      call.append(f"{f[2]}_out->typed_data()")
    elif k == "ioopt":
      args.append((f[1], f[2]))
      rets.append((f[1], f[2] + "_out"))
      call.append(f"{f[2]}.element_count() ? {f[2]}_out->typed_data() : nullptr")
    elif k == "out":
      rets.append((f[1], f[2]))
      call.append(f"{f[2]}->typed_data()")
    elif k == "ws":
      args.append(("U8", f[1]))
      call += [f"{f[1]}.untyped_data()", f"{f[1]}.size_bytes()"]
    elif k == "a":
      attrs.append((CT[f[1]], f[2]))
      call.append(f"{CAST[f[1]]}{f[2]}")
    elif k == "span":
      attrs.append(("ffi::Span<const int64_t>", f[1]))
      call += [f"{f[1]}.begin()", f"(int){f[1]}.size()"]
    elif k == "x":
      call.append(":".join(f[1:]))
    elif k == "null":
      call.append("nullptr")
    elif k == "glm":
      attrs += [(CT[t], n) for t, n in GLM_ATTRS]
      pre.append("  const sgmc_glm_spec spec{family, d, w_off, aux_off, prior, prior_off, "
                 "prior_size, prior_scale,\n                           potential_temperature, "
                 "x_absmax};")
      call.append("&spec")
    elif k == "mlp":
      attrs += [("ffi::Span<const int64_t>", "layer_sizes"),
                ("ffi::Span<const int64_t>", "w_off"), ("ffi::Span<const int64_t>", "b_off")]
      attrs += [(CT[t], n) for t, n in MLP_ATTRS]
      pre.append(
          "  sgmc_mlp_spec spec{};\n"
          "  spec.n_layers = (int32_t)w_off.size();\n"
          "  if (spec.n_layers > SGMC_MLP_MAX_LAYERS || layer_sizes.size() != w_off.size() + 1 ||\n"
          "      b_off.size() != w_off.size())\n"
          "    return ffi::Error(ffi::ErrorCode::kInvalidArgument, \"bad MLP layout\");\n"
          "  for (int l = 0; l <= spec.n_layers; ++l) spec.sizes[l] = (int32_t)layer_sizes.begin()[l];\n"
          "  for (int l = 0; l < spec.n_layers; ++l) {\n"
          "    spec.w_off[l] = w_off.begin()[l];\n"
          "    spec.b_off[l] = b_off.begin()[l];\n"
          "  }\n"
          "  spec.activation = activation; spec.prior = prior; spec.prior_off = prior_off;\n"
          "  spec.prior_size = prior_size; spec.prior_scale = prior_scale;\n"
          "  spec.temperature = potential_temperature;")
      call.append("&spec")
    elif k == "cnn":
      attrs += [("int32_t", "height"), ("int32_t", "width"), ("int32_t", "n_classes"),
                ("ffi::Span<const int64_t>", "channels"), ("ffi::Span<const int64_t>", "strides"),
                ("ffi::Span<const int64_t>", "w_off"), ("ffi::Span<const int64_t>", "b_off")]
      attrs += [(CT[t], n) for t, n in MLP_ATTRS if n != "activation"]
      pre.append(
          "  sgmc_cnn_spec spec{};\n"
          "  spec.n_conv = (int32_t)strides.size();\n"
          "  if (spec.n_conv < 1 || spec.n_conv > SGMC_CNN_MAX_CONV ||\n"
          "      channels.size() != strides.size() + 1 || w_off.size() != strides.size() + 1 ||\n"
          "      b_off.size() != w_off.size())\n"
          "    return ffi::Error(ffi::ErrorCode::kInvalidArgument, \"bad CNN layout\");\n"
          "  spec.height = height; spec.width = width; spec.n_classes = n_classes;\n"
          "  for (int l = 0; l <= spec.n_conv; ++l) {\n"
          "    spec.channels[l] = (int32_t)channels.begin()[l];\n"
          "    spec.w_off[l] = w_off.begin()[l];\n"
          "    spec.b_off[l] = b_off.begin()[l];\n"
          "  }\n"
          "  for (int l = 0; l < spec.n_conv; ++l) spec.stride[l] = (int32_t)strides.begin()[l];\n"
          "  spec.prior = prior; spec.prior_off = prior_off; spec.prior_size = prior_size;\n"
          "  spec.prior_scale = prior_scale; spec.temperature = potential_temperature;")
      call.append("&spec")
    else:
      raise ValueError(p)
  camel = "".join(w.capitalize() for w in name.split("_")) + "Impl"
  sig = ["cudaStream_t stream"]
  sig += [f"ffi::Buffer<ffi::{dt}> {n}" for dt, n in args]
  sig += [f"{t} {n}" for t, n in attrs]
  sig += [f"ffi::ResultBuffer<ffi::{dt}> {n}" for dt, n in rets]
  out = [f"// sgmc_{name}"]
  out.append(f"static ffi::Error {camel}(" + ",\n    ".join(sig) + ") {")
  out += pre
  out.append(f"  return Status(sgmc_{name}(" + ", ".join(call) + "));")
  out.append("}")
  bind = ["    ffi::Ffi::Bind()", "        .Ctx<ffi::PlatformStream<cudaStream_t>>()"]
  bind += [f"        .Arg<ffi::Buffer<ffi::{dt}>>()" for dt, _ in args]
  bind += [f"        .Attr<{t}>(\"{n}\")" for t, n in attrs]
  bind += [f"        .Ret<ffi::Buffer<ffi::{dt}>>()" for dt, _ in rets]
  out.append(f"XLA_FFI_DEFINE_HANDLER_SYMBOL(\n    sgmc_ffi_{name}, {camel},\n" + "\n".join(bind) + ");")
  return "\n".join(out) + "\n"


def render() -> str:
  return HEADER + "\n" + "\n".join(emit(n, p) for n, p in HANDLERS) + "\n" + FOOTER


if __name__ == "__main__":
  text = render()
  if "--check" in sys.argv:
    sys.exit(0 if open(OUT).read() == text else 1)
  with open(OUT, "w") as f:
    f.write(text)
  print(f"wrote {OUT}: {len(HANDLERS)} handlers")
