// XLA-FFI adapters for libsgmc_b200 (jax.ffi custom calls).
//
// NOT part of the default build: jax / the XLA FFI headers are not installable
// in the build image, so this file is compiled only where
//   python -c "import jax; print(jax.ffi.include_dir())"
// works:
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax;print(jax.ffi.include_dir())") \
//       -I/usr/local/cuda/include ffi_shim.cc -o libsgmc_b200_ffi.so \
//       -L_C -lsgmc_b200 -Wl,-rpath,'$ORIGIN/_C'
// It contains no logic: each handler unpacks an XLA call frame (stream from the
// execution context, device buffers, scalar attributes) and forwards to the
// C-ABI launcher of the same name declared in include/sgmc_b200.h.  The
// launchers only enqueue work on the given stream and never allocate, so the
// calls are legal inside jit / lax.scan and XLA command buffers.
// INTEGRATION.md shows the Python side (jax.ffi.register_ffi_target /
// jax.ffi.ffi_call with input_output_aliases for the in-place buffers).
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include <cstdint>
#include <vector>

#include "../../include/sgmc_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Status(int rc) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, sgmc_last_error());
}

// theta f32[C,P], grad f32[C,P], keys u32[C,2] -> theta' (aliased), keys'
static ffi::Error SgldUpdateImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> theta,
                                 ffi::Buffer<ffi::F32> grad, ffi::Buffer<ffi::U32> keys,
                                 ffi::Span<const int64_t> leaf_sizes, float step_size,
                                 float temperature, int32_t prng_layout,
                                 ffi::ResultBuffer<ffi::F32> theta_out,
                                 ffi::ResultBuffer<ffi::U32> keys_out) {
  // theta_out aliases theta (input_output_aliases={0: 0}); the kernel updates in place
  const int64_t C = theta.dimensions()[0];
  return Status(sgmc_sgld_update(stream, theta_out->typed_data(), grad.typed_data(),
                                 keys.typed_data(), keys_out->typed_data(), C,
                                 leaf_sizes.begin(), (int)leaf_sizes.size(), step_size,
                                 temperature, nullptr, prng_layout));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    sgmc_ffi_sgld_update, SgldUpdateImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::U32>>()
        .Attr<ffi::Span<const int64_t>>("leaf_sizes")
        .Attr<float>("step_size")
        .Attr<float>("temperature")
        .Attr<int32_t>("prng_layout")
        .Ret<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::U32>>());

// pSGLD: theta, v aliased in/out
static ffi::Error SgldRmsUpdateImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> theta,
                                    ffi::Buffer<ffi::F32> v, ffi::Buffer<ffi::F32> grad,
                                    ffi::Buffer<ffi::U32> keys,
                                    ffi::Span<const int64_t> leaf_sizes, float step_size,
                                    float temperature, float alpha, float lmbd,
                                    int32_t prng_layout,
                                    ffi::ResultBuffer<ffi::F32> theta_out,
                                    ffi::ResultBuffer<ffi::F32> v_out,
                                    ffi::ResultBuffer<ffi::U32> keys_out) {
  const int64_t C = theta.dimensions()[0];
  return Status(sgmc_sgld_rms_update(stream, theta_out->typed_data(), v_out->typed_data(),
                                     grad.typed_data(), keys.typed_data(),
                                     keys_out->typed_data(), C, leaf_sizes.begin(),
                                     (int)leaf_sizes.size(), step_size, temperature,
                                     nullptr, alpha, lmbd, prng_layout));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    sgmc_ffi_sgld_rms_update, SgldRmsUpdateImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::U32>>()
        .Attr<ffi::Span<const int64_t>>("leaf_sizes")
        .Attr<float>("step_size")
        .Attr<float>("temperature")
        .Attr<float>("alpha")
        .Attr<float>("lmbd")
        .Attr<int32_t>("prng_layout")
        .Ret<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::U32>>());

// integrator.random_tree: keys u32[C,2] -> noise f32[C,P]
static ffi::Error NormalLikeImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> keys,
                                 ffi::Span<const int64_t> leaf_sizes, int32_t prng_layout,
                                 ffi::ResultBuffer<ffi::F32> noise) {
  return Status(sgmc_normal_like(stream, keys.typed_data(), noise->typed_data(),
                                 keys.dimensions()[0], leaf_sizes.begin(),
                                 (int)leaf_sizes.size(), prng_layout));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    sgmc_ffi_normal_like, NormalLikeImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::U32>>()
        .Attr<ffi::Span<const int64_t>>("leaf_sizes")
        .Attr<int32_t>("prng_layout")
        .Ret<ffi::Buffer<ffi::F32>>());

// GLM potential + gradient (logistic family shown; the spec travels as attrs)
static ffi::Error GlmPotentialGradImpl(
    cudaStream_t stream, ffi::Buffer<ffi::F32> theta, ffi::Buffer<ffi::F32> X,
    ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::S32> idx, ffi::Buffer<ffi::U8> workspace,
    int32_t family, int32_t w_off, int32_t aux_off, int32_t prior, int32_t prior_off,
    int32_t prior_size, float prior_scale, float temperature, int64_t observation_count,
    int32_t path, ffi::ResultBuffer<ffi::F32> potential, ffi::ResultBuffer<ffi::F32> variance,
    ffi::ResultBuffer<ffi::F32> grad) {
  sgmc_glm_spec spec{family, (int32_t)X.dimensions()[1], w_off, aux_off, prior, prior_off,
                     prior_size, prior_scale, temperature, /*x_absmax=*/0.0f};
  return Status(sgmc_glm_potential_grad(
      stream, &spec, theta.typed_data(), theta.dimensions()[0], theta.dimensions()[1],
      X.typed_data(), y.typed_data(), idx.typed_data(), nullptr, idx.dimensions()[0],
      observation_count, potential->typed_data(), variance->typed_data(),
      grad->typed_data(), nullptr, workspace.untyped_data(), workspace.size_bytes(), path));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    sgmc_ffi_glm_potential_grad, GlmPotentialGradImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::Buffer<ffi::U8>>()
        .Attr<int32_t>("family")
        .Attr<int32_t>("w_off")
        .Attr<int32_t>("aux_off")
        .Attr<int32_t>("prior")
        .Attr<int32_t>("prior_off")
        .Attr<int32_t>("prior_size")
        .Attr<float>("prior_scale")
        .Attr<float>("temperature")
        .Attr<int64_t>("observation_count")
        .Attr<int32_t>("path")
        .Ret<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>());

// sghmc_begin / sghmc_step / obabo_pass_a / obabo_pass_b / minibatch_draw /
// resgld_decide / resgld_ladder_step follow the same pattern (stream from the
// context, buffers -> typed_data(), scalars as attributes, in-place buffers
// declared with input_output_aliases on the Python side).
#endif  // __has_include("xla/ffi/api/ffi.h")
