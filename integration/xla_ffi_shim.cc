// XLA FFI handlers over the C ABI of include/phlash_b200.h: the device-buffer path that removes the
// jax.pure_callback host round trip of the reference (src/phlash/gpu.py:441-465).
//
// COMPILE-GATED: needs the XLA FFI headers that ship with jaxlib (jax.ffi.include_dir()); neither jax
// nor those headers exist in this repository's build image, so this file is NOT part of
// __graft_entry__.build() and has not been compiled here.  Build on a machine with jax:
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include -I/usr/local/cuda/include integration/xla_ffi_shim.cc \
//       -L phlash_b200/_lib -lphlash_b200 -o phlash_b200/_lib/libphlash_b200_xla.so
// Python side: INTEGRATION.md section 3.
#include <cstdint>

#include <cuda_runtime_api.h>

#include "phlash_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// value + gradient, shared parameter rows [B, 6, M], per-pair pi [B, S, M]
static ffi::Error LoglikGradImpl(cudaStream_t stream, int64_t handle, ffi::Buffer<ffi::F32> params6,
                                 ffi::Buffer<ffi::F32> pi, ffi::Buffer<ffi::S64> inds,
                                 ffi::ResultBuffer<ffi::F64> ll, ffi::ResultBuffer<ffi::F32> dlog) {
    const auto pd = params6.dimensions();
    if (pd.size() != 3 || pd[1] != 6) return ffi::Error::InvalidArgument("params6 must be [B, 6, M]");
    const int64_t B = pd[0], M = pd[2], S = inds.dimensions()[0];
    auto *k = reinterpret_cast<phb_kernel *>(handle);
    if (phb_M(k) != M) return ffi::Error::InvalidArgument("M does not match the kernel object");
    const int rc = phb_loglik_device(k, params6.typed_data(), 6 * M, 0, pi.typed_data(), S * M, M, inds.typed_data(),
                                     B, S, /*want_grad=*/1, ll->typed_data(), dlog->typed_data(), stream);
    return rc == PHB_OK ? ffi::Error::Success() : ffi::Error::Internal(phb_last_error());
}

// fused warm-up: per-particle rows [B, 7, M] (pi row = stationary pi), kernel built on full chunks
static ffi::Error WarmupLoglikGradImpl(cudaStream_t stream, int64_t handle, int64_t overlap,
                                       ffi::Buffer<ffi::F32> params7, ffi::Buffer<ffi::S64> inds,
                                       ffi::ResultBuffer<ffi::F64> ll, ffi::ResultBuffer<ffi::F32> dlog) {
    const auto pd = params7.dimensions();
    if (pd.size() != 3 || pd[1] != 7) return ffi::Error::InvalidArgument("params7 must be [B, 7, M]");
    auto *k = reinterpret_cast<phb_kernel *>(handle);
    const int rc = phb_loglik_warmup_device(k, params7.typed_data(), inds.typed_data(), pd[0], inds.dimensions()[0],
                                            overlap, /*want_grad=*/1, ll->typed_data(), dlog->typed_data(), stream);
    return rc == PHB_OK ? ffi::Error::Success() : ffi::Error::Internal(phb_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(PhbLoglikGrad, LoglikGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(PhbWarmupLoglikGrad, WarmupLoglikGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<int64_t>("overlap")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// the whole HMM term of log_density and its gradient w.r.t. the flattened particles x [B, P]
// (INTEGRATION.md section 4); `widths` is the parsed pattern (util.py:8-37), one int32 per epoch
static ffi::Error HmmTermImpl(cudaStream_t stream, int64_t handle, int64_t overlap, double theta, double weight,
                              ffi::Span<const int32_t> widths, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::S64> inds,
                              ffi::ResultBuffer<ffi::F64> value, ffi::ResultBuffer<ffi::F64> grad_x) {
    const auto xd = x.dimensions();
    if (xd.size() != 2 || xd[1] != int64_t(widths.size()) + 3) return ffi::Error::InvalidArgument("x must be [B, 2 + n_epochs + 1]");
    auto *k = reinterpret_cast<phb_kernel *>(handle);
    const int rc = phb_hmm_term_device(k, x.typed_data(), xd[0], widths.begin(), int(widths.size()), theta, inds.typed_data(),
                                       inds.dimensions()[0], overlap, weight, value->typed_data(), grad_x->typed_data(), stream);
    return rc == PHB_OK ? ffi::Error::Success() : ffi::Error::Internal(phb_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(PhbHmmTerm, HmmTermImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<int64_t>("overlap")
                                  .Attr<double>("theta")
                                  .Attr<double>("weight")
                                  .Attr<ffi::Span<const int32_t>>("widths")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
