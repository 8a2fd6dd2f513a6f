// XLA FFI (jax.ffi) registration of the hot-path entry points -- the "thin jax.ffi / C-ABI custom-call layer" of
// BASELINE.json's north_star, for a maintainer who wants to call libnkfb200.so from inside the reference's jitted
// functions (dispatch points: netket_fidelity/infidelity/overlap_U/expect.py:40-47, overlap/expect.py:33-40).
//
// COMPILE-GUARDED: this file needs `xla/ffi/api/ffi.h` from a jaxlib installation (jaxlib/include), which this image
// does not have, so it is NOT part of `make` and has not been compiled or run here.  Build where jaxlib exists:
//   g++ -std=c++17 -shared -fPIC -DNKFB_WITH_XLA_FFI -I$(python -c 'import jaxlib,os;print(os.path.join(os.path.dirname(jaxlib.__file__),"include"))') \
//       -I../../include -I/usr/local/cuda/include xla_ffi_shim.cc -o libnkfb200_ffi.so -L../lib -lnkfb200
// and register from Python:
//   import ctypes, jax
//   ffi = ctypes.CDLL("libnkfb200_ffi.so"); ffi.nkfb_ffi_init(jax.devices()[0].id)
//   for name in ("NkfbLogPsi", "NkfbInfidelityFwd", "NkfbInfidelityGrad"):
//       jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(ffi, name)), platform="CUDA")
//   lp = jax.ffi.ffi_call("NkfbLogPsi", jax.ShapeDtypeStruct((B,), jnp.complex64))(W, b, a, x_packed)
// Every handler forwards raw device pointers and the XLA stream to the C ABI of include/nkfb200.h; gates travel as
// attributes (site index + the 2x2 table of matrix elements the reference computes at
// operator/singlequbit_gates.py:101-103,196-199,282-286), local states as two doubles.
#ifdef NKFB_WITH_XLA_FFI

#include <cuda_runtime.h>

#include <cmath>
#include <cstring>

#include "nkfb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static nkfb_handle_t g_handle = nullptr;

extern "C" int nkfb_ffi_init(int device) { return g_handle ? 0 : nkfb_create(device, &g_handle); }

static nkfb_rbm_t rbm_of(ffi::Buffer<ffi::C64> W, ffi::Buffer<ffi::C64> b, ffi::Buffer<ffi::C64> a) {
  nkfb_rbm_t net;
  std::memset(&net, 0, sizeof(net));
  net.n_visible = (int)W.dimensions()[0];
  net.n_hidden = (int)W.dimensions()[1];
  net.dtype = NKFB_F32;
  net.kernel = W.typed_data();
  net.bias = b.element_count() ? b.typed_data() : nullptr;
  net.visible_bias = a.element_count() ? a.typed_data() : nullptr;
  return net;
}

// gate attribute: idx >= 0 and 8 doubles mel[bit][conn][re/im]; idx < 0 means "no operator"
static const nkfb_op_t *gate_of(nkfb_op_t *op, int32_t idx, ffi::Span<const double> mel) {
  if (idx < 0) return nullptr;
  std::memset(op, 0, sizeof(*op));
  op->kind = NKFB_OP_GATE;
  op->idx = idx;
  for (int q = 0; q < 8 && q < (int)mel.size(); ++q) op->gate_mel[q] = mel[q];
  return op;
}

static ffi::Error status(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(nkfb_last_error(g_handle));
}

// log psi(x) or log(U psi)(x): utils/sampling_Ustate.py:34-47
static ffi::Error LogPsiImpl(cudaStream_t s, ffi::Buffer<ffi::C64> W, ffi::Buffer<ffi::C64> b, ffi::Buffer<ffi::C64> a,
                             ffi::Buffer<ffi::U32> x, ffi::ResultBuffer<ffi::C64> out, double s0, double s1,
                             int32_t gate_idx, ffi::Span<const double> gate_mel) {
  nkfb_rbm_t net = rbm_of(W, b, a);
  nkfb_op_t op;
  const double ls[2] = {s0, s1};
  return status(nkfb_logpsi(g_handle, &net, gate_of(&op, gate_idx, gate_mel), x.typed_data(), x.dimensions()[0], ls,
                            out->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(NkfbLogPsi, LogPsiImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::C64>>()
                                  .Attr<double>("s0").Attr<double>("s1")
                                  .Attr<int32_t>("gate_idx").Attr<ffi::Span<const double>>("gate_mel"));

// pass A: local estimator + control variates + sums (overlap_U/expect.py:101-120, overlap/expect.py:78-89)
static ffi::Error InfidelityFwdImpl(cudaStream_t s, ffi::Buffer<ffi::C64> Wp, ffi::Buffer<ffi::C64> bp,
                                    ffi::Buffer<ffi::C64> ap, ffi::Buffer<ffi::C64> Wt, ffi::Buffer<ffi::C64> bt,
                                    ffi::Buffer<ffi::C64> at, ffi::Buffer<ffi::U32> sigma, ffi::Buffer<ffi::U32> eta,
                                    ffi::ResultBuffer<ffi::C64> log_val, ffi::ResultBuffer<ffi::F32> L,
                                    ffi::ResultBuffer<ffi::C64> rho1, ffi::ResultBuffer<ffi::F64> sums, double s0,
                                    double s1, double cv, int32_t u_idx, ffi::Span<const double> u_mel, int32_t ud_idx,
                                    ffi::Span<const double> ud_mel) {
  nkfb_rbm_t psi = rbm_of(Wp, bp, ap), phi = rbm_of(Wt, bt, at);
  nkfb_op_t u, ud;
  const double ls[2] = {s0, s1};
  cudaMemsetAsync(sums->typed_data(), 0, 8 * sizeof(double), s);
  return status(nkfb_infidelity_fwd(g_handle, &psi, &phi, gate_of(&u, u_idx, u_mel), gate_of(&ud, ud_idx, ud_mel), nullptr,
                                    sigma.typed_data(), eta.typed_data(), sigma.dimensions()[0], ls, cv,
                                    log_val->typed_data(), L->typed_data(), rho1->typed_data(), sums->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(NkfbInfidelityFwd, InfidelityFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>()
                                  .Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>()
                                  .Arg<ffi::Buffer<ffi::U32>>().Arg<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::C64>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::C64>>().Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("s0").Attr<double>("s1").Attr<double>("cv")
                                  .Attr<int32_t>("u_idx").Attr<ffi::Span<const double>>("u_mel")
                                  .Attr<int32_t>("ud_idx").Attr<ffi::Span<const double>>("ud_mel"));

// pass B: score-function gradient accumulator (nkjax.expect + nkjax.vjp(conjugate=True) at overlap_U/expect.py:134-156)
static ffi::Error InfidelityGradImpl(cudaStream_t s, ffi::Buffer<ffi::C64> Wp, ffi::Buffer<ffi::C64> bp,
                                     ffi::Buffer<ffi::C64> ap, ffi::Buffer<ffi::U32> sigma, ffi::Buffer<ffi::U32> eta,
                                     ffi::Buffer<ffi::C64> log_val, ffi::Buffer<ffi::C64> rho1, ffi::Buffer<ffi::F64> sums,
                                     ffi::ResultBuffer<ffi::F64> grad_acc, double s0, double s1, double cv,
                                     int32_t ud_idx, ffi::Span<const double> ud_mel) {
  nkfb_rbm_t psi = rbm_of(Wp, bp, ap);
  nkfb_op_t ud;
  const double ls[2] = {s0, s1};
  cudaMemsetAsync(grad_acc->typed_data(), 0, grad_acc->element_count() * sizeof(double), s);
  return status(nkfb_infidelity_grad(g_handle, &psi, gate_of(&ud, ud_idx, ud_mel), sigma.typed_data(), eta.typed_data(),
                                     sigma.dimensions()[0], ls, cv, log_val.typed_data(), rho1.typed_data(),
                                     sums.typed_data(), grad_acc->typed_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(NkfbInfidelityGrad, InfidelityGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>()
                                  .Arg<ffi::Buffer<ffi::U32>>().Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::C64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("s0").Attr<double>("s1").Attr<double>("cv")
                                  .Attr<int32_t>("ud_idx").Attr<ffi::Span<const double>>("ud_mel"));

#endif  // NKFB_WITH_XLA_FFI
