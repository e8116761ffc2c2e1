// XLA FFI handlers over the C ABI of include/b200opt.h - the binding a maintainer of patrick-kidger/optimistix would
// build to call this library from inside `jax.jit` / `jax.vmap`.
//
// NOT COMPILED OR TESTED IN THIS REPOSITORY: neither jax nor the xla/ffi headers exist in the build image.
// Build (where jax is installed):
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I../include \
//       b200opt_xla.cc -L../optimistix_b200 -lb200opt -o libb200opt_xla.so
// Every handler only forwards raw device pointers and the XLA stream: the C entry points enqueue on that stream, never
// synchronise and allocate nothing, which is what XLA requires of a GPU custom call.
#include <cstdint>

#include "b200opt.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error status(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, b200opt_last_error());
}
static int32_t dtype_of(ffi::DataType t) { return t == ffi::F64 ? B200OPT_F64 : B200OPT_F32; }
static b200opt_options_t options(double rtol, double atol, int32_t norm, int32_t max_steps) {
  b200opt_options_t o;
  b200opt_default_options(&o);
  o.rtol = rtol, o.atol = atol, o.norm = norm, o.max_steps = max_steps;
  return o;
}

// ---- optx.least_squares(registered family, LevenbergMarquardt | GaussNewton): one persistent kernel for the batch
static ffi::Error LeastSquaresImpl(cudaStream_t stream, int32_t solver, int32_t model_id, double rtol, double atol, int32_t norm,
                                   int32_t max_steps, int32_t linear_solver /*B200OPT_LINSOLVE_*: the solver's linear_solver= field*/,
                                   int32_t data_dtype /*B200OPT_DATA_*: 16-bit data rows*/, ffi::AnyBuffer y0 /*[B,n]*/, ffi::AnyBuffer shared /*[m]*/,
                                   ffi::AnyBuffer prob /*[B,m]*/, ffi::Result<ffi::AnyBuffer> y /*[B,n]*/,
                                   ffi::Result<ffi::Buffer<ffi::S32>> num_steps, ffi::Result<ffi::Buffer<ffi::S32>> result) {
  const int64_t B = y0.dimensions()[0];
  const int32_t n = (int32_t)y0.dimensions()[1], m = (int32_t)prob.dimensions()[1];
  b200opt_options_t o = options(rtol, atol, norm, max_steps);
  o.linear_solver = linear_solver, o.data_dtype = data_dtype;
  o.scratch = nullptr;  // or 8 bytes from the handler's ffi::ScratchAllocator (include/b200opt.h: b200opt_options_t.scratch)
  return status(b200opt_least_squares(solver, model_id, dtype_of(y0.element_type()), B, n, m, y0.untyped_data(), shared.untyped_data(),
                                      prob.untyped_data(), m, &o, y->untyped_data(), num_steps->typed_data(), result->typed_data(),
                                      nullptr, nullptr, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(B200optLeastSquares, LeastSquaresImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int32_t>("solver")
                                  .Attr<int32_t>("model_id")
                                  .Attr<double>("rtol")
                                  .Attr<double>("atol")
                                  .Attr<int32_t>("norm")
                                  .Attr<int32_t>("max_steps")
                                  .Attr<int32_t>("linear_solver")
                                  .Attr<int32_t>("data_dtype")
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- one AbstractGaussNewton.step for an arbitrary fn: r and the materialised J come from jax.linearize in the caller's
// while_loop; y, y_eval, state, num_steps, result are the loop carry (donated / aliased input-output buffers).
static ffi::Error LsqStepImpl(cudaStream_t stream, int32_t solver, double rtol, double atol, int32_t norm, int32_t max_steps,
                              int32_t root_near_zero, ffi::AnyBuffer r /*[B,m]*/, ffi::AnyBuffer J /*[B,m,n]*/,
                              ffi::Result<ffi::AnyBuffer> y, ffi::Result<ffi::AnyBuffer> y_eval, ffi::Result<ffi::AnyBuffer> state,
                              ffi::Result<ffi::Buffer<ffi::S32>> num_steps, ffi::Result<ffi::Buffer<ffi::S32>> result,
                              ffi::Result<ffi::Buffer<ffi::S32>> active) {
  const int64_t B = J.dimensions()[0];
  const int32_t m = (int32_t)J.dimensions()[1], n = (int32_t)J.dimensions()[2];
  b200opt_options_t o = options(rtol, atol, norm, max_steps);
  o.root_near_zero = root_near_zero;
  return status(b200opt_lsq_step(solver, dtype_of(J.element_type()), B, n, m, r.untyped_data(), J.untyped_data(), y->untyped_data(),
                                 y_eval->untyped_data(), state->untyped_data(), &o, num_steps->typed_data(), result->typed_data(),
                                 active->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(B200optLsqStep, LsqStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int32_t>("solver")
                                  .Attr<double>("rtol")
                                  .Attr<double>("atol")
                                  .Attr<int32_t>("norm")
                                  .Attr<int32_t>("max_steps")
                                  .Attr<int32_t>("root_near_zero")
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- one Newton / Chord step for an arbitrary fn (newton_chord.py:103-196): f and the n x n Jacobian from the caller
static ffi::Error RootStepImpl(cudaStream_t stream, double rtol, double atol, double kappa, int32_t norm, int32_t max_steps,
                               int32_t cauchy_termination, ffi::AnyBuffer f /*[B,n]*/, ffi::AnyBuffer J /*[B,n,n]*/,
                               ffi::Result<ffi::AnyBuffer> y, ffi::Result<ffi::AnyBuffer> state,
                               ffi::Result<ffi::Buffer<ffi::S32>> num_steps, ffi::Result<ffi::Buffer<ffi::S32>> result,
                               ffi::Result<ffi::Buffer<ffi::S32>> active) {
  const int64_t B = f.dimensions()[0];
  const int32_t n = (int32_t)f.dimensions()[1];
  b200opt_options_t o = options(rtol, atol, norm, max_steps);
  o.kappa = kappa, o.cauchy_termination = cauchy_termination;
  return status(b200opt_root_step(dtype_of(f.element_type()), B, n, f.untyped_data(), J.untyped_data(), nullptr, nullptr,
                                  y->untyped_data(), state->untyped_data(), &o, num_steps->typed_data(), result->typed_data(),
                                  active->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(B200optRootStep, RootStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<double>("rtol")
                                  .Attr<double>("atol")
                                  .Attr<double>("kappa")
                                  .Attr<int32_t>("norm")
                                  .Attr<int32_t>("max_steps")
                                  .Attr<int32_t>("cauchy_termination")
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- one quasi-Newton step for an arbitrary fn (quasi_newton.py:204-300): f and grad from the caller
static ffi::Error BfgsStepImpl(cudaStream_t stream, int32_t update, int32_t use_inverse, double rtol, double atol, int32_t norm,
                               int32_t max_steps, ffi::AnyBuffer f /*[B]*/, ffi::AnyBuffer g /*[B,n]*/, ffi::Result<ffi::AnyBuffer> y,
                               ffi::Result<ffi::AnyBuffer> y_eval, ffi::Result<ffi::AnyBuffer> state,
                               ffi::Result<ffi::Buffer<ffi::S32>> num_steps, ffi::Result<ffi::Buffer<ffi::S32>> result,
                               ffi::Result<ffi::Buffer<ffi::S32>> active) {
  const int64_t B = g.dimensions()[0];
  const int32_t n = (int32_t)g.dimensions()[1];
  const b200opt_options_t o = options(rtol, atol, norm, max_steps);
  return status(b200opt_bfgs_step(update, use_inverse, dtype_of(g.element_type()), B, n, f.untyped_data(), g.untyped_data(), nullptr,
                                  y->untyped_data(), y_eval->untyped_data(), state->untyped_data(), &o, num_steps->typed_data(),
                                  result->typed_data(), active->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(B200optBfgsStep, BfgsStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int32_t>("update")
                                  .Attr<int32_t>("use_inverse")
                                  .Attr<double>("rtol")
                                  .Attr<double>("atol")
                                  .Attr<int32_t>("norm")
                                  .Attr<int32_t>("max_steps")
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- the linear solve of ImplicitAdjoint (_ad.py:104-110)
static ffi::Error LuSolveImpl(cudaStream_t stream, ffi::AnyBuffer A /*[B,n,n]*/, ffi::AnyBuffer b /*[B,n]*/, ffi::Result<ffi::AnyBuffer> x,
                              ffi::Result<ffi::Buffer<ffi::S32>> st) {
  const int64_t B = b.dimensions()[0];
  const int32_t n = (int32_t)b.dimensions()[1];
  return status(b200opt_lu_solve(dtype_of(b.element_type()), B, n, A.untyped_data(), b.untyped_data(), x->untyped_data(), st->typed_data(),
                                 stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(B200optLuSolve, LuSolveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::Buffer<ffi::S32>>());
