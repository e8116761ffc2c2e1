// Device-resident iteration state of the single large Gauss-Newton problem (BASELINE configs[4]; b200opt_gn_iterate).
// It is the carry of optimistix/_iterate.py:194-268 for AbstractGaussNewton with LearningRate(1.0) + NewtonDescent
// (gauss_newton.py:216-334): (y, y_eval, num_steps, terminate, result) plus what the two kernels of an iteration hand to
// each other.  Both kernels start by looking at it: once the solve has terminated (or failed, or used up max_steps) every
// further enqueued iteration is a no-op, so the host can run ahead of the device and read the state back only now and then.
#pragma once
#include "common.cuh"

namespace b200opt {

struct GnState {
  float y[256];       // accepted iterate (the carry's y)
  float y_eval[256];  // trial iterate the next residual / Gram pass is evaluated at
  int32_t steps;      // num_steps; its parity also selects the residual ping-pong buffer
  int32_t terminate;  // cauchy_termination on (y, f) of the last step (_misc.py:291-310)
  int32_t result;     // RESULTS code of the last step (the linear solve's; nonfinite; < 0: B200OPT_EXCHANGE_TIMEOUT)
  int32_t max_steps;
  float rtol, atol;
  float rr;           // r^T r at the last evaluated point (all ranks)
  float fstat;        // max_i |r_i - r_old_i| / (atol + rtol |r_i|) at the last evaluated point (all ranks)
  int32_t timeout_ms; // bound of the in-kernel wait for the peers' flags
  int32_t solver;     // 0: fp32 blocked Cholesky + fp64 refinement (fp64 fallback when it is not accurate), 1: fp64 only
  int32_t fallbacks;  // how many solves took the fp64 fallback (diagnostic)
  int32_t pad;
};
static_assert(sizeof(GnState) == 2048 + 48, "GnState layout is part of the C ABI (b200opt_gn_state_size)");

__device__ __forceinline__ bool gn_frozen(const GnState* s) {
  return s->terminate != 0 || s->result != RES_OK || s->steps >= s->max_steps;
}

}  // namespace b200opt
