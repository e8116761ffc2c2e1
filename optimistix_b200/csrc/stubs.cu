// Entry points declared in include/b200opt.h whose kernels are not built yet return B200OPT_ENOTIMPL
// (loudly - the Python layer raises).  Each stub disappears when its kernel lands.
#include "common.cuh"
using namespace b200opt;
#define NOTIMPL(name) \
  set_error(name ": not implemented in this build"); \
  return B200OPT_ENOTIMPL;
extern "C" {
int64_t b200opt_gn_gram_workspace_size(int32_t) { return -1; }
int32_t b200opt_gn_gram(const float*, const float*, int64_t, int32_t, float*, void*, void*) { NOTIMPL("gn_gram") }
int32_t b200opt_gn_solve(const float*, int32_t, double, float*, int32_t*, void*) { NOTIMPL("gn_solve") }
int32_t b200opt_linear_residual(const float*, const float*, const float*, int64_t, int32_t, float*, void*) {
  NOTIMPL("linear_residual")
}
}
