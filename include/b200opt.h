/*
 * b200opt.h - C ABI of the B200-native Optimistix hot path.
 *
 * The reference (patrick-kidger/optimistix 0.1.0) is pure Python traced into one XLA program and has
 * NO plugin/FFI interface of its own.  The seam this library sits behind is
 * the body of the jit'd entry points, i.e. what `jax.vmap(optx.least_squares)(...)` lowers to:
 *
 *   b200opt_least_squares  replaces  optimistix/_least_squares.py:44-135  -> _iterate.py:194-268 driving
 *                                    _solver/levenberg_marquardt.py:256-294 / _solver/gauss_newton.py:361-395
 *   b200opt_root_find      replaces  optimistix/_root_find.py:119-241 (and _fixed_point.py:50-129 through
 *                                    `fn(y) - y`)  -> _solver/newton_chord.py:59-251
 *   b200opt_minimise       replaces  optimistix/_minimise.py:37-119  -> _solver/quasi_newton.py:418-458 (BFGS)
 *   b200opt_lsq_step       replaces  one AbstractGaussNewton.step (gauss_newton.py:243-334) for an arbitrary
 *                                    user `fn`: the caller (XLA / torch) evaluates r and J, the kernel does
 *                                    search + accept/reject + termination + damped solve
 *   b200opt_gn_gram        replaces  the `lx.linear_solve(J, r)` of NewtonDescent (gauss_newton.py:64-78) for
 *                                    the single large row-sharded problem: per-shard J^T J, J^T r, r^T r
 *
 * Conventions
 *   - every pointer in the non-`_host` entry points is a DEVICE pointer; the call only enqueues work on
 *     `stream` (a cudaStream_t) and never synchronises, so it can be called from an XLA FFI handler.
 *   - the `_host` entry points take HOST pointers, copy in/out themselves (chunked, double-buffered over
 *     two streams) and return when the results are in the output buffers.
 *   - the caller owns every buffer.  The device-pointer entry points keep no state and allocate nothing.
 *     The `_host` variants keep per-host-thread staging (two streams, device slabs, page-locked bounce
 *     buffers; grow-only) between calls; b200opt_host_cache_release() frees it for the calling thread.
 *   - return value: 0 on success, a negative B200OPT_E* code otherwise; b200opt_last_error() gives the
 *     message for the calling thread.  Per-problem outcomes go to `result[]` using the member names of
 *     optimistix.RESULTS (optimistix/_solution.py:11-20; compare by NAME, b200opt_result_name()).
 *   - batched arrays are row-major with the problem index leading: y0[batch][n], per-problem args
 *     [batch][problem_args_stride].
 */
#ifndef B200OPT_H_
#define B200OPT_H_

#ifdef __CUDACC_RTC__ /* NVRTC (run-time compiled model families) has no host headers */
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define B200OPT_VERSION 100 /* 0.1.0, tracks the reference version restated */

/* error codes (return values) */
#define B200OPT_OK 0
#define B200OPT_EINVAL (-1)   /* bad argument (unknown model/dtype/solver, n or m not supported by the model) */
#define B200OPT_ECUDA (-2)    /* CUDA runtime error (launch failure, out of memory in a _host call) */
#define B200OPT_ENOTIMPL (-3) /* combination not compiled into this build */
#define B200OPT_EXCHANGE_TIMEOUT (-4) /* device-side status: a peer of the row-sharded exchange never published its block */

/* dtype of y / args / outputs: the arithmetic type of the whole solve (jax_enable_x64 analogue) */
#define B200OPT_F32 0
#define B200OPT_F64 1

/* norms: optimistix.max_norm / two_norm / rms_norm  (optimistix/_misc.py:41-45) */
#define B200OPT_NORM_MAX 0
#define B200OPT_NORM_TWO 1
#define B200OPT_NORM_RMS 2

/* per-problem result codes; names via b200opt_result_name(). lineax members first (lx.RESULTS), then
 * optimistix/_solution.py:11-20 */
#define B200OPT_RESULT_SUCCESSFUL 0
#define B200OPT_RESULT_MAX_STEPS_REACHED 1          /* lineax (iterative linear solvers; unused here) */
#define B200OPT_RESULT_SINGULAR 2                   /* lineax: non-finite linear solution */
#define B200OPT_RESULT_BREAKDOWN 3                  /* lineax (unused here) */
#define B200OPT_RESULT_STAGNATION 4                 /* lineax (unused here) */
#define B200OPT_RESULT_NONLINEAR_MAX_STEPS_REACHED 5
#define B200OPT_RESULT_NONLINEAR_DIVERGENCE 6
#define B200OPT_RESULT_NONFINITE 7

/* solver ids */
#define B200OPT_SOLVER_LEVENBERG_MARQUARDT 0 /* ClassicalTrustRegion + DampedNewtonDescent */
#define B200OPT_SOLVER_GAUSS_NEWTON 1        /* LearningRate(1.0) + NewtonDescent */
#define B200OPT_SOLVER_NEWTON 2
#define B200OPT_SOLVER_CHORD 3
#define B200OPT_SOLVER_BFGS 4                /* use_inverse=True, BacktrackingArmijo */

/* model kinds */
#define B200OPT_KIND_LEAST_SQUARES 0 /* residual model: y[n] -> r[m] */
#define B200OPT_KIND_ROOT 1          /* y[n] -> F[n] */
#define B200OPT_KIND_MINIMISE 2      /* y[n] -> scalar */

/* Solver options: the constructor fields of the reference solver classes
 * (levenberg_marquardt.py:281-294, gauss_newton.py:382-395, newton_chord.py:59-65, quasi_newton.py:443-458)
 * plus the `max_steps` keyword of the entry points (default 256, _least_squares.py:53). */
typedef struct b200opt_options {
  double rtol;
  double atol;
  int32_t norm;               /* B200OPT_NORM_* */
  int32_t max_steps;          /* >= 0 */
  int32_t cauchy_termination; /* Newton/Chord only: 1 (default) or 0 = Hairer-Wanner test, newton_chord.py:179-192 */
  int32_t root_near_zero;     /* 1 = also require norm(f) < atol (root_find through LM/BFGS, _root_find.py:45-80) */
  double kappa;               /* Newton/Chord non-Cauchy mode, default 1e-2 */
  int32_t linear_solver;      /* least squares only: B200OPT_LINSOLVE_* (the `linear_solver=` field of LevenbergMarquardt /
                               * GaussNewton, levenberg_marquardt.py:286, gauss_newton.py:387) */
  int32_t data_dtype;         /* least squares only: B200OPT_DATA_* - storage type of the per-problem data rows */
  void* scratch;              /* optional: >= 8 bytes of DEVICE memory, 8-byte aligned, that this call may use (the ticket counter of
                               * the dynamic problem queue; zeroed by the library in stream order, free again when the call's work on
                               * the stream is done).  NULL: the library takes the bytes from the stream-ordered allocator
                               * (cudaMallocAsync / cudaFreeAsync) - correct, but the default pool gives its memory back at every
                               * synchronisation, so the first launch after a sync then pays for a fresh allocation (0.5 - 10 ms).
                               * An XLA FFI handler passes its ScratchAllocator's buffer, PyTorch callers a caching-allocator tensor. */
} b200opt_options_t;

/* Storage type of `problem_args` of b200opt_least_squares(_host).  AS_COMPUTE: the compute dtype (the default).  F16 / U16 /
 * I16: 16-bit values, `problem_args_stride` counted in 16-bit elements; every value is converted exactly to the compute dtype
 * when the row is loaded, so the result equals the solve of the converted array bit for bit.  Halves the bytes per problem
 * over PCIe and HBM (measured curves usually ARE 16-bit detector counts). */
#define B200OPT_DATA_AS_COMPUTE 0
#define B200OPT_DATA_F16 1
#define B200OPT_DATA_U16 2
#define B200OPT_DATA_I16 3

/* Linear solve of the LM / GN descent.  NORMAL_CHOLESKY (default; BASELINE north_star): Cholesky of J^T J + lambda I, J never
 * materialised.  QR: the reference's own operations - Householder QR of the augmented system [J; sqrt(lambda) I] p = [r; 0]
 * (levenberg_marquardt.py:66-75) and the predicted reduction from J d + r (trust_region.py:182-186); J of the accepted
 * point is rebuilt in shared memory every step.  Use it when kappa(J)^2 eps is not small. */
#define B200OPT_LINSOLVE_NORMAL_CHOLESKY 0
#define B200OPT_LINSOLVE_QR 1

/* Residual-model plugin registry: the compiled-in model families (device residual/Jacobian functions). */
typedef struct b200opt_model_info {
  int32_t id;
  int32_t kind;            /* B200OPT_KIND_* */
  int32_t n;               /* number of parameters; 0 = any n accepted up to n_max */
  int32_t n_max;
  int32_t m_max;           /* least squares: maximum residual count; 0 = not applicable */
  int32_t shared_args;     /* number of shared-argument elements as a function of (n,m): see shared_layout */
  const char* name;        /* e.g. "exp_decay" */
  const char* params;      /* human-readable parameter layout */
  const char* shared_layout;  /* what `shared_args` points to */
  const char* problem_layout; /* what each row of `problem_args` holds */
} b200opt_model_info_t;

#ifndef __CUDACC_RTC__ /* host entry points: not visible to run-time compiled device code */
int32_t b200opt_version(void);
const char* b200opt_last_error(void);
const char* b200opt_result_name(int32_t result_code);
int32_t b200opt_model_count(void);
int32_t b200opt_model_info_get(int32_t model_id, b200opt_model_info_t* out);
int32_t b200opt_model_id(const char* name); /* -1 if unknown */
void b200opt_default_options(b200opt_options_t* out);

/* ---- run-time registration of a least-squares "curve" family (the user side of the model registry) -----------------
 * struct_body: CUDA C++ text, the members of
 *     template <typename T> struct UserModel { static constexpr int N = n; static constexpr int NAUX = naux; <struct_body> };
 * which must provide, exactly like the built-in families in optimistix_b200/csrc/models.cuh,
 *     struct Pre {...};  static __device__ Pre prepare(const T (&y)[N]);
 *     static __device__ T    residual(const T (&y)[N], const Pre&, const CurveRow<T>& row, T (&aux)[NAUX]);   (row.x abscissa, row.d data)
 *     static __device__ void jac_row (const T (&y)[N], const Pre&, const CurveRow<T>& row, const T (&aux)[NAUX], T (&j)[N]);
 * (the role of the reference's `fn(y, args)` and of jax.linearize on it, gauss_newton.py:158-181).  It is compiled with NVRTC
 * for sm_100a against the library's own kernel headers in `csrc_dir` (optimistix_b200/csrc) and then runs in the same
 * persistent one-warp-per-problem LM / GN kernel as the built-in families (fp32 and fp64).  Needs a CUDA device.
 * Returns the new model id (>= B200OPT_JIT_MODEL_BASE) for b200opt_least_squares[_host], or a negative error code;
 * a compilation error is reported through b200opt_last_error(). */
#define B200OPT_JIT_MODEL_BASE 1000
int32_t b200opt_register_curve_family(const char* name, int32_t n, int32_t naux, const char* struct_body,
                                      const char* csrc_dir);
/* The same for a root-finding family (Newton / Chord persistent kernel, b200opt_root_find[_host]): struct_body provides
 *     static __device__ void eval_row(const T* sh, const T* ys, const Ctx& c, int lane, T& f, T (&row)[N], bool want_jac);
 * lane `lane` evaluates F_lane(y) into f and, if want_jac, row `lane` of the Jacobian into row[0..N); sh = the first
 * shared_elems elements of shared_args (staged in shared memory), ys = the whole iterate y[0..N) (shared memory),
 * c.prob = this problem's row of problem_args.  n must be a multiple of 4, n <= 32. */
int32_t b200opt_register_root_family(const char* name, int32_t n, int32_t shared_elems, const char* struct_body,
                                     const char* csrc_dir);

/* ---- batched nonlinear least squares (LM / GN), one warp per problem, persistent kernel ---------------
 * y0           [batch][n]
 * shared_args  model-specific, shared by all problems (e.g. the abscissa grid t[m])
 * problem_args [batch][problem_args_stride] (e.g. the data d[m])
 * outputs: y_out [batch][n]; num_steps[batch]; result[batch];
 *          optional (may be NULL): step_size_out[batch] (final trust-region radius, dtype),
 *                                  num_accepted_out[batch] (gauss_newton.py:153-155 counter) */
int32_t b200opt_least_squares(int32_t solver, int32_t model_id, int32_t dtype, int64_t batch, int32_t n, int32_t m,
                              const void* y0, const void* shared_args, const void* problem_args,
                              int64_t problem_args_stride, const b200opt_options_t* opts, void* y_out,
                              int32_t* num_steps, int32_t* result, void* step_size_out, int32_t* num_accepted_out,
                              void* stream);

/* ---- batched root finding (Newton / Chord), one warp per problem --------------------------------------
 * lower/upper: optional [n] clip bounds shared by all problems (options["lower"/"upper"], newton_chord.py:112-135) */
int32_t b200opt_root_find(int32_t solver, int32_t model_id, int32_t dtype, int64_t batch, int32_t n, const void* y0,
                          const void* shared_args, const void* problem_args, int64_t problem_args_stride,
                          const void* lower, const void* upper, const b200opt_options_t* opts, void* y_out,
                          int32_t* num_steps, int32_t* result, void* stream);

/* ---- batched minimisation (BFGS, dense inverse Hessian on chip), one CTA per problem -------------------
 * optional outputs (may be NULL): f_out[batch] (dtype), hessian_inv_out[batch][n][n] (dtype), num_accepted_out */
int32_t b200opt_minimise(int32_t solver, int32_t model_id, int32_t dtype, int64_t batch, int32_t n, const void* y0,
                         const void* shared_args, const void* problem_args, int64_t problem_args_stride,
                         const b200opt_options_t* opts, void* y_out, int32_t* num_steps, int32_t* result,
                         void* f_out, void* hessian_inv_out, int32_t* num_accepted_out, void* stream);

/* ---- HOST-buffer variants: same meaning, pointers are host memory; chunked H2D / solve / D2H pipeline.
 * `device` is the CUDA device ordinal; chunk = problems per pipeline stage (0 = library default). */
int32_t b200opt_least_squares_host(int32_t device, int32_t solver, int32_t model_id, int32_t dtype, int64_t batch,
                                   int32_t n, int32_t m, const void* y0, const void* shared_args,
                                   int64_t shared_args_count, const void* problem_args, int64_t problem_args_stride,
                                   const b200opt_options_t* opts, void* y_out, int32_t* num_steps, int32_t* result,
                                   int64_t chunk);
int32_t b200opt_root_find_host(int32_t device, int32_t solver, int32_t model_id, int32_t dtype, int64_t batch,
                               int32_t n, const void* y0, const void* shared_args, int64_t shared_args_count,
                               const void* problem_args, int64_t problem_args_stride, const b200opt_options_t* opts,
                               void* y_out, int32_t* num_steps, int32_t* result, int64_t chunk);
/* As b200opt_root_find_host, with the per-problem arguments given as `n_parts` (<= 4) separate host arrays - part p is
 * [batch][part_elems[p]], contiguous - which the family's layout concatenates per problem (implicit_euler: y_prev [n], dt [1];
 * the arguments of the reference's fn(y, args) stay the arrays the caller already has).  The pipeline copies the parts as they
 * are and a copy kernel lays them out as [chunk][problem_args_stride] on the device (unused trailing elements of a row are
 * zero): no packing pass over the host arrays (71 MB per call for configs[2]). */
int32_t b200opt_root_find_host_parts(int32_t device, int32_t solver, int32_t model_id, int32_t dtype, int64_t batch,
                                     int32_t n, const void* y0, const void* shared_args, int64_t shared_args_count,
                                     const void* const* parts, const int32_t* part_elems, int32_t n_parts,
                                     int64_t problem_args_stride, const b200opt_options_t* opts, void* y_out,
                                     int32_t* num_steps, int32_t* result, int64_t chunk);
void b200opt_host_cache_release(void);
int32_t b200opt_minimise_host(int32_t device, int32_t solver, int32_t model_id, int32_t dtype, int64_t batch, int32_t n,
                              const void* y0, const void* shared_args, int64_t shared_args_count,
                              const void* problem_args, int64_t problem_args_stride, const b200opt_options_t* opts,
                              void* y_out, int32_t* num_steps, int32_t* result, int64_t chunk);

/* ---- generic per-iteration step for an arbitrary user `fn` (J and r computed by the caller) ------------
 * One call = one AbstractGaussNewton.step + the driver bookkeeping of _iterate.py:234-241 for every
 * still-running problem; n <= 64 (n <= 16, fp64: n <= 8: one warp per problem, Gram in registers; up to n = 64: one CTA per problem, Gram
 * in shared memory).  `state` is an opaque per-problem carry of b200opt_lsq_step_state_size(dtype,n,m)
 * bytes per problem, zero-initialised by b200opt_lsq_step_init.
 *   r_eval [batch][m], J_eval [batch][m][n]: residual and Jacobian at y_eval (as written by the last call)
 *   y_eval [batch][n] in/out: the next trial point;  y [batch][n] in/out: last accepted point
 *   active_count (device int32, optional): number of problems still running after this step */
int64_t b200opt_lsq_step_state_size(int32_t dtype, int32_t n, int32_t m);
int32_t b200opt_lsq_step_init(int32_t dtype, int64_t batch, int32_t n, int32_t m, const void* y0, void* y, void* y_eval,
                              void* state, int32_t* num_steps, int32_t* result, void* stream);
int32_t b200opt_lsq_step(int32_t solver, int32_t dtype, int64_t batch, int32_t n, int32_t m, const void* r_eval,
                         const void* J_eval, void* y, void* y_eval, void* state, const b200opt_options_t* opts,
                         int32_t* num_steps, int32_t* result, int32_t* active_count, void* stream);
/* after the loop: result successful & not terminated -> nonlinear_max_steps_reached (_iterate.py:253-257) */
int32_t b200opt_lsq_step_finish(int32_t dtype, int64_t batch, int32_t n, int32_t m, const void* state, int32_t* result,
                                void* stream);

/* ---- generic per-iteration step, root finding (Newton / Chord) for an arbitrary user `fn` -------------------
 * One call = one _AbstractNewtonChord.step + terminate (optimistix/_solver/newton_chord.py:103-196) for every
 * still-running problem; n <= 32: one warp per problem, row-per-lane LU; 32 < n <= 160 (fp64) / 224 (fp32): one CTA per problem,
 * the Jacobian in shared memory.  f_eval [batch][n] = fn(y), J_eval [batch][n][n] = d fn / d y (row-major) at the
 * CURRENT y for Newton; for Chord the caller passes the Jacobian at y0 on every call (newton_chord.py:80-87).
 * lower / upper: optional [n] clip bounds.  y [batch][n] is updated in place.  `state`: opaque carry of
 * b200opt_root_step_state_size bytes per problem.  opts: rtol, atol, norm, max_steps, cauchy_termination, kappa. */
int64_t b200opt_root_step_state_size(int32_t dtype, int32_t n);
int32_t b200opt_root_step_init(int32_t dtype, int64_t batch, int32_t n, const void* y0, void* y, void* state,
                               int32_t* num_steps, int32_t* result, void* stream);
int32_t b200opt_root_step(int32_t dtype, int64_t batch, int32_t n, const void* f_eval, const void* J_eval,
                          const void* lower, const void* upper, void* y, void* state, const b200opt_options_t* opts,
                          int32_t* num_steps, int32_t* result, int32_t* active_count, void* stream);
int32_t b200opt_root_step_finish(int32_t dtype, int64_t batch, int32_t n, const void* state, int32_t* result,
                                 void* stream);

/* ---- batched dense solve A x = b, LU with partial pivoting: n <= 32 one warp per system, n <= 160 (fp64) / 224 (fp32) one CTA ---
 * The linear solve of ImplicitAdjoint (optimistix/_ad.py:83-118: tangents / cotangents of a solution are
 * -(d rewrite / d y)^-1 d rewrite / d args).  A [batch][n][n] row-major, b and x [batch][n]; status (optional)
 * [batch]: successful, or singular when the solution is not finite. */
int32_t b200opt_lu_solve(int32_t dtype, int64_t batch, int32_t n, const void* A, const void* b, void* x, int32_t* status,
                         void* stream);

/* ---- generic per-iteration step, minimisation (BFGS, inverse Hessian, backtracking Armijo) -------------------
 * One call = one AbstractQuasiNewton.step (optimistix/_solver/quasi_newton.py:204-300, :342-415; backtracking.py:
 * 79-99) for every still-running problem; n <= 1024.  f_eval [batch] = fn(y_eval), g_eval [batch][n] = its gradient.
 * root_norm: optional [batch], norm of the root function at the pre-step y: adds the `norm(f) < atol` condition of
 * root_find-through-a-minimiser (_root_find.py:64-73).  The carry (H^-1 n x n, gradient, direction, f, step size,
 * flags) lives in `state` (b200opt_bfgs_step_state_size bytes per problem), initialised by _init.
 * update: B200OPT_UPDATE_BFGS or B200OPT_UPDATE_DFP (quasi_newton.py:487-573).  use_inverse != 0: the carry is the
 * inverse Hessian approximation and the direction is a mat-vec; use_inverse == 0: the carry is the Hessian
 * approximation and the direction comes from a Cholesky solve (n <= 64), non-finite solution => `singular`. */
#define B200OPT_UPDATE_BFGS 0
#define B200OPT_UPDATE_DFP 1
int64_t b200opt_bfgs_step_state_size(int32_t dtype, int32_t n);
int32_t b200opt_bfgs_step_init(int32_t dtype, int64_t batch, int32_t n, const void* y0, void* y, void* y_eval,
                               void* state, int32_t* num_steps, int32_t* result, void* stream);
int32_t b200opt_bfgs_step(int32_t update, int32_t use_inverse, int32_t dtype, int64_t batch, int32_t n,
                          const void* f_eval, const void* g_eval, const void* root_norm, void* y, void* y_eval,
                          void* state, const b200opt_options_t* opts, int32_t* num_steps, int32_t* result,
                          int32_t* active_count, void* stream);
int32_t b200opt_bfgs_step_finish(int32_t dtype, int64_t batch, int32_t n, const void* state, int32_t* result,
                                 void* stream);

/* ---- large row-sharded Gauss-Newton Gram on tcgen05 tensor cores (TF32 in, FP32 accumulate in TMEM) ----
 * J [m_local][n] row-major fp32 with 4 <= n <= 256, n % 4 == 0; r [m_local] fp32.  The kernels work on J padded with zero
 * columns to N = 256 (the padding is never read from memory: whole 32-column groups beyond n are zeroed once in shared memory, a
 * partial group is zero-filled by the TMA): every n-indexed vector of this family (y, p, the state) has 256 entries, the ones
 * beyond n are zero, and `out` keeps the 256-wide layout below with a unit diagonal for the padded unknowns.
 * out [N*N + N + 1] fp32: G = J^T J (row-major, full symmetric), then g = J^T r, then r^T r.
 * The caller all-reduces `out` across ranks (NCCL sum) and calls b200opt_gn_solve. `workspace` must hold
 * b200opt_gn_gram_workspace_size(n) bytes. */
int64_t b200opt_gn_gram_workspace_size(int32_t n);
int32_t b200opt_gn_gram(const float* J, const float* r, int64_t m_local, int32_t n, float* out, void* workspace,
                        void* stream);
/* Fused variant for the "linear_rows" family: forms r = J y - d from the same shared-memory tile the tensor core is
 * multiplying, so J is read from HBM ONCE per Gauss-Newton iteration. Writes r_out [m_local]; if r_old is not NULL
 * also the f-space termination statistic (see b200opt_linear_residual) into *fmax_bits (caller-zeroed). */
int32_t b200opt_gn_gram_linear(const float* J, const float* y, const float* d, const float* r_old, float* r_out,
                               uint32_t* fmax_bits, float rtol, float atol, int64_t m_local, int32_t n, float* out,
                               void* workspace, void* stream);
/* Solve (G + lambda I) p = g in fp64 on one CTA; gram_out is the (all-reduced) output of b200opt_gn_gram;
 * p_out [n] fp32; status_out: device int32 result code (singular if the solution is non-finite); `workspace` is
 * the same buffer that was given to b200opt_gn_gram. */
int32_t b200opt_gn_solve(const float* gram_out, int32_t n, double lambda, float* p_out, int32_t* status_out, void* workspace,
                         void* stream);
/* Multi-rank form: `parts` holds the ALL-GATHERED row-shard partials, nparts blocks of part_stride >= n*n + n + 2 floats,
 * each [G | g | r^T r | f-statistic (IEEE bits of a non-negative float, as written by b200opt_gn_gram_linear)].  The
 * kernel sums them in rank order in fp64 (bit-identical on every rank) before factorising, so one all-gather replaces
 * the sum all-reduce of the Gram and the max all-reduce of the termination statistic.  summary_out (optional, [2]):
 * total r^T r and the maximum f-statistic.  nparts == 1 needs no f-statistic slot unless summary_out is given. */
int32_t b200opt_gn_solve_parts(const float* parts, int32_t nparts, int64_t part_stride, int32_t n, double lambda,
                               float* p_out, int32_t* status_out, float* summary_out, void* workspace, void* stream);
/* ---- peer-memory exchange of the row-sharded problem: the collective INSIDE our own kernels (no NCCL on the data path) ----
 * Every rank owns one exchange buffer (b200opt_xchg_bytes bytes, cudaMalloc'ed by b200opt_xchg_alloc, which also returns
 * its 64-byte CUDA IPC handle).  The ranks swap handles out of band (torch.distributed all_gather_object), map each other's
 * buffers with b200opt_xchg_open and put the `world` base pointers (own buffer = the local pointer) into a device array.
 * b200opt_gn_gram_linear_xchg is b200opt_gn_gram_linear whose reduction kernel also stores this rank's
 * [G | g | r^T r | f-statistic] block into its slot of EVERY peer's buffer (P2P stores over NVLink / NVSwitch) and, from the
 * last CTA, raises this rank's flag there (sequence number `seq`, slot parity `parity` = seq & 1).  b200opt_gn_solve_xchg is
 * b200opt_gn_solve_parts on the local buffer after a bounded wait for all `world` flags; on time-out *status_out becomes
 * B200OPT_EXCHANGE_TIMEOUT.  `out` must have n*n+n+2 floats and fmax_bits must point at its last slot. */
int64_t b200opt_xchg_bytes(int32_t world, int32_t n);
int32_t b200opt_xchg_alloc(int64_t bytes, void** ptr_out, void* ipc_handle_out);
int32_t b200opt_xchg_open(const void* ipc_handle, void** ptr_out);
int32_t b200opt_xchg_close(void* ptr);
int32_t b200opt_xchg_free(void* ptr);
int32_t b200opt_gn_gram_linear_xchg(const float* J, const float* y, const float* d, const float* r_old, float* r_out,
                                    uint32_t* fmax_bits, float rtol, float atol, int64_t m_local, int32_t n, float* out,
                                    void* workspace, const void* const* xchg_bases_dev, int32_t world, int32_t rank,
                                    int32_t parity, uint32_t seq, void* stream);
int32_t b200opt_gn_solve_xchg(const void* xchg_local, int32_t world, int32_t parity, uint32_t seq, int32_t n, double lambda,
                              float* p_out, int32_t* status_out, float* summary_out, void* workspace, void* stream);
/* ---- state-driven iteration of the large problem: the whole Gauss-Newton loop on the device ----------------------------
 * Replaces, for the "linear_rows" family, the body of the reference's while loop (optimistix/_iterate.py:194-268 around
 * AbstractGaussNewton.step, _solver/gauss_newton.py:243-334, with LearningRate(1.0) and NewtonDescent): the carry
 * (y, y_eval, num_steps, terminate, result) lives in a device `state` of b200opt_gn_state_size() bytes.
 *   b200opt_gn_init     writes the initial carry (y = y_eval = y0; nonfinite y0 => result nonfinite), zero-fills the residual
 *                       buffer the first f-space test compares against (r_old0 = the `r1` later given to _iterate) and the
 *                       counters in `workspace`.  solver: 0 = equilibrated fp32 blocked Cholesky + fp64 refinement with an
 *                       fp64 fallback when the refinement step is not small, 1 = fp64 Cholesky only.  timeout_ms bounds the
 *                       in-kernel wait for the peers (<= 0: 30 s).
 *   b200opt_gn_iterate  enqueues ONE iteration = two (three) kernels: (1) the fused residual + tcgen05 Gram pass at state.y_eval
 *                       (J read once), the cross-CTA reduction and - when xchg_bases_dev / xchg_local are given - the
 *                       publication of this rank's [G | g | r^T r | f-statistic] into every peer's exchange buffer
 *                       (fold_reduce != 0: all in one cooperative launch, the CTAs meet at a grid-wide counter;
 *                       fold_reduce == 0: a small reduction kernel follows the pass - measured equally fast, the default of the
 *                       Python layer); (2) one CTA that waits for every rank's flag, sums the blocks in rank order
 *                       in fp64, solves (G + lambda I) p = g, applies the Cauchy test on y and f (_misc.py:291-310), and
 *                       advances the carry (y <- y_eval, y_eval <- y_eval - p, num_steps + 1, result).
 *                       Once state.terminate is set, state.result != successful or num_steps == max_steps, both kernels
 *                       return at once: the host may enqueue iterations ahead and read `state` back every few calls.
 * Residual buffers r0, r1 [m_local] (16-byte aligned): step k writes r[k & 1] and compares against r[(k & 1) ^ 1].
 * `out` [n*n+n+2] fp32 receives this rank's block; `workspace` as for b200opt_gn_gram.  State layout (little-endian):
 * float y[256]; float y_eval[256]; int32 num_steps, terminate, result, max_steps; float rtol, atol, r^T r, f-statistic;
 * int32 timeout_ms, solver, fallbacks, pad. */
/* The solver of the state-driven iteration as a stand-alone call (n == 256; parts as for b200opt_gn_solve_parts, 16-byte
 * aligned, part_stride a multiple of 4).  solver 0: equilibrated fp32 blocked Cholesky in shared memory + fp64 refinement,
 * fp64 fallback; 1: fp64.  info_out (optional, device int32): 1 = the fp32 + refinement path produced the solution. */
int32_t b200opt_gn_solve_fast(const float* parts, int32_t nparts, int64_t part_stride, int32_t n, double lambda, int32_t solver,
                              float* p_out, int32_t* status_out, float* summary_out, int32_t* info_out, void* workspace,
                              void* stream);
int64_t b200opt_gn_state_size(void);
int32_t b200opt_gn_init(void* state, const float* y0, int32_t n, double rtol, double atol, int32_t max_steps, int32_t solver,
                        int32_t timeout_ms, float* r_old0, int64_t m_local, float* out, void* workspace, void* stream);
int32_t b200opt_gn_iterate(const float* J, const float* d, float* r0, float* r1, int64_t m_local, int32_t n, void* state,
                           float* out, void* workspace, const void* const* xchg_bases_dev, void* xchg_local, int32_t world,
                           int32_t rank, double rtol, double atol, double lambda, int32_t fold_reduce, void* stream);
/* r = J y - d for the "linear_rows" family (one coalesced pass over the row shard). If r_old is not NULL the
 * f-space half of cauchy_termination (optimistix/_misc.py:306-309) is fused in: *fmax_bits (device, caller-zeroed)
 * receives, as an IEEE bit pattern, max_i |r_i - r_old_i| / (atol + rtol |r_i|) over this shard. r may alias r_old. */
int32_t b200opt_linear_residual(const float* J, const float* y, const float* d, int64_t m_local, int32_t n, float* r,
                                const float* r_old, float rtol, float atol, uint32_t* fmax_bits, void* stream);

#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* B200OPT_H_ */
