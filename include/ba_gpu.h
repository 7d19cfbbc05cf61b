/* ba_gpu.h — C ABI of the B200-native Levenberg-Marquardt inner step for bundle adjustment.
 *
 * Drop-in boundary for the solver/functor concept the reference's LM loops call
 * (jasvob/BundleAdjustment_Benchmarks; there is no FFI in the reference, the boundary is the
 * compile-time `typename FunctorType::QRSolver` concept, src/Optimization/BAFunctor.h:98-119).
 * Each entry point cites the reference interface it replaces. Plain pointers and sizes only; every
 * scalar crosses the boundary as `double` and is narrowed on upload when the handle was created
 * with BA_F32 (the reference's `typedef float Scalar`, src/BATypeUtils.h:6-7).
 *
 * Conventions: every function returns 0 on success, <0 on error (BA_ERR_*), with a message in
 * ba_last_error(). A handle is single-threaded and host-blocking; device memory is owned by the
 * handle; host arrays are caller-owned. There is NO CPU fallback: without a CUDA device ba_create
 * fails with BA_ERR_CUDA.
 *
 * Parameter vector layout (BAFunctor.h:183-191,300-309): points first [3p, 3p+3), then cameras
 * 3M + 9c + {T:0-2, omega:3-5, f:6, k1:7, k2:8}. Observations MUST be sorted by (point, camera)
 * (the reference's row permutation relies on it, BacktrackLevMarqQRChol.h:291-309).
 */
#ifndef BA_GPU_H
#define BA_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ba_handle ba_handle;

enum ba_precision { BA_F32 = 0, BA_F64 = 1 };                      /* src/BATypeUtils.h:6-7 */
enum ba_variant { BA_QRKIT = 0, BA_QRCHOL = 1, BA_MOREQR = 2, BA_CHOLESKY = 3 }; /* src/CMakeLists.txt:109-160 */
enum ba_error { BA_OK = 0, BA_ERR_ARG = -1, BA_ERR_CUDA = -2, BA_ERR_NCCL = -3, BA_ERR_STATE = -4,
                BA_ERR_NUMERIC = -5 };

/* Last error message of the calling thread (never NULL). */
const char* ba_last_error(void);

/* Library / build info string ("ba_b200 <version> sm_100a ..."). */
const char* ba_version(void);

/* ≙ BAFunctor::BAFunctor (src/Optimization/BAFunctor.cpp:5-19) + initQRSolver/initQRSolverInner
 * (BAFunctor.cpp:64-78): uploads the observation structure once, sizes the per-point block solver and
 * the reduced camera system. N = cameras (global), M, K = points / observations OWNED BY THIS HANDLE
 * (all of them on one GPU; one shard when ba_comm_init is used). view[K] in [0,N), point[K] in [0,M)
 * local indices sorted by (point, view); meas = 2K doubles (x0,y0,x1,y1,...), as Matrix2XX column-major.
 * device = CUDA ordinal. */
int ba_create(ba_handle** out, int N, int M, int K, const int* view, const int* point,
              const double* meas, double inlier_threshold, int precision, int variant, int device);
int ba_destroy(ba_handle* h);

/* Multi-GPU: one handle per process per GPU, points sharded across ranks; the per-GPU reduced
 * camera contributions and scalars are summed with ncclAllReduce inside ba_compute/ba_eval/... .
 * ba_comm_unique_id fills 128 bytes on rank 0; broadcast them by any host means, then every rank
 * calls ba_comm_init. bw_blocks = global block half-bandwidth of the reduced camera matrix
 * (max over ranks of ba_bandwidth), so every rank uses the same layout. */
int ba_comm_unique_id(void* id128);
int ba_comm_init(ba_handle* h, int rank, int nranks, const void* id128);
int ba_bandwidth(ba_handle* h, int* bw_blocks);
int ba_set_bandwidth(ba_handle* h, int bw_blocks);

/* ≙ InputType upload (BAFunctor.h:39-51): R[9N] row-major 3x3 per camera, T[3N], f[N] (= K(0,0),
 * negative for BAL), k1[N], k2[N], X[3M]. */
int ba_set_state(ba_handle* h, const double* R, const double* T, const double* f, const double* k1,
                 const double* k2, const double* X);
int ba_get_state(ba_handle* h, double* R, double* T, double* f, double* k1, double* k2, double* X);

/* ≙ functor(x, fvec); fvec.squaredNorm()  (BacktrackLevMarqQRChol.h:257-261; E_pos BAFunctor.h:160-178) */
int ba_eval(ba_handle* h, double* energy);

/* ≙ functor.df(x, J); JtRes; column norms (QRChol.h:264-280; More.h:268-291; Cholesky.h:247-265).
 * The Jacobian is never materialised for QRKIT/QRCHOL/CHOLESKY (re-evaluated inside ba_compute); for
 * MOREQR this also runs stage 1 (QR of the UN-damped point blocks, More.h:288-291: R0_j, Q0, c0 kept per point /
 * observation), so that every lambda trial of ba_compute only re-triangularises the 6x3 blocks [R0_j; sqrt(lambda) I3]
 * (More.h:293-348). Points with more than 32 observations, and inputs with single-observation points, re-factor the
 * damped block per trial instead (same step up to rounding). max_colnorm2 = max_c |J(:,c)|^2,
 * max_colnorm = its square root (blueNorm rule of More.h:277); either may be NULL to skip the pass. */
int ba_linearize(ba_handle* h, double* energy, double* max_colnorm2, double* max_colnorm);

/* ≙ row permutation + [J; sqrt(lambda) I] + m_solver.compute() + right-block compute
 * (QRChol.h:291-339; More.h:299-328; Cholesky.h:274-278): per-point block QR (or 3x3 LDLT for
 * CHOLESKY), Q^T on the camera columns and residual, accumulation of the reduced camera system,
 * all-reduce across ranks, factorisation of the reduced camera block (LDLT or Householder QR). */
int ba_compute(ba_handle* h, double lambda);

/* ≙ Q^T b, right-block solve, triangular back-substitution, column un-permutation
 * (QRChol.h:322-360; More.h:331-348; Cholesky.h:281-285) fused with
 * xTest = x; increment_in_place(&xTest, dx); functor(xTest, rTest); rTest.squaredNorm()
 * (QRChol.h:363-371; update_params BAFunctor.h:299-342).
 * Outputs: |dx|_2, dx^T(lambda dx + JtRes) (the rho denominator, QRChol.h:375) and the test energy. */
int ba_solve_try(ba_handle* h, double* dx_norm, double* rho_denominator, double* energy_test);

/* Numerical health of the last ba_solve_try. The reference never checks SimplicialLDLT::info(): a singular reduced
 * system yields a non-finite test energy, which `energyTest < m_energy` (QRChol.h:374) treats as a rejected trial.
 * Same here: on a zero/NaN pivot or a non-finite step ba_solve_try still returns BA_OK but reports energy_test = NaN,
 * and *info says why: 0 = fine, r > 0 = zero or NaN pivot at (1-based) row r of the reduced camera system,
 * -1 = non-finite step. (With the two-sided / separator-split factorisations r counts rows inside the block that was being
 * eliminated - a chain, a middle block or the separator - not rows of the whole system.) After ba_set_strict_numeric(h, 1) such a
 * trial makes ba_solve_try fail with BA_ERR_NUMERIC. */
int ba_numeric_status(ba_handle* h, int* info);
int ba_set_strict_numeric(ba_handle* h, int enable);

/* One LM trial from host state to host step in ONE call, for a caller that keeps x on the host the way the reference's LM loop does
 * (x handed to functor(x) / functor.df(x) and to solver.compute / solve on every trial, QRChol.h:257-372): ba_set_state +
 * ba_linearize (energy only) + ba_compute(lambda) + ba_solve_try + ba_get_dx, with the host<->device copies pipelined against the
 * point stage: the point coordinates go up in chunks while the point-factor kernel works on the chunks that have arrived, and dx
 * comes down in chunks behind the back-substitution kernel. Results are bit-identical to the separate calls. dx: 3M+9N doubles
 * (may be NULL: no download). The host buffers should be page-locked (otherwise the copies do not overlap) and must stay valid
 * until the call returns; it blocks like ba_solve_try. Float build / MOREQR two-stage: runs the separate calls in sequence.
 * R = T = f = k1 = k2 = X = NULL: the state already on the device is used (energy + compute + solve_try with one synchronisation). */
int ba_step_streamed(ba_handle* h, const double* R, const double* T, const double* f, const double* k1, const double* k2, const double* X,
                     double lambda, double* dx, double* energy, double* dx_norm, double* rho_den, double* energy_test);

/* ≙ x = xTest (QRChol.h:428) / discarding xTest on a rejected trial. */
int ba_accept(ba_handle* h);
int ba_reject(ba_handle* h);

/* Host-side plan of the separator split of the band LDL^T (no reference counterpart: the elimination order is the solver's own
 * business, QRChol.h:197-206 only asks for the solution). Pure arithmetic, needs no GPU and no handle: for a reduced system of n
 * rows and half-bandwidth kd, mode 0 / 1 / 2 (off / only when the chains are long enough / whenever possible) and 1..4 chain
 * segments, out[0..23] = { split used, separator rows w, first separator row s0, first row of part 1, rows of part 0, rows of part 1,
 * panels per chain of part 0 / part 1, middle-block rows of part 0 / 1, middle-block panels of part 0 / 1, spike panels of part 0 / 1,
 * segments, first panel of every segment of part 0's chains (segments + 1 values), 0... }. */
int ba_split_plan(int n, int kd, int mode, int segments, int* out, int out_len);

/* ≙ Utils::showErrorStatistics + Utils::showObjective (src/Utils.h:15-68; bundle_adjustment_large.cpp:130-131,170-171) as one GPU
 * reduction at the device-resident state: sums[0] = sum_k avg_f |p_k - m_k|, sums[1] = the same over the inliers
 * (error <= inlier_threshold), sums[2] = number of inliers, sums[3] = "True objective" = sum_k psi(thr^2, avg_f^2 |p_k - m_k|)
 * (the norm, not its square, as the reference computes it). Sums over all ranks' observations. The caller divides by K / by the
 * inlier count and prints. */
int ba_error_statistics(ba_handle* h, double avg_focal_length, double inlier_threshold, double* sums);

/* Diagnostics for parity tests (not on the hot path). dx: 3M+9N doubles; residuals: 2K;
 * reduced system of the last ba_compute BEFORE factorisation: S dense symmetric (9N)^2 row-major, g 9N. */
int ba_get_dx(ba_handle* h, double* dx);
int ba_get_residuals(ba_handle* h, double* r);
int ba_get_reduced_system(ba_handle* h, double* S, double* g);
int ba_keep_reduced_system(ba_handle* h, int enable);
/* Evaluate the analytic Jacobian blocks (dE_pos, BAFunctor.h:181-297): Jc 18K (2x9 row-major per
 * observation, camera columns T,omega,f,k1,k2), Jp 6K (2x3). */
int ba_get_jacobian(ba_handle* h, double* Jc, double* Jp);

/* Counters: kernels launched by this handle since creation; device milliseconds of the last
 * ba_compute / ba_solve_try by stage (CUDA events on the handle's stream).
 * stage_ms[8]: 0 per-point factor (Jacobian + point QR + records), 1 reduced-system gather (diagonal +
 * off-diagonal blocks), 2 all-reduce, 3 factor, 4 reduced solve, 5 camera update,
 * 6 back-substitution + test energy, 7 reductions. */
int ba_launch_count(ba_handle* h, long long* launches);
int ba_stage_ms(ba_handle* h, double* stage_ms8);
int ba_set_profiling(ba_handle* h, int enable);
/* Debug: 16 device cycle counters of the last dense-stage kernel (phase split; see csrc/ba_dense.cuh). */
int ba_debug_counters(ba_handle* h, long long* out16);
int ba_debug_counters_n(ba_handle* h, long long* out, int count);  /* up to 256 (tick builds: per-CTA timelines) */
/* Test hook (not on the hot path): factor + solve S y = g for an arbitrary symmetric band matrix (S dense
 * row-major n x n, half-bandwidth kd) with this handle's reduced-camera-block solver (LDL^T or QR by variant). */
int ba_debug_band_solve(ba_handle* h, int n, int kd, const double* S, const double* g, double* y);
/* Whole-region device timing with CUDA events recorded on the handle's own stream (the stream every
 * kernel of this handle is launched on): start, run any number of calls, stop -> elapsed ms. */
int ba_timer_start(ba_handle* h);
int ba_timer_stop(ba_handle* h, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* BA_GPU_H */
