// libba_b200.so — C ABI (include/ba_gpu.h) over hand-written sm_100a kernels. No CPU fallback.
// Host orchestration of one LM trial:
//   ba_compute : init S | k_point_factor_warp (Jacobian + point block factor -> P/D/point records) | k_schur_diag +
//                k_schur_gather (reduced camera system from the records, no atomics) | all-reduce | factor
//   ba_solve_try: reduced solve | camera update | k_backsub_eval (back-substitution + update + test energy) | reduce
// Reference call sites replaced: src/Eigen_ext/BacktrackLevMarqQRChol.h:257-371 (and the MOREQR /
// CHOLESKY counterparts), src/Optimization/BAFunctor.{h,cpp}.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ba_gpu.h"
#include "ba_dense.cuh"
#include "ba_ldlt2.cuh"
#include "ba_split.cuh"
#include "ba_qr.cuh"
#include "ba_tile.cuh"

namespace {

thread_local std::string g_err = "";
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}
#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) return fail(BA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------ NCCL
// libnccl.so.2 is dlopen'ed and nccl.h is not needed to build: the handful of ABI-stable types and enum values
// used here (nccl.h of NCCL 2.x: ncclUniqueId = 128 opaque bytes; ncclSuccess = 0; ncclSum = 0, ncclMax = 2;
// ncclFloat = 7, ncclDouble = 8) are declared locally.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
typedef int ncclRedOp_t;
constexpr ncclResult_t ncclSuccess = 0;
constexpr ncclRedOp_t ncclSum = 0, ncclMax = 2, ncclMin = 3;
constexpr ncclDataType_t ncclInt = 2, ncclFloat = 7, ncclDouble = 8;
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy && GetErrorString;
  }
};
NcclApi g_nccl;
#define NK(call)                                                                                     \
  do {                                                                                               \
    ncclResult_t r_ = (call);                                                                        \
    if (r_ != ncclSuccess) return fail(BA_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
  } while (0)

// ------------------------------------------------------------------------------- small kernels
using namespace ba;

// mode 0: energy partials; 1: + residuals out; 2: Jacobian blocks out (column norms: k_colnorm_pt / k_colnorm_cam below)
template <class T, int MODE>
__global__ void __launch_bounds__(256) k_obs(int K, const int* __restrict__ view, const int* __restrict__ point, const T* __restrict__ meas,
                                             const T* __restrict__ cams, const T* __restrict__ X, T tau2, int M,
                                             double* __restrict__ partials, T* __restrict__ out0, T* __restrict__ out1) {
  __shared__ double sred[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (i < K) {
    const int cidx = __ldg(view + i), pj = __ldg(point + i);
    Cam<T> c; load_cam<T>(cams, cidx, c);
    const T X0 = __ldg(X + 3 * (size_t)pj), X1 = __ldg(X + 3 * (size_t)pj + 1), X2 = __ldg(X + 3 * (size_t)pj + 2);
    const T m0 = __ldg(meas + 2 * (size_t)i), m1 = __ldg(meas + 2 * (size_t)i + 1);
    T e0, e1;
    if (MODE <= 1) {
      obs_residual<T>(c, X0, X1, X2, m0, m1, tau2, e0, e1);
      if (MODE == 1) { out0[2 * (size_t)i] = e0; out0[2 * (size_t)i + 1] = e1; }
    } else {
      T jc[18], jp[6];
      obs_jacobian<T>(c, X0, X1, X2, m0, m1, tau2, e0, e1, jc, jp);
      if (MODE == 2) {
#pragma unroll
        for (int k = 0; k < 18; ++k) out0[18 * (size_t)i + k] = jc[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) out1[6 * (size_t)i + k] = jp[k];
      }
    }
    acc = (double)(e0 * e0 + e1 * e1);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sred[w];
    partials[blockIdx.x] = s;
  }
}

// Squared column norms of J (QRChol.h:267-280), without atomics so that lambda_0 - and with it a whole LM run - is
// bit-reproducible: one thread per point walks the point's (contiguous) observations, one CTA per camera walks the
// camera's camera-major slots; fixed summation order in both.
template <class T>
__global__ void __launch_bounds__(256) k_colnorm_pt(int M, const int* __restrict__ pt_start, const int* __restrict__ view, const T* __restrict__ meas,
                                                    const T* __restrict__ cams, const T* __restrict__ X, T tau2, T* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const T X0 = __ldg(X + 3 * (size_t)j), X1 = __ldg(X + 3 * (size_t)j + 1), X2 = __ldg(X + 3 * (size_t)j + 2);
  T s0 = T(0), s1 = T(0), s2 = T(0);
  for (int i = __ldg(pt_start + j); i < __ldg(pt_start + j + 1); ++i) {
    Cam<T> c; load_cam<T>(cams, __ldg(view + i), c);
    T e0, e1, jc[18], jp[6];
    obs_jacobian<T>(c, X0, X1, X2, __ldg(meas + 2 * (size_t)i), __ldg(meas + 2 * (size_t)i + 1), tau2, e0, e1, jc, jp);
    s0 += jp[0] * jp[0] + jp[3] * jp[3]; s1 += jp[1] * jp[1] + jp[4] * jp[4]; s2 += jp[2] * jp[2] + jp[5] * jp[5];
  }
  out[3 * (size_t)j] = s0; out[3 * (size_t)j + 1] = s1; out[3 * (size_t)j + 2] = s2;
}

template <class T>
__global__ void __launch_bounds__(128) k_colnorm_cam(const int* __restrict__ cam_start, const int* __restrict__ obs_of_slot, const int* __restrict__ point,
                                                     const T* __restrict__ meas, const T* __restrict__ cams, const T* __restrict__ X, T tau2,
                                                     T* __restrict__ out /* 9 per camera */) {
  __shared__ T part[4][9];
  const int cam = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Cam<T> c; load_cam<T>(cams, cam, c);
  T s[9];
#pragma unroll
  for (int b = 0; b < 9; ++b) s[b] = T(0);
  for (int q = __ldg(cam_start + cam) + threadIdx.x; q < __ldg(cam_start + cam + 1); q += 128) {
    const int i = __ldg(obs_of_slot + q), pj = __ldg(point + i);
    T e0, e1, jc[18], jp[6];
    obs_jacobian<T>(c, __ldg(X + 3 * (size_t)pj), __ldg(X + 3 * (size_t)pj + 1), __ldg(X + 3 * (size_t)pj + 2), __ldg(meas + 2 * (size_t)i),
                    __ldg(meas + 2 * (size_t)i + 1), tau2, e0, e1, jc, jp);
#pragma unroll
    for (int b = 0; b < 9; ++b) s[b] += jc[b] * jc[b] + jc[9 + b] * jc[9 + b];
  }
#pragma unroll
  for (int b = 0; b < 9; ++b) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s[b] += __shfl_down_sync(0xffffffffu, s[b], off);
    if (lane == 0) part[warp][b] = s[b];
  }
  __syncthreads();
  if (threadIdx.x < 9) out[9 * (size_t)cam + threadIdx.x] = ((part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x]) + part[3][threadIdx.x];
}

// deterministic final reduction: block b sums partials[b*count .. (b+1)*count) -> out[b]
__global__ void __launch_bounds__(1024) k_reduce(const double* __restrict__ partials, int count, double* __restrict__ out) {
  __shared__ double s[1024];
  const double* p = partials + (size_t)blockIdx.x * count;
  // four independent running sums per thread keep four loads in flight (fixed order: deterministic)
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const int bd = blockDim.x;
  int i = threadIdx.x;
  for (; i + 3 * bd < count; i += 4 * bd) { a0 += p[i]; a1 += p[i + bd]; a2 += p[i + 2 * bd]; a3 += p[i + 3 * bd]; }
  for (; i < count; i += bd) a0 += p[i];
  const double acc = (a0 + a1) + (a2 + a3);
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) { if ((int)threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off]; __syncthreads(); }
  if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

// Reporting (Utils::showErrorStatistics / showObjective, src/Utils.h:15-68) as a reduction over the observations at the
// device-resident state: per-block partial sums of the reprojection error avg_f |p - m|, of the inliers' errors, of the
// inlier count and of the "true objective" psi(tau^2, avg_f^2 |p - m|) (the norm, not its square: quirk Q7 of the survey).
template <class T>
__global__ void __launch_bounds__(256) k_stats(int K, const int* __restrict__ view, const int* __restrict__ point, const T* __restrict__ meas,
                                               const T* __restrict__ cams, const T* __restrict__ X, T avg_f, T thr, double* __restrict__ partials, int nb) {
  __shared__ double sred[4][8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  if (i < K) {
    const int cidx = __ldg(view + i), pj = __ldg(point + i);
    Cam<T> c; load_cam<T>(cams, cidx, c);
    const T X0 = __ldg(X + 3 * (size_t)pj), X1 = __ldg(X + 3 * (size_t)pj + 1), X2 = __ldg(X + 3 * (size_t)pj + 2);
    const T xx = c.R[0] * X0 + c.R[1] * X1 + c.R[2] * X2 + c.t[0];
    const T yy = c.R[3] * X0 + c.R[4] * X1 + c.R[5] * X2 + c.t[1];
    const T zz = c.R[6] * X0 + c.R[7] * X1 + c.R[8] * X2 + c.t[2];
    const T xu0 = xx / zz, xu1 = yy / zz, r2u = xu0 * xu0 + xu1 * xu1, kr = T(1) + c.k1 * r2u + c.k2 * r2u * r2u;
    const T r0 = c.f * (kr * xu0) - __ldg(meas + 2 * (size_t)i), r1 = c.f * (kr * xu1) - __ldg(meas + 2 * (size_t)i + 1);
    const T en = tsqrt(r0 * r0 + r1 * r1), e = avg_f * en;
    v[0] = (double)e;
    if (e <= thr) { v[1] = (double)e; v[2] = 1.0; }
    const T tau2 = thr * thr, q2 = avg_f * avg_f * en, q4 = q2 * q2;
    v[3] = (double)((q2 < tau2) ? q2 * (T(3.0) - T(3.0) * q2 / tau2 + q4 / (tau2 * tau2)) / T(6.0) : tau2 / T(6.0));
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], off);
    if ((threadIdx.x & 31) == 0) sred[q][threadIdx.x >> 5] = v[q];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sred[threadIdx.x][w];
    partials[(size_t)threadIdx.x * nb + blockIdx.x] = t;
  }
}

template <class T>
__global__ void __launch_bounds__(256) k_max(const T* __restrict__ v, int n, double* __restrict__ out) {
  __shared__ double s[256];
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) m = fmax(m, (double)v[i]);
  s[threadIdx.x] = m;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) { if ((int)threadIdx.x < off) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + off]); __syncthreads(); }
  if (threadIdx.x == 0) out[0] = s[0];
}

// cams_test = update(cams, dx_cam) (update_params, BAFunctor.h:311-333), |dx_cam|^2 and dx_cam . gJ. One camera per thread,
// 128 per CTA; the CTA that arrives last sums the per-CTA partial sums in index order (deterministic).
constexpr int CAMUP_THREADS = 128;
template <class T>
__global__ void __launch_bounds__(CAMUP_THREADS) k_cam_update(int N, const T* __restrict__ cams, const T* __restrict__ dx_cam, const T* __restrict__ gJ,
                                                              T* __restrict__ cams_test, double* __restrict__ out_norm2, double* __restrict__ out_dot,
                                                              double* __restrict__ part, unsigned int* __restrict__ counter) {
  __shared__ double s[CAMUP_THREADS];
  __shared__ double s2[CAMUP_THREADS];
  __shared__ bool last;
  double acc = 0.0, acc2 = 0.0;
  const int c = blockIdx.x * CAMUP_THREADS + threadIdx.x;
  if (c < N) {
    const T* ci = cams + (size_t)c * CAM_STRIDE;
    T* co = cams_test + (size_t)c * CAM_STRIDE;
    T d[9];
#pragma unroll
    for (int b = 0; b < 9; ++b) { d[b] = dx_cam[9 * (size_t)c + b]; acc += (double)(d[b] * d[b]); acc2 += (double)(d[b] * gJ[9 * (size_t)c + b]); }
    T Rin[9], Rout[9];
#pragma unroll
    for (int b = 0; b < 9; ++b) Rin[b] = ci[b];
    rodrigues_left<T>(d[3], d[4], d[5], Rin, Rout);
#pragma unroll
    for (int b = 0; b < 9; ++b) co[b] = Rout[b];
    co[9] = ci[9] + d[0]; co[10] = ci[10] + d[1]; co[11] = ci[11] + d[2];
    co[12] = ci[12] + d[6]; co[13] = ci[13] + d[7]; co[14] = ci[14] + d[8]; co[15] = T(0);
  }
  s[threadIdx.x] = acc; s2[threadIdx.x] = acc2;
  __syncthreads();
  for (int off = CAMUP_THREADS / 2; off > 0; off >>= 1) { if ((int)threadIdx.x < off) { s[threadIdx.x] += s[threadIdx.x + off]; s2[threadIdx.x] += s2[threadIdx.x + off]; } __syncthreads(); }
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = s[0]; part[2 * blockIdx.x + 1] = s2[0];
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double a = 0.0, a2 = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) { a += __ldcg(part + 2 * b); a2 += __ldcg(part + 2 * b + 1); }
    out_norm2[0] = a; out_dot[0] = a2;
    *counter = 0u;
  }
}

template <class T> struct DevBuf {
  T* p = nullptr; size_t n = 0;
  cudaError_t alloc(size_t count) { free(); n = count; return cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)); }
  void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DevBuf() { free(); }
};

}  // namespace

// =============================================================================================
struct ba_handle {
  virtual ~ba_handle() {}
  virtual int set_state(const double*, const double*, const double*, const double*, const double*, const double*) = 0;
  virtual int get_state(double*, double*, double*, double*, double*, double*) = 0;
  virtual int step_streamed(const double*, const double*, const double*, const double*, const double*, const double*, double, double*, double*, double*, double*, double*) = 0;
  virtual int eval(double*) = 0;
  virtual int linearize(double*, double*, double*) = 0;
  virtual int compute(double) = 0;
  virtual int solve_try(double*, double*, double*) = 0;
  virtual int accept() = 0;
  virtual int reject() = 0;
  virtual int get_dx(double*) = 0;
  virtual int get_residuals(double*) = 0;
  virtual int error_statistics(double, double, double*) = 0;
  virtual int get_reduced(double*, double*) = 0;
  virtual int get_jacobian(double*, double*) = 0;
  virtual int comm_init(int, int, const void*) = 0;
  virtual int set_bandwidth(int) = 0;
  virtual int timer_start() = 0;
  virtual int timer_stop(double*) = 0;
  virtual int debug_counters(long long*, int) = 0;
  virtual int debug_band_solve(int, int, const double*, const double*, double*) = 0;
  int bw = 0;
  int numeric_info = 0;        // last ba_solve_try: 0 ok, > 0 zero/NaN pivot at that (1-based) row of the reduced system, -1 non-finite step
  bool strict_numeric = false; // ba_set_strict_numeric: return BA_ERR_NUMERIC instead of reporting a NaN test energy
  bool keep_reduced = false;
  bool force_grid_ldlt = false;
  bool profiling = false;
  long long launches = 0;
  double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

template <class T>
struct Impl : ba_handle {
  int N = 0, M = 0, K = 0, device = 0, variant = 0, ntiles = 0;
  int n = 0, kd = 0, ldsv = 0;  // ldsv: row stride of the band view, kd padded to 16 bytes
  T tau2 = T(0.25);
  double lambda = 0.0;
  bool computed = false, tried = false, linearized = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[9];
  cudaEvent_t tev[2] = {nullptr, nullptr};
  DevBuf<double> d_camup; DevBuf<unsigned int> d_camup_cnt;   // k_cam_update: per-CTA partial sums, arrival counter
  int camup_blocks() const { return (N + CAMUP_THREADS - 1) / CAMUP_THREADS; }
  cudaStream_t stream2 = nullptr;   // separator split: the spike kernel runs beside the next chain segment / the middle blocks
  cudaEvent_t sev[10] = {};
  cudaEvent_t lev[5] = {};          // launch completion events of the chain segments / middle blocks
  int split_segments = 3;
  bool tl_deferred = false; int tl_calls = 0;
  bool split_timeline = false; cudaEvent_t tlv[24] = {}; int tl_n = 0; const char* tl_name[24] = {};
  bool tl_s2only = false;
  void tl(const char* name, cudaStream_t st) { if (tl_s2only && st == stream && std::strncmp(name, "split", 5) != 0) return; if (split_timeline && tl_n < 24) { if (!tlv[tl_n]) cudaEventCreate(&tlv[tl_n]); cudaEventRecord(tlv[tl_n], st); tl_name[tl_n++] = name; } }
  // ba_step_streamed: the point coordinates are uploaded in chunks on stream2 while the point-factor kernel already works on
  // the chunks that have arrived, and dx is downloaded in chunks behind the back-substitution kernel
  static constexpr int SX_MAX = 16;
  int sx_chunks = 8, sx_n = 0, sx_unit[SX_MAX + 1] = {};
  bool sx_active = false;
  double* dx_sink = nullptr;
  cudaEvent_t xev[SX_MAX + 2] = {}, bev[SX_MAX + 2] = {};
  std::vector<int> h_unit_last_pt, h_tile_first;
  DevBuf<int> d_view, d_point, d_pt_start, d_tile_pt, d_info;
  DevBuf<int> d_slot, d_cam_start, d_blk_a, d_blk_b, d_blk_start, d_counter;  // static structure of the deterministic Schur gather
  DevBuf<int2> d_pairs;
  DevBuf<int> d_seg, d_obs_of_slot, d_unit_pt, d_big_pt, d_huge_pt;  // warp units (points with <= 32 observations) / tiles of the larger points / points with > TILE observations
  int nunits = 0, nbig = 0, nhuge = 0, huge_max = 0;
  DevBuf<T> d_P, d_D, d_Pt;  // per-observation (P, D) / per-point records written by k_point_factor*
  DevBuf<T> d_J, d_Pt0;      // MOREQR stage 1 (un-damped point QR, once per outer iteration): J records, point records
  bool two_stage = false, stage1_valid = false;
  // QRKIT / MOREQR right block: LDL^T of S + corrected semi-normal refinement through J2bot (k_csne_*), or (BA_QR_HOUSEHOLDER=1)
  // the Householder QR of the square S of round 1
  bool qr_csne = false; int csne_steps = 1, nlong = 0;
  DevBuf<T> d_u, d_delta;
  DevBuf<int> d_long_pt;
  int nblocks = 0, gather_grid = 0;
  DevBuf<T> d_meas, d_cams, d_cams_test, d_X, d_X_test, d_dx_pt, d_dx_cam, d_red /* S band | g | gJ */, d_keep, d_dvec, d_tmp;
  DevBuf<T> d_qr;  // general band copy for the Householder QR of S
  DevBuf<T> d_W;   // L_kk^-1 tiles kept by the cluster LDLT for the backward pass
  DevBuf<T> d_rev, d_dvec2, d_W2, d_y2, d_WM;  // two-sided factorisation: reversed bottom system [S' | g'], its D, W, y; middle W
  bool two_sided = true;
  int cluster_size = 16;  // non-portable size; falls back to 8 when 16 CTAs of this footprint cannot be co-scheduled
  bool solved_in_factor = false;
  int ldlt_roww = 6;  // row-tile warps per chain CTA of the cluster LDLT (BA_LDLT_ROWW=3|6)
  int ldlt_split = 1;       // separator split, four elimination chains (ba_split.cuh): BA_LDLT_SPLIT=0 off, 1 when the chains are long enough to pay for it, 2 whenever possible
  struct SplitPart {        // one part of the separator split: index-reversed half, factors, spike
    DevBuf<T> full, rev, dvec, dvec2, W, W2, y, y2, E;
  } sp[2];
  DevBuf<T> d_sep, d_sep_vec, d_sep_W;  // separator block (dense), [g | D | y], W
  int sep_w = -1;
  bool ldlt_v2 = false; // BA_LDLT_V2=1: forward elimination of the two-sided scheme by the owner-computes kernel (ba_ldlt2.cuh; correct, not yet faster)
  DevBuf<double> d_partials, d_scal;
  DevBuf<long long> d_dbg;
  double* h_scal = nullptr;  // pinned
  size_t red_count = 0;      // elements of S band storage (+ g behind it)
  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int coop_grid = 0, coop_grid_qr = 0, sm_count = 148;

  ~Impl() override {
    cudaSetDevice(device);
    if (comm) g_nccl.CommDestroy(comm);
    if (h_scal) cudaFreeHost(h_scal);
    if (h_stage) cudaFreeHost(h_stage);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : tev) if (e) cudaEventDestroy(e);
    for (auto& e : sev) if (e) cudaEventDestroy(e);
    for (auto& e : lev) if (e) cudaEventDestroy(e);
    for (auto& e : xev) if (e) cudaEventDestroy(e);
    for (auto& e : bev) if (e) cudaEventDestroy(e);
    if (stream2) cudaStreamDestroy(stream2);
    if (stream) cudaStreamDestroy(stream);
  }

  T* Sraw() { return d_red.p; }
  T* Sv() { return d_red.p + ldsv; }
  size_t lds() const { return (size_t)ldsv; }
  static int pad_lds(int kd_) { const int a = 16 / (int)sizeof(T); return (kd_ + a - 1) / a * a; }
  T* gvec() { return d_red.p + red_count; }
  T* gJvec() { return d_red.p + red_count + n; }  // sum_i Jc_i^T e_i per camera (rho denominator)
  BandMat<T> band() { return BandMat<T>{Sv(), lds(), n, kd}; }
  static constexpr ncclDataType_t nccl_t() { return sizeof(T) == 8 ? ncclDouble : ncclFloat; }

  int alloc_reduced() {
    kd = std::min(9 * N - 1, 9 * bw + 8);
    if (kd < 1) kd = 1;
    n = 9 * N;
    ldsv = pad_lds(kd);
    red_count = (size_t)n * (ldsv + 1);
    CK(d_red.alloc(red_count + 2 * (size_t)n));
    CK(d_dvec.alloc((size_t)n + NB));  // D is read/written in whole 32-wide panels
    if (keep_reduced) CK(d_keep.alloc(red_count + n));
    if (qr_csne) CK(d_delta.alloc(n));
    d_qr.free();
    return BA_OK;
  }

  int init(int N_, int M_, int K_, const int* view, const int* point, const double* meas, double tau, int variant_, int device_) {
    N = N_; M = M_; K = K_; variant = variant_; device = device_;
    tau2 = T(tau) * T(tau);
    for (auto& e : ev) e = nullptr;
    if (N <= 0 || M <= 0 || K <= 0) return fail(BA_ERR_ARG, "N, M, K must be positive");
    // host-side structure: sortedness, ranges, CSR offsets, tiles, bandwidth
    std::vector<int> pt_start(M + 1, 0);
    for (int i = 0; i < K; ++i) {
      if (view[i] < 0 || view[i] >= N || point[i] < 0 || point[i] >= M) return fail(BA_ERR_ARG, "observation %d: index out of range", i);
      if (i > 0 && (point[i] < point[i - 1] || (point[i] == point[i - 1] && view[i] <= view[i - 1])))
        return fail(BA_ERR_ARG, "observations must be sorted by (point, camera) without duplicates (observation %d)", i);
      pt_start[point[i] + 1]++;
    }
    for (int j = 0; j < M; ++j) {
      if (pt_start[j + 1] < 1) return fail(BA_ERR_ARG, "point %d has no observation", j);
      if (pt_start[j + 1] > BIG_POINT_MAX) return fail(BA_ERR_ARG, "point %d has %d observations; this build supports at most %d per point", j, pt_start[j + 1], BIG_POINT_MAX);
      pt_start[j + 1] += pt_start[j];
    }
    bw = 0;
    for (int j = 0; j < M; ++j) bw = std::max(bw, view[pt_start[j + 1] - 1] - view[pt_start[j]]);
    // back-substitution tiles: (first point, end point, first observation, observations) of consecutive points whose
    // observations fit TILE lanes; a point
    // with more than TILE observations ("huge": long tracks of real BAL files) gets its own CTA in separate kernels
    std::vector<int> tile_pt, huge_pt;
    {
      int first = 0, cur_obs = 0, cur_pts = 0;
      auto close = [&](int end) {
        if (end > first) { tile_pt.push_back(first); tile_pt.push_back(end); tile_pt.push_back(pt_start[first]); tile_pt.push_back(pt_start[end] - pt_start[first]); }
      };
      for (int j = 0; j < M; ++j) {
        const int nj = pt_start[j + 1] - pt_start[j];
        if (nj > TILE) { close(j); huge_pt.push_back(j); first = j + 1; cur_obs = 0; cur_pts = 0; continue; }
        if (cur_obs + nj > TILE || cur_pts + 1 > TILE) { close(j); first = j; cur_obs = 0; cur_pts = 0; }
        cur_obs += nj; cur_pts++;
      }
      close(M);
    }
    nhuge = (int)huge_pt.size();
    for (int j : huge_pt) huge_max = std::max(huge_max, pt_start[j + 1] - pt_start[j]);
    ntiles = (int)tile_pt.size() / 4;

    // point factor work units: one warp per run of consecutive points whose observations fit 32 lanes
    // (k_point_factor_warp); a point with more than 32 observations becomes its own shared-memory tile
    std::vector<int> unit_pt, big_pt;
    for (int j = 0; j < M;) {
      const int nj = pt_start[j + 1] - pt_start[j];
      if (nj > TILE) { ++j; continue; }  // huge point: k_point_factor_big
      if (nj > 32) { big_pt.push_back(j); big_pt.push_back(j + 1); ++j; continue; }
      int j1 = j, cnt = 0;
      while (j1 < M && pt_start[j1 + 1] - pt_start[j1] <= 32 && cnt + (pt_start[j1 + 1] - pt_start[j1]) <= 32) { cnt += pt_start[j1 + 1] - pt_start[j1]; ++j1; }
      unit_pt.push_back(pt_start[j]); unit_pt.push_back(cnt);  // (first observation, observation count) of the unit
      j = j1;
    }
    nunits = (int)unit_pt.size() / 2; nbig = (int)big_pt.size() / 2;
    h_unit_last_pt.resize(nunits);
    for (int u = 0; u < nunits; ++u) h_unit_last_pt[u] = point[unit_pt[2 * u] + unit_pt[2 * u + 1] - 1];
    h_tile_first.resize(ntiles);
    for (int t = 0; t < ntiles; ++t) h_tile_first[t] = tile_pt[4 * (size_t)t];
    std::vector<int> long_pt;
    for (int j = 0; j < M; ++j) if (pt_start[j + 1] - pt_start[j] > 32) long_pt.push_back(j);
    nlong = (int)long_pt.size();
    // double only: the refinement needs cond(S) eps < 1, which never holds in float on BA problems (measured: it makes the
    // float step worse); the float build keeps the Householder QR of S
    qr_csne = (variant == BA_QRKIT || variant == BA_MOREQR) && sizeof(T) == 8 && !std::getenv("BA_QR_HOUSEHOLDER");
    if (const char* rs = std::getenv("BA_QR_REFINE")) csne_steps = std::max(0, std::min(4, atoi(rs)));
    // MOREQR two-stage scheme (More.h:288-348) for the points handled by the warp kernel; a point with a single
    // observation has a rank-2 un-damped block (2 x 3), which the segmented Householder does not cover: such inputs
    // (none in the BAL files: every point has >= 2 observations) fall back to re-factoring the damped blocks per trial
    {
      bool single = false;
      for (int j = 0; j < M && !single; ++j) single = (pt_start[j + 1] - pt_start[j]) == 1;
      const char* ts = std::getenv("BA_MOREQR_TWOSTAGE");
      two_stage = (variant == BA_MOREQR) && !single && nunits > 0 && !(ts && atoi(ts) == 0);
    }
    // static structure of the Schur gather: camera-major record slots, non-empty camera-pair blocks, pair lists
    std::vector<int> cam_start(N + 1, 0), slot(K);
    for (int i = 0; i < K; ++i) cam_start[view[i] + 1]++;
    for (int c = 0; c < N; ++c) cam_start[c + 1] += cam_start[c];
    {
      std::vector<int> fill(cam_start.begin(), cam_start.end() - 1);
      for (int i = 0; i < K; ++i) slot[i] = fill[view[i]]++;
    }
    const size_t Wd = (size_t)bw + 1;
    if ((size_t)N * Wd > (size_t)1 << 30) return fail(BA_ERR_ARG, "camera-pair table too large (N = %d, bandwidth = %d)", N, bw);
    std::vector<int> cnt((size_t)N * Wd, 0);
    for (int j = 0; j < M; ++j)
      for (int ia = pt_start[j]; ia < pt_start[j + 1]; ++ia)
        for (int ib = pt_start[j]; ib < ia; ++ib) cnt[(size_t)view[ia] * Wd + (view[ia] - view[ib])]++;
    std::vector<int> blk_a, blk_b, blk_start, pos((size_t)N * Wd, -1);
    long long npairs_ll = 0;
    for (int a2 = 0; a2 < N; ++a2)
      for (int dl = 1; dl <= bw && dl <= a2; ++dl) {  // diagonal blocks are streamed by k_schur_diag
        const size_t key = (size_t)a2 * Wd + dl;
        if (cnt[key] > 0) {
          blk_a.push_back(a2); blk_b.push_back(a2 - dl); blk_start.push_back((int)npairs_ll);
          pos[key] = (int)npairs_ll; npairs_ll += cnt[key];
          if (npairs_ll > 2000000000LL) return fail(BA_ERR_ARG, "too many camera-pair contributions for 32-bit offsets");
        }
      }
    blk_start.push_back((int)npairs_ll);
    nblocks = (int)blk_a.size();
    std::vector<int2> pairs((size_t)std::max<long long>(npairs_ll, 1));
    for (int j = 0; j < M; ++j)
      for (int ia = pt_start[j]; ia < pt_start[j + 1]; ++ia)
        for (int ib = pt_start[j]; ib < ia; ++ib) {
          const size_t key = (size_t)view[ia] * Wd + (view[ia] - view[ib]);
          pairs[pos[key]++] = make_int2(ia, ib);  // P records live at the observation index
        }
    { std::vector<int>().swap(cnt); std::vector<int>().swap(pos); }
    // blocks are handed out in (a, b) order: concurrently processed blocks share cameras, so the P records of the
    // ~bw cameras in flight stay in L2 (largest-first order balanced slightly better but read 8 GB from HBM
    // instead of the 1.4 GB of records: profiles/r01_ncu_tiles_v2_summary.csv)

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BA_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    CK(cudaSetDevice(device));
    // When a chain segment (16-CTA clusters, each needs 16 free SMs in one GPC) and the spike kernel of the previous segment
    // (70 plain CTAs on stream2) become ready at the same moment, the clusters must be placed first: otherwise the spike's
    // CTAs scatter over the GPCs and the chains wait for them (measured: +0.2 ms per segment). The spike therefore also waits
    // for the LAUNCH COMPLETION event of the next chain launch (factor_reduced_split), and stream2 has the lower priority
    // (a higher-priority dependent may start before the launch completes).
    int prio_least = 0, prio_greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    CK(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_greatest));
    for (auto& e : ev) CK(cudaEventCreate(&e));
    CK(cudaStreamCreateWithPriority(&stream2, cudaStreamNonBlocking, prio_least));
    for (auto& e : sev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : lev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : xev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : bev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (const char* sc = std::getenv("BA_STREAM_CHUNKS")) sx_chunks = std::max(1, std::min(SX_MAX, atoi(sc)));
    CK(d_view.alloc(K)); CK(d_point.alloc(K)); CK(d_pt_start.alloc(M + 1)); CK(d_tile_pt.alloc(4 * (size_t)ntiles + 4)); CK(d_huge_pt.alloc(nhuge + 1)); CK(d_info.alloc(1));
    CK(d_meas.alloc(2 * (size_t)K));
    CK(d_cams.alloc((size_t)N * CAM_STRIDE)); CK(d_cams_test.alloc((size_t)N * CAM_STRIDE));
    CK(d_X.alloc(3 * (size_t)M)); CK(d_X_test.alloc(3 * (size_t)M));
    CK(d_dx_pt.alloc(3 * (size_t)M)); CK(d_dx_cam.alloc(9 * (size_t)N));
    const size_t npart = std::max<size_t>(3 * (size_t)(ntiles + nhuge), (size_t)(K + 255) / 256);
    CK(d_camup.alloc(2 * (size_t)camup_blocks())); CK(d_camup_cnt.alloc(1)); CK(cudaMemset(d_camup_cnt.p, 0, sizeof(unsigned int)));
    CK(d_partials.alloc(npart)); CK(d_scal.alloc(16)); CK(d_dbg.alloc(256)); CK(cudaMemset(d_dbg.p, 0, 256 * sizeof(long long)));
    CK(cudaMallocHost(&h_scal, 16 * sizeof(double)));
    CK(cudaMemcpyAsync(d_view.p, view, K * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_point.p, point, K * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_pt_start.p, pt_start.data(), (M + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (ntiles) CK(cudaMemcpyAsync(d_tile_pt.p, tile_pt.data(), tile_pt.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (nhuge) {
      CK(cudaMemcpyAsync(d_huge_pt.p, huge_pt.data(), nhuge * sizeof(int), cudaMemcpyHostToDevice, stream));
      CK(cudaFuncSetAttribute(k_point_factor_big<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_point_smem_bytes<T>(huge_max)));
    }
    CK(d_unit_pt.alloc(unit_pt.size())); CK(d_big_pt.alloc(big_pt.size()));
    if (nunits) CK(cudaMemcpyAsync(d_unit_pt.p, unit_pt.data(), unit_pt.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (nbig) CK(cudaMemcpyAsync(d_big_pt.p, big_pt.data(), big_pt.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(d_slot.alloc(K)); CK(d_cam_start.alloc(N + 1)); CK(d_blk_a.alloc(nblocks)); CK(d_blk_b.alloc(nblocks)); CK(d_blk_start.alloc(nblocks + 1));
    CK(d_counter.alloc(1));
    CK(d_pairs.alloc(pairs.size()));
    CK(d_P.alloc((size_t)K * REC)); CK(d_D.alloc((size_t)K * REC)); CK(d_Pt.alloc((size_t)M * PREC));
    if (two_stage) { CK(d_J.alloc((size_t)K * REC)); CK(d_Pt0.alloc((size_t)M * PREC)); }
    if (qr_csne) {
      CK(d_u.alloc(2 * (size_t)K)); CK(d_long_pt.alloc(long_pt.size() + 1));
      if (nlong) CK(cudaMemcpyAsync(d_long_pt.p, long_pt.data(), nlong * sizeof(int), cudaMemcpyHostToDevice, stream));
      CK(cudaStreamSynchronize(stream));
    }
    CK(cudaMemcpyAsync(d_slot.p, slot.data(), K * sizeof(int), cudaMemcpyHostToDevice, stream));
    {
      std::vector<int> inv(K);   // observation stored at each camera-major slot (column norms of the camera columns)
      for (int i = 0; i < K; ++i) inv[slot[i]] = i;
      CK(d_obs_of_slot.alloc(K));
      CK(cudaMemcpy(d_obs_of_slot.p, inv.data(), K * sizeof(int), cudaMemcpyHostToDevice));
    }
    {
      std::vector<int> seg(K);   // per observation: index within its point | min(count, 255) << 8 (warp units: count <= 32)
      for (int j = 0; j < M; ++j) {
        const int nj = pt_start[j + 1] - pt_start[j];
        for (int q = 0; q < nj; ++q) seg[pt_start[j] + q] = std::min(q, 255) | (std::min(nj, 255) << 8);
      }
      CK(d_seg.alloc(K));
      CK(cudaMemcpy(d_seg.p, seg.data(), K * sizeof(int), cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpyAsync(d_cam_start.p, cam_start.data(), (N + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_blk_a.p, blk_a.data(), nblocks * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_blk_b.p, blk_b.data(), nblocks * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_blk_start.p, blk_start.data(), (nblocks + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_pairs.p, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice, stream));
    std::vector<T> m(2 * (size_t)K);
    for (size_t i = 0; i < m.size(); ++i) m[i] = (T)meas[i];
    CK(cudaMemcpyAsync(d_meas.p, m.data(), m.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
    CK(cudaMemsetAsync(d_dx_pt.p, 0, 3 * (size_t)M * sizeof(T), stream));
    CK(cudaMemsetAsync(d_dx_cam.p, 0, 9 * (size_t)N * sizeof(T), stream));
    CK(cudaStreamSynchronize(stream));
    int rc = alloc_reduced();
    if (rc) return rc;
    CK(cudaFuncSetAttribute(k_point_factor<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem<T>)));
    int occ = 0, sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    sm_count = sms;
    CK(cudaFuncSetAttribute(k_schur_gather<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gather_smem_bytes<T>()));
    CK(cudaFuncSetAttribute(k_schur_diag<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem_bytes<T>()));
    CK(cudaFuncSetAttribute(k_backsub_eval<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)backsub_smem_bytes<T>()));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_schur_gather<T>, GATHER_THREADS, gather_smem_bytes<T>()));
    gather_grid = std::max(1, std::min(std::max(occ, 1) * sms, (nblocks + GATHER_THREADS / 32 - 1) / (GATHER_THREADS / 32)));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_band_ldlt<T>, DENSE_THREADS, 0));
    coop_grid = std::max(1, std::min(occ, 2)) * sms;
    if (const char* cs = std::getenv("BA_CLUSTER_SIZE")) cluster_size = std::max(1, std::min(16, atoi(cs)));
    if (std::getenv("BA_FORCE_GRID_LDLT")) force_grid_ldlt = true;
    if (const char* rw = std::getenv("BA_LDLT_ROWW")) ldlt_roww = atoi(rw) == 3 ? 3 : 6;
    if (const char* wa = std::getenv("BA_LDLT_W_AFTER")) { if (atoi(wa) != 0) ldlt_roww |= 0x100; }
    if (const char* ra = std::getenv("BA_LDLT_ROWS_AFTER")) { if (atoi(ra) != 0) ldlt_roww |= 0x200; }
    if (const char* ts = std::getenv("BA_LDLT_TWOSIDED")) two_sided = atoi(ts) != 0;
    if (const char* v2 = std::getenv("BA_LDLT_V2")) ldlt_v2 = atoi(v2) != 0;
    if (const char* v3 = std::getenv("BA_LDLT_SPLIT")) ldlt_split = atoi(v3);
    if (const char* tlm = std::getenv("BA_SPLIT_TIMELINE")) { split_timeline = true; tl_deferred = atoi(tlm) >= 2; tl_s2only = atoi(tlm) == 3; }
    if (const char* v4 = std::getenv("BA_LDLT_SPLIT_SEGMENTS")) split_segments = std::max(1, std::min(4, atoi(v4)));
    CK(cudaFuncSetAttribute(k_spike, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpikeSmem)));
    CK(cudaFuncSetAttribute(k_sep_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM));
    CK(cudaFuncSetAttribute(k_band_ldlt_fwd2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Ldlt2Smem)));
    CK(cudaFuncSetAttribute(k_band_ldlt_fwd2, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(k_band_ldlt_cluster<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ClusterSmem<T>)));
    CK(cudaFuncSetAttribute(k_band_ldlt_cluster<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(k_band_qr_backsolve_cluster<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (; cluster_size > 1; cluster_size /= 2) {  // largest cluster the device can co-schedule
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cluster_size); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = sizeof(ClusterSmem<T>);
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, k_band_ldlt_cluster<T>, &cfg) == cudaSuccess && nclusters >= 1) break;
      cudaGetLastError();
    }
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_band_qr<T>, QR_THREADS, 0));
    coop_grid_qr = std::max(1, std::min(occ, 2)) * sms;
    return BA_OK;
  }

  // ---- timing helpers
  void mark(int i) { if (profiling) cudaEventRecord(ev[i], stream); }
  void collect(int first, int last, const int* stage_of) {
    if (!profiling) return;
    cudaEventSynchronize(ev[last]);
    for (int i = first; i < last; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); stage_ms[stage_of[i - first]] = ms; }
  }

  TileArgs<T> tile_args(T lam) {
    TileArgs<T> a;
    a.tile_pt = d_tile_pt.p; a.pt_start = d_pt_start.p; a.view = d_view.p; a.point = d_point.p; a.meas = d_meas.p;
    a.cams = d_cams.p; a.X = d_X.p; a.tau2 = tau2; a.lambda = lam;
    a.factor = (variant == BA_CHOLESKY) ? PF_NORMAL : PF_HOUSEHOLDER;
    return a;
  }

  // host staging (pinned): the camera records are packed here; in the double build points / steps move straight
  // between the caller's buffers and HBM (full PCIe rate when the caller's memory is pinned)
  T* h_stage = nullptr; size_t h_stage_n = 0;
  int stage_buf(size_t count) {
    if (h_stage_n >= count) return BA_OK;
    if (h_stage) cudaFreeHost(h_stage);
    h_stage = nullptr; h_stage_n = 0;
    CK(cudaMallocHost(&h_stage, count * sizeof(T)));
    h_stage_n = count;
    return BA_OK;
  }

  int set_state(const double* R, const double* Tt, const double* f, const double* k1, const double* k2, const double* X) override {
    CK(cudaSetDevice(device));
    const size_t nc = (size_t)N * CAM_STRIDE, nx = 3 * (size_t)M;
    const bool direct = sizeof(T) == sizeof(double);
    { int rc = stage_buf(nc + (direct ? 0 : nx)); if (rc) return rc; }
    T* c = h_stage;
    for (int i = 0; i < N; ++i) {
      T* o = c + (size_t)i * CAM_STRIDE;
      for (int b = 0; b < 9; ++b) o[b] = (T)R[9 * (size_t)i + b];
      for (int b = 0; b < 3; ++b) o[9 + b] = (T)Tt[3 * (size_t)i + b];
      o[12] = (T)f[i]; o[13] = (T)k1[i]; o[14] = (T)k2[i]; o[15] = T(0);
    }
    CK(cudaMemcpyAsync(d_cams.p, c, nc * sizeof(T), cudaMemcpyHostToDevice, stream));
    if (direct) {
      CK(cudaMemcpyAsync(d_X.p, X, nx * sizeof(T), cudaMemcpyHostToDevice, stream));
    } else {
      T* x = h_stage + nc;
      for (size_t i = 0; i < nx; ++i) x[i] = (T)X[i];
      CK(cudaMemcpyAsync(d_X.p, x, nx * sizeof(T), cudaMemcpyHostToDevice, stream));
    }
    CK(cudaStreamSynchronize(stream));
    computed = tried = linearized = false; stage1_valid = false;
    return BA_OK;
  }

  // One LM trial from host state to host step in one call: set_state + linearize (energy only) + compute + solve_try +
  // get_dx, with the copies pipelined against the point stage (see sx_* above). Same results, bit for bit, as the
  // separate calls. Host buffers should be page-locked for the copies to overlap; they must stay valid until the call returns.
  int step_streamed(const double* R, const double* Tt, const double* f, const double* k1, const double* k2, const double* X, double lam,
                    double* dx, double* energy, double* dx_norm, double* rho_den, double* energy_test) override {
    CK(cudaSetDevice(device));
    if (R && (!Tt || !f || !k1 || !k2 || !X)) return fail(BA_ERR_ARG, "ba_step_streamed: pass all six state arrays or none");
    if (sizeof(T) != sizeof(double) || two_stage) {
      int rc = R ? set_state(R, Tt, f, k1, k2, X) : BA_OK;
      if (!rc) rc = linearize(energy, nullptr, nullptr);
      if (!rc) rc = compute(lam);
      if (!rc) rc = solve_try(dx_norm, rho_den, energy_test);
      if (!rc && dx) rc = get_dx(dx);
      return rc;
    }
    if (!R) {   // state already resident on the device: energy + compute + solve_try (+ download) with a single synchronisation
      computed = tried = false;
      int rc = compute(lam);
      if (!rc) rc = energy_pass(0);
      if (rc) return rc;
      linearized = true;
      dx_sink = dx;
      rc = solve_try(dx_norm, rho_den, energy_test);
      dx_sink = nullptr;
      if (rc) return rc;
      if (energy) *energy = h_scal[0];
      return BA_OK;
    }
    const size_t nc = (size_t)N * CAM_STRIDE;
    { int rc = stage_buf(nc); if (rc) return rc; }
    T* c = h_stage;
    for (int i = 0; i < N; ++i) {
      T* o = c + (size_t)i * CAM_STRIDE;
      for (int b = 0; b < 9; ++b) o[b] = (T)R[9 * (size_t)i + b];
      for (int b = 0; b < 3; ++b) o[9 + b] = (T)Tt[3 * (size_t)i + b];
      o[12] = (T)f[i]; o[13] = (T)k1[i]; o[14] = (T)k2[i]; o[15] = T(0);
    }
    CK(cudaEventRecord(xev[SX_MAX], stream));          // whatever still reads the old state finishes first
    CK(cudaStreamWaitEvent(stream2, xev[SX_MAX], 0));
    CK(cudaMemcpyAsync(d_cams.p, c, nc * sizeof(T), cudaMemcpyHostToDevice, stream2));
    sx_n = std::max(1, std::min(sx_chunks, std::max(nunits, 1)));
    int p0 = 0;
    sx_unit[0] = 0;
    for (int ch = 0; ch < sx_n; ++ch) {
      const int u1 = (int)((long long)nunits * (ch + 1) / sx_n);
      const int p1 = (ch == sx_n - 1) ? M : (u1 > 0 ? std::max(p0, h_unit_last_pt[u1 - 1] + 1) : p0);
      if (p1 > p0)
        CK(cudaMemcpyAsync(reinterpret_cast<double*>(d_X.p) + 3 * (size_t)p0, X + 3 * (size_t)p0, 3 * (size_t)(p1 - p0) * sizeof(double), cudaMemcpyHostToDevice, stream2));
      CK(cudaEventRecord(xev[ch], stream2));
      sx_unit[ch + 1] = u1;
      p0 = p1;
    }
    sx_active = true;
    computed = tried = linearized = false; stage1_valid = false;
    int rc = compute(lam);
    sx_active = false;
    if (rc) return rc;
    rc = energy_pass(0);                                // all of X has arrived: the last chunk's units waited for it
    if (rc) return rc;
    linearized = true;
    dx_sink = dx;
    rc = solve_try(dx_norm, rho_den, energy_test);
    dx_sink = nullptr;
    if (rc) return rc;
    if (energy) *energy = h_scal[0];
    return BA_OK;
  }

  int get_state(double* R, double* Tt, double* f, double* k1, double* k2, double* X) override {
    CK(cudaSetDevice(device));
    const size_t nc = (size_t)N * CAM_STRIDE, nx = 3 * (size_t)M;
    const bool direct = sizeof(T) == sizeof(double);
    { int rc = stage_buf(nc + (direct ? 0 : nx)); if (rc) return rc; }
    T* c = h_stage;
    CK(cudaMemcpyAsync(c, d_cams.p, nc * sizeof(T), cudaMemcpyDeviceToHost, stream));
    if (direct) CK(cudaMemcpyAsync(X, d_X.p, nx * sizeof(T), cudaMemcpyDeviceToHost, stream));
    else CK(cudaMemcpyAsync(h_stage + nc, d_X.p, nx * sizeof(T), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int i = 0; i < N; ++i) {
      const T* o = c + (size_t)i * CAM_STRIDE;
      for (int b = 0; b < 9; ++b) R[9 * (size_t)i + b] = (double)o[b];
      for (int b = 0; b < 3; ++b) Tt[3 * (size_t)i + b] = (double)o[9 + b];
      f[i] = (double)o[12]; k1[i] = (double)o[13]; k2[i] = (double)o[14];
    }
    if (!direct) for (size_t i = 0; i < nx; ++i) X[i] = (double)h_stage[nc + i];
    return BA_OK;
  }

  int allreduce_scal(int first, int count) {
    if (!comm) return BA_OK;
    NK(g_nccl.AllReduce(d_scal.p + first, d_scal.p + first, count, ncclDouble, ncclSum, comm, stream));
    return BA_OK;
  }

  int energy_pass(int slot) {
    const int nb = (K + 255) / 256;
    k_obs<T, 0><<<nb, 256, 0, stream>>>(K, d_view.p, d_point.p, d_meas.p, d_cams.p, d_X.p, tau2, M, d_partials.p, nullptr, nullptr);
    k_reduce<<<1, 1024, 0, stream>>>(d_partials.p, nb, d_scal.p + slot);
    launches += 2;
    CK(cudaGetLastError());
    return allreduce_scal(slot, 1);
  }

  int eval(double* energy) override {
    CK(cudaSetDevice(device));
    int rc = energy_pass(0);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_scal, d_scal.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (energy) *energy = h_scal[0];
    return BA_OK;
  }

  // MOREQR stage 1: QR of the un-damped point blocks at the current x (More.h:288-291), once per outer iteration
  int moreqr_stage1() {
    TileArgs<T> a0 = tile_args(T(0));
    a0.factor = PF_HOUSEHOLDER;
    k_point_factor_warp<T, true><<<(nunits + TILE / 32 - 1) / (TILE / 32), TILE, point_factor_warp_smem_bytes<T>(), stream>>>(a0, nunits, d_unit_pt.p, d_seg.p, d_slot.p,
                                                                                                                     d_J.p, nullptr, d_Pt0.p);
    launches++;
    CK(cudaGetLastError());
    stage1_valid = true;
    return BA_OK;
  }

  int linearize(double* energy, double* max_cn2, double* max_cn) override {
    CK(cudaSetDevice(device));
    const int nb = (K + 255) / 256;
    if (two_stage) { int rc1 = moreqr_stage1(); if (rc1) return rc1; }
    if (max_cn2 || max_cn) {
      const size_t np = 3 * (size_t)M + 9 * (size_t)N;
      if (d_tmp.n < np) CK(d_tmp.alloc(np));
      k_obs<T, 0><<<nb, 256, 0, stream>>>(K, d_view.p, d_point.p, d_meas.p, d_cams.p, d_X.p, tau2, M, d_partials.p, nullptr, nullptr);
      k_reduce<<<1, 1024, 0, stream>>>(d_partials.p, nb, d_scal.p + 0);
      k_colnorm_pt<T><<<(M + 255) / 256, 256, 0, stream>>>(M, d_pt_start.p, d_view.p, d_meas.p, d_cams.p, d_X.p, tau2, d_tmp.p);
      k_colnorm_cam<T><<<N, 128, 0, stream>>>(d_cam_start.p, d_obs_of_slot.p, d_point.p, d_meas.p, d_cams.p, d_X.p, tau2, d_tmp.p + 3 * (size_t)M);
      launches += 4;
      if (comm) {  // camera column norms are sums over all ranks' observations
        NK(g_nccl.AllReduce(d_tmp.p + 3 * (size_t)M, d_tmp.p + 3 * (size_t)M, 9 * (size_t)N, nccl_t(), ncclSum, comm, stream));
      }
      k_max<T><<<1, 256, 0, stream>>>(d_tmp.p, (int)np, d_scal.p + 5);
      launches += 1;
      CK(cudaGetLastError());
      int rc = allreduce_scal(0, 1);
      if (rc) return rc;
      if (comm) NK(g_nccl.AllReduce(d_scal.p + 5, d_scal.p + 5, 1, ncclDouble, ncclMax, comm, stream));
    } else {
      int rc = energy_pass(0);
      if (rc) return rc;
    }
    CK(cudaMemcpyAsync(h_scal, d_scal.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (energy) *energy = h_scal[0];
    if (max_cn2) *max_cn2 = h_scal[5];
    if (max_cn) *max_cn = (double)std::sqrt((T)h_scal[5]);
    linearized = true;
    computed = tried = false;
    return BA_OK;
  }

  int compute(double lam) override {
    CK(cudaSetDevice(device));
    lambda = lam;
    const T lamT = (T)lam;
    const T sl = (T)std::sqrt(lamT);
    const T diag = (variant == BA_CHOLESKY) ? lamT : sl * sl;  // QR variants square the sqrt(lambda) rows
    mark(0);
    CK(cudaMemsetAsync(d_red.p, 0, (red_count + 2 * (size_t)n) * sizeof(T), stream));
    if (two_stage) {
      // per lambda trial only the 6x3 blocks [R0_j; sqrt(lambda) I3] are re-triangularised (More.h:293-348)
      if (!stage1_valid) { int rc1 = moreqr_stage1(); if (rc1) return rc1; }
      k_moreqr_stage2<T><<<(nunits + TILE / 32 - 1) / (TILE / 32), TILE, point_factor_warp_smem_bytes<T>(), stream>>>(lamT, nunits, d_unit_pt.p, d_seg.p, d_slot.p, d_point.p,
                                                                                                                 d_J.p, d_Pt0.p, d_P.p, d_D.p, d_Pt.p);
      launches++;
    } else if (nunits && sx_active) {
      // streamed state upload: the units of a chunk start as soon as the points they read have arrived
      constexpr int UPB = TILE / 32;
      for (int c = 0; c < sx_n; ++c) {
        CK(cudaStreamWaitEvent(stream, xev[c], 0));
        const int u0 = sx_unit[c], nu = sx_unit[c + 1] - u0;
        if (nu > 0) {
          k_point_factor_warp<T><<<(nu + UPB - 1) / UPB, TILE, point_factor_warp_smem_bytes<T>(), stream>>>(tile_args(lamT), nu, d_unit_pt.p + 2 * (size_t)u0, d_seg.p, d_slot.p, d_P.p, d_D.p, d_Pt.p);
          launches++;
        }
      }
      sx_active = false;
    } else if (nunits) { k_point_factor_warp<T><<<(nunits + TILE / 32 - 1) / (TILE / 32), TILE, point_factor_warp_smem_bytes<T>(), stream>>>(tile_args(lamT), nunits, d_unit_pt.p, d_seg.p, d_slot.p, d_P.p, d_D.p, d_Pt.p); launches++; }
    if (sx_active) { for (int c = 0; c < sx_n; ++c) CK(cudaStreamWaitEvent(stream, xev[c], 0)); sx_active = false; }
    if (nbig) {
      TileArgs<T> ab = tile_args(lamT); ab.tile_pt = d_big_pt.p;
      k_point_factor<T><<<nbig, TILE, sizeof(TileSmem<T>), stream>>>(ab, 2, d_slot.p, d_P.p, d_D.p, d_Pt.p); launches++;
    }
    if (nhuge) {
      k_point_factor_big<T><<<nhuge, BIG_THREADS, big_point_smem_bytes<T>(huge_max), stream>>>(tile_args(lamT), d_huge_pt.p, d_slot.p, d_P.p, d_D.p, d_Pt.p);
      launches++;
    }
    CK(cudaGetLastError());
    mark(1);
    CK(cudaMemsetAsync(d_counter.p, 0, sizeof(int), stream));
    k_schur_diag<T><<<N, DIAG_THREADS, diag_smem_bytes<T>(), stream>>>(d_cam_start.p, d_D.p, Sv(), lds(), gvec(), gJvec(), rank == 0 ? diag : T(0));
    if (nblocks > 0)
      k_schur_gather<T><<<gather_grid, GATHER_THREADS, gather_smem_bytes<T>(), stream>>>(nblocks, d_blk_a.p, d_blk_b.p, d_blk_start.p, d_pairs.p, d_P.p,
                                                                     Sv(), lds(), d_counter.p);
    launches += 2;
    CK(cudaGetLastError());
    mark(2);
    if (comm) NK(g_nccl.AllReduce(d_red.p, d_red.p, red_count + 2 * (size_t)n, nccl_t(), ncclSum, comm, stream));
    if (keep_reduced) {
      if (d_keep.n < red_count + n) CK(d_keep.alloc(red_count + n));
      CK(cudaMemcpyAsync(d_keep.p, d_red.p, (red_count + n) * sizeof(T), cudaMemcpyDeviceToDevice, stream));
    }
    mark(3);
    { int rc = factor_reduced(); if (rc) return rc; }
    mark(4);
    static const int stages[] = {0, 1, 2, 3};
    collect(0, 4, stages);
    if (!profiling) { /* stay asynchronous: errors surface in solve_try */ }
    computed = true; tried = false;
    return BA_OK;
  }

  static cudaError_t launch_fwd2_impl(const cudaLaunchConfig_t& cfg, const LdltJob<double>& job, long long* dbg) { return cudaLaunchKernelEx(&cfg, k_band_ldlt_fwd2, job, dbg); }
  static cudaError_t launch_fwd2_impl(const cudaLaunchConfig_t&, const LdltJob<float>&, long long*) { return cudaErrorNotSupported; }

  // factorisation of the reduced camera block in d_red (LDL^T or Householder QR of S)
  // solve_only: S is already factored (same handle, same band), only a new right-hand side sits in gvec(): forward
  // substitution with the existing factor (do_fwd = 2) instead of factorisation + folded forward substitution
  // Separator split (ba_split.cuh): S = [part 0 | separator | part 1], four elimination chains side by side, spikes,
  // separator Schur complement, then the backward passes in the opposite order. Returns false when the system is too
  // small for it (the caller falls back to the two-sided scheme).
  template <class L> int factor_reduced_split(bool solve_only, L&& launch, bool& done) {
    done = false;
    if constexpr (sizeof(T) == 8) {
      const int fwd = solve_only ? 2 : 1;
      const int bt = (kd + NB - 1) / NB;
      const SplitPlan pl = split_plan(n, kd, ldlt_split);
      if (!pl.ok) return BA_OK;
      const int w = pl.w, s0 = pl.s0, p1 = pl.p1;
      const int npart[2] = {pl.npart[0], pl.npart[1]};
      int q[2], r0[2], nm[2], nph[2], ntm[2], npE[2], ldE[2];
      for (int p = 0; p < 2; ++p) { q[p] = pl.q[p]; r0[p] = pl.r0[p]; nm[p] = pl.nm[p]; nph[p] = pl.nph[p]; ntm[p] = pl.ntm[p]; npE[p] = pl.npE[p]; ldE[p] = pl.ldE[p]; }
      BandMat<T> A = band();
      const size_t ldv = lds();
      T* Xv[2]; T* gX[2]; T* yX[2]; T* Rv[2]; T* gr[2];
      for (int p = 0; p < 2; ++p) {
        SplitPart& P = sp[p];
        const size_t revc = (size_t)nph[p] * (ldv + 1);
        const int ntp = (npart[p] + NB - 1) / NB + 1, nth = (nph[p] + NB - 1) / NB + 1;
        if (P.rev.n < revc + nph[p]) CK(P.rev.alloc(revc + nph[p]));
        if (P.dvec.n < (size_t)npart[p] + 2 * NB) CK(P.dvec.alloc((size_t)npart[p] + 2 * NB));
        if (P.dvec2.n < (size_t)nph[p] + 2 * NB) CK(P.dvec2.alloc((size_t)nph[p] + 2 * NB));
        if (P.W.n < (size_t)ntp * NB * NB) CK(P.W.alloc((size_t)ntp * NB * NB));
        if (P.W2.n < (size_t)nth * NB * NB) CK(P.W2.alloc((size_t)nth * NB * NB));
        if (P.y2.n < (size_t)nph[p]) CK(P.y2.alloc(nph[p]));
        if (P.E.n < (size_t)w * ldE[p]) CK(P.E.alloc((size_t)w * ldE[p]));
        Rv[p] = P.rev.p + ldv; gr[p] = P.rev.p + revc;
      }
      {
        const size_t fullc = (size_t)npart[0] * (ldv + 1);
        if (sp[0].full.n < fullc + npart[0]) CK(sp[0].full.alloc(fullc + npart[0]));
        if (sp[0].y.n < (size_t)npart[0]) CK(sp[0].y.alloc(npart[0]));
        Xv[0] = sp[0].full.p + ldv; gX[0] = sp[0].full.p + fullc; yX[0] = sp[0].y.p;
        Xv[1] = Sv() + (size_t)p1 * ldv + p1; gX[1] = gvec() + p1; yX[1] = d_dx_cam.p + p1;
      }
      const int kds = w - 1, ldw = pad_lds(kds), nts = (w + NB - 1) / NB;
      const size_t sepc = (size_t)(w + 2 * NB) * (ldw + 1);
      if (d_sep.n < sepc) CK(d_sep.alloc(sepc));
      if (sep_w != w) { CK(cudaMemsetAsync(d_sep.p, 0, d_sep.n * sizeof(T), stream)); sep_w = w; }   // nothing stale outside the lower triangle
      if (d_sep_vec.n < (size_t)3 * (w + 2 * NB)) CK(d_sep_vec.alloc((size_t)3 * (w + 2 * NB)));
      if (d_sep_W.n < (size_t)(nts + 1) * NB * NB) CK(d_sep_W.alloc((size_t)(nts + 1) * NB * NB));
      T* Sd = d_sep.p + ldw; T* gs = d_sep_vec.p; T* ds = gs + (w + 2 * NB); T* ys = ds + (w + 2 * NB);
      const int ab = 4 * sm_count;
      // part 0 as an index-reversed copy of its own (the separator then sits above row 0 of either part)
      BandMat<T> A0{Sv(), ldv, npart[0], kd};
      if (split_timeline && tl_deferred && tl_n > 0 && !solve_only) {   // the previous trial's events (complete: its solve_try synchronised)
        if (cudaEventQuery(tlv[tl_n - 1]) == cudaSuccess && ++tl_calls % 8 == 0)
          for (int i = 1; i < tl_n; ++i) { float ms = 0; cudaEventElapsedTime(&ms, tlv[0], tlv[i]); fprintf(stderr, "[split timeline] %-18s %8.3f ms\n", tl_name[i], ms); }
        tl_n = 0;
      }
      tl("split begin", stream);
      if (solve_only) k_rhs_reverse<T><<<64, 256, 0, stream>>>(gvec(), gX[0], npart[0], npart[0], npart[0]);
      BandMat<T> X[2], Ar[2], Am[2];
      LdltJob<T> job = {}, mid = {};
      job.sign = T(-1); mid.sign = T(-1);
      for (int p = 0; p < 2; ++p) {
        X[p] = BandMat<T>{Xv[p], ldv, npart[p], kd};
        Ar[p] = BandMat<T>{Rv[p], ldv, nph[p], kd};
        Am[p] = BandMat<T>{Xv[p] + (size_t)r0[p] * ldv + r0[p], ldv, nm[p], std::min(kd, nm[p] - 1)};
        if (solve_only) k_rhs_reverse<T><<<64, 256, 0, stream>>>(gX[p], gr[p], npart[p], nph[p], r0[p]);
        job.p[2 * p] = LdltProblem<T>{X[p], sp[p].dvec.p, sp[p].W.p, gX[p], yX[p], d_info.p, q[p], 0, fwd, 0};
        job.p[2 * p + 1] = LdltProblem<T>{Ar[p], sp[p].dvec2.p, sp[p].W2.p, gr[p], sp[p].y2.p, d_info.p, q[p], 0, fwd, 0};
        mid.p[p] = LdltProblem<T>{Am[p], sp[p].dvec.p + r0[p], sp[p].W.p + (size_t)q[p] * NB * NB, gX[p] + r0[p], yX[p] + r0[p], d_info.p, ntm[p], 0, fwd, 0};
      }
      if (!solve_only) {
        // the reversed half of (index-reversed) part 0 is the top of part 0 as it stands; all three copies in one launch
        RevJob<T> rj0{A0, gvec(), Xv[0], gX[0], npart[0], npart[0], 1};
        RevJob<T> rj1{A0, gvec(), Rv[0], gr[0], nph[0], r0[0], 0};
        RevJob<T> rj2{X[1], gX[1], Rv[1], gr[1], nph[1], r0[1], 1};
        k_band_reverse3<T><<<dim3(8 * sm_count, 3), 256, 0, stream>>>(rj0, rj1, rj2);
        tl("reverse end", stream);
        CK(cudaEventRecord(sev[9], stream));
        CK(cudaStreamWaitEvent(stream2, sev[9], 0));
        for (int p = 0; p < 2; ++p) k_spike_init<<<ab / 2, 256, 0, stream2>>>(A, sp[p].E.p, ldE[p], w, (bt + 1) * NB, s0, p1, npart[p], p);
        launches += 3;
      }
      // The chains run in split_segments launches (a later segment is the same elimination on the view that starts at
      // its first panel); the spike of a finished segment is formed on stream2 beside the next segment, the last one
      // beside the middle blocks.
      const int nseg = solve_only ? 1 : split_segments;
      const int ncolE[2] = {r0[0] + nm[0], r0[1] + nm[1]};
      const int nstrips = (w + SPK_STRIP - 1) / SPK_STRIP;
      // after_spike: recorded between the spike and its SYRK; before_syrk: the SYRK waits for it (the SYRKs accumulate into the
      // same block and must stay ordered)
      auto spike = [&](int seg0, int seg1, bool middle, cudaStream_t st, cudaEvent_t after_spike = nullptr, cudaEvent_t before_syrk = nullptr) -> cudaError_t {
        SpikeJob sj[2];
        for (int p = 0; p < 2; ++p) {
          const int kb = middle ? q[p] : seg_bound(q[p], seg0, nseg), ke = middle ? npE[p] : seg_bound(q[p], seg1, nseg);
          sj[p] = SpikeJob{BandMat<double>{Xv[p], ldv, ncolE[p], kd}, sp[p].dvec.p, sp[p].W.p, sp[p].E.p, ldE[p], kb, ke};
        }
        tl(middle ? "spike_mid begin" : "spike begin", st);
        k_spike<<<2 * nstrips, SPK_THREADS, sizeof(SpikeSmem), st>>>(sj[0], sj[1], w, d_dbg.p);
        tl(middle ? "spike_mid end" : "spike end", st);
        if (after_spike) { cudaError_t e = cudaEventRecord(after_spike, st); if (e != cudaSuccess) return e; }
        if (before_syrk) { cudaError_t e = cudaStreamWaitEvent(st, before_syrk, 0); if (e != cudaSuccess) return e; }
        // the separator block receives the Schur complement of the same panels right away
        k_sep_syrk<<<nts * (nts + 1) / 2, 256, SYRK_SMEM, st>>>(A, s0, w, Sd, ldw, SyrkSide{sp[0].E.p, ldE[0], sp[0].dvec.p, sj[0].k_begin, sj[0].k_end},
                                                              SyrkSide{sp[1].E.p, ldE[1], sp[1].dvec.p, sj[1].k_begin, sj[1].k_end}, (!middle && seg0 == 0) ? 1 : 0);
        launches += 2;
        return cudaSuccess;
      };
      for (int sg = 0; sg < nseg; ++sg) {
        LdltJob<T> seg = job;
        for (int c = 0; c < 4; ++c) {
          const int p = c / 2, a = seg_bound(q[p], sg, nseg), b = seg_bound(q[p], sg + 1, nseg);
          LdltProblem<T>& P = seg.p[c];
          P.A = BandMat<T>{P.A.v + (size_t)a * NB * (ldv + 1), ldv, P.A.n - a * NB, kd};
          P.dvec += (size_t)a * NB; P.Wbuf += (size_t)a * NB * NB; P.rhs += (size_t)a * NB; P.y += (size_t)a * NB;
          P.np_fwd = b - a;
        }
        tl("chain seg begin", stream);
        const bool beside = !solve_only && sg > 0;
        CK(launch(seg, 4, beside ? lev[sg] : nullptr));
        tl("chain seg end", stream);
        if (!solve_only) {
          CK(cudaEventRecord(sev[sg], stream));
          if (sg > 0) {               // spike of segment sg - 1 beside chain segment sg, once its clusters have been placed
            CK(cudaStreamWaitEvent(stream2, sev[sg - 1], 0));
            if (beside) CK(cudaStreamWaitEvent(stream2, lev[sg], 0));
            CK(spike(sg - 1, sg, false, stream2));
          }
        }
      }
      k_band_combine2<T><<<dim3(solve_only ? 8 : 64, 2), 256, 0, stream>>>(CombJob<T>{X[0], gX[0], Rv[0], gr[0], r0[0], nm[0]},
                                                                         CombJob<T>{X[1], gX[1], Rv[1], gr[1], r0[1], nm[1]}, solve_only ? 1 : 0);
      CK(launch(mid, 2, !solve_only ? lev[0] : nullptr));
      tl("mid end", stream);
      if (!solve_only) {
        CK(cudaStreamWaitEvent(stream2, sev[nseg - 1], 0));
        CK(cudaStreamWaitEvent(stream2, lev[0], 0));
        CK(spike(nseg - 1, nseg, false, stream2, sev[7]));   // beside the middle blocks
        CK(cudaEventRecord(sev[8], stream2));
        CK(cudaStreamWaitEvent(stream, sev[7], 0));          // the middle panels' spike needs the last segment's spike,
        CK(spike(0, 0, true, stream, nullptr, sev[8]));      // their SYRK the last segment's SYRK
      }
      k_sep_rhs<<<w, 256, 0, stream>>>(gvec(), s0, w, gs, RhsSide{sp[0].E.p, ldE[0], sp[0].dvec.p, gX[0], ncolE[0]},
                                                 RhsSide{sp[1].E.p, ldE[1], sp[1].dvec.p, gX[1], ncolE[1]});
      LdltJob<T> sep = {};
      sep.sign = T(-1);
      sep.p[0] = LdltProblem<T>{BandMat<T>{Sd, (size_t)ldw, w, kds}, ds, d_sep_W.p, gs, ys, d_info.p, nts, nts, fwd, 1};
      tl("sep begin", stream);
      CK(launch(sep, 1));
      tl("sep end", stream);
      k_spike_correct<<<dim3((std::max(ncolE[0], ncolE[1]) + 31) / 32, 2), 256, 0, stream>>>(CorrSide{sp[0].E.p, ldE[0], ncolE[0], gX[0]}, CorrSide{sp[1].E.p, ldE[1], ncolE[1], gX[1]}, w, ys, -1.0);
      for (int p = 0; p < 2; ++p) { mid.p[p].do_fwd = 0; mid.p[p].do_bwd = 1; mid.p[p].kb_bwd = ntm[p]; }
      CK(launch(mid, 2));
      k_flip_copy2<T><<<dim3(8, 2), 256, 0, stream>>>(FlipJob<T>{sp[0].y2.p, yX[0], npart[0], r0[0], nph[0]}, FlipJob<T>{sp[1].y2.p, yX[1], npart[1], r0[1], nph[1]});
      for (int p = 0; p < 2; ++p)
        for (int c = 0; c < 2; ++c) { job.p[2 * p + c].do_fwd = 0; job.p[2 * p + c].do_bwd = 1; job.p[2 * p + c].kb_bwd = q[p]; }
      CK(launch(job, 4));
      k_split_assemble<T><<<64, 256, 0, stream>>>(d_dx_cam.p, n, s0, p1, yX[0], sp[0].y2.p, r0[0] + nm[0], ys, sp[1].y2.p, r0[1] + nm[1]);
      launches += 10;
      CK(cudaGetLastError());
      tl("split end", stream);
      if (split_timeline && !solve_only && !tl_deferred) {
        CK(cudaStreamSynchronize(stream)); CK(cudaStreamSynchronize(stream2));
        for (int i = 1; i < tl_n; ++i) { float ms = 0; cudaEventElapsedTime(&ms, tlv[0], tlv[i]); fprintf(stderr, "[split timeline] %-18s %8.3f ms\n", tl_name[i], ms); }
      }
      if (!tl_deferred) tl_n = 0;
      done = true;
    }
    return BA_OK;
  }

  int factor_reduced(bool solve_only = false) {
    if (variant == BA_QRCHOL || variant == BA_CHOLESKY || qr_csne) {
      if (!solve_only) CK(cudaMemsetAsync(d_info.p, 0, sizeof(int), stream));
      const int fwd = solve_only ? 2 : 1;
      BandMat<T> A = band();
      const int nt = (n + NB - 1) / NB, bt = (kd + NB - 1) / NB;
      const double flops = (double)n * kd * kd;
      if (flops < 2e11 && bt <= CL_MAX_BT && !force_grid_ldlt) {
        // latency-bound regime: thread-block clusters, forward solve folded in, backward pass on the cluster
        // started: recorded once all CTAs of the launch have begun execution (launch completion event; lets the spike kernel on
        // the second stream start only after the clusters of the next chain segment have been placed)
        auto launch = [&](const LdltJob<T>& job, int nclusters, cudaEvent_t started = nullptr) -> cudaError_t {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(cluster_size * nclusters); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = sizeof(ClusterSmem<T>); cfg.stream = stream;
          cudaLaunchAttribute attr[2];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr; cfg.numAttrs = 1;
          if (started) {
            attr[1].id = cudaLaunchAttributeLaunchCompletionEvent;
            attr[1].val.launchCompletionEvent.event = started; attr[1].val.launchCompletionEvent.flags = 0;
            cfg.numAttrs = 2;
          }
          launches++;
#ifdef BA_L2_TICKS
          return cudaLaunchKernelEx(&cfg, k_band_ldlt_cluster<T>, job, (long long*)nullptr, ldlt_roww);  // keep fwd2's counters
#else
          return cudaLaunchKernelEx(&cfg, k_band_ldlt_cluster<T>, job, d_dbg.p, ldlt_roww);
#endif
        };
        auto launch_fwd2 = [&](const LdltJob<T>& job) -> cudaError_t {  // owner-computes forward elimination (ba_ldlt2.cuh)
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(16 * 2); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = sizeof(Ldlt2Smem); cfg.stream = stream;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr; cfg.numAttrs = 1;
          launches++;
          return launch_fwd2_impl(cfg, job, d_dbg.p);
        };
        if (d_W.n < (size_t)nt * NB * NB) CK(d_W.alloc((size_t)nt * NB * NB));
        const int q = (n - mid_rows_min(kd)) / (2 * NB);  // panels eliminated from either end by the two-sided scheme
        bool split_done = false;
        if (two_sided && ldlt_split && cluster_size == 16) { int rc = factor_reduced_split(solve_only, launch, split_done); if (rc) return rc; }
        if (split_done) {
        } else if (!two_sided || q < bt + 2) {
          LdltJob<T> job = {};
          job.sign = T(-1);
          job.p[0] = LdltProblem<T>{A, d_dvec.p, d_W.p, gvec(), d_dx_cam.p, d_info.p, nt, nt, fwd, 1};
          CK(launch(job, 1));
        } else {
          // Two-sided (twisted) factorisation: the chain of n dependent pivots is cut in two. Cluster 0 eliminates rows
          // [0, r0) top-down, cluster 1 rows [r0 + nm, n) bottom-up (= top-down on the index-reversed copy S'); both
          // leave their Schur complement on the middle block [r0, r0 + nm), nm >= kd + 2 panels, which is summed,
          // factored and solved by one cluster; the two backward passes then run side by side again.
          const int r0 = q * NB, nm = n - 2 * q * NB, np = n - r0;          // np rows of S' = bottom + middle
          const int ntp = (np + NB - 1) / NB, ntm = (nm + NB - 1) / NB;
          const size_t rev_count = (size_t)np * (ldsv + 1);
          if (d_rev.n < rev_count + np) CK(d_rev.alloc(rev_count + np));
          if (d_dvec2.n < (size_t)np + NB) CK(d_dvec2.alloc((size_t)np + NB));
          if (d_W2.n < (size_t)ntp * NB * NB) CK(d_W2.alloc((size_t)ntp * NB * NB));
          if (d_y2.n < (size_t)np) CK(d_y2.alloc(np));
          if (d_WM.n < (size_t)ntm * NB * NB) CK(d_WM.alloc((size_t)ntm * NB * NB));
          T* Rv = d_rev.p + ldsv; T* gr = d_rev.p + rev_count;
          BandMat<T> Ar{Rv, lds(), np, kd};
          BandMat<T> Am{Sv() + (size_t)r0 * ldsv + r0, lds(), nm, std::min(kd, nm - 1)};
          const int ab = 4 * sm_count;
          if (solve_only) k_rhs_reverse<T><<<64, 256, 0, stream>>>(gvec(), gr, n, np, q * NB);
          else k_band_reverse<T><<<ab, 256, 0, stream>>>(A, gvec(), Rv, gr, np, q * NB);
          LdltJob<T> job = {};
          job.sign = T(-1);
          job.p[0] = LdltProblem<T>{A, d_dvec.p, d_W.p, gvec(), d_dx_cam.p, d_info.p, q, 0, fwd, 0};
          job.p[1] = LdltProblem<T>{Ar, d_dvec2.p, d_W2.p, gr, d_y2.p, d_info.p, q, 0, fwd, 0};
          if (!solve_only && ldlt_v2 && sizeof(T) == 8 && cluster_size == 16 && bt <= L2_MAX_BT) CK(launch_fwd2(job));
          else CK(launch(job, 2));
          if (solve_only) k_rhs_combine<T><<<8, 256, 0, stream>>>(gvec(), gr, n, r0, nm);
          else k_band_combine<T><<<64, 256, 0, stream>>>(A, gvec(), Rv, gr, r0, nm);
          LdltJob<T> mid = {};
          mid.sign = T(-1);
          mid.p[0] = LdltProblem<T>{Am, d_dvec.p + r0, d_WM.p, gvec() + r0, d_dx_cam.p + r0, d_info.p, ntm, ntm, fwd, 1};
          CK(launch(mid, 1));
          k_flip_copy<T><<<8, 256, 0, stream>>>(d_y2.p, d_dx_cam.p, n, q * NB, np);  // y'(i') = y(n-1-i') on the middle rows
          job.p[0].do_fwd = 0; job.p[0].do_bwd = 1; job.p[0].kb_bwd = q;
          job.p[1].do_fwd = 0; job.p[1].do_bwd = 1; job.p[1].kb_bwd = q;
          CK(launch(job, 2));
          k_flip_copy<T><<<32, 256, 0, stream>>>(d_dx_cam.p, d_y2.p, n, r0 + nm, n);  // y(i) = y'(n-1-i) on the bottom rows
          launches += 4;
          CK(cudaGetLastError());
        }
        solved_in_factor = true;
      } else {
        T* dv = d_dvec.p; int* info = d_info.p;
        void* args[] = {&A, &dv, &info};
        const int useful = std::max(1, std::min(bt, nt) * (std::min(bt, nt) + 1) / 2);
        const int grid = std::max(1, std::min(coop_grid, useful));
        if (!solve_only) CK(cudaLaunchCooperativeKernel((void*)k_band_ldlt<T>, dim3(grid), dim3(DENSE_THREADS), args, 0, stream));
        solved_in_factor = false;
      }
      launches++;
    } else {
      int rc = qr_factor();
      if (rc) return rc;
    }
    return BA_OK;
  }

  // Householder QR of the square symmetric band matrix (QRKIT / MOREQR right block)
  int qr_factor() {
    const int ku = std::min(n - 1, 2 * kd);
    const size_t ld = (size_t)kd + ku + 1;
    if (d_qr.n < (size_t)n * ld + 2 * (size_t)n + 128) CK(d_qr.alloc((size_t)n * ld + 2 * (size_t)n + 128));
    T* G = d_qr.p; T* tauv = d_qr.p + (size_t)n * ld; T* rhs = tauv + n; T* Tg2 = rhs + n;  // Tg2: T of the current / next panel
    k_band_expand<T><<<std::min<size_t>(((size_t)n * ld + 255) / 256, 65535), 256, 0, stream>>>(band(), G, ld, ku, gvec(), rhs);
    launches++;
    CK(cudaGetLastError());
    QRMat<T> Q{G, ld, n, kd, ku};
    long long* dbgp = d_dbg.p;
    void* args[] = {&Q, &tauv, &rhs, &Tg2, &dbgp};
    if (kd + QR_PB <= 32 * QR_MAXR) {  // banded case: reflectors in shared memory, columns in registers
      const size_t smem = QrSmem<T>::bytes(kd);
      CK(cudaFuncSetAttribute(k_band_qr_reg<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int occ = 0, sms = 0;
      CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_band_qr_reg<T>, QR_THREADS, smem));
      const int want = 1 + (std::min(n - 1, 2 * kd) + 1 + QR_THREADS / 32 - 1) / (QR_THREADS / 32);  // panel CTA + one warp per trailing column
      const int grid = std::max(1, std::min(std::max(occ, 1) * sms, want));
      CK(cudaLaunchCooperativeKernel((void*)k_band_qr_reg<T>, dim3(grid), dim3(QR_THREADS), args, smem, stream));
    } else if (kd + QR_PB <= 2560 && QrTallSmem<T>::bytes(kd) <= 200 * 1024 && !std::getenv("BA_QR_GLOBAL")) {
      // columns too tall for registers (dense S of the larger bundled problems): streamed compact-WY updates
      const size_t smem = QrTallSmem<T>::bytes(kd);
      void* fn = (kd + QR_PB <= 1536) ? (void*)k_band_qr_tall<T, 6> : (void*)k_band_qr_tall<T, 10>;
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int occ = 0, sms = 0;
      CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, QR_THREADS, smem));
      const int want = 1 + (std::min(n - 1, 2 * kd) / 2 + 1 + QR_THREADS / 32 - 1) / (QR_THREADS / 32);  // panel CTA + one warp per two trailing columns
      const int grid = std::max(1, std::min(std::max(occ, 1) * sms, want));
      void* targs[] = {&Q, &tauv, &rhs, &Tg2};
      CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(QR_THREADS), targs, smem, stream));
    } else {
      CK(cudaLaunchCooperativeKernel((void*)k_band_qr<T>, dim3(coop_grid_qr), dim3(QR_THREADS), args, 0, stream));
    }
    launches++;
    return BA_OK;
  }

  // dx_cam = -S^-1 g from the factor above
  int solve_reduced() {
    if (variant == BA_QRCHOL || variant == BA_CHOLESKY || qr_csne) {
      // QR variants: y = S^-1 g, dx_cam = -y. CHOLESKY: g already holds b_c - W V^-1 b_p up to sign (see k_schur), same sign rule.
      if (!solved_in_factor) {
        k_band_ldlt_solve<T><<<1, SOLVE_THREADS, 0, stream>>>(band(), d_dvec.p, gvec(), d_dx_cam.p, T(-1));
        launches++;
      }
    } else {
      const int ku = std::min(n - 1, 2 * kd);
      const size_t ld = (size_t)kd + ku + 1;
      T* G = d_qr.p; T* rhs = d_qr.p + (size_t)n * ld + n;
      QRMat<T> Q{G, ld, n, kd, ku};
      if (cluster_size >= 2 && !std::getenv("BA_QR_SOLVE_1CTA")) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cluster_size); cfg.blockDim = dim3(QRS_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, k_band_qr_backsolve_cluster<T>, Q, rhs, d_dx_cam.p, T(-1)));
      } else {
        k_band_qr_backsolve<T><<<1, QR_SOLVE_THREADS, 0, stream>>>(Q, rhs, d_dx_cam.p, T(-1));
      }
      launches++;
    }
    return BA_OK;
  }

  // QRKIT / MOREQR: corrected semi-normal refinement of dx_cam = -y0 (see k_csne_point): r = J2bot^T (d - J2bot y0) through
  // the D records, S delta = r by forward + backward substitution with the factor of S already in place, dx_cam -= delta
  int csne_refine(T lamT) {
    const T sl = (T)std::sqrt(lamT);
    const T diag = sl * sl;
    for (int it = 0; it < csne_steps; ++it) {
      if (nunits) k_csne_point<T><<<(nunits + TILE / 32 - 1) / (TILE / 32), TILE, 0, stream>>>(nunits, d_unit_pt.p, d_seg.p, d_slot.p, d_view.p, d_D.p, d_dx_cam.p, d_u.p);
      if (nlong) k_csne_point_long<T><<<(nlong + TILE / 32 - 1) / (TILE / 32), TILE, 0, stream>>>(nlong, d_long_pt.p, d_pt_start.p, d_slot.p, d_view.p, d_D.p, d_dx_cam.p, d_u.p);
      k_csne_cam<T><<<N, 128, 0, stream>>>(d_cam_start.p, d_D.p, d_u.p, d_dx_cam.p, rank == 0 ? diag : T(0), gvec());
      launches += 2 + (nlong ? 1 : 0);
      CK(cudaGetLastError());
      if (comm) NK(g_nccl.AllReduce(gvec(), gvec(), (size_t)n, nccl_t(), ncclSum, comm, stream));
      std::swap(d_dx_cam.p, d_delta.p);                 // the solvers write -S^-1 rhs into d_dx_cam
      int rc = factor_reduced(true);                    // S is factored: forward + backward substitution only
      if (!rc) rc = solve_reduced();
      std::swap(d_dx_cam.p, d_delta.p);
      if (rc) return rc;
      k_axpy1<T><<<std::max(1, std::min(64, (n + 255) / 256)), 256, 0, stream>>>(n, d_delta.p, d_dx_cam.p);
      launches++;
      CK(cudaGetLastError());
    }
    return BA_OK;
  }

  int solve_try(double* dx_norm, double* rho_den, double* energy_test) override {
    CK(cudaSetDevice(device));
    if (!computed) return fail(BA_ERR_STATE, "ba_solve_try called before ba_compute");
    const T lamT = (T)lambda;
    mark(4);
    { int rc = solve_reduced(); if (rc) return rc; }
    CK(cudaGetLastError());
    if (qr_csne) { int rc = csne_refine(lamT); if (rc) return rc; }
    mark(5);
    k_cam_update<T><<<camup_blocks(), CAMUP_THREADS, 0, stream>>>(N, d_cams.p, d_dx_cam.p, gJvec(), d_cams_test.p, d_scal.p + 4, d_scal.p + 6, d_camup.p, d_camup_cnt.p);
    launches++;
    mark(6);
    const int ntot = ntiles + nhuge;
    if (nhuge)
      k_backsub_big<T><<<nhuge, BIG_THREADS, 0, stream>>>(tile_args(lamT), d_huge_pt.p, d_P.p, d_Pt.p, d_dx_cam.p, d_cams_test.p, d_dx_pt.p, d_X_test.p,
                                                          d_partials.p, ntiles, ntot);
    if (ntiles && dx_sink && sizeof(T) == sizeof(double)) {
      // streamed download: the step of the points of a chunk of tiles goes to the host while the next chunk is computed
      const int C = std::max(1, std::min(sx_chunks, ntiles));
      CK(cudaEventRecord(bev[SX_MAX], stream));
      CK(cudaStreamWaitEvent(stream2, bev[SX_MAX], 0));
      CK(cudaMemcpyAsync(dx_sink + 3 * (size_t)M, d_dx_cam.p, 9 * (size_t)N * sizeof(T), cudaMemcpyDeviceToHost, stream2));
      int p0 = 0;
      for (int c = 0; c < C; ++c) {
        const int t0 = (int)((long long)ntiles * c / C), t1 = (int)((long long)ntiles * (c + 1) / C);
        if (t1 <= t0) continue;
        TileArgs<T> ta = tile_args(lamT); ta.tile_pt = d_tile_pt.p + 4 * (size_t)t0;
        k_backsub_eval<T><<<t1 - t0, TILE, backsub_smem_bytes<T>(), stream>>>(ta, d_P.p, d_Pt.p, d_dx_cam.p, d_cams_test.p, d_dx_pt.p, d_X_test.p,
                                                                            d_partials.p + t0, ntot);
        launches++;
        const int p1 = (t1 == ntiles) ? M : h_tile_first[t1];
        CK(cudaEventRecord(bev[c], stream));
        CK(cudaStreamWaitEvent(stream2, bev[c], 0));
        CK(cudaMemcpyAsync(dx_sink + 3 * (size_t)p0, d_dx_pt.p + 3 * (size_t)p0, 3 * (size_t)(p1 - p0) * sizeof(T), cudaMemcpyDeviceToHost, stream2));
        p0 = p1;
      }
      CK(cudaEventRecord(bev[SX_MAX + 1], stream2));
      CK(cudaStreamWaitEvent(stream, bev[SX_MAX + 1], 0));
    } else if (ntiles) {
      k_backsub_eval<T><<<ntiles, TILE, backsub_smem_bytes<T>(), stream>>>(tile_args(lamT), d_P.p, d_Pt.p, d_dx_cam.p, d_cams_test.p, d_dx_pt.p, d_X_test.p,
                                                      d_partials.p, ntot);
      launches++;
    }
    launches += (nhuge ? 1 : 0);

    CK(cudaGetLastError());
    mark(7);
    k_reduce<<<3, 1024, 0, stream>>>(d_partials.p, ntot, d_scal.p + 1);
    launches++;
    int rc = allreduce_scal(1, 3);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_scal, d_scal.p, 7 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h_scal + 8, d_info.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    mark(8);
    CK(cudaStreamSynchronize(stream));
    static const int stages[] = {4, 5, 6, 7};
    collect(4, 8, stages);
    const double dx2 = h_scal[2] + h_scal[4];
    if (dx_norm) *dx_norm = std::sqrt(dx2);
    if (rho_den) *rho_den = lambda * dx2 - h_scal[3] - h_scal[6];  // dx^T(lambda dx + JtRes), JtRes = -J^T r
    if (energy_test) *energy_test = h_scal[1];
    tried = true;
    // A singular / NaN reduced system (zero or NaN pivot recorded by the LDL^T kernels, or a non-finite step out of
    // either right-block solver) is never reported as an ordinary trial: the test energy becomes NaN, which the
    // reference's `energyTest < m_energy` (QRChol.h:374) treats as a rejection, and ba_numeric_status says why;
    // with ba_set_strict_numeric the call fails with BA_ERR_NUMERIC instead.
    int bad = 0;
    std::memcpy(&bad, h_scal + 8, sizeof(int));
    numeric_info = bad > 0 ? bad : ((std::isfinite(dx2) && std::isfinite(h_scal[1])) ? 0 : -1);
    if (numeric_info != 0) {
      if (energy_test) *energy_test = std::nan("");
      if (strict_numeric) {
        if (numeric_info > 0) return fail(BA_ERR_NUMERIC, "reduced camera system: zero or NaN pivot at row %d (lambda = %g)", numeric_info - 1, lambda);
        return fail(BA_ERR_NUMERIC, "non-finite step from the reduced camera system (lambda = %g)", lambda);
      }
    }
    return BA_OK;
  }

  int accept() override {
    if (!tried) return fail(BA_ERR_STATE, "ba_accept called without a trial step");
    std::swap(d_cams.p, d_cams_test.p);
    std::swap(d_X.p, d_X_test.p);
    computed = tried = linearized = false; stage1_valid = false;
    return BA_OK;
  }
  int reject() override { tried = false; return BA_OK; }

  template <class A> int d2h(const A* dev, double* host, size_t count) {
    if (sizeof(A) == sizeof(double)) {  // double build: straight into the caller's buffer
      CK(cudaMemcpyAsync(host, dev, count * sizeof(A), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      return BA_OK;
    }
    std::vector<A> tmp(count);
    CK(cudaMemcpyAsync(tmp.data(), dev, count * sizeof(A), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (size_t i = 0; i < count; ++i) host[i] = (double)tmp[i];
    return BA_OK;
  }

  int get_dx(double* dx) override {
    CK(cudaSetDevice(device));
    int rc = d2h(d_dx_pt.p, dx, 3 * (size_t)M);
    if (rc) return rc;
    return d2h(d_dx_cam.p, dx + 3 * (size_t)M, 9 * (size_t)N);
  }

  int error_statistics(double avg_f, double thr, double* out4) override {
    CK(cudaSetDevice(device));
    const int nb = (K + 255) / 256;
    if (d_partials.n < 4 * (size_t)nb) { CK(cudaStreamSynchronize(stream)); CK(d_partials.alloc(4 * (size_t)nb)); }
    k_stats<T><<<nb, 256, 0, stream>>>(K, d_view.p, d_point.p, d_meas.p, d_cams.p, d_X.p, (T)avg_f, (T)thr, d_partials.p, nb);
    k_reduce<<<4, 1024, 0, stream>>>(d_partials.p, nb, d_scal.p + 10);
    launches += 2;
    CK(cudaGetLastError());
    int rc = allreduce_scal(10, 4);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_scal + 10, d_scal.p + 10, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int q = 0; q < 4; ++q) out4[q] = h_scal[10 + q];
    return BA_OK;
  }

  int get_residuals(double* r) override {
    CK(cudaSetDevice(device));
    if (d_tmp.n < 2 * (size_t)K) CK(d_tmp.alloc(2 * (size_t)K));
    const int nb = (K + 255) / 256;
    k_obs<T, 1><<<nb, 256, 0, stream>>>(K, d_view.p, d_point.p, d_meas.p, d_cams.p, d_X.p, tau2, M, d_partials.p, d_tmp.p, nullptr);
    launches++;
    CK(cudaGetLastError());
    return d2h(d_tmp.p, r, 2 * (size_t)K);
  }

  int get_jacobian(double* Jc, double* Jp) override {
    CK(cudaSetDevice(device));
    if (d_tmp.n < 24 * (size_t)K) CK(d_tmp.alloc(24 * (size_t)K));
    const int nb = (K + 255) / 256;
    k_obs<T, 2><<<nb, 256, 0, stream>>>(K, d_view.p, d_point.p, d_meas.p, d_cams.p, d_X.p, tau2, M, d_partials.p, d_tmp.p, d_tmp.p + 18 * (size_t)K);
    launches++;
    CK(cudaGetLastError());
    int rc = d2h(d_tmp.p, Jc, 18 * (size_t)K);
    if (rc) return rc;
    return d2h(d_tmp.p + 18 * (size_t)K, Jp, 6 * (size_t)K);
  }

  int get_reduced(double* S, double* g) override {
    CK(cudaSetDevice(device));
    if (!keep_reduced || d_keep.n < red_count + n) return fail(BA_ERR_STATE, "enable ba_keep_reduced_system before ba_compute");
    std::vector<T> tmp(red_count + n);
    CK(cudaMemcpyAsync(tmp.data(), d_keep.p, tmp.size() * sizeof(T), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (size_t i = 0; i < (size_t)n * n; ++i) S[i] = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = std::max(0, i - kd); j <= i; ++j) {
        const double v = (double)tmp[(size_t)ldsv + (size_t)i * ldsv + j];
        S[(size_t)i * n + j] = v; S[(size_t)j * n + i] = v;
      }
    for (int i = 0; i < n; ++i) g[i] = (double)tmp[red_count + i];
    return BA_OK;
  }

  int comm_init(int rank_, int nranks_, const void* id128) override {
    CK(cudaSetDevice(device));
    if (!g_nccl.load()) return fail(BA_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    rank = rank_; nranks = nranks_;
    NK(g_nccl.CommInitRank(&comm, nranks, id, rank));
    // every rank must use the same band layout (the all-reduce of [S | g | gJ] is element-wise): take the largest
    // block half-bandwidth over the ranks, whatever ba_set_bandwidth was or was not called with
    int* dbw = reinterpret_cast<int*>(d_dbg.p);
    CK(cudaMemcpyAsync(dbw, &bw, sizeof(int), cudaMemcpyHostToDevice, stream));
    NK(g_nccl.AllReduce(dbw, dbw, 1, ncclInt, ncclMax, comm, stream));
    int bw_all = bw;
    CK(cudaMemcpyAsync(&bw_all, dbw, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaMemsetAsync(d_dbg.p, 0, sizeof(long long), stream));
    if (bw_all != bw) { bw = std::min(bw_all, N - 1); computed = tried = false; return alloc_reduced(); }
    return BA_OK;
  }

  int timer_start() override {
    CK(cudaSetDevice(device));
    for (auto& e : tev) if (!e) CK(cudaEventCreate(&e));
    CK(cudaEventRecord(tev[0], stream));
    return BA_OK;
  }
  int timer_stop(double* ms) override {
    CK(cudaSetDevice(device));
    if (!tev[0]) return fail(BA_ERR_STATE, "ba_timer_stop without ba_timer_start");
    CK(cudaEventRecord(tev[1], stream));
    CK(cudaEventSynchronize(tev[1]));
    float f = 0;
    CK(cudaEventElapsedTime(&f, tev[0], tev[1]));
    *ms = f;
    return BA_OK;
  }

  int debug_counters(long long* out, int count) override {
    CK(cudaSetDevice(device));
    if (count < 1 || count > 256) return fail(BA_ERR_ARG, "debug counters: 1..256");
    CK(cudaMemcpy(out, d_dbg.p, count * sizeof(long long), cudaMemcpyDeviceToHost));
    return BA_OK;
  }

  // Test hook: factor + solve an arbitrary symmetric band system with this handle's reduced-system
  // solver (LDLT variants: band LDL^T; QR variants: Householder QR of S). S dense row-major n x n.
  int debug_band_solve(int n_, int kd_, const double* S, const double* g, double* yout) override {
    CK(cudaSetDevice(device));
    const int n0 = n, kd0 = kd, lds0 = ldsv; const size_t rc0 = red_count;
    const bool csne0 = qr_csne; qr_csne = false;    // the hook exercises the Householder band QR itself for the QR variants
    n = n_; kd = std::max(1, std::min(kd_, n_ - 1)); ldsv = pad_lds(kd); red_count = (size_t)n * (ldsv + 1);
    DevBuf<T> red, yv, dv, dxc;
    CK(red.alloc(red_count + 2 * (size_t)n)); CK(yv.alloc(n)); CK(dv.alloc((size_t)n + NB)); CK(dxc.alloc(n));
    std::vector<T> hb(red_count + 2 * (size_t)n, T(0));
    for (int i = 0; i < n; ++i)
      for (int j = std::max(0, i - kd); j <= i; ++j) hb[(size_t)ldsv + (size_t)i * ldsv + j] = (T)S[(size_t)i * n + j];
    for (int i = 0; i < n; ++i) hb[red_count + i] = (T)g[i];
    std::swap(d_red.p, red.p); std::swap(d_dvec.p, dv.p); std::swap(d_dx_cam.p, dxc.p);
    int rc = BA_OK;
    cudaError_t ce = cudaMemcpyAsync(d_red.p, hb.data(), hb.size() * sizeof(T), cudaMemcpyHostToDevice, stream);
    if (ce == cudaSuccess) {
      rc = factor_reduced();
      if (!rc) rc = solve_reduced();
      if (!rc) rc = d2h(d_dx_cam.p, yout, (size_t)n);
      for (int i = 0; i < n && !rc; ++i) yout[i] = -yout[i];  // the solvers return dx_cam = -y
    }
    cudaStreamSynchronize(stream);
    std::swap(d_red.p, red.p); std::swap(d_dvec.p, dv.p); std::swap(d_dx_cam.p, dxc.p);
    n = n0; kd = kd0; ldsv = lds0; red_count = rc0; d_qr.free(); qr_csne = csne0;
    if (ce != cudaSuccess) return fail(BA_ERR_CUDA, "debug_band_solve upload: %s", cudaGetErrorString(ce));
    return rc;
  }

  int set_bandwidth(int bw_) override {
    CK(cudaSetDevice(device));
    if (bw_ < bw) return fail(BA_ERR_ARG, "bandwidth %d smaller than this shard's own %d", bw_, bw);
    bw = std::min(bw_, N - 1);
    computed = tried = false;
    return alloc_reduced();
  }
};

}  // namespace

// ============================================================================================ C ABI
extern "C" {

const char* ba_last_error(void) { return g_err.c_str(); }
const char* ba_version(void) { return "ba_b200 0.4 (sm_100a; kernels: k_point_factor_warp/_big, k_moreqr_stage2, k_schur_diag/gather, k_band_ldlt_cluster, k_band_ldlt_fwd2, k_band_qr_reg/_tall, k_backsub_eval/_big)"; }

int ba_create(ba_handle** out, int N, int M, int K, const int* view, const int* point, const double* meas,
              double inlier_threshold, int precision, int variant, int device) {
  if (!out || !view || !point || !meas) return fail(BA_ERR_ARG, "null argument");
  if (variant < BA_QRKIT || variant > BA_CHOLESKY) return fail(BA_ERR_ARG, "unknown variant %d", variant);
  *out = nullptr;
  int rc;
  if (precision == BA_F32) { auto* p = new Impl<float>(); rc = p->init(N, M, K, view, point, meas, inlier_threshold, variant, device); if (rc) { delete p; return rc; } *out = p; }
  else if (precision == BA_F64) { auto* p = new Impl<double>(); rc = p->init(N, M, K, view, point, meas, inlier_threshold, variant, device); if (rc) { delete p; return rc; } *out = p; }
  else return fail(BA_ERR_ARG, "unknown precision %d", precision);
  return BA_OK;
}
int ba_destroy(ba_handle* h) { delete h; return BA_OK; }
#define H_CHECK if (!h) return fail(BA_ERR_ARG, "null handle")
int ba_comm_unique_id(void* id128) {
  if (!g_nccl.load()) return fail(BA_ERR_NCCL, "cannot load libnccl.so.2");
  ncclUniqueId id;
  NK(g_nccl.GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  return BA_OK;
}
int ba_comm_init(ba_handle* h, int rank, int nranks, const void* id128) { H_CHECK; return h->comm_init(rank, nranks, id128); }
int ba_bandwidth(ba_handle* h, int* bw) { H_CHECK; *bw = h->bw; return BA_OK; }
int ba_set_bandwidth(ba_handle* h, int bw) { H_CHECK; return h->set_bandwidth(bw); }
int ba_set_state(ba_handle* h, const double* R, const double* T, const double* f, const double* k1, const double* k2, const double* X) { H_CHECK; return h->set_state(R, T, f, k1, k2, X); }
int ba_get_state(ba_handle* h, double* R, double* T, double* f, double* k1, double* k2, double* X) { H_CHECK; return h->get_state(R, T, f, k1, k2, X); }
int ba_eval(ba_handle* h, double* energy) { H_CHECK; return h->eval(energy); }
int ba_linearize(ba_handle* h, double* energy, double* a, double* b) { H_CHECK; return h->linearize(energy, a, b); }
int ba_compute(ba_handle* h, double lambda) { H_CHECK; return h->compute(lambda); }
int ba_solve_try(ba_handle* h, double* dx_norm, double* rho_den, double* energy_test) { H_CHECK; return h->solve_try(dx_norm, rho_den, energy_test); }
int ba_accept(ba_handle* h) { H_CHECK; return h->accept(); }
int ba_reject(ba_handle* h) { H_CHECK; return h->reject(); }
int ba_get_dx(ba_handle* h, double* dx) { H_CHECK; return h->get_dx(dx); }
int ba_step_streamed(ba_handle* h, const double* R, const double* T, const double* f, const double* k1, const double* k2, const double* X, double lambda,
                     double* dx, double* energy, double* dx_norm, double* rho_den, double* energy_test) {
  H_CHECK; return h->step_streamed(R, T, f, k1, k2, X, lambda, dx, energy, dx_norm, rho_den, energy_test);
}
int ba_split_plan(int n, int kd, int mode, int segments, int* out, int out_len) {
  if (!out || out_len < 24 || n <= 0 || kd <= 0) return BA_ERR_ARG;
  const SplitPlan pl = split_plan(n, kd, mode);
  const int nseg = std::max(1, std::min(4, segments));
  int v[24] = {pl.ok, pl.w, pl.s0, pl.p1, pl.npart[0], pl.npart[1], pl.q[0], pl.q[1], pl.nm[0], pl.nm[1], pl.ntm[0], pl.ntm[1], pl.npE[0], pl.npE[1], nseg};
  for (int j = 0; j <= nseg; ++j) v[15 + j] = pl.ok ? seg_bound(pl.q[0], j, nseg) : 0;
  for (int i = 0; i < 24; ++i) out[i] = v[i];
  return BA_OK;
}
int ba_error_statistics(ba_handle* h, double avg_focal_length, double inlier_threshold, double* sums) { H_CHECK; if (!sums) return BA_ERR_ARG; return h->error_statistics(avg_focal_length, inlier_threshold, sums); }
int ba_get_residuals(ba_handle* h, double* r) { H_CHECK; return h->get_residuals(r); }
int ba_get_reduced_system(ba_handle* h, double* S, double* g) { H_CHECK; return h->get_reduced(S, g); }
int ba_keep_reduced_system(ba_handle* h, int enable) { H_CHECK; h->keep_reduced = enable != 0; return BA_OK; }
int ba_get_jacobian(ba_handle* h, double* Jc, double* Jp) { H_CHECK; return h->get_jacobian(Jc, Jp); }
int ba_launch_count(ba_handle* h, long long* launches) { H_CHECK; *launches = h->launches; return BA_OK; }
int ba_stage_ms(ba_handle* h, double* s) { H_CHECK; for (int i = 0; i < 8; ++i) s[i] = h->stage_ms[i]; return BA_OK; }
int ba_set_profiling(ba_handle* h, int enable) { H_CHECK; h->profiling = enable != 0; return BA_OK; }
int ba_debug_counters(ba_handle* h, long long* out16) { H_CHECK; return h->debug_counters(out16, 16); }
int ba_debug_counters_n(ba_handle* h, long long* out, int count) { H_CHECK; return h->debug_counters(out, count); }
int ba_timer_start(ba_handle* h) { H_CHECK; return h->timer_start(); }
int ba_debug_band_solve(ba_handle* h, int n, int kd, const double* S, const double* g, double* y) { H_CHECK; return h->debug_band_solve(n, kd, S, g, y); }
int ba_timer_stop(ba_handle* h, double* ms) { H_CHECK; return h->timer_stop(ms); }
int ba_numeric_status(ba_handle* h, int* info) { H_CHECK; if (info) *info = h->numeric_info; return BA_OK; }
int ba_set_strict_numeric(ba_handle* h, int enable) { H_CHECK; h->strict_numeric = enable != 0; return BA_OK; }

}  // extern "C"
