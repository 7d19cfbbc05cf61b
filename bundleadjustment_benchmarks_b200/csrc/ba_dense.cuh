// Reduced camera block on sm_100a: blocked LDL^T of the symmetric (block-)banded matrix S and the
// triangular solves. Replaces the right-block solver of the reference,
//   SimplicialLDLT<JacobianType, Lower>::compute / solve      BacktrackLevMarqQRChol.h:339-341,
//                                                              BacktrackLevMarqCholesky.h:278-282
// (stock Eigen, NOT IN TREE). Un-pivoted LDL^T like SimplicialLDLT (D may be negative), natural
// camera order; the band (co-visibility window) plays the role of the sparsity pattern.
//
// Storage: lower band, entry (i,j), 0 <= i-j <= kd, at v[i*lds + j] (lds = kd, v = base + kd), i.e.
// LAPACK 'L' band storage viewed row-wise, so a 32x32 tile is an ordinary strided matrix.
// One cooperative persistent kernel: per 32-column panel every CTA refactors the 32x32 diagonal tile
// redundantly in shared memory (no broadcast needed), CTAs split the triangular solves of the row
// tiles, grid.sync, CTAs split the trailing tile updates, grid.sync.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

namespace ba {
namespace cg = cooperative_groups;

constexpr int NB = 32;
constexpr int DENSE_THREADS = 256;

template <class T> struct BandMat { T* v; size_t lds; int n; int kd; };

template <class T> __device__ __forceinline__ bool band_ok(const BandMat<T>& A, int i, int j) {
  return i < A.n && j <= i && (i - j) <= A.kd;
}

template <class T>
__global__ void __launch_bounds__(DENSE_THREADS) k_band_ldlt(BandMat<T> A, T* __restrict__ dvec, int* __restrict__ info) {
  cg::grid_group grid = cg::this_grid();
  __shared__ T sD[NB][NB + 1];
  __shared__ T sW[NB][NB + 1];
  __shared__ T sA[NB][NB + 1];
  __shared__ T sB[NB][NB + 1];
  __shared__ T sd[NB];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;  // 8 warps
  const int n = A.n, kd = A.kd;
  const int nt = (n + NB - 1) / NB;
  const int bt = (kd + NB - 1) / NB;  // row tiles below the diagonal tile that a panel can touch
  for (int k = 0; k < nt; ++k) {
    const int k0 = k * NB;
    for (int r = ty; r < NB; r += 8) {
      const int gi = k0 + r, gj = k0 + tx;
      sD[r][tx] = (gi < n) ? (band_ok(A, gi, gj) ? A.v[(size_t)gi * A.lds + gj] : T(0)) : (r == tx ? T(1) : T(0));
    }
    // un-blocked right-looking LDL^T of the tile; column scaling deferred
    for (int j = 0; j < NB; ++j) {
      __syncthreads();
      const T dj = sD[j][j];
      const T inv = T(1) / dj;
      for (int i = j + 1 + ty; i < NB; i += 8) {
        if (tx > j && tx <= i) sD[i][tx] -= sD[i][j] * sD[tx][j] * inv;
      }
    }
    __syncthreads();
    if (ty == 0) {
      const T d = sD[tx][tx];
      sd[tx] = d;
      if (blockIdx.x == 0 && k0 + tx < n && (d == T(0) || !(d == d))) atomicCAS(info, 0, k0 + tx + 1);
    }
    __syncthreads();
    for (int r = ty; r < NB; r += 8) if (tx < r) sD[r][tx] = sD[r][tx] / sd[tx];
    __syncthreads();
    // W = inverse of the unit lower L_kk; lane tx builds column tx by forward substitution
    if (ty == 0) {
      T x[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        T s = (i == tx) ? T(1) : T(0);
#pragma unroll
        for (int m = 0; m < i; ++m) s -= (m >= tx) ? sD[i][m] * x[m] : T(0);
        x[i] = (i >= tx) ? s : T(0);
        sW[i][tx] = x[i];
      }
    }
    __syncthreads();
    // triangular solves of the row tiles: L_ik = A_ik * L_kk^-T * D^-1
    const int last = min(nt - 1, k + bt);
    for (int it = k + 1 + blockIdx.x; it <= last; it += gridDim.x) {
      for (int r = ty; r < NB; r += 8) {
        const int gi = it * NB + r, gj = k0 + tx;
        sA[r][tx] = band_ok(A, gi, gj) ? A.v[(size_t)gi * A.lds + gj] : T(0);
      }
      __syncthreads();
      T acc[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll 8
      for (int m = 0; m < NB; ++m) {
        const T w = sW[tx][m];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += sA[ty * 4 + q][m] * w;
      }
      const T idc = T(1) / sd[tx];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = it * NB + ty * 4 + q, gj = k0 + tx;
        if (band_ok(A, gi, gj)) A.v[(size_t)gi * A.lds + gj] = acc[q] * idc;
      }
      __syncthreads();
    }
    grid.sync();
    if (blockIdx.x == 0) {  // write back the factored diagonal tile (nobody reads it again here)
      for (int r = ty; r < NB; r += 8) {
        const int gi = k0 + r, gj = k0 + tx;
        if (band_ok(A, gi, gj)) A.v[(size_t)gi * A.lds + gj] = (r == tx) ? sd[tx] : sD[r][tx];
      }
      if (ty == 0 && k0 + tx < n) dvec[k0 + tx] = sd[tx];
    }
    // trailing update A_ij -= L_ik D_k L_jk^T for k < j <= i <= last
    const int nb = last - k;
    const int npairs = nb * (nb + 1) / 2;
    for (int pidx = blockIdx.x; pidx < npairs; pidx += gridDim.x) {
      int ii = (int)((sqrtf(8.0f * (float)pidx + 1.0f) - 1.0f) * 0.5f);
      while ((ii + 1) * (ii + 2) / 2 <= pidx) ++ii;
      while (ii * (ii + 1) / 2 > pidx) --ii;
      const int jj = pidx - ii * (ii + 1) / 2;
      const int it = k + 1 + ii, jt = k + 1 + jj;
      if ((it - jt) * NB - (NB - 1) > kd) continue;  // tile entirely outside the band
      for (int r = ty; r < NB; r += 8) {
        const int gi = it * NB + r, gj2 = jt * NB + r, gc = k0 + tx;
        sA[r][tx] = band_ok(A, gi, gc) ? A.v[(size_t)gi * A.lds + gc] : T(0);
        sB[r][tx] = band_ok(A, gj2, gc) ? A.v[(size_t)gj2 * A.lds + gc] * sd[tx] : T(0);
      }
      __syncthreads();
      T acc[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll 8
      for (int m = 0; m < NB; ++m) {
        const T b = sB[tx][m];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += sA[ty * 4 + q][m] * b;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = it * NB + ty * 4 + q, gj = jt * NB + tx;
        if (band_ok(A, gi, gj)) A.v[(size_t)gi * A.lds + gj] -= acc[q];
      }
      __syncthreads();
    }
    grid.sync();
  }
}

// y = S^-1 g from the factor above: L z = g, w = D^-1 z, L^T y = w. Single CTA, 16 warps; the
// 32x32 diagonal systems are solved by warp 0 with shuffles, the band updates are split over warps.
constexpr int SOLVE_THREADS = 512;

template <class T>
__global__ void __launch_bounds__(SOLVE_THREADS) k_band_ldlt_solve(BandMat<T> A, const T* __restrict__ dvec, const T* __restrict__ g,
                                                                   T* __restrict__ y, T sign) {
  __shared__ T sL[NB][NB + 1];
  __shared__ T sb[NB];
  __shared__ T sacc[SOLVE_THREADS / 32][NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = SOLVE_THREADS / 32;
  const int n = A.n, kd = A.kd, nt = (n + NB - 1) / NB;
  for (int i = tid; i < n; i += SOLVE_THREADS) y[i] = g[i];
  __syncthreads();
  // forward
  for (int k = 0; k < nt; ++k) {
    const int k0 = k * NB;
    for (int r = warp; r < NB; r += nw) {
      const int gi = k0 + r, gj = k0 + lane;
      sL[r][lane] = (gj < gi && band_ok(A, gi, gj)) ? A.v[(size_t)gi * A.lds + gj] : T(0);
    }
    __syncthreads();
    if (warp == 0) {
      T b = (k0 + lane < n) ? y[k0 + lane] : T(0);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const T bj = __shfl_sync(0xffffffffu, b, j);
        if (lane > j) b -= sL[lane][j] * bj;
      }
      sb[lane] = b;
      if (k0 + lane < n) y[k0 + lane] = b;
    }
    __syncthreads();
    const int r1 = min(n - 1, k0 + NB - 1 + kd);
    for (int r = k0 + NB + warp; r <= r1; r += nw) {
      const int gj = k0 + lane;
      T v = band_ok(A, r, gj) ? A.v[(size_t)r * A.lds + gj] * sb[lane] : T(0);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if (lane == 0) y[r] -= v;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += SOLVE_THREADS) y[i] = y[i] / dvec[i];
  __syncthreads();
  // backward
  for (int k = nt - 1; k >= 0; --k) {
    const int k0 = k * NB;
    for (int r = warp; r < NB; r += nw) {
      const int gi = k0 + r, gj = k0 + lane;
      sL[r][lane] = (gj < gi && band_ok(A, gi, gj)) ? A.v[(size_t)gi * A.lds + gj] : T(0);
    }
    T acc = T(0);
    const int r1 = min(n - 1, k0 + NB - 1 + kd);
    for (int r = k0 + NB + warp; r <= r1; r += nw) {
      const int gj = k0 + lane;
      if (band_ok(A, r, gj)) acc += A.v[(size_t)r * A.lds + gj] * y[r];
    }
    sacc[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      T b = (k0 + lane < n) ? y[k0 + lane] : T(0);
      for (int w = 0; w < nw; ++w) b -= sacc[w][lane];
#pragma unroll
      for (int j = NB - 1; j >= 0; --j) {
        const T bj = __shfl_sync(0xffffffffu, b, j);
        if (lane < j) b -= sL[j][lane] * bj;
      }
      if (k0 + lane < n) y[k0 + lane] = b;
    }
    __syncthreads();
  }
  if (sign != T(1)) { for (int i = tid; i < n; i += SOLVE_THREADS) y[i] = sign * y[i]; }
}

// ---------------------------------------------------------------------------------------------
// Cluster-resident variant (v2): ONE thread-block cluster (16 CTAs, 8 portable) factors the band
// matrix AND solves. The per-panel work of a banded factorisation is small (<= bt(bt+1)/2 32^3 tile
// updates), so the chain of n/32 dependent panels is latency-bound; the design removes every
// avoidable latency from that chain (measured constants: tools/ubench, profiles/):
//   * diagonal tile: ONE warp per CTA (redundantly, so no broadcast is needed) eliminates the
//     32x32 tile in registers, lane = row; per step the pivot travels by warp shuffle (it feeds
//     the 71-cycle FP64 reciprocal, the critical chain) and the pivot column by one shared-memory
//     round trip (vector broadcast loads); no block barriers inside the 32 steps; the right-hand
//     side rides along as a 33rd column (forward substitution L z = g).
//   * row tiles below (L_ik = A_ik L_kk^-T D^-1): one warp per tile, lane = row, forward
//     substitution in registers against L_kk^T broadcast from shared memory; the rows are
//     prefetched into registers while the diagonal tile is being factored. W_k = L_kk^-1 (for the
//     backward pass) is the same substitution applied to identity rows.
//   * trailing update A_ij -= L_ik D_k L_jk^T: FP64 tensor cores (mma.sync m8n8k4 = DMMA), four
//     warps per 32x32 tile, operands staged global->shared with cp.async (double-buffered, no
//     registers), stride-36 rows -> conflict-free fragment loads, C fragments straight from/to L2.
//   * cluster barriers (~0.2 us) instead of grid-wide barriers (~4 us).
//   * backward pass on the whole cluster: per step the far tiles (i >= k+2) are reduced one step
//     ahead by all CTAs into CTA 0's shared memory through DSMEM; the critical chain (CTA 0, warp 0)
//     is two 32x32 mat-vecs: the (k+1,k) tile with y_{k+1}, then W_k^T.
// Code size matters: the steady-state loop must stay inside the instruction cache (straight-line
// code beyond ~128 KB runs 3x slower, tools/ubench), hence vector loads and shared helpers.
// ---------------------------------------------------------------------------------------------
// Masked tile elements are loaded from this zero word (address select instead of value select): a value
// select right after each load makes ptxas serialise load -> select -> load at the 310-cycle L2 latency.
__device__ double ba_zero_word[2];

constexpr int CL_THREADS = 256;  // 8 warps: 255 registers per thread keep the register-resident factor/substitution spill-free
constexpr int CL_WARPS = CL_THREADS / 32;
constexpr int CL_GROUPS = CL_THREADS / 128;
constexpr int TS = NB + 4;
constexpr int CL_MAX_BT = 144;  // far-tile slots of the backward pass alias the operand staging area

template <class T> struct VecOf;
template <> struct VecOf<double> { using V2 = double2; };
template <> struct VecOf<float> { using V2 = float2; };

template <class T> struct PanelSmem {
  alignas(16) T sLT[NB][NB];      // sLT[m][c] = L_kk[c][m] (0 for c <= m)
  alignas(16) T sCol[2][2 * NB];  // pivot column broadcast (tail zero: window positions past column 31)
  alignas(16) T sd[NB];
  alignas(16) T sinvd[NB];
  alignas(16) T sz[NB];
  int progress;                   // columns of L_kk published so far (factor warp -> substitution warps)
};

template <class T> struct ClusterSmem {
  PanelSmem<T> pn;                // L_kk^T, D, 1/D, z of the panel being factored (chain team)
  alignas(16) T sL[NB][NB + 2];   // staged diagonal tile
  alignas(16) T pad_[NB];         // reads past sLT's last row (dead window positions) stay inside the struct
  alignas(16) T sdU[2][NB];       // D of panel k (parity k&1) for the trailing updates, which lag the chain by one panel
  alignas(16) T gA[CL_GROUPS][2][2][NB][TS];  // [team][buffer][tile of the 2x2 block] row-operand tiles
  alignas(16) T gB[CL_GROUPS][2][2][NB][TS];  // column-operand tiles
  alignas(16) T sLk[2][NB][NB + 1];
  alignas(16) T sWk[2][NB][NB + 1];
  alignas(16) T sw[2][NB];
  int work[2];                    // cluster-wide block-op counters (CTA 0's copy is used), alternating by panel parity
  int grab[CL_GROUPS][2];
  long long tc[16];
};
static_assert(2 * CL_MAX_BT * NB <= 2 * 2 * 2 * CL_GROUPS * NB * TS, "backward slots must fit the staging area");

__device__ __forceinline__ void group_barrier(int group) { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(128) : "memory"); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Shared-memory vector load the compiler may not sink to its use: ptxas otherwise funnels a run of
// independent broadcast loads through one register quad and serialises load -> FMA -> load at the
// 30-cycle LDS latency (measured in tools/ubench2: 7.4k instead of 1.2k cycles per 32x32 substitution).
__device__ __forceinline__ double2 lds_v2(const double* p) {
  double2 v; const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_v2(const float* p) {
  float2 v; const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}

template <class T>
__device__ __forceinline__ T dep0(T x, T last) { return fma(last, T(0), x); }

// Reciprocal of a normal-range pivot: MUFU.RCP64H seed (~20 bits) + one cubic correction (3 dependent
// FMAs instead of the 5 + range checks of the IEEE division): relative error ~2^-60, not correctly rounded.
__device__ __forceinline__ double pivot_rcp(double d) {
  double x0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(d));
  const double e = fma(-d, x0, 1.0);
  const double p = fma(e, e, e);
  return fma(x0, p, x0);
}
__device__ __forceinline__ float pivot_rcp(float d) { return 1.0f / d; }

__device__ __forceinline__ void st_release_cta(int* p, int v) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta(const int* p) {
  int v; const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}

// ---- register-resident 32x32 kernels of the panel chain --------------------------------------------
// Both are ROLLED loops over blocks of 4 columns with the lane's row window rotated by 4 registers per
// block, so that all register indices are static while the code stays a few KB: the fully unrolled
// forms (45 KB + 32 KB) pushed the panel loop past the instruction cache and ran 9x slower inside the
// kernel than in isolation (tools/ubench2, profiles/r01_dense_notes.md). Window width 32 for the first
// 16 columns, 16 for the rest.

// 4 elimination steps of the LDL^T, lane = row, a[p] = A(row, jb + p). Critical chain per step: first
// update of the next column (9 cycles) -> pivot shuffle (30) -> reciprocal (~55) -> multiplier (9).
// The pivot column travels by one shared-memory round trip issued as ONE batch of broadcast vector loads
// (`dep0` ties the bulk FMAs to the last load, otherwise ptxas serialises load/FMA pairs through one
// register); the bulk of step j is issued after step j+1's shuffles (shfl.sync is a code-motion barrier).
template <class T, int W>
__device__ __forceinline__ void ldlt_block4(T (&a)[NB], T& z, T& dj, T& zj, const int lane, const int jb, PanelSmem<T>& sm) {
  using V2 = typename VecOf<T>::V2;
  constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int j = jb + s;
    const T* col = sm.sCol[s & 1];
    V2 t[W / 2];
#pragma unroll
    for (int h = 0; h < W / 2; ++h)
      if (2 * h + 1 > s) t[h] = lds_v2(col + jb + 2 * h);  // A(jb+2h, j), A(jb+2h+1, j), un-scaled
    const T inv = pivot_rcp(dj);
    const T l = a[s] * inv;
    if (lane > j) z -= l * zj;
    sm.sLT[j][lane] = (lane > j) ? l : T(0);
    if (lane == j) { sm.sd[j] = dj; sm.sinvd[j] = inv; sm.sz[j] = zj; }
    // column j+1 first, handed to step j+1 before the bulk of step j
    a[s + 1] -= l * (((s + 1) & 1) ? t[(s + 1) / 2].y : t[(s + 1) / 2].x);
    sm.sCol[(s + 1) & 1][lane] = a[s + 1];
    dj = __shfl_sync(FULL, a[s + 1], (j + 1) & 31);
    zj = __shfl_sync(FULL, z, (j + 1) & 31);
    __syncwarp();
    const T ld = dep0(l, t[W / 2 - 1].y);
#pragma unroll
    for (int h = 0; h < W / 2; ++h) {
      if (2 * h > s + 1) a[2 * h] -= ld * t[h].x;
      if (2 * h + 1 > s + 1) a[2 * h + 1] -= ld * t[h].y;
    }
  }
#pragma unroll
  for (int p = 0; p < NB; ++p) a[p] = (p + 4 < W) ? a[p + 4] : T(0);
}

// lane = row, a[c] = A(row, c) for c <= row (zero above). Results go to shared memory: sLT, sd, sinvd and
// sz = L^-1 (incoming sz).
template <class T>
__device__ __forceinline__ void warp_ldlt32(T (&a)[NB], T z, const int lane, PanelSmem<T>& sm) {
  constexpr unsigned FULL = 0xffffffffu;
  sm.sCol[0][lane] = a[0];
  T dj = __shfl_sync(FULL, a[0], 0);
  T zj = __shfl_sync(FULL, z, 0);
  __syncwarp();
  // after every block of 4 columns the substitution warps may consume them (they run one block behind)
#pragma unroll 1
  for (int jb = 0; jb < NB / 2; jb += 4) { ldlt_block4<T, NB>(a, z, dj, zj, lane, jb, sm); __syncwarp(); if (lane == 0) st_release_cta(&sm.progress, jb + 4); }
#pragma unroll 1
  for (int jb = NB / 2; jb < NB; jb += 4) { ldlt_block4<T, NB / 2>(a, z, dj, zj, lane, jb, sm); __syncwarp(); if (lane == 0) st_release_cta(&sm.progress, jb + 4); }
}

// 4 columns of the substitution X L^T = A, lane = row, a[p] = A(row, jb + p). x_m (scaled by 1/d_m when
// `scale`) is written to out[m * ostride] for m >= mmin; dot accumulates sum_m l_m z_m.
template <class T, int W>
__device__ __forceinline__ void trsm_block4(T (&a)[NB], const int jb, const PanelSmem<T>& sm, T& dot, T* __restrict__ out, const int ostride,
                                            const bool scale, const int mmin) {
  using V2 = typename VecOf<T>::V2;
  V2 t[2][W / 2];
#pragma unroll
  for (int h = 0; h < W / 2; ++h) t[0][h] = lds_v2(&sm.sLT[jb][jb + 2 * h]);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int m = jb + s;
    if (s + 1 < 4) {
#pragma unroll
      for (int h = 0; h < W / 2; ++h)
        if (2 * h + 1 > s + 1) t[(s + 1) & 1][h] = lds_v2(&sm.sLT[m + 1][jb + 2 * h]);
    }
    const T xm = dep0(a[s], t[s & 1][W / 2 - 1].y);
#pragma unroll
    for (int h = 0; h < W / 2; ++h) {
      if (2 * h > s) a[2 * h] -= xm * t[s & 1][h].x;
      if (2 * h + 1 > s) a[2 * h + 1] -= xm * t[s & 1][h].y;
    }
    const T l = scale ? xm * sm.sinvd[m] : xm;
    dot += l * sm.sz[m];
    if (m >= mmin) out[(size_t)m * ostride] = l;
  }
#pragma unroll
  for (int p = 0; p < NB; ++p) a[p] = (p + 4 < W) ? a[p + 4] : T(0);
}

template <class T>
__device__ __forceinline__ T warp_trsm32(T (&a)[NB], const PanelSmem<T>& sm, T* __restrict__ out, const int ostride, const bool scale, const int mmin) {
  T dot = T(0);
#pragma unroll 1
  for (int jb = 0; jb < NB / 2; jb += 4) { while (ld_acquire_cta(&sm.progress) < jb + 4) {} trsm_block4<T, NB>(a, jb, sm, dot, out, ostride, scale, mmin); }
#pragma unroll 1
  for (int jb = NB / 2; jb < NB; jb += 4) { while (ld_acquire_cta(&sm.progress) < jb + 4) {} trsm_block4<T, NB / 2>(a, jb, sm, dot, out, ostride, scale, mmin); }
  return dot;
}

// acc0/acc1 (fragment layout, even/odd k-steps) += sA * (-D sB)^T over the 32-wide panel. All 40 fragment
// loads are issued as one batch (volatile asm keeps them ahead of the DMMAs, which are volatile asm too);
// two accumulator sets keep 8 independent DMMA chains in flight per warp (16 issue cycles each).
__device__ __forceinline__ double lds_f64(const double* p) {
  double v; const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void tile_mma(const double (&sA)[NB][TS], const double (&sB)[NB][TS], const double* sd, int gt, double (&acc)[8]) {
  const int w4 = (gt >> 5) & 3, lane = gt & 31, lr = lane >> 2, lc = lane & 3;
  const int ra = (w4 >> 1) * 16 + lr, rb = (w4 & 1) * 16 + lr;
  double a0[NB / 4], a1[NB / 4], b0[NB / 4], b1[NB / 4], nd[NB / 4];
#pragma unroll
  for (int kk = 0; kk < NB / 4; ++kk) {
    a0[kk] = lds_f64(&sA[ra][kk * 4 + lc]); a1[kk] = lds_f64(&sA[ra + 8][kk * 4 + lc]);
    b0[kk] = lds_f64(&sB[rb][kk * 4 + lc]); b1[kk] = lds_f64(&sB[rb + 8][kk * 4 + lc]);
    nd[kk] = lds_f64(sd + kk * 4 + lc);
  }
  double accb[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) accb[e] = 0.0;
#pragma unroll
  for (int kk = 0; kk < NB / 4; ++kk) {
    const double m = -nd[kk];
    const double p0 = b0[kk] * m, p1 = b1[kk] * m;
    if (kk & 1) {
      dmma884(accb[0], accb[1], a0[kk], p0); dmma884(accb[2], accb[3], a0[kk], p1);
      dmma884(accb[4], accb[5], a1[kk], p0); dmma884(accb[6], accb[7], a1[kk], p1);
    } else {
      dmma884(acc[0], acc[1], a0[kk], p0); dmma884(acc[2], acc[3], a0[kk], p1);
      dmma884(acc[4], acc[5], a1[kk], p0); dmma884(acc[6], acc[7], a1[kk], p1);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] += accb[e];
}
__device__ __forceinline__ void tile_mma(const float (&sA)[NB][TS], const float (&sB)[NB][TS], const float* sd, int gt, float (&acc)[8]) {
  const int c = gt & 31, r0 = (gt >> 5) & 3;
#pragma unroll 8
  for (int m = 0; m < NB; ++m) {
    const float b = -sB[c][m] * sd[m];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += sA[r0 + 4 * q][m] * b;
  }
}

// One warp = one 32x32 tile of a 2x2 block: acc (16 m8n8 sub-tiles, e = 2 (4 mi + ni) + h) += sA (-D sB)^T.
// Per k-step 8 fragment loads feed 16 independent DMMAs (16 accumulator chains hide the DMMA latency).
__device__ __forceinline__ void block_mma(const double (&sA)[NB][TS], const double (&sB)[NB][TS], const double* sd, int lane, double (&acc)[32]) {
  const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int kk = 0; kk < NB / 4; ++kk) {
    const double m = -sd[kk * 4 + lc];
    double af[4], bf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { af[i] = sA[i * 8 + lr][kk * 4 + lc]; bf[i] = sB[i * 8 + lr][kk * 4 + lc] * m; }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) dmma884(acc[2 * (4 * mi + ni)], acc[2 * (4 * mi + ni) + 1], af[mi], bf[ni]);
  }
}
__device__ __forceinline__ void block_mma(const float (&sA)[NB][TS], const float (&sB)[NB][TS], const float* sd, int lane, float (&acc)[32]) {
  // float build: plain FMAs in the same fragment layout (row = 8 mi + lane/4, col = 8 ni + 2 (lane%4) + h)
  const int lr = lane >> 2, lc = lane & 3;
#pragma unroll 4
  for (int m = 0; m < NB; ++m) {
    const float d = -sd[m];
    float af[4], bf[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { af[i] = sA[i * 8 + lr][m]; bf[2 * i] = sB[i * 8 + 2 * lc][m] * d; bf[2 * i + 1] = sB[i * 8 + 2 * lc + 1][m] * d; }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) { acc[2 * (4 * mi + ni)] += af[mi] * bf[2 * ni]; acc[2 * (4 * mi + ni) + 1] += af[mi] * bf[2 * ni + 1]; }
  }
}

// element e (0..7) of a thread's C fragment -> tile-local (row, col). double: m8n8k4 accumulator
// layout of the warp's 16x16 quadrant (e = 4 mi + 2 ni + h); float: plain 4 x 32 thread grid.
template <class T>
__device__ __forceinline__ void frag_rc(int gt, int e, int& r, int& c) {
  if (sizeof(T) == 8) {
    const int w4 = (gt >> 5) & 3, lane = gt & 31;
    r = (w4 >> 1) * 16 + (e >> 2) * 8 + (lane >> 2);
    c = (w4 & 1) * 16 + ((e >> 1) & 1) * 8 + 2 * (lane & 3) + (e & 1);
  } else {
    r = ((gt >> 5) & 3) + 4 * e;
    c = gt & 31;
  }
}

// One band system per cluster (grid = nclusters x cluster size). The forward part factors panels [0, np_fwd) only
// (their updates still flow into the rows below: the Schur complement onto the remaining rows), the backward part
// starts at panel kb_bwd - 1 with y given for the rows of the tiles >= kb_bwd. A whole system is np_fwd = kb_bwd =
// number of tiles; the partial forms serve the two-sided factorisation (ba_gpu.cu).
template <class T> struct LdltProblem {
  BandMat<T> A; T* dvec; T* Wbuf; T* rhs; T* y; int* info; int np_fwd; int kb_bwd; int do_fwd; int do_bwd;
};
template <class T> struct LdltJob { LdltProblem<T> p[4]; T sign; };  // one problem per cluster of the launch

// Kernel. Two teams of 4 warps per CTA:
//   chain team (warps 0-3): per panel k stages the diagonal tile, warp 0 factors it (every CTA redundantly),
//     warps 1-3 solve the CTA's row tiles (and form W_k); then joins the update team.
//   update team (warps 4-7): 32x32 tile updates on the FP64 tensor cores.
// Per panel two phases separated by cluster barriers (look-ahead of one panel):
//   phase B(k): the column-(k+1) tiles receive panel k's update (all teams, <= 2 tile-ops per CTA);
//   phase A(k+1): chain(k+1) runs while the remaining tiles (columns >= k+2) receive panel k's update;
//     tile-ops of a phase are handed out per CTA through a shared-memory counter, so the chain team picks
//     up whatever is left when it is done.
template <class T>
__global__ void __launch_bounds__(CL_THREADS, 1) k_band_ldlt_cluster(const LdltJob<T> job, long long* __restrict__ dbg, int roww_arg) {
  using V2 = typename VecOf<T>::V2;
  cg::cluster_group cluster = cg::this_cluster();
  const LdltProblem<T>& P = job.p[blockIdx.x / cluster.num_blocks()];
  const BandMat<T> A = P.A;
  T* const dvec = P.dvec; T* const Wbuf = P.Wbuf; T* const rhs = P.rhs; T* const y = P.y; int* const info = P.info;
  const T sign = job.sign;
  const int np_fwd = P.np_fwd, kb_bwd = P.kb_bwd;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int EPC = 16 / (int)sizeof(T);   // elements per 16-byte chunk
  constexpr int CPR = NB / EPC;              // chunks per tile row
  constexpr int NCP = NB * CPR / 128;        // cp.async per thread per tile
  const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  extern __shared__ __align__(16) unsigned char cl_smem_raw[];
  ClusterSmem<T>& sm = *reinterpret_cast<ClusterSmem<T>*>(cl_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, team = warp >> 2, gt = tid & 127, tw = warp & 3;
  const int n = A.n, kd = A.kd, lds = (int)A.lds;
  T* const Av = A.v;
  const T* const zp = reinterpret_cast<const T*>(ba_zero_word);
  const int nt = (n + NB - 1) / NB, bt = (kd + NB - 1) / NB;
  long long t0 = clock64();
  if (tid < 16) sm.tc[tid] = 0;
  if (tid < 2 * NB) { sm.pn.sCol[0][NB + (tid & 31)] = T(0); sm.pn.sCol[1][NB + (tid & 31)] = T(0); }
  if (tid == 0) { sm.work[0] = 0; sm.work[1] = 0; }
  // chain CTAs (ranks < NC) run the panel chain with all 8 warps (warp 0 factors, warps 1..7 own the row tiles);
  // the other CTAs only update, so that the chain never shares an FP64 pipe with bulk DMMA work (measured:
  // the register factorisation took 6.4 us instead of 2.7 us per panel next to an update team, profiles/)
  // row tiles live on the six warps that do not share an SMSP with the factoring warp 0 (warps 1-3, 5-7);
  // warp 4 (same SMSP as warp 0) only forms W_k, after the factorisation
  const int ROWW = ((roww_arg & 0xff) == 3) ? 3 : CL_WARPS - 2;  // 3: warps 1-3 only (one row warp per SMSP, more chain CTAs)
  // W_k = L_kk^-1 is only needed by the backward pass: formed AFTER the factorisation (bit 8), so that its warp does not
  // share SMSP 0 with the factor warp while the pivot chain runs; bit 9: the row-tile substitutions start after the
  // factorisation too (experiment)
  const bool w_after = (roww_arg & 0x100) != 0, rows_after = (roww_arg & 0x200) != 0;
  const int NC = min(max(1, (ROWW == 3) ? C / 2 : C / 4), (bt + ROWW - 1) / ROWW);
  int* const work0 = cluster.map_shared_rank(&sm.work[0], 0);
#ifdef BA_DENSE_TICKS
#define TICK(i) { if (threadIdx.x == 128) { const long long t1_ = clock64(); sm.tc[i] += t1_ - t0; t0 = t1_; } __syncwarp(); }
#else
#define TICK(i) {}
#endif
#ifdef BA_DENSE_TICKS
  long long t2 = clock64();
#define TICKT(th, i) { if (threadIdx.x == (th)) { const long long t1_ = clock64(); sm.tc[i] += t1_ - t2; t2 = t1_; } __syncwarp(); }
#define TICKC(i) { if (threadIdx.x == 32) { const long long t1_ = clock64(); sm.tc[i] += t1_ - t2; t2 = t1_; } __syncwarp(); }
#else
#define TICKC(i) {}
#define TICKT(th, i) {}
#endif
  // per-thread constants of the tile-op path
  int aoff[NCP], soff[NCP];
#pragma unroll
  for (int q = 0; q < NCP; ++q) { const int id = gt + 128 * q, r = id / CPR, ch = id % CPR; aoff[q] = r * lds + ch * EPC; soff[q] = r * TS + ch * EPC; }
  int coff[4];
#pragma unroll
  for (int e = 0; e < 8; e += 2) { int r, c; frag_rc<T>(gt, e, r, c); coff[e >> 1] = r * lds + c; }

  // ---- tile (t, k) -> shared memory
  auto stage_operand = [&](T (&dst)[NB][TS], const int t, const int k0) {
    const int row0 = t * NB;
    const bool interior = (row0 + NB - 1 < n) && (row0 + NB - 1 - k0 <= kd);
    const T* tp = Av + (size_t)row0 * lds + k0;
    if (interior) {
#pragma unroll
      for (int q = 0; q < NCP; ++q) cp_async16(&dst[0][0] + soff[q], tp + aoff[q]);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int idx = gt + 128 * q, r = idx >> 5, c = idx & 31, gi = row0 + r;
        const bool ok = gi < n && gi - (k0 + c) <= kd;
        dst[r][c] = *(ok ? tp + r * lds + c : zp);
      }
    }
  };
  // ---- one panel of the chain (chain team only; team barriers): the column-k tiles first receive panel
  // k-1's update (the diagonal tile by the four warps together, every CTA redundantly; each row tile by the
  // warp that will solve it, one warp = one 32x32 DMMA tile product, while warp 0 factors), then
  // diagonal factor, row-tile substitutions, W_k, publication. No cluster-level phase for column k.
  auto stage_tile_warp = [&](T (&dst)[NB][TS], const int t, const int c0) {  // tile (t, c0/NB) by ONE warp
    const int row0 = t * NB;
    const bool interior = (row0 + NB - 1 < n) && (row0 + NB - 1 - c0 <= kd);
    const T* tp = Av + (size_t)row0 * lds + c0;
    if (interior) {
#pragma unroll
      for (int q = 0; q < NB * CPR / 32; ++q) { const int id = lane + 32 * q, r = id / CPR, ch = id % CPR; cp_async16(&dst[r][ch * EPC], tp + r * lds + ch * EPC); }
    } else {
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        const int gi = row0 + lane;
        const bool ok = gi < n && gi - (c0 + c) <= kd;
        dst[lane][c] = *(ok ? tp + lane * lds + c : zp);
      }
    }
  };
  auto chain = [&](const int k) {
    const int k0 = k * NB, last = min(nt - 1, k + bt);
    const bool upd = k > 0;
    const int kp0 = (k - 1) * NB;
    const T* sdp = sm.sdU[(k - 1) & 1];
    T (&sBop)[NB][TS] = sm.gB[0][0][0];            // L(k, k-1): column operand of every column-k update
    T (&wbuf)[NB][TS] = *(reinterpret_cast<T (*)[NB][TS]>(&sm.gA[0][0][0][0][0]) + warp);  // per-warp tile buffer (8 of the 8+8 slots)
    const int lr = lane >> 2, lc = lane & 3;
    T (*tiles)[NB][TS] = reinterpret_cast<T (*)[NB][TS]>(&sm.gA[0][0][0][0][0]);  // 16 tile slots (gA then gB); slot 8 = sBop
    auto stage_tile_cta = [&](T (&dst)[NB][TS], const int t, const int c0) {  // tile (t, c0/NB) by the whole CTA
      const int row0 = t * NB;
      const bool interior = (row0 + NB - 1 < n) && (row0 + NB - 1 - c0 <= kd);
      const T* tp = Av + (size_t)row0 * lds + c0;
      if (interior) {
#pragma unroll
        for (int q = 0; q < NB * CPR / CL_THREADS; ++q) { const int id = tid + CL_THREADS * q, r = id / CPR, ch = id % CPR; cp_async16(&dst[r][ch * EPC], tp + r * lds + ch * EPC); }
      } else {
#pragma unroll
        for (int q = 0; q < NB * NB / CL_THREADS; ++q) {
          const int idx = tid + CL_THREADS * q, r = idx >> 5, c = idx & 31, gi = row0 + r;
          const bool ok = gi < n && gi - (c0 + c) <= kd;
          dst[r][c] = *(ok ? tp + r * lds + c : zp);
        }
      }
    };
    // Column-k tiles of this CTA: item 0 = the diagonal tile (every chain CTA, redundantly), items 1..ROWW = the first
    // tiles of its row warps. They receive panel k-1's update as 4-warp quadrant DMMA products BEFORE the factorisation
    // starts (two teams, alternating items), so that no tensor work shares the SM with the pivot chain; results go to
    // shared memory: the diagonal tile to sL, row tile j to the tile buffer of its row warp.
    const bool w_warp = (warp == 4) && (rank == k % NC);
    const bool row_warp = (warp & 3) != 0 && (ROWW == 6 || warp < 4);
    const int ri = (warp < 4) ? warp - 1 : warp - 2;  // 0..5 over warps 1,2,3,5,6,7
    int it = row_warp ? k + 1 + rank + NC * ri : last + 1;
    bool pre = row_warp && (it <= last);
    TICKT(0, 15) TICKT(32, 14)
    if (tid == 0) sm.pn.progress = 0;
    if (upd) {
      stage_tile_cta(sBop, k, kp0);
      for (int j = 1; j <= ROWW; ++j) {
        const int tj = k + 1 + rank + NC * (j - 1);
        if (tj <= last && tj <= k - 1 + bt) stage_tile_cta(tiles[CL_WARPS + j], tj, kp0);
      }
    }
    cp_async_commit();
    T cc[4][8];  // C fragments (quadrant layout) of this team's items j = team, team + 2, ...
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = team + 2 * u;
      const int tj = (j == 0) ? k : k + 1 + rank + NC * (j - 1);
      const bool have = j <= ROWW && (j == 0 || tj <= last);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int r, c; frag_rc<T>(gt, e, r, c);
        const int gi = tj * NB + r, gj = k0 + c;
        const bool ok = have && gi < n && gj <= gi && gi - gj <= kd;
        cc[u][e] = *(ok ? Av + (size_t)gi * lds + gj : zp);
      }
    }
    const T zr = *((tid < NB && k0 + tid < n) ? rhs + k0 + tid : zp);
    cp_async_wait_all();
    __syncthreads();
    TICKT(0, 8)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = team + 2 * u;
      const int tj = (j == 0) ? k : k + 1 + rank + NC * (j - 1);
      if (j <= ROWW && (j == 0 || tj <= last)) {
        T dacc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dacc[e] = T(0);
        if (upd && (j == 0 || tj <= k - 1 + bt)) tile_mma(j == 0 ? sBop : tiles[CL_WARPS + j], sBop, sdp, gt, dacc);
        const int rw = (j - 1 < 3) ? j : j + 1;  // row warp of item j >= 1: ri = j-1 -> warps 1,2,3,5,6,7
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          int r, c; frag_rc<T>(gt, e, r, c);
          if (j == 0) sm.sL[r][c] = (k0 + r >= n && r == c) ? T(1) : cc[u][e] + dacc[e];
          else tiles[rw][r][c] = cc[u][e] + dacc[e];
        }
      }
    }
    if (tid < NB) sm.pn.sz[tid] = zr;
    __syncthreads();
    TICKT(0, 9) TICKT(32, 14)
    T a[NB];
    auto fetch_issue = [&](const int t, bool& use_mma) {  // later passes only: start fetching row tile (t, k)
      use_mma = upd && (t <= k - 1 + bt);
      if (use_mma) {
        stage_tile_warp(wbuf, t, kp0);
      } else {
        const int gi = t * NB + lane;
        const T* rp = Av + (size_t)gi * lds + k0;
#pragma unroll
        for (int c = 0; c < NB; ++c) { const bool ok = gi < n && gi - (k0 + c) <= kd; a[c] = *(ok ? rp + c : zp); }
      }
    };
    auto fetch_finish = [&](const int t) {  // operands landed (caller synchronised): tile product, then rows into registers
      T acc[32], prod[32];
      const T* cp = Av + (size_t)(t * NB) * lds + k0;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int r = (e >> 3) * 8 + lr, c = ((e >> 1) & 3) * 8 + 2 * lc + (e & 1), gi = t * NB + r;
        const bool ok = gi < n && gi - (k0 + c) <= kd;
        acc[e] = *(ok ? cp + r * lds + c : zp);
        prod[e] = T(0);
      }
      block_mma(wbuf, sBop, sdp, lane, prod);
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int r = (e >> 3) * 8 + lr, c = ((e >> 1) & 3) * 8 + 2 * lc + (e & 1);
        wbuf[r][c] = acc[e] + prod[e];
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < NB; ++c) a[c] = wbuf[lane][c];
      __syncwarp();
    };
    if (warp == 0) {
      const T z = sm.pn.sz[lane];
#pragma unroll
      for (int c = 0; c < NB; ++c) a[c] = (c <= lane) ? sm.sL[lane][c] : T(0);
      warp_ldlt32<T>(a, z, lane, sm.pn);
      __syncwarp();
      const T d = sm.pn.sd[lane];
      sm.sdU[k & 1][lane] = d;
      if (rank == 0 && k0 + lane < n && (d == T(0) || !(d == d))) atomicCAS(info, 0, k0 + lane + 1);
      TICKT(0, 10)
    } else if (pre) {
#pragma unroll
      for (int c = 0; c < NB; ++c) a[c] = wbuf[lane][c];
      __syncwarp();
    }
    // no block barrier here: the substitution warps follow the factorisation one 4-column block behind (sm.pn.progress)
    if (warp == 0 && rank == 0) {  // idle after the factorisation: publish the factored diagonal tile, D and w_k = D^-1 z_k
#pragma unroll 4
      for (int c = 0; c < NB; ++c) {
        const int gi = k0 + lane, gj = k0 + c;  // lane = row: sLT[c][lane] is conflict-free
        if (gi < n && gj <= gi && gi - gj <= kd) Av[(size_t)gi * lds + gj] = (c == lane) ? sm.pn.sd[c] : sm.pn.sLT[c][lane];
      }
      dvec[k0 + lane] = sm.pn.sd[lane];
      if (k0 + lane < n) rhs[k0 + lane] = sm.pn.sz[lane] * sm.pn.sinvd[lane];
    }
    if (warp >= 1) {
      bool do_w = w_warp;
      while (do_w || it <= last) {
        const int gi = it * NB + lane;
        T rold = T(0);
        if (do_w) {
#pragma unroll
          for (int c = 0; c < NB; ++c) a[c] = (c == lane) ? T(1) : T(0);
        } else {
          if (gi < n) rold = rhs[gi];
          if (!pre) {
            bool um = false;
            fetch_issue(it, um);
            cp_async_commit();
            cp_async_wait_all();
            __syncwarp();
            if (um) fetch_finish(it);
          }
        }
        pre = false;
        if ((do_w && w_after) || (!do_w && rows_after)) { while (ld_acquire_cta(&sm.pn.progress) < NB) {} }
        // one substitution call site (code size). W mode: identity rows, x[c] = W(c, lane), un-scaled, straight to
        // Wbuf. Tile mode: X L_kk^T = A_ik, L_ik = X D^-1 into the warp's tile buffer (row = lane), then written
        // out with coalesced 16-byte stores; g_i -= L_ik z_k.
        T* out = do_w ? (Wbuf + (size_t)k * NB * NB + lane) : &wbuf[lane][0];
        const int ostride = do_w ? NB : 1;
        const T s = warp_trsm32<T>(a, sm.pn, out, ostride, !do_w, 0);
        TICKT(32, 11)
        if (do_w) do_w = false;
        else {
          if (gi < n) rhs[gi] = rold - s;
          __syncwarp();
          T* tp = Av + (size_t)(it * NB) * lds + k0;
          const bool interior = (it * NB + NB - 1 < n) && (it * NB + NB - 1 - k0 <= kd);
          if (interior) {
#pragma unroll
            for (int q = 0; q < NB * CPR / 32; ++q) {
              const int id = lane + 32 * q, r = id / CPR, ch = id % CPR;
              *reinterpret_cast<V2*>(tp + r * lds + ch * EPC) = *reinterpret_cast<const V2*>(&wbuf[r][ch * EPC]);
              if (EPC == 4) *reinterpret_cast<V2*>(tp + r * lds + ch * EPC + 2) = *reinterpret_cast<const V2*>(&wbuf[r][ch * EPC + 2]);
            }
          } else {
            for (int c = 0; c < NB; ++c) {
              const int g2 = it * NB + lane;
              if (g2 < n && g2 - (k0 + c) <= kd) tp[lane * lds + c] = wbuf[lane][c];
            }
          }
          __syncwarp();
          it += NC * ROWW;
        }
      }
    }
    TICKT(32, 12)
    __syncthreads();  // everyone is done with pn and the tile buffers before they are reused
    TICKT(0, 13) TICKT(32, 14)
  };

  // ---- the remaining tiles (columns >= k+2) receive panel k's update in 2x2 blocks of tiles: block (I, J),
  // J <= I, covers tile rows k+2+2I(+1) and tile columns k+2+2J(+1); each warp of the team owns one of the
  // four 32x32 tiles (skipped when above the diagonal, past the last row tile, or outside the band); the four
  // operand tiles are staged once per block with cp.async (double-buffered). Items p = rank + C*t.
  auto block_phase = [&](const int k) {
    const int k0 = k * NB, last = min(nt - 1, k + bt), nb = last - k;
    const int nbk = nb / 2;  // ceil((nb - 1) / 2) block rows
    const int count = nbk * (nbk + 1) / 2;
    const T* sd = sm.sdU[k & 1];
    auto decode = [&](int p, int& bi, int& bj) {
      int ii = (int)((__fsqrt_rn(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
      while ((ii + 1) * (ii + 2) / 2 <= p) ++ii;
      while (ii * (ii + 1) / 2 > p) --ii;
      bi = ii; bj = p - ii * (ii + 1) / 2;
    };
    auto stage_block = [&](const int b, const int bi, const int bj) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ta = k + 2 + 2 * bi + h, tb = k + 2 + 2 * bj + h;
        if (ta <= last) stage_operand(sm.gA[team][b][h], ta, k0);
        if (tb <= last) stage_operand(sm.gB[team][b][h], tb, k0);
      }
      cp_async_commit();
    };
    int slot = 0, buf = 0, bi = 0, bj = 0;
    int* const wk = work0 + (k & 1);
    // the counter lives in CTA 0 (DSMEM atomics, ~0.3 us round trip): items are fetched TWO ahead so that the
    // round trip never sits between two block-ops
    if (gt == 0) { sm.grab[team][0] = atomicAdd(wk, 1); sm.grab[team][1] = atomicAdd(wk, 1); }
    group_barrier(team);
    int p = sm.grab[team][0];
    int pn = sm.grab[team][1];
    if (p < count) { decode(p, bi, bj); stage_block(0, bi, bj); }
    const int lr = lane >> 2, lc = lane & 3;
    while (p < count) {
      const int ci = k + 2 + 2 * bi + (tw >> 1), cj = k + 2 + 2 * bj + (tw & 1);  // this warp's tile
      int gnext = 0;
      if (gt == 0 && pn < count) gnext = atomicAdd(wk, 1);  // item after next (in flight during this block-op)
      // C fragments straight into the accumulators (in flight while the operands land)
      const bool active = ci <= last && cj <= ci && (ci - cj) * NB - (NB - 1) <= kd;
      const bool interior = active && (ci != cj) && (ci * NB + NB - 1 < n) && (ci * NB + NB - 1 - cj * NB <= kd);
      T* cp = Av + (size_t)(ci * NB) * lds + cj * NB;
      T cc[32];  // C fragments: loaded here, added after the DMMAs (their latency hides behind the tile product)
      if (interior) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const V2 v = *reinterpret_cast<const V2*>(cp + (mi * 8 + lr) * lds + ni * 8 + 2 * lc);
            cc[2 * (4 * mi + ni)] = v.x; cc[2 * (4 * mi + ni) + 1] = v.y;
          }
      } else if (active) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int r = (e >> 3) * 8 + lr, c = ((e >> 1) & 3) * 8 + 2 * lc + (e & 1), gi = ci * NB + r, gj = cj * NB + c;
          const bool ok = gi < n && gj <= gi && gi - gj <= kd;
          cc[e] = *(ok ? cp + r * lds + c : zp);
        }
      }
      cp_async_wait_all();
      group_barrier(team);  // operands landed; next item visible; everyone is done with the other buffer
      p = pn;
      if (p < count) { decode(p, bi, bj); stage_block(buf ^ 1, bi, bj); }
      if (active) {
        T acc[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = T(0);
        block_mma(sm.gA[team][buf][tw >> 1], sm.gB[team][buf][tw & 1], sd, lane, acc);
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] += cc[e];
        if (interior) {
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
              V2 v; v.x = acc[2 * (4 * mi + ni)]; v.y = acc[2 * (4 * mi + ni) + 1];
              *reinterpret_cast<V2*>(cp + (mi * 8 + lr) * lds + ni * 8 + 2 * lc) = v;
            }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int r = (e >> 3) * 8 + lr, c = ((e >> 1) & 3) * 8 + 2 * lc + (e & 1), gi = ci * NB + r, gj = cj * NB + c;
            if (gi < n && gj <= gi && gi - gj <= kd) cp[r * lds + c] = acc[e];
          }
        }
      }
      // hand the prefetched item index to the team (read after the next iteration's barrier)
      slot ^= 1;
      if (gt == 0) sm.grab[team][slot] = (pn < count) ? gnext : count;
      group_barrier(team);
      pn = sm.grab[team][slot];
      buf ^= 1;
    }
  };

  // ---- partial factorisation only: the tiles (i, k+1) get panel k's update outside a chain (one warp per tile,
  // plain loads; runs once per problem)
  auto col_phase = [&](const int k) {
    const int k0 = k * NB, last = min(nt - 1, k + bt), nb = last - k;
    if (tid < NB) sm.sdU[k & 1][tid] = dvec[k0 + tid];
    __syncthreads();
    const T* sd = sm.sdU[k & 1];
    T (*tiles)[NB][TS] = reinterpret_cast<T (*)[NB][TS]>(&sm.gA[0][0][0][0][0]);  // 8 + 8 tile slots (gA, gB)
    T (&bufA)[NB][TS] = tiles[warp];
    T (&bufB)[NB][TS] = tiles[CL_WARPS + warp];
    const int lr = lane >> 2, lc = lane & 3;
    for (int p = rank * CL_WARPS + warp; p < nb; p += C * CL_WARPS) {
      const int ti = k + 1 + p, tj = k + 1;
      const int ga = ti * NB + lane, gb = tj * NB + lane;
      for (int c = 0; c < NB; ++c) {
        const bool oka = ga < n && ga - (k0 + c) <= kd, okb = gb < n && gb - (k0 + c) <= kd;
        bufA[lane][c] = *(oka ? Av + (size_t)ga * lds + k0 + c : zp);
        bufB[lane][c] = *(okb ? Av + (size_t)gb * lds + k0 + c : zp);
      }
      __syncwarp();
      T acc[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = T(0);
      block_mma(bufA, bufB, sd, lane, acc);
      T* cp = Av + (size_t)(ti * NB) * lds + tj * NB;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int r = (e >> 3) * 8 + lr, c = ((e >> 1) & 3) * 8 + 2 * lc + (e & 1), gi = ti * NB + r, gj = tj * NB + c;
        if (gi < n && gj <= gi && gi - gj <= kd) cp[r * lds + c] += acc[e];
      }
      __syncwarp();
    }
  };

  // ================= factorisation with one panel of look-ahead: per panel ONE cluster barrier.
  // Iteration k: chain(k+1) (which first gives column k+1 panel k's update) runs next to the update of
  // the columns >= k+2 with panel k; the chain team joins the update when it is done.
  __syncthreads();
  if (P.do_fwd == 2) {
    // ---------------------------------------------------------------- forward substitution with the EXISTING factor
    // (refinement solves of the QR variants: S is already factored, only a new right-hand side comes in): z = L^-1 g on the
    // panels [0, np_fwd), rhs <- D^-1 z there, g_i -= sum_k L(i,k) z_k on the rows below. Mirror image of the backward pass:
    // z_k = W_k (g_k - sum_{i<k} L(k,i) z_i); the far tiles (k, i), i <= k-2, are reduced one step ahead by all CTAs into
    // CTA 0 through DSMEM, the chain (CTA 0, warp 0) is two 32x32 mat-vecs: the tile (k, k-1) with z_{k-1}, then W_k.
    T* fslots = &sm.gA[0][0][0][0][0];                  // [2][bt][NB]
    T* fslots0 = cluster.map_shared_rank(fslots, 0);
    constexpr int FRB = 256;
    T* fring = &sm.gB[0][0][0][0][0];                   // z_k ring in CTA 0
    const T* fring0 = cluster.map_shared_rank(fring, 0);
    auto row_tile_times_z = [&](const int ti, const int tk) -> T {   // lane = row: sum_c L(32 ti + lane, 32 tk + c) z_tk[c]
      const int gi = ti * NB + lane;
      const T zv = fring0[(tk % FRB) * NB + lane];
      const T* tp = Av + (size_t)gi * lds + tk * NB;
      T v[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) { const bool ok = gi < n && gi - (tk * NB + c) <= kd; v[c] = *(ok ? tp + c : zp); }
      T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
#pragma unroll
      for (int c = 0; c < NB; c += 4) {
        a0 += v[c] * __shfl_sync(FULL, zv, c); a1 += v[c + 1] * __shfl_sync(FULL, zv, c + 1);
        a2 += v[c + 2] * __shfl_sync(FULL, zv, c + 2); a3 += v[c + 3] * __shfl_sync(FULL, zv, c + 3);
      }
      return (a0 + a1) + (a2 + a3);
    };
    auto fstage = [&](const int km) {                   // everything step km needs except z_{km-1}
      const int lo = max(0, km - bt);
      const int nfar = (km - 2) - lo + 1;               // tiles (km, i), i = km-2 .. lo
      if (warp >= 1 && warp <= 3) {
        for (int sI = rank + C * (warp - 1); sI < nfar; sI += C * 3)
          fslots0[((size_t)(km & 1) * bt + sI) * NB + lane] = row_tile_times_z(km, km - 2 - sI);
      }
      if (rank == 0 && warp >= 4) {
        const int p = km & 1, km0 = km * NB;
        T lv[NB * NB / 128], wv[NB * NB / 128];
#pragma unroll
        for (int q = 0; q < NB * NB / 128; ++q) {
          const int idx = (warp - 4) * 32 + lane + 128 * q, r = idx >> 5, c = idx & 31, gi = km0 + r, gj = km0 - NB + c;
          const bool ok = km > 0 && gi < n && gi - gj <= kd;
          lv[q] = *(ok ? (Av + (size_t)gi * lds + gj) : zp);
          wv[q] = Wbuf[(size_t)km * NB * NB + idx];
        }
        const T swv = (warp == 4) ? *((km0 + lane < n) ? rhs + km0 + lane : zp) : T(0);
#pragma unroll
        for (int q = 0; q < NB * NB / 128; ++q) {
          const int idx = (warp - 4) * 32 + lane + 128 * q, r = idx >> 5, c = idx & 31;
          sm.sLk[p][r][c] = lv[q];
          sm.sWk[p][r][c] = wv[q];
        }
        if (warp == 4) sm.sw[p][lane] = swv;
      }
    };
    if (np_fwd > 0) fstage(0);
    cluster.sync();
    T zprev = T(0);
    for (int k = 0; k < np_fwd; ++k) {
      if (rank == 0 && warp == 0) {
        const int p = k & 1, lo = max(0, k - bt), nfar = (k - 2) - lo + 1;
        T b0 = sm.sw[p][lane], b1 = T(0), b2 = T(0), b3 = T(0);
        for (int sI = 0; sI < nfar; ++sI) b0 -= fslots[((size_t)p * bt + sI) * NB + lane];
        if (k > 0) {
#pragma unroll 8
          for (int c = 0; c < NB; c += 4) {
            b0 -= sm.sLk[p][lane][c] * __shfl_sync(FULL, zprev, c);
            b1 -= sm.sLk[p][lane][c + 1] * __shfl_sync(FULL, zprev, c + 1);
            b2 -= sm.sLk[p][lane][c + 2] * __shfl_sync(FULL, zprev, c + 2);
            b3 -= sm.sLk[p][lane][c + 3] * __shfl_sync(FULL, zprev, c + 3);
          }
        }
        const T b = (b0 + b1) + (b2 + b3);
        T z0 = T(0), z1 = T(0), z2 = T(0), z3 = T(0);
#pragma unroll 8
        for (int m = 0; m < NB; m += 4) {
          z0 += sm.sWk[p][lane][m] * __shfl_sync(FULL, b, m);
          z1 += sm.sWk[p][lane][m + 1] * __shfl_sync(FULL, b, m + 1);
          z2 += sm.sWk[p][lane][m + 2] * __shfl_sync(FULL, b, m + 2);
          z3 += sm.sWk[p][lane][m + 3] * __shfl_sync(FULL, b, m + 3);
        }
        zprev = (k * NB + lane < n) ? ((z0 + z1) + (z2 + z3)) : T(0);
        fring[(k % FRB) * NB + lane] = zprev;
      } else {
        // D^-1 z of the block solved in the previous step goes to global memory now (a whole step to complete)
        if (rank == 0 && warp == 1 && k >= 1 && (k - 1) * NB + lane < n) rhs[(k - 1) * NB + lane] = fring[((k - 1) % FRB) * NB + lane] / dvec[(k - 1) * NB + lane];
        if (k + 1 < np_fwd) fstage(k + 1);
      }
      cluster.sync();
    }
    if (rank == 0 && warp == 1 && np_fwd > 0 && (np_fwd - 1) * NB + lane < n)
      rhs[(np_fwd - 1) * NB + lane] = fring[((np_fwd - 1) % FRB) * NB + lane] / dvec[(np_fwd - 1) * NB + lane];
    // rows below the eliminated panels: g_i -= sum_k L(i, k) z_k, one warp per row tile (fixed order over k)
    const int ilast = min(nt - 1, np_fwd - 1 + bt);
    for (int i = np_fwd + rank * CL_WARPS + warp; i <= ilast; i += C * CL_WARPS) {
      T acc = T(0);
      for (int kk = max(0, i - bt); kk < np_fwd; ++kk) acc += row_tile_times_z(i, kk);
      if (i * NB + lane < n) rhs[i * NB + lane] -= acc;
    }
    cluster.sync();
  }
  if (P.do_fwd == 1) {
    if (rank < NC && np_fwd > 0) chain(0);
    cluster.sync();
    for (int k = 0; k < np_fwd; ++k) {
      TICK(1)
      if (tid < NB) sm.sdU[k & 1][tid] = dvec[k * NB + tid];        // D_k (published by rank 0) for the updates
      if (rank == 0 && tid == 0) sm.work[(k + 1) & 1] = 0;          // next panel's counter (idle during this one)
      __syncthreads();
      if (rank < NC && k + 1 < np_fwd) chain(k + 1);
      TICK(2)
      block_phase(k);
      TICK(3)
      cluster.sync();
      TICK(4)
    }
    if (np_fwd < nt) {  // partial factorisation: the first remaining column still owes panel np_fwd-1's update
      col_phase(np_fwd - 1);
      cluster.sync();
    }
  }
  if (!P.do_bwd) return;
  // ------------------------------------------------------------------ backward pass, whole cluster
  T* slots = &sm.gA[0][0][0][0][0];                    // [2][bt][NB] far-tile partial sums, written through DSMEM
  T* slots0 = cluster.map_shared_rank(slots, 0);    // CTA 0's copy
  // y_k of the blocks solved in this launch lives in a ring in CTA 0's shared memory (read by the other CTAs through
  // DSMEM); it is written to global memory one step later by another warp, so that no cluster barrier ever waits for
  // the round trip of a global store (measured: 0.72 us of the 1.9 us per step)
  constexpr int RB = 256;                           // ring blocks (>= CL_MAX_BT + 2), in the idle column-operand area
  static_assert(RB * NB <= 2 * 2 * CL_GROUPS * NB * TS && RB >= CL_MAX_BT + 2, "y ring must fit the staging area");
  T* ring = &sm.gB[0][0][0][0][0];
  const T* ring0 = cluster.map_shared_rank(ring, 0);
  auto stage = [&](int km) {  // everything step km needs except y_{km+1}: executed one step ahead
    const int km0 = km * NB;
    const int nfar = min(nt - 1, km + bt) - (km + 2) + 1;
    if (warp >= 1 && warp <= 3) {
      for (int s = rank + C * (warp - 1); s < nfar; s += C * 3) {
        const int i = km + 2 + s, gj = km0 + lane;
        const T yv = (i < kb_bwd) ? ring0[(i % RB) * NB + lane] : sign * *((i * NB + lane < n) ? y + i * NB + lane : zp);  // y holds sign * solution
        const T* tp = Av + (size_t)(i * NB) * lds + gj;
        T v[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) {
          const int g0 = i * NB + r;
          const bool ok0 = g0 < n && gj < n && g0 - gj <= kd;
          v[r] = *(ok0 ? tp + r * lds : zp);
        }
        T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
#pragma unroll
        for (int r = 0; r < NB; r += 4) {
          acc0 += v[r] * __shfl_sync(FULL, yv, r);
          acc1 += v[r + 1] * __shfl_sync(FULL, yv, r + 1);
          acc2 += v[r + 2] * __shfl_sync(FULL, yv, r + 2);
          acc3 += v[r + 3] * __shfl_sync(FULL, yv, r + 3);
        }
        slots0[((size_t)(km & 1) * bt + s) * NB + lane] = (acc0 + acc1) + (acc2 + acc3);  // the ring holds the solution itself (unsigned)
        // the tile this warp will need two steps from now (factored milliseconds ago, possibly evicted): pull it into L2
        if (km >= 2 && (i - 2) * NB + lane < n) {
          const T* pf = Av + (size_t)((i - 2) * NB + lane) * lds + (km0 - 2 * NB);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + NB - 1));
        }
      }
    }
    if (rank == 0 && warp >= 4) {
      const int p = km & 1;
      // all 16 loads of a thread in flight together (a rolled loop with the shared stores in between serialised
      // eight L2 round trips: 1.9 us per step, the longest leg of the whole backward step)
      T lv[NB * NB / 128], wv[NB * NB / 128];
#pragma unroll
      for (int q = 0; q < NB * NB / 128; ++q) {
        const int idx = (warp - 4) * 32 + lane + 128 * q, r = idx >> 5, c = idx & 31, gi = km0 + NB + r, gj = km0 + c;
        const bool ok = gi < n && gj < n && gi - gj <= kd;
        lv[q] = *(ok ? (Av + (size_t)gi * lds + gj) : zp);
        wv[q] = Wbuf[(size_t)km * NB * NB + idx];
      }
      const T swv = (warp == 4) ? *((km0 + lane < n) ? rhs + km0 + lane : zp) : T(0);
#pragma unroll
      for (int q = 0; q < NB * NB / 128; ++q) {
        const int idx = (warp - 4) * 32 + lane + 128 * q, r = idx >> 5, c = idx & 31;
        sm.sLk[p][r][c] = lv[q];
        sm.sWk[p][r][c] = wv[q];
      }
      if (warp == 4) sm.sw[p][lane] = swv;
    }
  };
  stage(kb_bwd - 1);
  cluster.sync();
  T yprev = (rank == 0 && warp == 0 && kb_bwd < nt && kb_bwd * NB + lane < n) ? sign * y[kb_bwd * NB + lane] : T(0);
  for (int k = kb_bwd - 1; k >= 0; --k) {
    if (rank == 0 && warp == 0) {
      const int p = k & 1, k0 = k * NB;
      const int nfar = min(nt - 1, k + bt) - (k + 2) + 1;
      T b0 = sm.sw[p][lane], b1 = T(0), b2 = T(0), b3 = T(0);
      int s = 0;
      for (; s + 4 <= nfar; s += 4) {
        b0 -= slots[((size_t)p * bt + s) * NB + lane]; b1 -= slots[((size_t)p * bt + s + 1) * NB + lane];
        b2 -= slots[((size_t)p * bt + s + 2) * NB + lane]; b3 -= slots[((size_t)p * bt + s + 3) * NB + lane];
      }
      for (; s < nfar; ++s) b0 -= slots[((size_t)p * bt + s) * NB + lane];
      if (k + 1 < nt) {
#pragma unroll 8
        for (int r = 0; r < NB; r += 4) {
          b0 -= sm.sLk[p][r][lane] * __shfl_sync(FULL, yprev, r);
          b1 -= sm.sLk[p][r + 1][lane] * __shfl_sync(FULL, yprev, r + 1);
          b2 -= sm.sLk[p][r + 2][lane] * __shfl_sync(FULL, yprev, r + 2);
          b3 -= sm.sLk[p][r + 3][lane] * __shfl_sync(FULL, yprev, r + 3);
        }
      }
      const T b = (b0 + b1) + (b2 + b3);
      T y0 = T(0), y1 = T(0), y2 = T(0), y3 = T(0);
#pragma unroll 8
      for (int m = 0; m < NB; m += 4) {
        y0 += sm.sWk[p][m][lane] * __shfl_sync(FULL, b, m);
        y1 += sm.sWk[p][m + 1][lane] * __shfl_sync(FULL, b, m + 1);
        y2 += sm.sWk[p][m + 2][lane] * __shfl_sync(FULL, b, m + 2);
        y3 += sm.sWk[p][m + 3][lane] * __shfl_sync(FULL, b, m + 3);
      }
      yprev = (k0 + lane < n) ? ((y0 + y1) + (y2 + y3)) : T(0);
      ring[(k % RB) * NB + lane] = yprev;
      TICKT(0, 8)
    } else {
      // the block solved in the previous step goes to global memory now: its store has a whole step to complete
      if (rank == 0 && warp == 1 && k + 1 < kb_bwd && (k + 1) * NB + lane < n) y[(k + 1) * NB + lane] = sign * ring[((k + 1) % RB) * NB + lane];
      if (k > 0) stage(k - 1);
      TICKT(32, 11)
    }
    cluster.sync();
    TICKT(0, 9) TICKT(32, 12)
  }
  if (rank == 0 && warp == 1 && kb_bwd > 0 && lane < n) y[lane] = sign * ring[lane];   // block 0
  TICK(5)
  if (dbg && rank == C - 1 && tid == 128) for (int i = 0; i < 8; ++i) dbg[i] = sm.tc[i];   // an update CTA
  if (dbg && rank == 0 && tid == 0) for (int i = 8; i < 16; ++i) dbg[i] = sm.tc[i];          // chain CTA 0: warp 0 / warp 1 ticks
#undef TICK
#undef TICKC
#undef TICKT
}

// ---- helpers of the two-sided factorisation (ba_gpu.cu): the bottom part of S is eliminated from the last row
// upwards, i.e. as an ordinary top-down factorisation of the index-reversed matrix S'(i', j') = S(n-1-j', n-1-i').
// rows0 = first row of S that S' covers (S' has np = n - rows0 rows); the block of the rows >= mrow0 of S' (the
// middle block, which both eliminations update) starts from zero and is added to S afterwards.
template <class T>
__global__ void k_band_reverse(BandMat<T> A, const T* __restrict__ g, T* __restrict__ Rv, T* __restrict__ gr, int np, int mrow0) {
  const int n = A.n, kd = A.kd;
  const size_t lds = A.lds, total = (size_t)np * (kd + 1), stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int ip = (int)(idx / (kd + 1)), d = (int)(idx - (size_t)ip * (kd + 1)), jp = ip - d;
    if (jp < 0) continue;
    T v = T(0);
    if (!(ip >= mrow0 && jp >= mrow0)) v = A.v[(size_t)(n - 1 - jp) * lds + (n - 1 - ip)];
    Rv[(size_t)ip * lds + jp] = v;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)np; i += stride) gr[i] = ((int)i < mrow0) ? g[n - 1 - i] : T(0);
}
// S(i, j) += S'(n-1-j, n-1-i), g(i) += g'(n-1-i) for the middle rows [r0, r0 + nm)
template <class T>
__global__ void k_band_combine(BandMat<T> A, T* __restrict__ g, const T* __restrict__ Rv, const T* __restrict__ gr, int r0, int nm) {
  const int n = A.n, kd = A.kd;
  const size_t lds = A.lds, total = (size_t)nm * (kd + 1), stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int i = r0 + (int)(idx / (kd + 1)), d = (int)(idx % (kd + 1)), j = i - d;
    if (j < r0) continue;
    A.v[(size_t)i * lds + j] += Rv[(size_t)(n - 1 - j) * lds + (n - 1 - i)];
  }
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < (size_t)nm; t += stride) g[r0 + t] += gr[n - 1 - (r0 + (int)t)];
}
// right-hand side only (solves with an existing two-sided factor): g'(i') = g(n-1-i') on the rows of the reversed system
// outside the middle block, 0 inside; and g(i) += g'(n-1-i) on the middle rows
template <class T>
__global__ void k_rhs_reverse(const T* __restrict__ g, T* __restrict__ gr, int n, int np, int mrow0) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) gr[i] = (i < mrow0) ? g[n - 1 - i] : T(0);
}
template <class T>
__global__ void k_rhs_combine(T* __restrict__ g, const T* __restrict__ gr, int n, int r0, int nm) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nm; t += gridDim.x * blockDim.x) g[r0 + t] += gr[n - 1 - (r0 + t)];
}
// dst[i] = src[n-1-i] for i in [i0, i1)  (dst and src indexed from their own origins: dst_off/src_off)
template <class T>
__global__ void k_flip_copy(T* __restrict__ dst, const T* __restrict__ src, int n, int i0, int i1) {
  for (int i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += gridDim.x * blockDim.x) dst[i] = src[n - 1 - i];
}

}  // namespace ba
