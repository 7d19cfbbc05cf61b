// Forward elimination of the band LDL^T, second generation: OWNER-COMPUTES on one 16-CTA cluster.
//
// Same arithmetic as k_band_ldlt_cluster's forward part (ba_dense.cuh; replaces SimplicialLDLT::compute and the
// forward substitution of ::solve, BacktrackLevMarqQRChol.h:339-341), different execution model. The first-generation
// kernel streams every trailing tile from L2, updates it and writes it back once per panel, and its chain CTAs do
// staging, column products, factorisation and substitution one after the other (15.4 us per 32-column panel,
// profiles/r01_dense_notes.md). Here
//   * every 32x32 tile (i, j) of the sliding window (k < j <= i <= k + bt) has ONE owner CTA, (i mod 4, j mod 4) on the
//     4 x 4 grid of the cluster, and stays in that CTA's shared memory, in DMMA accumulator order, from its first
//     update to its last: the trailing update never touches L2 for C; per panel a CTA stages only the <= 5 + 5
//     operand tiles L(i, k), i = r (mod 4), and L(j, k), j = c (mod 4), it needs (cp.async, XOR-swizzled rows ->
//     conflict-free fragment loads without padding);
//   * the triangular solves of a panel are GEMMs: L(i, k) = T(i, k) W_k^T D_k^-1 with W_k = L_kk^-1 (which the backward
//     pass needs anyway), one DMMA tile product per row tile, spread over the four CTAs that own column k;
//   * the pivot chain is one warp: the owner of the diagonal tile gives it panel k's update first (one DMMA product of
//     the lower blocks), factors it in registers (warp_ldlt32) while every other warp of the cluster applies panel
//     k to the rest of the window, and a second warp forms W_{k+1} four columns behind.
// One panel of look-ahead, ONE cluster barrier per panel: in iteration k the window receives panel k's update (column
// k+1 first) while the chain warp factors panel k+1; the four CTAs that own column k+1 then wait for the chain's flag
// (release/acquire through distributed shared memory) and solve their row tiles; cluster barrier.
// Only the forward part lives here; the middle block and the backward pass use k_band_ldlt_cluster. double only.
#pragma once
#include "ba_dense.cuh"

namespace ba {

constexpr int L2_MAX_BT = 18;        // row tiles below the diagonal a panel may touch (kd <= 576)
constexpr int L2_NSLOT = 15;         // resident C tiles per CTA (bt = 18: 5 + 4 + 3 + 2 + 1 over the five diagonals = r - c mod 4)
constexpr int L2_NOP = 5;            // operand tiles per role per CTA
constexpr int L2_UW = 5;             // update warps of the CTA that runs the chain: 1, 2, 3, 6, 7
constexpr int L2_UT = 32 * L2_UW;
constexpr int L2_TILE = NB * NB;

struct Ldlt2Smem {
  PanelSmem<double> pn;                          // L_kk^T (doubles as staging of the diagonal tile), pivot column, D, 1/D, z
  alignas(16) double slot[L2_NSLOT][L2_TILE];    // resident C tiles, accumulator order: pair p of lane l at [p * 64 + 2 l]
  alignas(16) double opA[L2_NOP][L2_TILE];       // row-operand tiles L(i, k) (swizzled rows); T phase: per-warp scratch
  alignas(16) double opB[L2_NOP][L2_TILE];       // column-operand tiles L(j, k); T phase: opB[0] = W_k
  alignas(16) double sdU[NB];                    // D_k
  alignas(16) double sinvdT[NB];
  alignas(16) double szT[NB];
  int sbase[8];                                  // first slot of diagonal class dq (d = d0 + 4 dq)
  int snb[8];                                    // slots of that class
  long long tc[16];                              // phase cycle counters (-DBA_L2_TICKS)
  int chain_done;                                // this CTA's chain warp published panel p: p + 1 (warp 0 -> warp 4)
  int col_ready;                                 // the owner of this CTA's helper column has finished its column tiles of panel p: p + 1 (written remotely by that owner)
  int prio_cnt;                                  // column tiles that have received their last update (running count, all panels)
  int panel_ready;                               // W_p, D_p, z_p of panel p are in global memory: p + 1 (written by the diagonal owner's warp 4 into the four CTAs that own column p)
};

// element (r, c) of a swizzled 32 x 32 tile: 16-byte chunk c/2 of row r sits at chunk (c/2) ^ (2 (r & 3))
__device__ __forceinline__ int l2_swz(int r, int c) { return r * NB + ((((c >> 1) ^ ((r & 3) << 1)) << 1) | (c & 1)); }
__device__ __forceinline__ void l2_st_release_cluster(int* p, int v) { asm volatile("st.release.cluster.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int l2_ld_acquire_cluster(const int* p) { int v; asm volatile("ld.acquire.cluster.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int l2_atom_add_acq_rel_cta(int* p, int v) {
  int old; const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("atom.acq_rel.cta.shared.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void l2_bar_update(const int nthr) { asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory"); }

// acc (accumulator order, e = 2 (4 mi + ni) + h <-> row 8 mi + lane/4, column 8 ni + 2 (lane%4) + h) += A (ms B)^T.
// MODE 0: all 16 blocks; 1: lower blocks only (ni <= mi: symmetric diagonal tile); 2: B lower triangular (W_k): k-steps
// beyond the block's last column contribute nothing.
template <int MODE>
__device__ __forceinline__ void l2_mma(const double* __restrict__ sA, const double* __restrict__ sB, const double (&ms)[NB / 4], const int lane,
                                       double (&acc)[32]) {
  const int lr = lane >> 2, lc = lane & 3, xr = (lr & 3) << 1;
  const double* pa = sA + lr * NB;
  const double* pb = sB + lr * NB;
#pragma unroll
  for (int kk = 0; kk < NB / 4; ++kk) {
    const int off = ((((kk << 1) | (lc >> 1)) ^ xr) << 1) | (lc & 1);
    double af[4], bf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { af[i] = pa[i * 8 * NB + off]; bf[i] = pb[i * 8 * NB + off] * ms[kk]; }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        if (MODE == 1 && ni > mi) continue;
        if (MODE == 2 && kk > 2 * ni + 1) continue;
        dmma884(acc[2 * (4 * mi + ni)], acc[2 * (4 * mi + ni) + 1], af[mi], bf[ni]);
      }
  }
}

__device__ __forceinline__ void l2_load_slot(const double* s, const int lane, double (&acc)[32]) {
#pragma unroll
  for (int p = 0; p < 16; ++p) { const double2 v = *reinterpret_cast<const double2*>(s + p * 64 + 2 * lane); acc[2 * p] = v.x; acc[2 * p + 1] = v.y; }
}
__device__ __forceinline__ void l2_store_slot(double* s, const int lane, const double (&acc)[32]) {
#pragma unroll
  for (int p = 0; p < 16; ++p) *reinterpret_cast<double2*>(s + p * 64 + 2 * lane) = make_double2(acc[2 * p], acc[2 * p + 1]);
}

// tile (ti, tj) of the band matrix <-> accumulator-order registers; entries outside the band / above the diagonal /
// past row n read as zero and are never written
__device__ __forceinline__ void l2_load_tile(const BandMat<double>& A, const int ti, const int tj, const int lane, double (&acc)[32]) {
  const int lr = lane >> 2, lc = lane & 3, row0 = ti * NB, col0 = tj * NB, lds = (int)A.lds;
  const double* tp = A.v + (size_t)row0 * lds + col0;
  const bool interior = (ti != tj) && (row0 + NB - 1 < A.n) && (row0 + NB - 1 - col0 <= A.kd);
  if (interior) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const double2 v = *reinterpret_cast<const double2*>(tp + (mi * 8 + lr) * lds + ni * 8 + 2 * lc);
        acc[2 * (4 * mi + ni)] = v.x; acc[2 * (4 * mi + ni) + 1] = v.y;
      }
  } else {
    const double* zp = ba_zero_word;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int r = (e >> 3) * 8 + lr, c = ((e >> 1) & 3) * 8 + 2 * lc + (e & 1), gi = row0 + r, gj = col0 + c;
      const bool ok = gi < A.n && gj <= gi && gi - gj <= A.kd;
      acc[e] = *(ok ? tp + r * lds + c : zp);
    }
  }
}
__device__ __forceinline__ void l2_store_tile(const BandMat<double>& A, const int ti, const int tj, const int lane, const double (&acc)[32]) {
  const int lr = lane >> 2, lc = lane & 3, row0 = ti * NB, col0 = tj * NB, lds = (int)A.lds;
  double* tp = A.v + (size_t)row0 * lds + col0;
  const bool interior = (ti != tj) && (row0 + NB - 1 < A.n) && (row0 + NB - 1 - col0 <= A.kd);
  if (interior) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
        *reinterpret_cast<double2*>(tp + (mi * 8 + lr) * lds + ni * 8 + 2 * lc) = make_double2(acc[2 * (4 * mi + ni)], acc[2 * (4 * mi + ni) + 1]);
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int r = (e >> 3) * 8 + lr, c = ((e >> 1) & 3) * 8 + 2 * lc + (e & 1), gi = row0 + r, gj = col0 + c;
      if (gi < A.n && gj <= gi && gi - gj <= A.kd) tp[r * lds + c] = acc[e];
    }
  }
}
// accumulator-order registers -> swizzled row-major tile in shared memory (one warp)
__device__ __forceinline__ void l2_acc_to_smem(double* s, const int lane, const double (&acc)[32]) {
  const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
      *reinterpret_cast<double2*>(s + l2_swz(mi * 8 + lr, ni * 8 + 2 * lc)) = make_double2(acc[2 * (4 * mi + ni)], acc[2 * (4 * mi + ni) + 1]);
}
// tile (t, column offset c0) of the band matrix -> swizzled shared-memory tile, by `nthr` threads (this one is `tix`).
// Tiles whose rows all exist travel as 16-byte cp.async (caller commits / waits), INCLUDING the positions outside the
// band (they alias other rows of the band storage): l2_fix_tile zeroes those afterwards, each thread the chunks it
// copied itself, so no barrier is needed in between. Tiles reaching past row n take masked loads + shared stores.
__device__ __forceinline__ void l2_stage_tile(double* dst, const BandMat<double>& A, const int t, const int c0, const int tix, const int nthr) {
  const int row0 = t * NB, lds = (int)A.lds;
  const double* tp = A.v + (size_t)row0 * lds + c0;
  if (row0 + NB - 1 < A.n && row0 >= c0 + NB) {
    for (int q = tix; q < NB * NB / 2; q += nthr) {
      const int r = q >> 4, ch = q & 15;
      cp_async16(dst + r * NB + ((ch ^ ((r & 3) << 1)) << 1), tp + r * lds + 2 * ch);
    }
  } else {
    const double* zp = ba_zero_word;
    for (int q = tix; q < NB * NB; q += nthr) {
      const int r = q >> 5, c = q & 31, gi = row0 + r, gj = c0 + c;
      const bool ok = gi < A.n && gj <= gi && gi - gj <= A.kd;
      dst[l2_swz(r, c)] = *(ok ? tp + r * lds + c : zp);
    }
  }
}
__device__ __forceinline__ void l2_fix_tile(double* dst, const BandMat<double>& A, const int t, const int c0, const int tix, const int nthr) {
  const int row0 = t * NB;
  if (row0 + NB - 1 < A.n && row0 >= c0 + NB && row0 + NB - 1 - c0 > A.kd) {
    for (int q = tix; q < NB * NB / 2; q += nthr) {
      const int r = q >> 4, ch = q & 15, dlt = row0 + r - (c0 + 2 * ch);
      double* e = dst + r * NB + ((ch ^ ((r & 3) << 1)) << 1);
      if (dlt > A.kd) e[0] = 0.0;
      if (dlt - 1 > A.kd) e[1] = 0.0;
    }
  }
}
// swizzled shared-memory tile -> tile (t, c0) of the band matrix with coalesced 16-byte stores (one warp)
__device__ __forceinline__ void l2_smem_to_tile(const double* s, const BandMat<double>& A, const int t, const int c0, const int lane) {
  const int row0 = t * NB, lds = (int)A.lds;
  double* tp = A.v + (size_t)row0 * lds + c0;
  const bool interior = (row0 + NB - 1 < A.n) && (row0 + NB - 1 - c0 <= A.kd) && (row0 >= c0 + NB);
#pragma unroll 4
  for (int q = lane; q < NB * NB / 2; q += 32) {
    const int r = q >> 4, ch = q & 15;
    const double2 v = *reinterpret_cast<const double2*>(s + r * NB + ((ch ^ ((r & 3) << 1)) << 1));
    if (interior) {
      *reinterpret_cast<double2*>(tp + r * lds + 2 * ch) = v;
    } else {
      const int gi = row0 + r, gj = c0 + 2 * ch;
      if (gi < A.n && gj <= gi && gi - gj <= A.kd) tp[r * lds + 2 * ch] = v.x;
      if (gi < A.n && gj + 1 <= gi && gi - gj - 1 <= A.kd) tp[r * lds + 2 * ch + 1] = v.y;
    }
  }
}

// ---- half tiles (rows 16 h .. 16 h + 15: accumulator blocks mi = 2 h, 2 h + 1) for the T phase, where the <= 5 row
// tiles of a CTA are spread over all eight warps
__device__ __forceinline__ void l2_load_slot_half(const double* s, const int lane, const int h, double (&acc)[16]) {
#pragma unroll
  for (int p = 0; p < 8; ++p) { const double2 v = *reinterpret_cast<const double2*>(s + (8 * h + p) * 64 + 2 * lane); acc[2 * p] = v.x; acc[2 * p + 1] = v.y; }
}
__device__ __forceinline__ void l2_load_tile_half(const BandMat<double>& A, const int ti, const int tj, const int lane, const int h, double (&acc)[16]) {
  const int lr = lane >> 2, lc = lane & 3, row0 = ti * NB, col0 = tj * NB, lds = (int)A.lds;
  const double* tp = A.v + (size_t)row0 * lds + col0;
  const double* zp = ba_zero_word;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int r = (2 * h + (e >> 3)) * 8 + lr, c = ((e >> 1) & 3) * 8 + 2 * lc + (e & 1), gi = row0 + r, gj = col0 + c;
    const bool ok = gi < A.n && gj <= gi && gi - gj <= A.kd;
    acc[e] = *(ok ? tp + r * lds + c : zp);
  }
}
__device__ __forceinline__ void l2_acc_to_smem_half(double* s, const int lane, const int h, const double (&acc)[16]) {
  const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int ml = 0; ml < 2; ++ml)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
      *reinterpret_cast<double2*>(s + l2_swz((2 * h + ml) * 8 + lr, ni * 8 + 2 * lc)) = make_double2(acc[2 * (4 * ml + ni)], acc[2 * (4 * ml + ni) + 1]);
}
// acc (two row blocks) += A_half W^T, W lower triangular
__device__ __forceinline__ void l2_mma_half_w(const double* __restrict__ sA, const double* __restrict__ sW, const int lane, const int h, double (&acc)[16]) {
  const int lr = lane >> 2, lc = lane & 3, xr = (lr & 3) << 1;
  const double* pa = sA + (16 * h + lr) * NB;
  const double* pb = sW + lr * NB;
#pragma unroll
  for (int kk = 0; kk < NB / 4; ++kk) {
    const int off = ((((kk << 1) | (lc >> 1)) ^ xr) << 1) | (lc & 1);
    double af[2], bf[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) af[i] = pa[i * 8 * NB + off];
#pragma unroll
    for (int i = 0; i < 4; ++i) bf[i] = pb[i * 8 * NB + off];
#pragma unroll
    for (int ml = 0; ml < 2; ++ml)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        if (kk > 2 * ni + 1) continue;
        dmma884(acc[2 * (4 * ml + ni)], acc[2 * (4 * ml + ni) + 1], af[ml], bf[ni]);
      }
  }
}
// rows 16 h .. 16 h + 15 of a swizzled shared-memory tile -> tile (t, c0) of the band matrix (one warp, 16-byte stores)
__device__ __forceinline__ void l2_smem_to_tile_half(const double* s, const BandMat<double>& A, const int t, const int c0, const int lane, const int h) {
  const int row0 = t * NB, lds = (int)A.lds;
  double* tp = A.v + (size_t)row0 * lds + c0;
  const bool interior = (row0 + NB - 1 < A.n) && (row0 + NB - 1 - c0 <= A.kd) && (row0 >= c0 + NB);
#pragma unroll 4
  for (int q = lane; q < NB * NB / 4; q += 32) {
    const int r = 16 * h + (q >> 4), ch = q & 15;
    const double2 v = *reinterpret_cast<const double2*>(s + r * NB + ((ch ^ ((r & 3) << 1)) << 1));
    if (interior) {
      *reinterpret_cast<double2*>(tp + r * lds + 2 * ch) = v;
    } else {
      const int gi = row0 + r, gj = c0 + 2 * ch;
      if (gi < A.n && gj <= gi && gi - gj <= A.kd) tp[r * lds + 2 * ch] = v.x;
      if (gi < A.n && gj + 1 <= gi && gi - gj - 1 <= A.kd) tp[r * lds + 2 * ch + 1] = v.y;
    }
  }
}

// warp_ldlt32 / warp_trsm32 with an epoch in the progress word: the substitution warp of panel k waits for
// base + columns, base = 64 (k + 1), so the word never has to be reset between panels
__device__ __forceinline__ void l2_warp_ldlt32(double (&a)[NB], double z, const int lane, PanelSmem<double>& sm, const int base) {
  constexpr unsigned FULL = 0xffffffffu;
  sm.sCol[0][lane] = a[0];
  double dj = __shfl_sync(FULL, a[0], 0);
  double zj = __shfl_sync(FULL, z, 0);
  __syncwarp();
#pragma unroll 1
  for (int jb = 0; jb < NB / 2; jb += 4) { ldlt_block4<double, NB>(a, z, dj, zj, lane, jb, sm); __syncwarp(); if (lane == 0) st_release_cta(&sm.progress, base + jb + 4); }
#pragma unroll 1
  for (int jb = NB / 2; jb < NB; jb += 4) { ldlt_block4<double, NB / 2>(a, z, dj, zj, lane, jb, sm); __syncwarp(); if (lane == 0) st_release_cta(&sm.progress, base + jb + 4); }
}
__device__ __forceinline__ void l2_warp_w32(double (&a)[NB], const PanelSmem<double>& sm, double* __restrict__ out, const int base) {
  double dot = 0.0;
#pragma unroll 1
  for (int jb = 0; jb < NB / 2; jb += 4) { while (ld_acquire_cta(&sm.progress) < base + jb + 4) {} trsm_block4<double, NB>(a, jb, sm, dot, out, NB, false, 0); }
#pragma unroll 1
  for (int jb = NB / 2; jb < NB; jb += 4) { while (ld_acquire_cta(&sm.progress) < base + jb + 4) {} trsm_block4<double, NB / 2>(a, jb, sm, dot, out, NB, false, 0); }
}

// One band system per cluster (grid = nclusters x 16): panels [0, np_fwd) of P.A are eliminated, their updates flow
// into the rows below (Schur complement on the remaining rows, partial forward substitution of P.rhs). On return the
// band storage holds L (strictly lower) and D for the eliminated panels and the updated remainder, dvec = D,
// Wbuf[k] = L_kk^-1, rhs = D^-1 L^-1 g on the eliminated rows (what the backward pass of k_band_ldlt_cluster expects).
// P.y is used as scratch for z_k = L_kk^-1 g_k on the eliminated rows.
#ifdef BA_L2_TICKS
#define L2TICK(i) { if (lane == 0) { const long long t1_ = clock64(); sm.tc[i] += t1_ - tprev; tprev = t1_; } __syncwarp(); }
// timeline of ONE iteration (k == BA_L2_EVK) of cluster 0: event e of warp w of CTA `rank`, cycles since that CTA left the barrier
#ifndef BA_L2_EVK
#define BA_L2_EVK 101
#endif
#define L2EV(e) { if (dbg && lane == 0 && k == BA_L2_EVK && blockIdx.x < 16 && (warp == 0 || warp == 1 || warp == 5)) dbg[16 + rank * 12 + (warp == 0 ? 0 : warp == 1 ? 4 : 8) + (e)] = clock64() - tev0; }
#else
#define L2TICK(i) {}
#define L2EV(e) {}
#endif
__global__ void __launch_bounds__(CL_THREADS, 1) k_band_ldlt_fwd2(const LdltJob<double> job, long long* __restrict__ dbg) {
  cg::cluster_group cluster = cg::this_cluster();
  const LdltProblem<double>& P = job.p[blockIdx.x / cluster.num_blocks()];
  const BandMat<double> A = P.A;
  double* const dvec = P.dvec; double* const Wbuf = P.Wbuf; double* const rhs = P.rhs; double* const zscr = P.y; int* const info = P.info;
  const int np_fwd = P.np_fwd;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char l2_smem_raw[];
  Ldlt2Smem& sm = *reinterpret_cast<Ldlt2Smem*>(l2_smem_raw);
  const int rank = (int)cluster.block_rank(), r = rank >> 2, c = rank & 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // In the CTA that runs the chain: warp 0 factors and has SMSP 0 to itself (warp 4 idles: the W-forming warp next to it
  // on the same SMSP made the factorisation take 4.9 us instead of ~3), warp 5 forms W_k, warps 1, 2, 3, 6, 7 update.
  const bool upd_warp = (warp & 3) != 0 && warp != 5;
  const int u = (warp < 4) ? warp - 1 : warp - 3;      // 0..4 over warps 1, 2, 3, 6, 7
  const int ut = u * 32 + lane;                         // thread index among the update warps
  const int n = A.n, kd = A.kd, lds = (int)A.lds;
  const int nt = (n + NB - 1) / NB, bt = (kd + NB - 1) / NB;
  const double* const zp = ba_zero_word;
  if (tid < 2 * NB) { sm.pn.sCol[0][NB + (tid & 31)] = 0.0; sm.pn.sCol[1][NB + (tid & 31)] = 0.0; }
  if (tid < 16) sm.tc[tid] = 0;
  if (tid == 0) {
    sm.pn.progress = 0; sm.chain_done = 0; sm.panel_ready = 0; sm.prio_cnt = 0; sm.col_ready = 0;
    const int d0 = (r - c + 4) & 3;
    int b = 0;
    for (int dq = 0; dq < 8; ++dq) {
      const int d = d0 + 4 * dq;
      const int cnt = (d < bt) ? (bt - d + 3) / 4 : 0;
      sm.sbase[dq] = b; sm.snb[dq] = cnt; b += cnt;
    }
  }
  __syncthreads();
#ifdef BA_L2_TICKS
  long long tprev = clock64();
#endif
  auto slot_of = [&](const int i, const int j) -> double* {
    const int dq = (i - j) >> 2;
    return sm.slot[sm.sbase[dq] + ((j >> 2) % sm.snb[dq])];
  };
  auto diag_owner = [](const int k) { return 5 * (k & 3); };
  // The chain of panel k+1 (iteration k) does NOT run on the owner of the diagonal tile: that CTA, like every owner of
  // column k+1, is at the peak of its load in iteration k (its first column is k+1: 15 tiles for the classes r - c = 0, 1).
  // It runs on the CTA whose first column is k+4, class r - c = 3: six tiles at bt = 18, the lightest of the cluster, which
  // also leaves slots 13 and 14 of that CTA free: the diagonal tile is pushed there through distributed shared memory by
  // its owner when it receives panel k-1's update, the operand L(k+1, k) is staged into slot 13.
  auto chain_cta = [](const int k) { return 4 * ((k + 3) & 3) + (k & 3); };
  constexpr int SLOT_DIAG = L2_NSLOT - 1, SLOT_OPND = L2_NSLOT - 2;

  // ---- factorisation of the diagonal tile of panel k by warp 0 (registers a[] = rows, z = right-hand side), publication
  auto factor_publish = [&](double (&a)[NB], const double z, const int k) {
    const int k0 = k * NB;
    l2_warp_ldlt32(a, z, lane, sm.pn, 64 * (k + 1));
    __syncwarp();
    const double d = sm.pn.sd[lane];
    if (k0 + lane < n && (d == 0.0 || !(d == d))) atomicCAS(info, 0, k0 + lane + 1);
#pragma unroll 4
    for (int cc = 0; cc < NB; ++cc) {
      const int gi = k0 + lane, gj = k0 + cc;
      if (gi < n && gj <= gi && gi - gj <= kd) A.v[(size_t)gi * lds + gj] = (cc == lane) ? sm.pn.sd[cc] : sm.pn.sLT[cc][lane];
    }
    dvec[k0 + lane] = d;
    if (k0 + lane < n) { rhs[k0 + lane] = sm.pn.sz[lane] * sm.pn.sinvd[lane]; zscr[k0 + lane] = sm.pn.sz[lane]; }
    __syncwarp();
    if (lane == 0) st_release_cta(&sm.chain_done, k + 1);
  };
  auto form_w = [&](const int k) {
    double a[NB];
#pragma unroll
    for (int cc = 0; cc < NB; ++cc) a[cc] = (cc == lane) ? 1.0 : 0.0;
    l2_warp_w32(a, sm.pn, Wbuf + (size_t)k * NB * NB + lane, 64 * (k + 1));
    __syncwarp();
    // W_k is out; once the chain warp has published D_k, z_k and the diagonal tile too, tell the owners of column k
    if (lane < 4) {
      while (ld_acquire_cta(&sm.chain_done) < k + 1) {}
      l2_st_release_cluster(cluster.map_shared_rank(&sm.panel_ready, 4 * lane + (k == 0 ? 0 : ((k - 1) & 3))), k + 1);
    }
    __syncwarp();
  };

  // ---- T(k): L(i, k) = T(i, k) W_k^T D_k^-1 on the CTAs that own column k, once the chain has published panel k (flag)
  // and the column's tiles have received their last update (prio_cnt >= prio_target). Every warp works on its own: unit
  // q = (row tile q / 2, half q % 2) goes to participating warp q % nw; the row operand comes straight from the tile's
  // slot (accumulator order, read in fragment order), W_k / D_k / z_k straight from L2, the result straight to the band
  // storage: no shared scratch, no CTA barrier, so the rest of the window update goes on around it.
  // slot of tile (i, j) in a CTA of diagonal class d0c (the table in shared memory is this CTA's own class)
  auto slot_index_cls = [&](const int d0c, const int i, const int j) -> int {
    const int dq = (i - j) >> 2;
    int b = 0;
    for (int q = 0; q < dq; ++q) { const int d = d0c + 4 * q; b += (d < bt) ? (bt - d + 3) / 4 : 0; }
    const int nb = (bt - (d0c + 4 * dq) + 3) / 4;
    return b + ((j >> 2) % nb);
  };
  auto solve_column = [&](const int k, const int wi, const int nw, const int owner, const bool wait_col) {
    const int k0 = k * NB, hi = min(k + bt, nt - 1);
    const int ia0 = (k + 1) + ((r - (k + 1)) & 3);
    const int ntile = (hi >= ia0) ? ((hi - ia0) >> 2) + 1 : 0;
    if (wi >= 2 * ntile) return;
    if (lane == 0) {
      while (l2_ld_acquire_cluster(&sm.panel_ready) < k + 1) {}
      if (wait_col) { while (l2_ld_acquire_cluster(&sm.col_ready) < k + 1) {} }
    }
    __syncwarp();
    (void)l2_ld_acquire_cluster(&sm.panel_ready);
    if (wait_col) (void)l2_ld_acquire_cluster(&sm.col_ready);
    const int d0o = (r - (k & 3)) & 3;                              // class of the column's owner (r, k mod 4)
    const int lr = lane >> 2, lc = lane & 3;
    const double* Wk = Wbuf + (size_t)k * NB * NB;
    // column operands, once per warp: W_k (lower triangular), 1 / D_k and z_k of this lane's eight columns
    double bf[4][NB / 4], invd[8], zk[8];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int kk = 0; kk < NB / 4; ++kk) bf[ni][kk] = (kk <= 2 * ni + 1) ? Wk[(ni * 8 + lr) * NB + kk * 4 + lc] : 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = (e >> 1) * 8 + 2 * lc + (e & 1);
      invd[e] = pivot_rcp(dvec[k0 + col]);
      zk[e] = *((k0 + col < n) ? zscr + k0 + col : zp);
    }
    for (int q = wi; q < 2 * ntile; q += nw) {
      const int t = q >> 1, h = q & 1, i = ia0 + 4 * t;
      double rold[2];
#pragma unroll
      for (int ml = 0; ml < 2; ++ml) { const int gi = i * NB + (2 * h + ml) * 8 + lr; rold[ml] = *((lc == 0 && gi < n) ? rhs + gi : zp); }
      // row operand in fragment order: A(8 (2h + ml) + lr, 4 kk + lc)
      double af[2][NB / 4];
      if (k > max(0, i - bt)) {                                   // received at least one update: lives in its owner's slot
        const double* sl = (owner < 0) ? slot_of(i, k) : cluster.map_shared_rank(&sm.slot[slot_index_cls(d0o, i, k)][0], owner);
#pragma unroll
        for (int ml = 0; ml < 2; ++ml)
#pragma unroll
          for (int kk = 0; kk < NB / 4; ++kk)
            af[ml][kk] = sl[(4 * (2 * h + ml) + (kk >> 1)) * 64 + (lr * 4 + (kk & 1) * 2 + (lc >> 1)) * 2 + (lc & 1)];
      } else {
        const double* tp = A.v + (size_t)(i * NB) * lds + k0;
#pragma unroll
        for (int ml = 0; ml < 2; ++ml)
#pragma unroll
          for (int kk = 0; kk < NB / 4; ++kk) {
            const int rr = (2 * h + ml) * 8 + lr, cc = kk * 4 + lc, gi = i * NB + rr, gj = k0 + cc;
            const bool ok = gi < n && gj <= gi && gi - gj <= kd;
            af[ml][kk] = *(ok ? tp + rr * lds + cc : zp);
          }
      }
      double x[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) x[e] = 0.0;
#pragma unroll
      for (int kk = 0; kk < NB / 4; ++kk)
#pragma unroll
        for (int ml = 0; ml < 2; ++ml)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            if (kk > 2 * ni + 1) continue;
            dmma884(x[2 * (4 * ml + ni)], x[2 * (4 * ml + ni) + 1], af[ml][kk], bf[ni][kk]);
          }
      double part[2] = {0.0, 0.0};
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int ce = ((e >> 1) & 3) * 2 + (e & 1);                // index into this lane's eight columns
        x[e] *= invd[ce];
        part[e >> 3] += x[e] * zk[ce];
      }
#pragma unroll
      for (int ml = 0; ml < 2; ++ml) {
        part[ml] += __shfl_xor_sync(FULL, part[ml], 1);
        part[ml] += __shfl_xor_sync(FULL, part[ml], 2);
        const int gi = i * NB + (2 * h + ml) * 8 + lr;
        if (lc == 0 && gi < n) rhs[gi] = rold[ml] - part[ml];
      }
      double* tp = A.v + (size_t)(i * NB) * lds + k0;
      const bool interior = (i * NB + NB - 1 < n) && (i * NB + NB - 1 - k0 <= kd);
#pragma unroll
      for (int ml = 0; ml < 2; ++ml)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int rr = (2 * h + ml) * 8 + lr, cc = ni * 8 + 2 * lc;
          if (interior) {
            *reinterpret_cast<double2*>(tp + rr * lds + cc) = make_double2(x[2 * (4 * ml + ni)], x[2 * (4 * ml + ni) + 1]);
          } else {
            const int gi = i * NB + rr, gj = k0 + cc;
            if (gi < n && gi - gj <= kd) tp[rr * lds + cc] = x[2 * (4 * ml + ni)];
            if (gi < n && gi - gj - 1 <= kd) tp[rr * lds + cc + 1] = x[2 * (4 * ml + ni) + 1];
          }
        }
    }
  };

  // ---- F(0), T(0)
  if (np_fwd > 0 && rank == diag_owner(0)) {
    if (warp == 0) {
      double a[NB];
      const int gi = lane;
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) {
        const bool ok = gi < n && cc <= gi && gi - cc <= kd;
        a[cc] = *(ok ? A.v + (size_t)gi * lds + cc : zp);
        if (gi >= n && cc == gi) a[cc] = 1.0;
      }
      const double z = *((gi < n) ? rhs + gi : zp);
      factor_publish(a, z, 0);
    } else if (warp == 5) {
      form_w(0);
    }
  }
  if (np_fwd > 0 && c == 0) solve_column(0, warp, CL_WARPS, -1, false);
  cluster.sync();

  int prio_target = 0;   // running number of column tiles this CTA has to finish before it may solve the column (same on every warp)
  for (int k = 0; k < np_fwd; ++k) {
    const int k0 = k * NB, hi = min(k + bt, nt - 1);
    const int ia0 = (k + 1) + ((r - (k + 1)) & 3);     // first tile row >= k+1 owned by this CTA row
    const int jb0 = (k + 1) + ((c - (k + 1)) & 3);     // first tile column >= k+1 owned by this CTA column
#ifdef BA_L2_TICKS
    const long long tev0 = clock64();
#endif
    // ======================= U(k): the window receives panel k's update; F(k+1) on the owner of the next diagonal tile
    const bool prio = (k + 1 < np_fwd);                 // the tile (k+1, k+1) goes to the chain warp
    const bool last = (k == np_fwd - 1);                // last panel: updated tiles also return to the band storage
    const bool chain_here = prio && rank == chain_cta(k);        // warps 0 and 4 of this CTA run the chain of panel k+1
    const int nuw = chain_here ? L2_UW : CL_WARPS;               // otherwise they update like everybody else
    const int uw = chain_here ? (upd_warp ? u : -1) : warp;
    // Column k+1: its owners (column class (k+1) mod 4, at the peak of their load) only finish its tiles; the solves T(k+1)
    // run on their row neighbours with column class k mod 4 (first column k+4: the lightest CTAs), which read the tiles
    // from the owner's slots through distributed shared memory once the owner's flag is up.
    const bool col_cta = prio && c == ((k + 1) & 3);
    const bool hlp_cta = prio && c == (k & 3);
    int ncol = 0;                                                 // column tiles of row class r that receive panel k's update
    if (prio) { for (int i = (k + 2) + ((r - (k + 2)) & 3); i <= hi; i += 4) ++ncol; }
    if (col_cta) prio_target += ncol;
    if (uw >= 0) {
      const int nthr = 32 * nuw, utx = 32 * uw + lane;
      for (int t = 0, i = ia0; i <= hi; ++t, i += 4) {
        if (prio && i == k + 1) continue;               // only the diagonal tile uses this row: staged by the chain warp
        l2_stage_tile(sm.opA[t], A, i, k0, utx, nthr);
      }
      for (int t = 0, j = jb0; j <= hi; ++t, j += 4) l2_stage_tile(sm.opB[t], A, j, k0, utx, nthr);
      cp_async_commit();
      if (utx < NB) sm.sdU[utx] = dvec[k0 + utx];
      // the tiles of the row that enters the window (first update from panel k) come from the band storage into their
      // slots while the operands are in flight
      if (k >= 1) {
        const int ie = k + bt;
        if (ie <= nt - 1 && ((ie - r) & 3) == 0) {
          int tt = 0;
          for (int j = jb0; j <= ie; j += 4, ++tt)
            if (tt % nuw == uw && !(prio && ie == k + 1 && j == k + 1)) { double acc[32]; l2_load_tile(A, ie, j, lane, acc); l2_store_slot(slot_of(ie, j), lane, acc); }
        }
      }
      cp_async_wait_all();
      for (int t = 0, i = ia0; i <= hi; ++t, i += 4) {
        if (prio && i == k + 1) continue;
        l2_fix_tile(sm.opA[t], A, i, k0, utx, nthr);
      }
      for (int t = 0, j = jb0; j <= hi; ++t, j += 4) l2_fix_tile(sm.opB[t], A, j, k0, utx, nthr);
      l2_bar_update(nthr);
      if (warp == 1) L2TICK(6)
      L2EV(0)
      double ms[NB / 4];
#pragma unroll
      for (int kk = 0; kk < NB / 4; ++kk) ms[kk] = -sm.sdU[kk * 4 + (lane & 3)];
      // Column k+1 comes first (jb0 = k+1 on its owners); every such tile bumps prio_cnt when its update is stored, and a
      // warp solves its share of the column (T(k+1)) as soon as the chain's flag is up, between two tile updates.
      bool t_pending = hlp_cta;
      const int owner = 4 * r + ((k + 1) & 3);
      int t = 0;
      for (int jt = 0, j = jb0; j <= hi; ++jt, j += 4) {
        const int i0 = j + ((r - j) & 3);
        for (int i = i0; i <= hi; i += 4) {
          if (prio && i == k + 1 && j == k + 1) continue;
          if (t_pending && l2_ld_acquire_cluster(&sm.panel_ready) >= k + 2 && (ncol == 0 || l2_ld_acquire_cluster(&sm.col_ready) >= k + 2)) {
            solve_column(k + 1, uw, nuw, owner, ncol > 0);
            t_pending = false;
          }
          if (t % nuw == uw) {
            double acc[32];
            double* s = slot_of(i, j);
            const bool first = (k == 0);
            if (first) l2_load_tile(A, i, j, lane, acc); else l2_load_slot(s, lane, acc);
            if (i == j) l2_mma<1>(sm.opA[(i - ia0) >> 2], sm.opB[jt], ms, lane, acc);
            else l2_mma<0>(sm.opA[(i - ia0) >> 2], sm.opB[jt], ms, lane, acc);
            if (last) l2_store_tile(A, i, j, lane, acc); else l2_store_slot(s, lane, acc);
            if (i == k + 2 && j == k + 2 && k + 2 < np_fwd)          // next iteration's chain works on it: hand it over
              l2_store_slot(cluster.map_shared_rank(&sm.slot[SLOT_DIAG][0], chain_cta(k + 1)), lane, acc);
            if (col_cta && j == k + 1) {                               // the last one tells the helper CTA
              __syncwarp();
              if (lane == 0 && l2_atom_add_acq_rel_cta(&sm.prio_cnt, 1) + 1 == prio_target)
                l2_st_release_cluster(cluster.map_shared_rank(&sm.col_ready, 4 * r + (k & 3)), k + 2);
            }
          }
          ++t;
        }
      }
      if (warp == 1) L2TICK(7)
      L2EV(1)
      if (t_pending) solve_column(k + 1, uw, nuw, owner, ncol > 0);
      L2EV(2)
      if (warp == 1) L2TICK(5)
    } else {   // chain_here: warps 0 and 4
      const int k1 = k + 1;
      if (warp == 0) {
        double* const opnd = &sm.slot[SLOT_OPND][0];
        l2_stage_tile(opnd, A, k1, k0, lane, 32);       // L(k+1, k): both operands of the diagonal update
        cp_async_commit();
        double acc[32], ms[NB / 4];
        if (k == max(0, k1 - bt)) l2_load_tile(A, k1, k1, lane, acc); else l2_load_slot(&sm.slot[SLOT_DIAG][0], lane, acc);
#pragma unroll
        for (int kk = 0; kk < NB / 4; ++kk) ms[kk] = -dvec[k0 + kk * 4 + (lane & 3)];
        const int gi = k1 * NB + lane;
        const double z = *((gi < n) ? rhs + gi : zp);
        cp_async_wait_all();
        l2_fix_tile(opnd, A, k1, k0, lane, 32);
        __syncwarp();
        L2TICK(0)
        L2EV(0)
        l2_mma<1>(opnd, opnd, ms, lane, acc);
        L2TICK(1)
        // accumulator order -> one row per lane through the (idle) L_kk^T buffer, rows rotated by their index
        double* stg = &sm.pn.sLT[0][0];
        {
          const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int rr = (e >> 3) * 8 + lr, cc = ((e >> 1) & 3) * 8 + 2 * lc + (e & 1);
            stg[rr * NB + ((cc + rr) & 31)] = acc[e];
          }
        }
        __syncwarp();
        double a[NB];
#pragma unroll
        for (int cc = 0; cc < NB; ++cc) {
          const double v = stg[lane * NB + ((cc + lane) & 31)];
          a[cc] = (cc <= lane) ? v : 0.0;
          if (gi >= n) a[cc] = (cc == lane) ? 1.0 : 0.0;
        }
        __syncwarp();
        L2TICK(2)
        L2EV(1)
        factor_publish(a, z, k1);
        L2TICK(3)
        L2EV(2)
      } else if (warp == 5) {
        form_w(k1);
        L2EV(0)
      }
    }
    if (warp == 1) L2TICK(8)
    cluster.sync();
    L2EV(3)
    if (warp == 1) L2TICK(9)
    if (warp == 0) L2TICK(12)
  }
#ifdef BA_L2_TICKS
  if (dbg && blockIdx.x == 0 && tid < 16) dbg[tid] = sm.tc[tid];
#endif
}
#undef L2TICK
#undef L2EV

}  // namespace ba
