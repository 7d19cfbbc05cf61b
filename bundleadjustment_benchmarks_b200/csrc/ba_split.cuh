// Separator split of the band LDL^T (double precision, cluster path): the chain of n dependent pivots of the reduced
// camera system S is cut in FOUR instead of two.
//
//   rows of S:  [ part 0 : 0 .. s0 ) [ separator : s0 .. s0 + w ) [ part 1 : s0 + w .. n ),   w >= kd + 1
//
// The two parts do not touch each other (their distance exceeds the bandwidth), so with the elimination order
// [part 0, part 1, separator] each part is an independent band system that the two-sided scheme of ba_gpu.cu factors
// with two chains (four chains run side by side), and the separator block receives the Schur complement
//
//   S_sep' = S_sep - E_0 D_0 E_0^T - E_1 D_1 E_1^T,      E_p = B_p L_p^-T D_p^-1   (the "spike" of part p),
//
// where B_p = S[separator, part p] is non-zero only in the kd columns of the part next to the separator. With part 0
// held index-reversed, either part sees the separator just above its row 0, its top-down chain is the one next to the
// separator, and the spike is non-zero only on the columns of that chain and of the part's middle block: E_p has the
// same recurrence as a row tile of L that never leaves the band,
//
//   E_k = ( B_k - sum_{k' = k-bt}^{k-1} E_k' D_k' L(k, k')^T ) W_k^T D_k^-1,         W_k = L_kk^-1,
//
// and its rows are independent: k_spike gives every strip of 8 separator rows to one CTA (no cluster, no global
// synchronisation) which walks the panels with the last bt strips of E in shared memory; L, W and D are read from the
// finished factor. Right-hand side: g_sep' = g_sep - sum_p E_p z_p (z = L^-1 g), and before the backward passes
// w_p -= E_p^T y_sep. Reference interface: this replaces the same Eigen::SimplicialLDLT / band solve as the one- and
// two-sided kernels (QRChol.h:197-206); the elimination order is the solver's own business.
#pragma once
#include <cstdlib>
#include "ba_dense.cuh"

namespace ba {

constexpr int SPK_PW = 9;             // product warps: far products t = 1 + w and t = 10 + w of a panel
constexpr int SPK_FW = 2;             // finisher warps, one per 8-row block of the strip
constexpr int SPK_WARPS = SPK_PW + SPK_FW;
constexpr int SPK_THREADS = 32 * SPK_WARPS;
constexpr int SPK_MI = 2;             // 8-row DMMA blocks per strip
constexpr int SPK_STRIP = 8 * SPK_MI; // separator rows per CTA
constexpr int SPK_MAX_BT = 18;        // ring of bt + 1 strips; 2 * SPK_PW >= bt - 1 far products per panel
constexpr int SPK_LD = 36;            // padded row of the finishers' T tile: conflict-free A-fragment loads
constexpr int SPK_TILES = SPK_PW + SPK_FW;

__device__ double g_spk_zero[4] = {0.0, 0.0, 0.0, 0.0};   // target of masked loads (keeps them unconditional and in flight together)

struct SpikeJob {
  BandMat<double> X;                  // the part's factored matrix (chain + middle panels), separator above its row 0
  const double* dvec; const double* W;  // D and W_k = L_kk^-1 (32 x 32 row-major per panel) of the same panels
  double* E; int ldE;                 // spike, w x ldE row-major; holds B in its first (bt + 1) * 32 columns on entry
  int k_begin, k_end;                 // panels to walk; the strips of the bt panels before k_begin are read back from E
};

// E := B (zero outside the coupling) on the first ncol columns. mode 0: part kept index-reversed, column j' is row
// s0 - 1 - j' of S and E(s, j') = S(s0 + s, s0 - 1 - j'); mode 1: column j is row p1 + j, E(s, j) = S(p1 + j, s0 + s).
__global__ void k_spike_init(BandMat<double> A, double* __restrict__ E, int ldE, int w, int ncol, int s0, int p1, int npart, int mode) {
  const size_t total = (size_t)w * ncol, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int s = (int)(idx / ncol), j = (int)(idx - (size_t)s * ncol);
    double v = 0.0;
    if (j < npart) {
      if (mode == 0) { const int gi = s0 + s, gj = s0 - 1 - j; if (gi - gj <= A.kd) v = A.v[(size_t)gi * A.lds + gj]; }
      else { const int gi = p1 + j, gj = s0 + s; if (gi - gj <= A.kd) v = A.v[(size_t)gi * A.lds + gj]; }
    }
    E[(size_t)s * ldE + j] = v;
  }
}

struct SpikeSmem {
  double tile[SPK_TILES][NB * NB];                // L tiles staged by cp.async, 16-byte chunks XOR-swizzled (spk_tile)
  double ring[SPK_MAX_BT + 1][SPK_STRIP * NB];    // last bt + 1 strips of E, columns XOR-swizzled (spk_ring)
  double part[SPK_PW][SPK_STRIP * NB];            // far partial sums of the panel being finished (spk_ring layout)
  double tb[SPK_STRIP][SPK_LD];
};
static_assert(sizeof(SpikeSmem) <= 227 * 1024, "spike kernel shared memory");

__device__ __forceinline__ int spk_tile(int r, int c) { return r * NB + ((((c >> 1) ^ ((r & 3) << 1)) << 1) | (c & 1)); }
__device__ __forceinline__ int spk_ring(int r, int c) { return r * NB + (c ^ ((r & 3) << 2)); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void spk_bar_arrive() { asm volatile("bar.arrive 1, %0;" ::"n"(SPK_THREADS) : "memory"); }
__device__ __forceinline__ void spk_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SPK_THREADS) : "memory"); }

// half h (rows 16h .. 16h + 15) of the tile (rows k0.., columns kp0..) of the factored band matrix into shared memory,
// one warp: 8 cp.async of 16 bytes per lane, every instruction covers two rows of 256 contiguous bytes
__device__ __forceinline__ void spike_stage(double* dst, const BandMat<double>& X, int k0, int kp0, int h, int lane) {
  const double* tp = X.v + (size_t)k0 * X.lds + kp0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int r = 16 * h + 2 * q + (lane >> 4), ch = lane & 15;
    cp_async16(dst + r * NB + ((ch ^ ((r & 3) << 1)) << 1), tp + (size_t)r * X.lds + 2 * ch);
  }
  cp_async_commit();
}
// rows past the part / entries outside the band are zero (the staged copy holds whatever lies there)
__device__ __forceinline__ void spike_fix(double* dst, const BandMat<double>& X, int k0, int kp0, int h, int lane) {
  if (k0 + NB - 1 < X.n && k0 + NB - 1 - kp0 <= X.kd) return;
  const int r = 16 * h + (lane >> 1), gi = k0 + r;
  for (int c = (lane & 1) * 16; c < (lane & 1) * 16 + 16; ++c) if (gi >= X.n || gi - (kp0 + c) > X.kd) dst[spk_tile(r, c)] = 0.0;
  __syncwarp();
}
// acc[mi][ni] += E_k'(rows 8 mi .., 32 cols) * diag(D_k') * tile(rows 8 ni ..)^T for ni in half h of the tile
template <int NMI>
__device__ __forceinline__ void spike_mma(const double* es, const double* tl, const double (&dk)[8], int h, int lr, int lc, double (&acc)[NMI][4][2]) {
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    double a[NMI];
#pragma unroll
    for (int mi = 0; mi < NMI; ++mi) a[mi] = es[spk_ring(mi * 8 + lr, kk * 4 + lc)] * dk[kk];
#pragma unroll
    for (int nh = 0; nh < 2; ++nh) {
      const int ni = 2 * h + nh;
      const double b = tl[spk_tile(ni * 8 + lr, kk * 4 + lc)];
#pragma unroll
      for (int mi = 0; mi < NMI; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b);
    }
  }
}
// one far product of a product warp: tile (kn, kp) in two halves; the halves of the next product (kp2 >= 0) are staged
// into the buffer as soon as this product has read them
__device__ __forceinline__ void spike_product(double* buf, const double* es, const BandMat<double>& X, const double* __restrict__ dvec,
                                              int kn, int kp, int kp2, int lane, double (&acc)[SPK_MI][4][2]) {
  const int lr = lane >> 2, lc = lane & 3;
  double dk[8];
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) dk[kk] = dvec[kp * NB + kk * 4 + lc];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    cp_async_wait_1();
    __syncwarp();
    spike_fix(buf, X, kn * NB, kp * NB, h, lane);
    spike_mma<SPK_MI>(es, buf, dk, h, lr, lc, acc);
    __syncwarp();
    if (kp2 >= 0) spike_stage(buf, X, kn * NB, kp2 * NB, h, lane); else cp_async_commit();   // empty group keeps the count
  }
}

// One CTA per strip of 16 separator rows. Software pipeline with a skew of one panel: in iteration k the product warps
// form the far products of panel k + 1 (those with E_k', k' <= k - 1, already in the ring) while the two finisher
// warps (8 rows each) finish panel k: near product with E_(k-1), fixed-order sum of the partials, W-GEMM, E_k into the
// ring and to global memory. The L tiles are staged row-wise with cp.async: fragment-wise 8-byte loads from global
// memory touch 8 cache lines per instruction and saturate the L1 tag stage.
__global__ void __launch_bounds__(SPK_THREADS, 1) k_spike(const SpikeJob j0, const SpikeJob j1, int w, long long* __restrict__ dbg) {
  extern __shared__ __align__(16) unsigned char spk_raw[];
  SpikeSmem& sm = *reinterpret_cast<SpikeSmem*>(spk_raw);
  const int nstrips = (w + SPK_STRIP - 1) / SPK_STRIP;
  const SpikeJob& J = (blockIdx.x < nstrips) ? j0 : j1;
  const int strip = (blockIdx.x < nstrips) ? blockIdx.x : blockIdx.x - nstrips;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const BandMat<double> X = J.X;
  const int n = X.n, bt = (X.kd + NB - 1) / NB, R = bt + 1;
  const int kb = J.k_begin, ke = J.k_end;
  // strips of the panels before k_begin (a resumed walk)
  for (int kp = max(0, kb - bt); kp < kb; ++kp)
    for (int i = tid; i < SPK_STRIP * NB; i += SPK_THREADS) {
      const int r = i >> 5, c = i & 31, srow = strip * SPK_STRIP + r;
      sm.ring[kp % R][spk_ring(r, c)] = (srow < w) ? J.E[(size_t)srow * J.ldE + kp * NB + c] : 0.0;
    }
  __syncthreads();
  double acc[SPK_MI][4][2];            // product warps: far partial sums of the next panel, carried across the barrier
#pragma unroll
  for (int mi = 0; mi < SPK_MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  for (int k = (kb > 0 ? kb - 1 : 0); k < ke; ++k) {
    const int k0 = k * NB;
    const bool finish = k >= kb;      // the first trip of a resumed walk only forms the far products of panel k_begin
#ifdef BA_SPK_TICKS
    long long tk[8]; tk[0] = clock64();
#define SPT(i) tk[i] = clock64();
#else
#define SPT(i)
#endif
    if (warp < SPK_PW) {
      double* pp = sm.part[warp];
#pragma unroll
      for (int mi = 0; mi < SPK_MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          pp[spk_ring(mi * 8 + lr, ni * 8 + 2 * lc)] = acc[mi][ni][0];
          pp[spk_ring(mi * 8 + lr, ni * 8 + 2 * lc + 1)] = acc[mi][ni][1];
          acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        }
      spk_bar_arrive();
      SPT(1)
      // far products of panel kn = k + 1: t = 1 .. min(kn, bt) - 1 with E_(kn-1-t) from the ring
      const int kn = k + 1, nprev = (kn < ke) ? min(kn, bt) : 0;
      const int ta = 1 + warp, tb2 = ta + SPK_PW;
      double* buf = sm.tile[warp];
      if (ta < nprev) {
        const int kpa = kn - 1 - ta, kpb = (tb2 < nprev) ? kn - 1 - tb2 : -1;
        spike_stage(buf, X, kn * NB, kpa * NB, 0, lane);
        spike_stage(buf, X, kn * NB, kpa * NB, 1, lane);
        spike_product(buf, sm.ring[kpa % R], X, J.dvec, kn, kpa, kpb, lane, acc);
        SPT(2)
        if (kpb >= 0) spike_product(buf, sm.ring[kpb % R], X, J.dvec, kn, kpb, -1, lane, acc);
        cp_async_wait_all();
        SPT(3)
      }
    } else {
      // finisher of rows 8 * fm .. of panel k
      const int fm = warp - SPK_PW;
      const int srow = strip * SPK_STRIP + fm * 8 + lr;
      const bool rowok = srow < w;
      double* const Erow = J.E + (size_t)(rowok ? srow : 0) * J.ldE;
      double* bufn = sm.tile[SPK_PW + fm];
      if (finish) {
        if (k > 0) { spike_stage(bufn, X, k0, k0 - NB, 0, lane); spike_stage(bufn, X, k0, k0 - NB, 1, lane); }
        double wv[4][8], dinv[4][2], t[4][2], dk[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) dk[kk] = (k > 0) ? J.dvec[k0 - NB + kk * 4 + lc] : 0.0;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) wv[ni][kk] = (kk <= 2 * ni + 1) ? J.W[(size_t)k * NB * NB + (ni * 8 + lr) * NB + kk * 4 + lc] : 0.0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = ni * 8 + 2 * lc + h;
            dinv[ni][h] = J.dvec[k0 + c];
            t[ni][h] = *((k <= bt && rowok) ? Erow + k0 + c : g_spk_zero);
          }
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) { dinv[ni][0] = 1.0 / dinv[ni][0]; dinv[ni][1] = 1.0 / dinv[ni][1]; }
        double an[1][4][2];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) an[0][ni][0] = an[0][ni][1] = 0.0;
        if (k > 0) {
          const double* es = sm.ring[(k - 1) % R] + fm * 8 * NB;     // spk_ring(8 fm + r, c) = 8 fm NB + spk_ring(r, c)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 0) cp_async_wait_1(); else cp_async_wait_all();
            __syncwarp();
            spike_fix(bufn, X, k0, k0 - NB, h, lane);
            spike_mma<1>(es, bufn, dk, h, lr, lc, an);
          }
        }
        SPT(1)
        spk_bar_sync();               // the product warps have written this panel's far partial sums
        SPT(2)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = ni * 8 + 2 * lc + h;
            double v = t[ni][h] - an[0][ni][h];
#pragma unroll
            for (int q = 0; q < SPK_PW; ++q) v -= sm.part[q][spk_ring(fm * 8 + lr, c)];
            sm.tb[fm * 8 + lr][c] = v;
          }
        }
        __syncwarp();
        SPT(3)
        double o[4][2], o2[4][2];     // two accumulator sets halve the dependent DMMA chain
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) o[ni][0] = o[ni][1] = o2[ni][0] = o2[ni][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const double a = sm.tb[fm * 8 + lr][kk * 4 + lc];
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) if (kk <= 2 * ni + 1) { if (kk & 1) dmma884(o2[ni][0], o2[ni][1], a, wv[ni][kk]); else dmma884(o[ni][0], o[ni][1], a, wv[ni][kk]); }
        }
        double* ed = sm.ring[k % R];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = ni * 8 + 2 * lc + h;
            const double v = (k0 + c < n) ? (o[ni][h] + o2[ni][h]) * dinv[ni][h] : 0.0;
            ed[spk_ring(fm * 8 + lr, c)] = v;
            if (rowok) Erow[k0 + c] = v;
          }
        }
        SPT(4)
      } else {
        spk_bar_sync();
      }
    }
    __syncthreads();
    SPT(5)
#ifdef BA_SPK_TICKS
    if (dbg && blockIdx.x == 5 && k == 60 && lane == 0) for (int i = 0; i < 6; ++i) dbg[128 + warp * 8 + i] = tk[i];
#endif
  }
#undef SPT
}

// Separator block: Sd (w x w, row stride ldw, lower triangle) = S[sep, sep] - sum_p E_p D_p E_p^T, accumulated over panel
// ranges (init = 1 on the first range, which starts from S[sep, sep]). One CTA per 32 x 32 tile (ta >= tb); its 8 warps
// split the panels of both spikes and are summed in a fixed order.
struct SyrkSide { const double* E; int ldE; const double* dvec; int kb, ke; };   // panels [kb, ke) of the spike
__global__ void __launch_bounds__(256) k_sep_syrk(BandMat<double> A, int s0, int w, double* __restrict__ Sd, int ldw, SyrkSide e0, SyrkSide e1, int init) {
  extern __shared__ __align__(16) unsigned char syrk_raw[];
  double (*red)[NB][NB + 1] = reinterpret_cast<double (*)[NB][NB + 1]>(syrk_raw);
  const int nts = (w + NB - 1) / NB;
  int ta = 0, rem = blockIdx.x;
  while (rem > ta) { rem -= ta + 1; ++ta; }
  const int tb = rem;
  if (ta >= nts) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  const int n0 = e0.ke - e0.kb, ntot = n0 + e1.ke - e1.kb;
  for (int pidx = warp; pidx < ntot; pidx += 8) {
    const SyrkSide& S = (pidx < n0) ? e0 : e1;
    const int kp = (pidx < n0) ? e0.kb + pidx : e1.kb + pidx - n0;
    const double* Ea = S.E + (size_t)kp * NB;
    double dk[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) dk[kk] = S.dvec[kp * NB + kk * 4 + lc];
#pragma unroll
    for (int half = 0; half < 2; ++half) {      // two passes of 4 k-steps keep the fragment registers bounded
      double a[4][4], b[4][4];
#pragma unroll
      for (int kq = 0; kq < 4; ++kq) {
        const int kk = half * 4 + kq;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          const int ra = ta * NB + mi * 8 + lr, rb = tb * NB + mi * 8 + lr;
          a[mi][kq] = *((ra < w) ? Ea + (size_t)ra * S.ldE + kk * 4 + lc : g_spk_zero);
          b[mi][kq] = *((rb < w) ? Ea + (size_t)rb * S.ldE + kk * 4 + lc : g_spk_zero) * dk[kk];
        }
      }
#pragma unroll
      for (int kq = 0; kq < 4; ++kq)
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi][kq], b[ni][kq]);
    }
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      red[warp][mi * 8 + lr][ni * 8 + 2 * lc] = acc[mi][ni][0];
      red[warp][mi * 8 + lr][ni * 8 + 2 * lc + 1] = acc[mi][ni][1];
    }
  __syncthreads();
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx >> 5, c = idx & 31, gi = ta * NB + r, gj = tb * NB + c;
    if (gi >= w || gj > gi) continue;
    double v = init ? ((gi - gj <= A.kd) ? A.v[(size_t)(s0 + gi) * A.lds + (s0 + gj)] : 0.0) : Sd[(size_t)gi * ldw + gj];
#pragma unroll
    for (int q = 0; q < 8; ++q) v -= red[q][r][c];
    Sd[(size_t)gi * ldw + gj] = v;
  }
}

constexpr size_t SYRK_SMEM = sizeof(double) * 8 * NB * (NB + 1);

// g_sep'(s) = g(s0 + s) - sum_p sum_c E_p(s, c) D_p(c) w_p(c): one CTA per separator row, fixed summation order
struct RhsSide { const double* E; int ldE; const double* dvec; const double* wv; int ncol; };
__global__ void __launch_bounds__(256) k_sep_rhs(const double* __restrict__ g, int s0, int w, double* __restrict__ gs, RhsSide e0, RhsSide e1) {
  __shared__ double red[8];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double a0 = 0.0, a1 = 0.0;
  const double* r0 = e0.E + (size_t)s * e0.ldE;
  const double* r1 = e1.E + (size_t)s * e1.ldE;
  for (int c = tid; c < e0.ncol; c += 512) {
    a0 += r0[c] * (e0.dvec[c] * e0.wv[c]);
    if (c + 256 < e0.ncol) a1 += r0[c + 256] * (e0.dvec[c + 256] * e0.wv[c + 256]);
  }
  for (int c = tid; c < e1.ncol; c += 512) {
    a0 += r1[c] * (e1.dvec[c] * e1.wv[c]);
    if (c + 256 < e1.ncol) a1 += r1[c + 256] * (e1.dvec[c + 256] * e1.wv[c + 256]);
  }
  double sum = a0 + a1;
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int q = 0; q < 8; ++q) t += red[q];
    gs[s] = g[s0 + s] - t;
  }
}

// w_p(c) -= sum_s E_p(s, c) * (sign * ysep(s)): the separator's solution enters the parts' backward passes. One CTA per
// 32 columns, warp g sums the rows s = g mod 8; fixed order.
struct CorrSide { const double* E; int ldE; int ncol; double* wv; };
__global__ void __launch_bounds__(256) k_spike_correct(CorrSide c0, CorrSide c1, int w, const double* __restrict__ ysep, double sign) {
  __shared__ double red[8][32];
  const CorrSide& S = blockIdx.y ? c1 : c0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x * 32 + lane;
  if (blockIdx.x * 32 >= S.ncol) return;
  double a0 = 0.0, a1 = 0.0;
  if (c < S.ncol) {
    int s = warp;
    for (; s + 8 < w; s += 16) {
      a0 += S.E[(size_t)s * S.ldE + c] * ysep[s];
      a1 += S.E[(size_t)(s + 8) * S.ldE + c] * ysep[s + 8];
    }
    if (s < w) a0 += S.E[(size_t)s * S.ldE + c] * ysep[s];
  }
  red[warp][lane] = a0 + a1;
  __syncthreads();
  if (warp == 0 && c < S.ncol) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][lane];
    S.wv[c] -= sign * t;
  }
}

// The three index-reversed copies of the split in one launch (k_band_reverse for each job; they are independent: the
// reversed half of the index-reversed part 0 is the top of part 0 as it stands).
template <class T> struct RevJob { BandMat<T> A; const T* g; T* Rv; T* gr; int np, mrow0, flip; };   // flip = 0: plain copy of the top np rows
// blockIdx.y = job; blockIdx.x walks 32 x 32 tiles (tile row bi, tile diagonal dj) of the output band. A flipped tile is read
// along the input's rows (coalesced) and transposed through shared memory; the plain copy needs no transpose.
template <class T>
__global__ void __launch_bounds__(256) k_band_reverse3(RevJob<T> j0, RevJob<T> j1, RevJob<T> j2) {
  __shared__ T tile[32][33];
  const RevJob<T>& J = blockIdx.y == 0 ? j0 : (blockIdx.y == 1 ? j1 : j2);
  const int n = J.A.n, kd = J.A.kd, np = J.np, mrow0 = J.mrow0;
  const size_t lds = J.A.lds;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ntr = (np + 31) / 32, ndj = (kd + 31) / 32 + 1;
  for (int t = blockIdx.x; t < ntr * ndj; t += gridDim.x) {
    const int bi = t / ndj, dj = t - bi * ndj, bj = bi - dj;
    if (bj < 0) continue;
    const int i0 = 32 * bi, c0 = 32 * bj;
    if (J.flip) {
      // tile[c][r] = A(n-1-(c0+c), n-1-(i0+r)): consecutive r -> consecutive (descending) input addresses
      for (int c = ty; c < 32; c += 8) {
        const int ip = i0 + tx, jp = c0 + c;
        const bool ok = ip < np && jp <= ip && ip - jp <= kd && !(ip >= mrow0 && jp >= mrow0);
        tile[c][tx] = ok ? J.A.v[(size_t)(n - 1 - jp) * lds + (n - 1 - ip)] : T(0);
      }
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        const int ip = i0 + r, jp = c0 + tx;
        if (ip < np && jp <= ip && ip - jp <= kd) J.Rv[(size_t)ip * lds + jp] = tile[tx][r];
      }
      __syncthreads();
    } else {
      for (int r = ty; r < 32; r += 8) {
        const int ip = i0 + r, jp = c0 + tx;
        if (ip < np && jp <= ip && ip - jp <= kd) J.Rv[(size_t)ip * lds + jp] = (ip >= mrow0 && jp >= mrow0) ? T(0) : J.A.v[(size_t)ip * lds + jp];
      }
    }
  }
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)np; i += stride)
    J.gr[i] = ((int)i < mrow0) ? (J.flip ? J.g[n - 1 - i] : J.g[i]) : T(0);
}

// both parts' middle blocks in one launch (k_band_combine / k_rhs_combine per job, blockIdx.y = part)
template <class T> struct CombJob { BandMat<T> A; T* g; const T* Rv; const T* gr; int r0, nm; };
template <class T>
__global__ void k_band_combine2(CombJob<T> j0, CombJob<T> j1, int rhs_only) {
  const CombJob<T>& J = blockIdx.y ? j1 : j0;
  const int n = J.A.n, kd = J.A.kd, r0 = J.r0, nm = J.nm;
  const size_t lds = J.A.lds, total = rhs_only ? 0 : (size_t)nm * (kd + 1), stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int i = r0 + (int)(idx / (kd + 1)), d = (int)(idx % (kd + 1)), j = i - d;
    if (j < r0) continue;
    J.A.v[(size_t)i * lds + j] += J.Rv[(size_t)(n - 1 - j) * lds + (n - 1 - i)];
  }
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < (size_t)nm; t += stride) J.g[r0 + t] += J.gr[n - 1 - (r0 + (int)t)];
}
// y'(i') = y(n-1-i') on the middle rows of both parts' reversed halves (before the chains' backward passes)
template <class T> struct FlipJob { T* dst; const T* src; int n, i0, i1; };
template <class T>
__global__ void k_flip_copy2(FlipJob<T> j0, FlipJob<T> j1) {
  const FlipJob<T>& J = blockIdx.y ? j1 : j0;
  for (int i = J.i0 + blockIdx.x * blockDim.x + threadIdx.x; i < J.i1; i += gridDim.x * blockDim.x) J.dst[i] = J.src[J.n - 1 - i];
}
// Solution of the whole system from the pieces: part 0 lives index-reversed (its bottom rows in the reversed half's y'), the
// separator in ys, the bottom rows of part 1 in its reversed half's y'; the top of part 1 is already in place.
template <class T>
__global__ void k_split_assemble(T* __restrict__ out, int n, int s0, int p1, const T* __restrict__ y0, const T* __restrict__ y20, int top0,
                                 const T* __restrict__ ys, const T* __restrict__ y21, int top1) {
  const int n0 = s0, n1 = n - p1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < n0) { const int j = n0 - 1 - i; out[i] = (j < top0) ? y0[j] : y20[i]; }
    else if (i < p1) out[i] = ys[i - s0];
    else { const int j = i - p1; if (j >= top1) out[i] = y21[n1 - 1 - j]; }
  }
}

// ---- host-side plan of the separator split (pure arithmetic: testable without a GPU through ba_split_plan)
// Rows the middle block of the two-sided scheme must keep: the last panel of either chain updates the rows up to kd below it,
// and those must not have been eliminated by the other chain, so n - 2 q NB >= kd + 1 (BA_LDLT_MID_PANELS=1 restores the earlier,
// tile-granular (bt + 2) * NB).
inline int mid_rows_min(int kd) {
  static const bool wide = std::getenv("BA_LDLT_MID_PANELS") != nullptr;
  const int bt = (kd + NB - 1) / NB;
  return wide ? (bt + 2) * NB : kd + 1;
}
// first panel of chain segment j of nseg. With three or more segments the last one is a quarter of the chain: its spike and SYRK
// run beside the middle blocks and must not outlast them; the others share the rest evenly.
inline int seg_bound(int q, int j, int nseg) {
  if (j <= 0) return 0;
  if (j >= nseg) return q;
  if (nseg < 3) return (int)((long long)q * j / nseg);
  const int head = q - q / 4;
  return (int)((long long)head * j / (nseg - 1));
}
// S = [part 0 : 0..s0) [separator : s0..p1) [part 1 : p1..n): separator of w >= kd + 1 rows (even, so that part 1 starts 16-byte
// aligned), parts as equal as possible; per part q panels per chain, a middle block of nm rows (ntm panels), nph rows in the
// reversed half, npE panels (ldE columns) of spike. mode: 0 off, 1 only when both chains are long enough to pay for the extra
// stages (middle blocks, spike, separator: about 60 panel times at bt = 18), 2 whenever possible.
struct SplitPlan { int ok, w, s0, p1, npart[2], q[2], r0[2], nm[2], nph[2], ntm[2], npE[2], ldE[2]; };
inline SplitPlan split_plan(int n, int kd, int mode) {
  SplitPlan pl = {};
  if (mode <= 0 || n <= 0 || kd <= 0) return pl;
  const int bt = (kd + NB - 1) / NB;
  int w = kd + 1; if (w & 1) ++w;
  if (n <= w) return pl;
  pl.w = w; pl.s0 = ((n - w) / 2) & ~1; pl.p1 = pl.s0 + w;
  pl.npart[0] = pl.s0; pl.npart[1] = n - pl.p1;
  for (int p = 0; p < 2; ++p) {
    pl.q[p] = (pl.npart[p] - mid_rows_min(kd)) / (2 * NB);
    if (pl.npart[p] < mid_rows_min(kd) || pl.q[p] < bt + 2 || (mode < 2 && pl.q[p] < 2 * bt + 8)) return pl;
    pl.r0[p] = pl.q[p] * NB; pl.nm[p] = pl.npart[p] - 2 * pl.r0[p]; pl.nph[p] = pl.npart[p] - pl.r0[p];
    pl.ntm[p] = (pl.nm[p] + NB - 1) / NB; pl.npE[p] = pl.q[p] + pl.ntm[p]; pl.ldE[p] = pl.npE[p] * NB;
  }
  if (bt > SPK_MAX_BT || (w - 1 + NB - 1) / NB > CL_MAX_BT) return pl;
  pl.ok = 1;
  return pl;
}

}  // namespace ba
