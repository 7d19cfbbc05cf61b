// Per-observation camera model for sm_100a: projection residual (robustified) and the analytic 2x12
// Jacobian block. Replaces the reference's CPU evaluators
//   BAFunctor::E_pos / dE_pos / poseDerivatives / projectPoint  (src/Optimization/BAFunctor.h:126-297)
//   DistortionFunction::operator(), derivativeWrt*               (src/DistortionFunction.cpp:14-51)
//   CameraMatrix::transformPointIntoCameraSpace, getFocalLength  (src/CameraMatrix.cpp:207-209,259-261)
//   Math::createRotationMatrixRodrigues                          (src/MathUtils.h:66-82)
// Written from the model equations, templated on the reference's Scalar switch.
#pragma once
#include <cuda_runtime.h>

namespace ba {

// Camera record in HBM: 16 scalars, 16*sizeof(T)-byte aligned -> vector loads.
// [0..8] R row-major, [9..11] T, [12] f (= K00, negative for BAL), [13] k1, [14] k2, [15] pad
constexpr int CAM_STRIDE = 16;

template <class T> struct Cam { T R[9]; T t[3]; T f, k1, k2; };

template <class T> __device__ __forceinline__ void load_cam(const T* __restrict__ cams, int c, Cam<T>& o);

template <> __device__ __forceinline__ void load_cam<double>(const double* __restrict__ cams, int c, Cam<double>& o) {
  const double2* p = reinterpret_cast<const double2*>(cams + (size_t)c * CAM_STRIDE);
  double2 v0 = __ldg(p + 0), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
  double2 v4 = __ldg(p + 4), v5 = __ldg(p + 5), v6 = __ldg(p + 6), v7 = __ldg(p + 7);
  o.R[0] = v0.x; o.R[1] = v0.y; o.R[2] = v1.x; o.R[3] = v1.y; o.R[4] = v2.x; o.R[5] = v2.y;
  o.R[6] = v3.x; o.R[7] = v3.y; o.R[8] = v4.x; o.t[0] = v4.y; o.t[1] = v5.x; o.t[2] = v5.y;
  o.f = v6.x; o.k1 = v6.y; o.k2 = v7.x;
}
template <> __device__ __forceinline__ void load_cam<float>(const float* __restrict__ cams, int c, Cam<float>& o) {
  const float4* p = reinterpret_cast<const float4*>(cams + (size_t)c * CAM_STRIDE);
  float4 v0 = __ldg(p + 0), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
  o.R[0] = v0.x; o.R[1] = v0.y; o.R[2] = v0.z; o.R[3] = v0.w; o.R[4] = v1.x; o.R[5] = v1.y;
  o.R[6] = v1.z; o.R[7] = v1.w; o.R[8] = v2.x; o.t[0] = v2.y; o.t[1] = v2.z; o.t[2] = v2.w;
  o.f = v3.x; o.k1 = v3.y; o.k2 = v3.z;
}

template <class T> __device__ __forceinline__ T tsqrt(T x);
template <> __device__ __forceinline__ double tsqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float tsqrt<float>(float x) { return sqrtf(x); }
template <class T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }
template <class T> __device__ __forceinline__ T tabs(T a) { return a < T(0) ? -a : a; }

// psi / psi_weight: BAFunctor.h:147-148
template <class T> __device__ __forceinline__ T psi(T tau2, T r2) {
  return (r2 < tau2) ? r2 * (T(2.0) - r2 / tau2) / T(4.0) : tau2 / T(4.0);
}

// e = r_hat * sqrt(psi(|r|^2)), r = f * dist(pi(R X + T)) - m     (E_pos, BAFunctor.h:160-178)
template <class T>
__device__ __forceinline__ void obs_residual(const Cam<T>& c, T X0, T X1, T X2, T m0, T m1, T tau2, T& e0, T& e1) {
  const T eps = T(1e-15);  // BAFunctor.h:159
  const T xx = c.R[0] * X0 + c.R[1] * X1 + c.R[2] * X2 + c.t[0];
  const T yy = c.R[3] * X0 + c.R[4] * X1 + c.R[5] * X2 + c.t[1];
  const T zz = c.R[6] * X0 + c.R[7] * X1 + c.R[8] * X2 + c.t[2];
  const T xu0 = xx / zz, xu1 = yy / zz;
  const T r2u = xu0 * xu0 + xu1 * xu1, r4u = r2u * r2u;
  const T kr = T(1) + c.k1 * r2u + c.k2 * r4u;
  const T r0 = c.f * (kr * xu0) - m0, r1 = c.f * (kr * xu1) - m1;
  const T r2 = r0 * r0 + r1 * r1;
  const T sp = tsqrt(psi(tau2, r2));
  const T rn = T(1.0) / tmax(eps, tsqrt(r2));
  e0 = r0 * sp * rn;
  e1 = r1 * sp * rn;
}

// Residual + Jacobian block. Jc[18] = 2x9 row-major over camera columns (T0..2, w0..2, f, k1, k2),
// Jp[6] = 2x3 row-major over the point's xyz.   (dE_pos, BAFunctor.h:181-297)
template <class T>
__device__ __forceinline__ void obs_jacobian(const Cam<T>& c, T X0, T X1, T X2, T m0, T m1, T tau2,
                                             T& e0, T& e1, T* __restrict__ Jc, T* __restrict__ Jp) {
  const T eps = T(1e-15);
  const T xx = c.R[0] * X0 + c.R[1] * X1 + c.R[2] * X2 + c.t[0];
  const T yy = c.R[3] * X0 + c.R[4] * X1 + c.R[5] * X2 + c.t[1];
  const T zz = c.R[6] * X0 + c.R[7] * X1 + c.R[8] * X2 + c.t[2];
  // d(RX+T)/d omega = -[RX]_x (left-multiplicative update), poseDerivatives :126-142
  const T D0 = xx - c.t[0], D1 = yy - c.t[1], D2 = zz - c.t[2];
  const T iz = T(1.0) / zz;
  const T xu0 = xx / zz, xu1 = yy / zz;
  const T r2u = xu0 * xu0 + xu1 * xu1, r4u = r2u * r2u;
  const T kr = T(1) + c.k1 * r2u + c.k2 * r4u;
  const T xd0 = kr * xu0, xd1 = kr * xu1;
  // dxu/dXX (:219-221)
  const T a02 = -xx / (zz * zz), a12 = -yy / (zz * zz);
  // f * dxd/dxu (DistortionFunction.cpp:38-51)
  const T dkr = T(2) * c.k1 + T(4) * c.k2 * r2u;
  const T d00 = c.f * (kr + xu0 * xu0 * dkr), d01 = c.f * (xu0 * xu1 * dkr), d11 = c.f * (kr + xu1 * xu1 * dkr);
  // dp/dXX (2x3)
  const T p00 = d00 * iz, p01 = d01 * iz, p02 = d00 * a02 + d01 * a12;
  const T p10 = d01 * iz, p11 = d11 * iz, p12 = d01 * a02 + d11 * a12;
  // robust kernel: residual and outer derivative (:227-242)
  const T r0 = c.f * xd0 - m0, r1 = c.f * xd1 - m1;
  const T r2 = r0 * r0 + r1 * r1;
  const T W = tmax(T(0.0), T(1.0) - r2 / tau2);
  const T sp = tsqrt(psi(tau2, r2));
  const T rsp = T(1.0) / tmax(eps, sp);
  const T rcp_r2 = T(1.0) / tmax(eps, r2);
  const T nr = tsqrt(r2);
  const T rn = T(1.0) / tmax(eps, nr);
  e0 = r0 * sp * rn;
  e1 = r1 * sp * rn;
  const T rr00 = r0 * r0 * rn, rr01 = r0 * r1 * rn, rr11 = r1 * r1 * rn;
  const T ca = W / T(2.0) * rsp, cb = sp * rcp_r2;
  const T o00 = ca * rr00 + cb * (nr - rr00), o01 = ca * rr01 + cb * (T(0) - rr01), o11 = ca * rr11 + cb * (nr - rr11);
  // un-robustified 2x12 block
  T b0[12], b1[12];
  b0[0] = p00; b0[1] = p01; b0[2] = p02;
  b1[0] = p10; b1[1] = p11; b1[2] = p12;
  // cols 3-5: dp_dXX * (-[D]_x) with -[D]_x = [0 D2 -D1; -D2 0 D0; D1 -D0 0]
  b0[3] = -p01 * D2 + p02 * D1; b0[4] = p00 * D2 - p02 * D0; b0[5] = -p00 * D1 + p01 * D0;
  b1[3] = -p11 * D2 + p12 * D1; b1[4] = p10 * D2 - p12 * D0; b1[5] = -p10 * D1 + p11 * D0;
  b0[6] = xd0; b1[6] = xd1;
  b0[7] = c.f * (xu0 * r2u); b1[7] = c.f * (xu1 * r2u);
  b0[8] = c.f * (xu0 * r4u); b1[8] = c.f * (xu1 * r4u);
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    b0[9 + b] = p00 * c.R[b] + p01 * c.R[3 + b] + p02 * c.R[6 + b];
    b1[9 + b] = p10 * c.R[b] + p11 * c.R[3 + b] + p12 * c.R[6 + b];
  }
#pragma unroll
  for (int b = 0; b < 9; ++b) {
    Jc[b] = o00 * b0[b] + o01 * b1[b];
    Jc[9 + b] = o01 * b0[b] + o11 * b1[b];
  }
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    Jp[b] = o00 * b0[9 + b] + o01 * b1[9 + b];
    Jp[3 + b] = o01 * b0[9 + b] + o11 * b1[9 + b];
  }
}

// R <- Rodrigues(w) * R with the reference's hard |w| > 1e-6 cut-off (MathUtils.h:66-82, quirk Q2)
template <class T>
__device__ __forceinline__ void rodrigues_left(const T w0, const T w1, const T w2, const T* __restrict__ Rin, T* __restrict__ Rout) {
  const T theta = tsqrt(w0 * w0 + w1 * w1 + w2 * w2);
  T dR[9] = {T(1), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(1)};
  if (tabs(theta) > T(1e-6)) {
    const T J[9] = {T(0), -w2, w1, w2, T(0), -w0, -w1, w0, T(0)};
    T J2[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) J2[3 * a + b] = J[3 * a] * J[b] + J[3 * a + 1] * J[3 + b] + J[3 * a + 2] * J[6 + b];
    const T c1 = sin(theta) / theta;
    const T c2 = (T(1.0) - cos(theta)) / (theta * theta);
#pragma unroll
    for (int i = 0; i < 9; ++i) dR[i] = dR[i] + c1 * J[i] + c2 * J2[i];
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) Rout[3 * a + b] = dR[3 * a] * Rin[b] + dR[3 * a + 1] * Rin[3 + b] + dR[3 * a + 2] * Rin[6 + b];
}

}  // namespace ba
