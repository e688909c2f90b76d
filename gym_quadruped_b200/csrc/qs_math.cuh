// qs_math.cuh -- small fixed-size vector / quaternion / spatial-algebra helpers for the fused step kernel.
// Spatial vectors follow the engine's convention [rot(3), lin(3)], expressed in a world-aligned frame whose origin
// is the robot's subtree centre of mass ("com-based frame").
#pragma once
#include <math.h>
#include <stdint.h>

// QS_DEV marks code shared between the CUDA build and the host warp emulator (tests/emu), which supplies
// qs::syncwarp / shfl / shfl_xor / ballot with 32 lock-stepped host threads before including this header.
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define QS_DEV __device__ __forceinline__
#define QS_NOINLINE __device__ __noinline__
namespace qs {
QS_DEV void syncwarp() { __syncwarp(); }
template <typename T> QS_DEV T shfl_xor(T v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
template <typename T> QS_DEV T shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
QS_DEV unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
QS_DEV int popc(unsigned x) { return __popc(x); }
QS_DEV int ctz(unsigned x) { return __ffs(int(x)) - 1; }
QS_DEV uint32_t umulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
}  // namespace qs
#else
#ifndef QS_DEV
#error "host build must include the emulator shim (tests/emu/warp_emu.h) first"
#endif
#endif

namespace qs {

template <typename T> struct Num;
template <> struct Num<float> {
  static QS_DEV float sqrt(float x) { return sqrtf(x); }
#if defined(__CUDA_ARCH__)
  // fp32 product path: MUFU-based reciprocal / rsqrt (<= 2 ulp) instead of the IEEE division subroutine; the fp64
  // instantiation (parity build) and the host emulator keep exact division.
  static QS_DEV float rsqrt(float x) { return rsqrtf(x); }
  static QS_DEV float rcp(float x) { return __fdividef(1.0f, x); }
  static QS_DEV float div(float a, float b) { return __fdividef(a, b); }
#else
  static QS_DEV float rsqrt(float x) { return 1.0f / sqrtf(x); }
  static QS_DEV float rcp(float x) { return 1.0f / x; }
  static QS_DEV float div(float a, float b) { return a / b; }
#endif
  static QS_DEV void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
  static QS_DEV float atan2(float y, float x) { return atan2f(y, x); }
  static QS_DEV float asin(float x) { return asinf(x); }
  static QS_DEV float pow(float x, float y) { return powf(x, y); }
  static QS_DEV float abs(float x) { return fabsf(x); }
  static QS_DEV float max(float a, float b) { return fmaxf(a, b); }
  static QS_DEV float min(float a, float b) { return fminf(a, b); }
  static QS_DEV float floor(float x) { return floorf(x); }
  static constexpr float minval = 1e-15f;
  static constexpr float big = 1e30f;
};
template <> struct Num<double> {
  static QS_DEV double sqrt(double x) { return ::sqrt(x); }
  static QS_DEV double rsqrt(double x) { return 1.0 / ::sqrt(x); }
  static QS_DEV double rcp(double x) { return 1.0 / x; }
  static QS_DEV double div(double a, double b) { return a / b; }
  static QS_DEV void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
  static QS_DEV double atan2(double y, double x) { return ::atan2(y, x); }
  static QS_DEV double asin(double x) { return ::asin(x); }
  static QS_DEV double pow(double x, double y) { return ::pow(x, y); }
  static QS_DEV double abs(double x) { return fabs(x); }
  static QS_DEV double max(double a, double b) { return fmax(a, b); }
  static QS_DEV double min(double a, double b) { return fmin(a, b); }
  static QS_DEV double floor(double x) { return ::floor(x); }
  static constexpr double minval = 1e-15;
  static constexpr double big = 1e300;
};

template <typename T> QS_DEV T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> QS_DEV void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
// r = M v, M row-major 3x3
template <typename T> QS_DEV void mul_mv(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
// r = M^T v
template <typename T> QS_DEV void mul_mtv(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> QS_DEV void quat_mul(T* r, const T* a, const T* b) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
template <typename T> QS_DEV void quat_normalize(T* q) {
  T n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2 < Num<T>::minval) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  T inv = T(1) / Num<T>::sqrt(n2);
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
template <typename T> QS_DEV void quat_to_mat(T* m, const T* q) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
template <typename T> QS_DEV void rot_vec_quat(T* r, const T* v, const T* q) {
  T m[9];
  quat_to_mat(m, q);
  mul_mv(r, m, v);
}
// 10-number spatial inertia [Ixx,Iyy,Izz,Ixy,Ixz,Iyz, m*ox, m*oy, m*oz, m] times motion vector
template <typename T> QS_DEV void mul_inert_vec(T* r, const T* I, const T* v) {
  r[0] = I[0] * v[0] + I[3] * v[1] + I[4] * v[2] - I[8] * v[4] + I[7] * v[5];
  r[1] = I[3] * v[0] + I[1] * v[1] + I[5] * v[2] + I[8] * v[3] - I[6] * v[5];
  r[2] = I[4] * v[0] + I[5] * v[1] + I[2] * v[2] - I[7] * v[3] + I[6] * v[4];
  r[3] = I[8] * v[1] - I[7] * v[2] + I[9] * v[3];
  r[4] = I[6] * v[2] - I[8] * v[0] + I[9] * v[4];
  r[5] = I[7] * v[0] - I[6] * v[1] + I[9] * v[5];
}
template <typename T> QS_DEV void cross_motion(T* r, const T* vel, const T* v) {
  T a[3], b[3], c[3];
  cross3(a, vel, v); cross3(b, vel, v + 3); cross3(c, vel + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
template <typename T> QS_DEV void cross_force(T* r, const T* vel, const T* f) {
  T a[3], b[3], c[3];
  cross3(a, vel, f); cross3(b, vel + 3, f + 3); cross3(c, vel, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

// warp reductions (full mask)
template <typename T> QS_DEV T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor(v, o);
  return v;
}
// several independent butterfly sums interleaved: same result per value as warp_sum, one dependent chain instead of k
template <typename T> QS_DEV void warp_sum2(T& a, T& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const T ta = shfl_xor(a, o), tb = shfl_xor(b, o); a += ta; b += tb; }
}
template <typename T> QS_DEV void warp_sum3(T& a, T& b, T& c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const T ta = shfl_xor(a, o), tb = shfl_xor(b, o), tc = shfl_xor(c, o); a += ta; b += tb; c += tc; }
}
template <typename T> QS_DEV void warp_sum4(T& a, T& b, T& c, T& d) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T ta = shfl_xor(a, o), tb = shfl_xor(b, o), tc = shfl_xor(c, o), td = shfl_xor(d, o);
    a += ta; b += tb; c += tc; d += td;
  }
}
template <typename T> QS_DEV T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { T w = shfl_xor(v, o); v = w > v ? w : v; }
  return v;
}

// counter-based RNG: Philox-4x32-10 keyed by (seed_lo, seed_hi), counter (env, stream, draw, 0)
QS_DEV void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
QS_DEV float u32_to_unit(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }  // [0,1)

}  // namespace qs
