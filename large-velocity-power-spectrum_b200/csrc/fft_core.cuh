// Register-resident line FFT used by all three passes of the 3-D transform.
//
// A line of L = R1*R2*R3 complex points is transformed by T = L/R1 threads, R1 points per thread (R1 = 16 for the
// power-of-two lines, 10 or 5 for the 2^a 5^b lines behind N = 250, 500, 1000; R2 and R3 divide R1):
//   stage 1: radix-R1 butterflies on points  j*T + t           (j = 0..R1-1)     -> twiddle W_L^(t*k1)
//   stage 2: radix-R2 butterflies, R1/R2 per thread                               -> twiddle W_L^(R1*d3*k2)
//   stage 3: radix-R3 butterflies, R1/R3 per thread (absent when R3 == 1)
// Stages exchange data through shared memory; the array index at every exchange is the mixed-radix
// number (d1,d2,d3) = d1*T + d2*R3 + d3 whose digits are replaced n -> k stage by stage, so the final
// register slot j of thread t holds output index  kout(j,t)  (see below).  The index maps, twiddles and kout are
// restated in numpy and checked against numpy.fft for every instantiated radix combination in
// tests/test_fft_index_model.py.
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// forward DFTs (e^{-2 pi i nk/R}), natural order in and out
__device__ __forceinline__ void dft2(float2& a0, float2& a1) {
  float2 t = a0;
  a0 = cadd(t, a1);
  a1 = csub(t, a1);
}
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
  a0 = cadd(t0, t2);
  a1 = cadd(t1, t3);
  a2 = csub(t0, t2);
  a3 = csub(t1, t3);
}

template <int R>
__device__ __forceinline__ void dft(float2* u);
template <>
__device__ __forceinline__ void dft<1>(float2*) {}
template <>
__device__ __forceinline__ void dft<2>(float2* u) { dft2(u[0], u[1]); }
template <>
__device__ __forceinline__ void dft<4>(float2* u) { dft4(u[0], u[1], u[2], u[3]); }
template <>
__device__ __forceinline__ void dft<8>(float2* u) {
  // n = 4*n1 + n' : radix-2 over n1, twiddle W8^(n'*k1), radix-4 over n';  X[k1 + 2k'] lands in y[k1][k']
  const float h = 0.70710678118654752440f;
  float2 y0[4], y1[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    y0[n] = cadd(u[n], u[n + 4]);
    y1[n] = csub(u[n], u[n + 4]);
  }
  y1[1] = cmul(y1[1], make_float2(h, -h));
  y1[2] = cmul_mi(y1[2]);
  y1[3] = cmul(y1[3], make_float2(-h, -h));
  dft4(y0[0], y0[1], y0[2], y0[3]);
  dft4(y1[0], y1[1], y1[2], y1[3]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    u[2 * k] = y0[k];
    u[2 * k + 1] = y1[k];
  }
}
template <>
__device__ __forceinline__ void dft<16>(float2* u) {
  // n = 4*n1 + n' : radix-4 over n1, twiddle W16^(n'*k1), radix-4 over n';  X[k1 + 4k'] = y[k1][k']
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  float2 y[4][4];  // [n'][k1] after the first step
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    y[n][0] = u[n];
    y[n][1] = u[n + 4];
    y[n][2] = u[n + 8];
    y[n][3] = u[n + 12];
    dft4(y[n][0], y[n][1], y[n][2], y[n][3]);
  }
  // W16^m = (cos(2 pi m/16), -sin(2 pi m/16)),  m = n'*k1
  y[1][1] = cmul(y[1][1], make_float2(c1, -s1));   // m=1
  y[1][2] = cmul(y[1][2], make_float2(h, -h));     // m=2
  y[1][3] = cmul(y[1][3], make_float2(s1, -c1));   // m=3
  y[2][1] = cmul(y[2][1], make_float2(h, -h));     // m=2
  y[2][2] = cmul_mi(y[2][2]);                      // m=4
  y[2][3] = cmul(y[2][3], make_float2(-h, -h));    // m=6
  y[3][1] = cmul(y[3][1], make_float2(s1, -c1));   // m=3
  y[3][2] = cmul(y[3][2], make_float2(-h, -h));    // m=6
  y[3][3] = cmul(y[3][3], make_float2(-c1, s1));   // m=9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    float2 a0 = y[0][k1], a1 = y[1][k1], a2 = y[2][k1], a3 = y[3][k1];
    dft4(a0, a1, a2, a3);
    u[k1] = a0;
    u[k1 + 4] = a1;
    u[k1 + 8] = a2;
    u[k1 + 12] = a3;
  }
}

template <>
__device__ __forceinline__ void dft<5>(float2* u) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f, s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
  const float2 t1 = cadd(u[1], u[4]), t2 = cadd(u[2], u[3]), t3 = csub(u[1], u[4]), t4 = csub(u[2], u[3]);
  const float2 a0 = u[0];
  const float2 m1 = make_float2(a0.x + c1 * t1.x + c2 * t2.x, a0.y + c1 * t1.y + c2 * t2.y);
  const float2 m2 = make_float2(a0.x + c2 * t1.x + c1 * t2.x, a0.y + c2 * t1.y + c1 * t2.y);
  const float2 n1 = cmul_mi(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));   // -i (s1 t3 + s2 t4)
  const float2 n2 = cmul_mi(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));   // -i (s2 t3 - s1 t4)
  u[0] = make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
  u[1] = cadd(m1, n1);
  u[4] = csub(m1, n1);
  u[2] = cadd(m2, n2);
  u[3] = csub(m2, n2);
}
template <>
__device__ __forceinline__ void dft<10>(float2* u) {
  // n = 2*n2 + n1: two 5-point transforms (even, odd inputs), X[k2] = E[k2] + W10^k2 O[k2], X[k2+5] = E[k2] - W10^k2 O[k2]
  float2 e[5] = {u[0], u[2], u[4], u[6], u[8]}, o[5] = {u[1], u[3], u[5], u[7], u[9]};
  dft<5>(e);
  dft<5>(o);
  o[1] = cmul(o[1], make_float2(0.80901699437494742410f, -0.58778525229247312917f));
  o[2] = cmul(o[2], make_float2(0.30901699437494742410f, -0.95105651629515357212f));
  o[3] = cmul(o[3], make_float2(-0.30901699437494742410f, -0.95105651629515357212f));
  o[4] = cmul(o[4], make_float2(-0.80901699437494742410f, -0.58778525229247312917f));
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    u[k] = cadd(e[k], o[k]);
    u[k + 5] = csub(e[k], o[k]);
  }
}

// Shared-memory placement of exchange index `idx` (in float2 units, before the column/line offset).
// A skew of one element per 16 keeps the strided stage-2/3 reads off a single bank group.
template <int ISTRIDE>
__device__ __forceinline__ int xphys(int idx) {
  return ISTRIDE == 1 ? idx + (idx >> 4) : idx * ISTRIDE + 2 * (idx >> 4);
}
template <int L, int ISTRIDE>
__host__ __device__ constexpr int xsize() {  // float2 elements needed for one line group
  return ISTRIDE == 1 ? L + (L >> 4) : L * ISTRIDE + 2 * (L >> 4);
}

template <int R1, int R2, int R3, int ISTRIDE>
struct LineFFT {
  static constexpr int P = R1;            // points per thread
  static constexpr int L = R1 * R2 * R3;
  static constexpr int T = R2 * R3;
  static constexpr int M2 = R1 / R2;
  static constexpr int M3 = R3 > 1 ? R1 / R3 : R1;
  static_assert(R1 % R2 == 0 && R1 % R3 == 0, "the later radices must divide the number of points per thread");

  // output index held in register slot j of thread t after run()
  __device__ __forceinline__ static int kout(int j, int t) {
    if (R3 == 1) return (t * M2 + j / R2) + R1 * (j % R2);
    return t + T * (j / R3) + R1 * R2 * (j % R3);
  }

  // v[j] = x[j*T + t] on entry.  sm points at this line's (or this column's) exchange area.
  // tw = table of W_L^m, m in [0,L).  Every thread of the CTA must call this (it contains barriers).
  __device__ __forceinline__ static void run(float2 (&v)[R1], int t, float2* sm, const float2* __restrict__ tw) {
    dft<R1>(v);
#pragma unroll
    for (int k1 = 1; k1 < R1; ++k1) v[k1] = cmul(v[k1], __ldg(tw + t * k1));
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) sm[xphys<ISTRIDE>(k1 * T + t)] = v[k1];
    __syncthreads();
#pragma unroll
    for (int b = 0; b < M2; ++b) {
      const int g = t * M2 + b, k1 = g / R3, d3 = g % R3;
      float2 u[R2];
#pragma unroll
      for (int a = 0; a < R2; ++a) u[a] = sm[xphys<ISTRIDE>(k1 * T + a * R3 + d3)];
      dft<R2>(u);
      if (R3 > 1) {
#pragma unroll
        for (int k2 = 1; k2 < R2; ++k2) u[k2] = cmul(u[k2], __ldg(tw + R1 * d3 * k2));
      }
#pragma unroll
      for (int a = 0; a < R2; ++a) v[a + R2 * b] = u[a];
    }
    if (R3 == 1) return;
    __syncthreads();
#pragma unroll
    for (int b = 0; b < M2; ++b) {
      const int g = t * M2 + b, k1 = g / R3, d3 = g % R3;
#pragma unroll
      for (int a = 0; a < R2; ++a) sm[xphys<ISTRIDE>(k1 * T + a * R3 + d3)] = v[a + R2 * b];
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < M3; ++b) {
      const int g = t + T * b, k1 = g % R1, k2 = g / R1;
      float2 u[R3 > 1 ? R3 : 1];
#pragma unroll
      for (int a = 0; a < R3; ++a) u[a] = sm[xphys<ISTRIDE>(k1 * T + k2 * R3 + a)];
      dft<R3>(u);
#pragma unroll
      for (int a = 0; a < R3; ++a) v[a + R3 * b] = u[a];
    }
  }
};
