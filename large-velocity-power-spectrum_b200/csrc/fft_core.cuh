// Register-resident line FFT used by all three passes of the 3-D transform.
//
// A line of L = 16*R2*R3 complex points is transformed by T = L/16 threads, 16 points per thread:
//   stage 1: radix-16 butterflies on points  j*T + t           (j = 0..15)       -> twiddle W_L^(t*k1)
//   stage 2: radix-R2 butterflies, 16/R2 per thread                               -> twiddle W_L^(16*d3*k2)
//   stage 3: radix-R3 butterflies, 16/R3 per thread (absent when R3 == 1)
// Stages exchange data through shared memory; the array index at every exchange is the mixed-radix
// number (d1,d2,d3) = d1*T + d2*R3 + d3 whose digits are replaced n -> k stage by stage, so the final
// register slot j of thread t holds output index  kout(j,t)  (see below).  The index maps are checked
// against numpy in tests/test_fft_index_model.py.
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

template <int R2, int R3, int ISTRIDE>
struct LineFFT {
  static constexpr int L = 16 * R2 * R3;
  static constexpr int T = R2 * R3;
  static constexpr int M2 = 16 / R2;
  static constexpr int M3 = R3 > 1 ? 16 / R3 : 16;

  // output index held in register slot j of thread t after run()
  __device__ __forceinline__ static int kout(int j, int t) {
    if (R3 == 1) return (t * M2 + j / R2) + 16 * (j % R2);
    return t + T * (j / R3) + 16 * R2 * (j % R3);
  }

  // v[j] = x[j*T + t] on entry.  sm points at this line's (or this column's) exchange area.
  // tw = table of W_L^m, m in [0,L).  Every thread of the CTA must call this (it contains barriers).
  __device__ __forceinline__ static void run(float2 (&v)[16], int t, float2* sm, const float2* __restrict__ tw) {
    dft<16>(v);
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], __ldg(tw + t * k1));
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) sm[xphys<ISTRIDE>(k1 * T + t)] = v[k1];
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
        for (int k2 = 1; k2 < R2; ++k2) u[k2] = cmul(u[k2], __ldg(tw + 16 * d3 * k2));
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
      const int g = t + T * b, k1 = g % 16, k2 = g / 16;
      float2 u[R3 > 1 ? R3 : 1];
#pragma unroll
      for (int a = 0; a < R3; ++a) u[a] = sm[xphys<ISTRIDE>(k1 * T + k2 * R3 + a)];
      dft<R3>(u);
#pragma unroll
      for (int a = 0; a < R3; ++a) v[a + R3 * b] = u[a];
    }
  }
};
