// Folding as an outer stage: the device forms of
//   BoxField.fold        vpower/interp.py:598-609   phi = fold_field(v * phase_beta, m) / m^1.5
//   _get_phase           vpower/interp.py:1215-1225 phase = exp(-i (2 pi / N) (beta . x))
//   fold_field           vpower/interp.py:1228-1252 sum of the m^3 sub-blocks of size N/m, (i, j, k) order
//   FoldedBox.fold_spctrm vpower/interp.py:755-791  P = 1/2 sum_c |a FFT_n(phi_c)|^2  (the |k| pairing with the beta shift and the
//                                                    shell histogram are vp_k_magnitude / vp_hist_weighted)
// The folded field is complex, so its transform is a plain c2c 3-D FFT of size n = N/m: three in-place line passes with the
// register/shared-memory LineFFT of fft_core.cuh for n = 64 .. 1024 (f32, like the main transform), and a direct f64 DFT per
// axis for any other n (small or odd folded sizes; O(n^4), exact to rounding).
#include <math.h>

#include <vector>

#include "common.cuh"
#include "fft_core.cuh"

namespace {

// One thread per folded node: phase (f64 sincos of the same product numpy forms), f64 accumulation in fold_field's order.
// out: [n1,n1,n1,ncomp] complex128, the reference's FoldedBox.f layout.
__global__ void __launch_bounds__(256) k_fold(const float* __restrict__ f0, const float* __restrict__ f1, const float* __restrict__ f2,
                                              int ncomp, int N, int m, int b0, int b1, int b2, double w, double div,
                                              double2* __restrict__ out) {
  const int n1 = N / m;
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= size_t(n1) * n1 * n1) return;
  const int z = int(t % n1);
  const size_t u = t / n1;
  const int y = int(u % n1), x = int(u / n1);
  double re[3] = {0.0, 0.0, 0.0}, im[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j)
      for (int k = 0; k < m; ++k) {
        const int X = x + i * n1, Y = y + j * n1, Z = z + k * n1;
        const long long d = (long long)b0 * X + (long long)b1 * Y + (long long)b2 * Z;
        double s, c;
        sincos(w * double(d), &s, &c);                       // exp(-i theta) = (cos theta, -sin theta), theta = (2 pi / N) * d
        const size_t at = (size_t(X) * N + Y) * N + Z;
        const float v[3] = {f0[at], ncomp > 1 ? f1[at] : 0.f, ncomp > 2 ? f2[at] : 0.f};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          re[q] += double(v[q]) * c;
          im[q] += double(v[q]) * -s;
        }
      }
  for (int q = 0; q < ncomp; ++q) out[t * ncomp + q] = make_double2(re[q] / div, im[q] / div);
}

// ---- generic axis pass: direct DFT in f64, one CTA per line (line element i at base + i * stride)
__global__ void __launch_bounds__(128) k_dft_axis_d(double2* __restrict__ z, int n, size_t stride, int n_inner, size_t outer_stride,
                                                    size_t inner_stride, const double2* __restrict__ tw) {
  extern __shared__ double2 dline[];
  const size_t base = size_t(blockIdx.x / n_inner) * outer_stride + size_t(blockIdx.x % n_inner) * inner_stride;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dline[i] = z[base + i * stride];
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double re = 0, im = 0;
    int mm = 0;
    for (int i = 0; i < n; ++i) {
      const double2 wv = tw[mm], a = dline[i];
      re += a.x * wv.x - a.y * wv.y;
      im += a.x * wv.y + a.y * wv.x;
      mm += k;
      if (mm >= n) mm -= n;
    }
    z[base + k * stride] = make_double2(re, im);
  }
}
__global__ void k_comp_to_z(const double2* __restrict__ f, int ncomp, int c, size_t n3, double2* __restrict__ z) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n3) z[i] = f[i * ncomp + c];
}
__global__ void k_accum_power_d(const double2* __restrict__ z, double* __restrict__ P, size_t n3, int first) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const double v = z[i].x * z[i].x + z[i].y * z[i].y;
  P[i] = first ? v : P[i] + v;
}

// ---- fast path: f32 c2c line passes with LineFFT
__global__ void k_comp_to_zf(const double2* __restrict__ f, int ncomp, int c, size_t n3, float2* __restrict__ z) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n3) { const double2 v = f[i * ncomp + c]; z[i] = make_float2(float(v.x), float(v.y)); }
}
__global__ void k_accum_power_f(const float2* __restrict__ z, double* __restrict__ P, size_t n3, int first) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const double v = double(z[i].x) * z[i].x + double(z[i].y) * z[i].y;
  P[i] = first ? v : P[i] + v;
}
// contiguous lines (z axis), 256/T lines per CTA
template <int R1, int R2, int R3>
__global__ void __launch_bounds__(256) k_c2c_z(float2* __restrict__ data, const float2* __restrict__ tw) {
  using F = LineFFT<R1, R2, R3, 1>;
  constexpr int L = F::L, T = F::T, LINES = 256 / T, XS = xsize<L, 1>();
  extern __shared__ float2 sm[];
  const int tid = threadIdx.x, ll = tid / T, t = tid % T;
  float2* g = data + (size_t(blockIdx.x) * LINES + ll) * L;
  float2 v[F::P];
#pragma unroll
  for (int j = 0; j < F::P; ++j) v[j] = g[j * T + t];
  F::run(v, t, sm + ll * XS, tw);
#pragma unroll
  for (int j = 0; j < F::P; ++j) g[F::kout(j, t)] = v[j];
}
// strided lines (y or x axis), C adjacent columns per CTA, in place:
//   element i of column c of block b:  (b / tiles) * outer + (b % tiles) * C + c + i * es
template <int R1, int R2, int R3, int C>
__global__ void __launch_bounds__(R2* R3* C) k_c2c_s(float2* __restrict__ data, int tiles, size_t outer, size_t es, const float2* __restrict__ tw) {
  using F = LineFFT<R1, R2, R3, C>;
  constexpr int T = F::T;
  extern __shared__ float2 sm[];
  const int tid = threadIdx.x, c = tid % C, t = tid / C;
  float2* base = data + size_t(blockIdx.x / tiles) * outer + size_t(blockIdx.x % tiles) * C + c;
  float2 v[F::P];
#pragma unroll
  for (int j = 0; j < F::P; ++j) v[j] = base[size_t(j * T + t) * es];
  F::run(v, t, sm + c, tw);
#pragma unroll
  for (int j = 0; j < F::P; ++j) base[size_t(F::kout(j, t)) * es] = v[j];
}

template <int R1, int R2, int R3, int C>
int c2c_pow2(vp_ctx* ctx, float2* z, int n, const float2* tw, cudaStream_t st) {
  using FZ = LineFFT<R1, R2, R3, 1>;
  using FS = LineFFT<R1, R2, R3, C>;
  constexpr int LINES = 256 / FZ::T;
  const size_t smz = size_t(LINES) * xsize<FZ::L, 1>() * sizeof(float2), sms = size_t(xsize<FS::L, C>()) * sizeof(float2);
  static bool attr = false;
  if (!attr) {
    VP_CUDA(cudaFuncSetAttribute(k_c2c_z<R1, R2, R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smz)));
    VP_CUDA(cudaFuncSetAttribute(k_c2c_s<R1, R2, R3, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sms)));
    attr = true;
  }
  const size_t nn = size_t(n) * n;
  vp_stage stage(ctx, "fold_c2c", st, 3, 3.0 * 16.0 * double(nn) * n);
  k_c2c_z<R1, R2, R3><<<unsigned(nn / LINES), 256, smz, st>>>(z, tw);
  const int tiles = n / C;
  k_c2c_s<R1, R2, R3, C><<<unsigned(n * tiles), FS::T * C, sms, st>>>(z, tiles, nn, size_t(n), tw);   // y lines: outer = x
  k_c2c_s<R1, R2, R3, C><<<unsigned(n * tiles), FS::T * C, sms, st>>>(z, tiles, size_t(n), nn, tw);   // x lines: outer = y
  VP_CHECK_LAUNCH();
  return VP_OK;
}

}  // namespace

extern "C" int vp_fold_field(vp_ctx* ctx, const float* const* field_d, int ncomp, int N, int m, const int* beta, double* folded_d,
                             void* stream) {
  VP_REQUIRE(ctx && field_d && beta && folded_d, "vp_fold_field: null argument");
  VP_REQUIRE(ncomp >= 1 && ncomp <= 3 && N >= 1 && m >= 1 && N % m == 0, "vp_fold_field: N=%d must be a multiple of the folding factor m=%d", N, m);
  for (int c = 0; c < ncomp; ++c) VP_REQUIRE(field_d[c], "vp_fold_field: null field %d", c);
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n1 = N / m;
  const size_t n3 = size_t(n1) * n1 * n1;
  // N^3 reals read per component, n1^3 complex128 written per component
  vp_stage stage(ctx, "fold_field", st, 1, 4.0 * double(N) * N * N * ncomp + 16.0 * double(n3) * ncomp);
  k_fold<<<unsigned((n3 + 255) / 256), 256, 0, st>>>(field_d[0], ncomp > 1 ? field_d[1] : nullptr, ncomp > 2 ? field_d[2] : nullptr, ncomp,
                                                   N, m, beta[0], beta[1], beta[2], 2.0 * M_PI / double(N), pow(double(m), 1.5),
                                                   reinterpret_cast<double2*>(folded_d));
  VP_CHECK_LAUNCH();
  return VP_OK;
}

extern "C" int vp_fold_power(vp_ctx* ctx, const double* folded_d, int ncomp, int n, double* P_d, void* stream) {
  VP_REQUIRE(ctx && folded_d && P_d && ncomp >= 1 && ncomp <= 3 && n >= 1, "vp_fold_power: bad argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n3 = size_t(n) * n * n, nn = size_t(n) * n;
  const unsigned nb = unsigned((n3 + 255) / 256);
  const double2* f = reinterpret_cast<const double2*>(folded_d);
  const bool fast = n == 64 || n == 128 || n == 256 || n == 512 || n == 1024;
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, vp_align256(n3 * (fast ? sizeof(float2) : sizeof(double2))) + vp_align256(sizeof(double2) * n) + 1024));
  if (fast) {
    float2* z = static_cast<float2*>(vp_arena_alloc(ctx, n3 * sizeof(float2)));
    float2* tw = static_cast<float2*>(vp_arena_alloc(ctx, sizeof(float2) * n));
    VP_REQUIRE(z && tw, "vp_fold_power: arena carve failed");
    std::vector<float2> twh(n);
    for (int q = 0; q < n; ++q) twh[q] = make_float2(float(cos(2.0 * M_PI * q / n)), float(-sin(2.0 * M_PI * q / n)));
    VP_CUDA(cudaMemcpyAsync(tw, twh.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, st));
    VP_CUDA(cudaStreamSynchronize(st));   // twh is a stack-lifetime host buffer
    for (int c = 0; c < ncomp; ++c) {
      k_comp_to_zf<<<nb, 256, 0, st>>>(f, ncomp, c, n3, z);
      switch (n) {
        case 64: VP_TRY((c2c_pow2<16, 4, 1, 32>(ctx, z, n, tw, st))); break;
        case 128: VP_TRY((c2c_pow2<16, 8, 1, 32>(ctx, z, n, tw, st))); break;
        case 256: VP_TRY((c2c_pow2<16, 16, 1, 16>(ctx, z, n, tw, st))); break;
        case 512: VP_TRY((c2c_pow2<16, 16, 2, 8>(ctx, z, n, tw, st))); break;
        default: VP_TRY((c2c_pow2<16, 16, 4, 8>(ctx, z, n, tw, st))); break;
      }
      k_accum_power_f<<<nb, 256, 0, st>>>(z, P_d, n3, c == 0);
      ctx->n_launch += 2;
    }
  } else {
    double2* z = static_cast<double2*>(vp_arena_alloc(ctx, n3 * sizeof(double2)));
    double2* tw = static_cast<double2*>(vp_arena_alloc(ctx, sizeof(double2) * n));
    VP_REQUIRE(z && tw, "vp_fold_power: arena carve failed");
    std::vector<double2> twh(n);
    for (int q = 0; q < n; ++q) twh[q] = make_double2(cos(2.0 * M_PI * q / n), -sin(2.0 * M_PI * q / n));
    VP_CUDA(cudaMemcpyAsync(tw, twh.data(), sizeof(double2) * n, cudaMemcpyHostToDevice, st));
    VP_CUDA(cudaStreamSynchronize(st));
    const size_t smem = sizeof(double2) * n;
    VP_REQUIRE(smem <= 48 * 1024, "vp_fold_power: folded size %d has no transform (use a power of two 64..1024, or n <= 3072)", n);
    vp_stage stage(ctx, "fold_dft", st, 5 * ncomp);
    for (int c = 0; c < ncomp; ++c) {
      k_comp_to_z<<<nb, 256, 0, st>>>(f, ncomp, c, n3, z);
      k_dft_axis_d<<<unsigned(nn), 128, smem, st>>>(z, n, 1, n, nn, size_t(n), tw);       // z lines
      k_dft_axis_d<<<unsigned(nn), 128, smem, st>>>(z, n, size_t(n), n, nn, 1, tw);       // y lines
      k_dft_axis_d<<<unsigned(nn), 128, smem, st>>>(z, n, nn, n, size_t(n), 1, tw);       // x lines
      k_accum_power_d<<<nb, 256, 0, st>>>(z, P_d, n3, c == 0);
    }
  }
  VP_CHECK_LAUNCH();
  VP_CUDA(cudaStreamSynchronize(st));     // the scratch is released on return
  return VP_OK;
}
