// Standalone pieces of the k-shell sampling stage (the fused fast path lives in fft3d.cu):
//   vp_k_magnitude   -- the |k| column of _pair_power (vpower/interp.py:1448-1460)
//   vp_hist_weighted -- _hist_sample's two np.histogram calls (interp.py:1474-1477) on arbitrary pairs
#include <math.h>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_kmag(const double* __restrict__ kx, const double* __restrict__ ky,
                                              const double* __restrict__ kz, int n, double* __restrict__ out) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(n) * n * n) return;
  int c = int(i % n);
  size_t t = i / n;
  int b = int(t % n), a = int(t / n);
  double s = __dadd_rn(__dadd_rn(__dmul_rn(kx[a], kx[a]), __dmul_rn(ky[b], ky[b])), __dmul_rn(kz[c], kz[c]));
  out[i] = __dsqrt_rn(s);
}

// numpy.histogram with explicit edges: bin j holds e[j] <= v < e[j+1]; the last bin also holds v == e[nb]
__global__ void __launch_bounds__(256) k_hist(const double* __restrict__ k, const double* __restrict__ w, int64_t n,
                                              const double* __restrict__ e, int nb, double* __restrict__ psum,
                                              unsigned long long* __restrict__ cnt) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = k[i];
  if (!(v >= e[0]) || !(v <= e[nb])) return;  // also drops NaN
  int lo = 0, hi = nb + 1;                     // number of edges <= v
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (e[mid] <= v) lo = mid + 1; else hi = mid;
  }
  int b = lo - 1;
  if (b == nb) b = nb - 1;
  atomicAdd(psum + b, w[i]);
  atomicAdd(cnt + b, 1ull);
}

}  // namespace

extern "C" int vp_k_magnitude(vp_ctx* ctx, const double* kx_h, const double* ky_h, const double* kz_h, int n, double* out_d,
                              void* stream) {
  VP_REQUIRE(ctx && kx_h && ky_h && kz_h && out_d && n > 0, "vp_k_magnitude: bad argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, size_t(3) * n * 8 + 1024));
  double* t = static_cast<double*>(vp_arena_alloc(ctx, size_t(3) * n * 8));
  VP_REQUIRE(t, "vp_k_magnitude: arena carve failed");
  VP_CUDA(cudaMemcpyAsync(t, kx_h, size_t(n) * 8, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaMemcpyAsync(t + n, ky_h, size_t(n) * 8, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaMemcpyAsync(t + 2 * n, kz_h, size_t(n) * 8, cudaMemcpyHostToDevice, st));
  size_t n3 = size_t(n) * n * n;
  ctx->n_launch += 1;
  k_kmag<<<unsigned((n3 + 255) / 256), 256, 0, st>>>(t, t + n, t + 2 * n, n, out_d);
  VP_CHECK_LAUNCH();
  VP_CUDA(cudaStreamSynchronize(st));  // the table block is released with the scope
  return VP_OK;
}

extern "C" int vp_hist_weighted(vp_ctx* ctx, const double* k_d, const double* w_d, int64_t n, const double* edges_h, int nbins,
                                double* psum_d, uint64_t* nsample_d, void* stream) {
  VP_REQUIRE(ctx && k_d && w_d && edges_h && psum_d && nsample_d && nbins >= 1 && n >= 0, "vp_hist_weighted: bad argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, size_t(nbins + 1) * 8 + 1024));
  double* e = static_cast<double*>(vp_arena_alloc(ctx, size_t(nbins + 1) * 8));
  VP_REQUIRE(e, "vp_hist_weighted: arena carve failed");
  VP_CUDA(cudaMemcpyAsync(e, edges_h, size_t(nbins + 1) * 8, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaMemsetAsync(psum_d, 0, size_t(nbins) * 8, st));
  VP_CUDA(cudaMemsetAsync(nsample_d, 0, size_t(nbins) * 8, st));
  if (n > 0) {
    ctx->n_launch += 1;
    k_hist<<<unsigned((n + 255) / 256), 256, 0, st>>>(k_d, w_d, n, e, nbins, psum_d, reinterpret_cast<unsigned long long*>(nsample_d));
    VP_CHECK_LAUNCH();
  }
  VP_CUDA(cudaStreamSynchronize(st));
  return VP_OK;
}
