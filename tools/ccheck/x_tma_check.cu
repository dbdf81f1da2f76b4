// Stand-alone C-ABI check of the TMA-fed x pass (no Python, starts in about a second): vp_pk_fields (blocked layout,
// k_fft_x_pow_tma) against vp_pk_dist_local + vp_pk_dist_final with one rank (row-major layout, per-thread-load x pass) on the
// same pseudo-random cubes -- mode counts identical, shell sums equal to 1e-12 (f64 atomics order).
// Build (from the repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/ccheck/x_tma_check tools/ccheck/x_tma_check.cu \
//        -Llarge-velocity-power-spectrum_b200 -l:libvpower_b200.so -Xlinker -rpath -Xlinker '$ORIGIN/../../large-velocity-power-spectrum_b200'
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../include/vpower_b200.h"

#define CK(x) do { int r_ = (x); if (r_ != 0) { printf("{\"check\": \"x_tma\", \"status\": \"FAIL\", \"call\": \"%s\", \"rc\": %d, \"err\": \"%s\"}\n", #x, r_, vp_last_error()); return 1; } } while (0)

__global__ void k_fill(float* a, size_t n, uint32_t seed) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = uint32_t(i) * 2654435761u + seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  a[i] = (x >> 8) * (1.0f / 16777216.0f) - 0.5f;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 256, ncomp = argc > 2 ? atoi(argv[2]) : 3;
  const size_t n3 = size_t(N) * N * N;
  std::vector<double> k(N), edges;
  for (int i = 0; i < N; ++i) k[i] = 2.0 * M_PI * (i < (N + 1) / 2 ? i : i - N);
  for (double e = 0.5; e < N / 2 + 1; e += 1.0) edges.push_back(e * 2.0 * M_PI);
  const int nbins = int(edges.size()) - 1;
  vp_ctx* ctx = nullptr;
  CK(vp_ctx_create(0, &ctx));
  vp_pk_plan *pa = nullptr, *pb = nullptr;
  CK(vp_pk_plan_create(ctx, N, k.data(), edges.data(), nbins, &pa));
  CK(vp_pk_plan_create_dist(ctx, N, 1, 0, k.data(), edges.data(), nbins, &pb));
  float *fa[3], *fb[3], *send[3];
  for (int c = 0; c < ncomp; ++c) {
    cudaMalloc(&fa[c], n3 * 4); cudaMalloc(&fb[c], n3 * 4); cudaMalloc(&send[c], n3 * 4);
    k_fill<<<unsigned((n3 + 255) / 256), 256>>>(fa[c], n3, 17u + c);
    cudaMemcpy(fb[c], fa[c], n3 * 4, cudaMemcpyDeviceToDevice);
  }
  double *psa, *psb; uint64_t *nsa, *nsb;
  cudaMalloc(&psa, nbins * 8); cudaMalloc(&psb, nbins * 8); cudaMalloc(&nsa, nbins * 8); cudaMalloc(&nsb, nbins * 8);
  CK(vp_pk_fields(pa, fa, ncomp, psa, nsa, nullptr));
  CK(vp_pk_dist_local(pb, fb, ncomp, send, nullptr));
  CK(vp_pk_dist_final(pb, send, ncomp, psb, nsb, nullptr));
  std::vector<double> ha(nbins), hb(nbins); std::vector<uint64_t> na(nbins), nb(nbins);
  cudaMemcpy(ha.data(), psa, nbins * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), psb, nbins * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(na.data(), nsa, nbins * 8, cudaMemcpyDeviceToHost); cudaMemcpy(nb.data(), nsb, nbins * 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  // (the per-CTA partial sums meet in f64 atomics: the last bits of a shell sum depend on the order)
  int bad_p = 0, bad_n = 0; double tot = 0; uint64_t modes = 0;
  for (int j = 0; j < nbins; ++j) { bad_p += !(fabs(ha[j] - hb[j]) <= 1e-12 * fabs(hb[j])); bad_n += na[j] != nb[j]; tot += ha[j]; modes += na[j]; }
  int64_t lay[24]; CK(vp_fft_x_layout(N, N / 2, lay));
  const bool ok = e == cudaSuccess && !bad_p && !bad_n && modes > 0 && tot > 0;
  printf("{\"check\": \"x_tma\", \"N\": %d, \"ncomp\": %d, \"tma_kernel\": %lld, \"bins\": %d, \"modes\": %llu, \"psum_total\": %.9e, \"bins_differing\": %d, \"counts_differing\": %d, \"cuda\": \"%s\", \"status\": \"%s\"}\n",
         N, ncomp, (long long)lay[5], nbins, (unsigned long long)modes, tot, bad_p, bad_n, cudaGetErrorString(e), ok ? "PASS" : "FAIL");
  return ok ? 0 : 1;
}
