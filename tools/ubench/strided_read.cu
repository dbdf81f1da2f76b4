// Microbenchmark: what does B200 deliver when a kernel reads (or writes) 64-byte pieces that are S bytes apart -- the access
// pattern of a line FFT along a non-contiguous axis (x pass: S = 4 MB at N = 1024; y pass: S = 4 KB)?
// Each CTA of 512 threads handles one "tile": 1024 pieces of 64 B at stride S (thread t reads 8 B of piece t/8 + 64*j, j = 0..15);
// consecutive CTAs take consecutive 64-byte columns, like the FFT kernels.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

template <bool WRITE>
__global__ void __launch_bounds__(512, 2) k_strided(float2* __restrict__ a, size_t stride_el, size_t tiles_per_row, float2* __restrict__ sink) {
  // tile = (row group g, column tile c): base = g * 1024 * stride ... we mimic: base = (blockIdx / tiles_per_row) * rowblock + (blockIdx % tiles_per_row) * 8
  const size_t g = blockIdx.x / tiles_per_row, c = blockIdx.x % tiles_per_row;
  const int t = threadIdx.x / 8, l = threadIdx.x % 8;
  float2* base = a + g * (stride_el * 1024) + c * 8 + l;     // assumes tiles_per_row * 8 <= stride_el
  float2 v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float2* p = base + size_t(j * 64 + t) * stride_el;
    if (WRITE) *p = make_float2(float(j), float(t)); else v[j] = *p;
  }
  if (!WRITE) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += v[j].x + v[j].y;
    if (s == 123.456f) sink[0] = make_float2(s, s);
  }
}

int main() {
  const size_t total = size_t(1) << 32;          // 4 GiB array
  float2* a;
  float2* sink;
  cudaMalloc(&a, total);
  cudaMalloc(&sink, 64);
  cudaMemset(a, 0, total);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t nel = total / 8;
  // stride S bytes; a "row" of the matrix is S bytes = S/64 column tiles; 1024 rows form a group of 1024*S bytes
  for (size_t S : {size_t(4096), size_t(65536), size_t(1) << 20, size_t(4) << 20}) {
    const size_t stride_el = S / 8;
    const size_t tiles_per_row = S / 64;
    const size_t groups = nel / (stride_el * 1024);
    const size_t ntiles = groups * tiles_per_row;
    for (int wr = 0; wr < 2; ++wr) {
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (wr) k_strided<true><<<unsigned(ntiles), 512>>>(a, stride_el, tiles_per_row, sink);
        else k_strided<false><<<unsigned(ntiles), 512>>>(a, stride_el, tiles_per_row, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2) printf("{\"test\": \"%s_64B_pieces\", \"stride_bytes\": %zu, \"ms\": %.3f, \"GBs\": %.0f}\n", wr ? "write" : "read", S, ms, total / ms * 1e-6);
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"status\": \"%s\"}\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
