// Microbenchmark behind the cell-list build design: what does it cost on B200 to move 32-byte records
//   (a) by a random gather (the round-1 k_permute),            (b) by a random scatter of whole 32-byte sectors,
//   (c) inside windows that fit L1 / L2,                        (d) as staged runs of R consecutive records.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_gather scatter_gather.cu ; run on one B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

struct __align__(32) R32 { uint4 a, b; };

__device__ __forceinline__ uint32_t bij(uint32_t x, uint32_t mask, int bits) {   // bijection on [0, mask]
  x = (x * 0x9E3779B1u) & mask;
  x ^= x >> (bits / 2 + 1);
  x = (x * 0x85EBCA6Bu + 0x1234567u) & mask;
  x ^= x >> (bits / 2);
  x = (x * 0xC2B2AE35u) & mask;
  return x;
}
// destination of record i: windows of 2^wbits records stay in place, records are permuted inside their window
__device__ __forceinline__ size_t dest(size_t i, int wbits) {
  const uint32_t mask = (wbits >= 32) ? 0xffffffffu : ((1u << wbits) - 1u);
  return (i & ~size_t(mask)) | bij(uint32_t(i) & mask, mask, wbits);
}

__device__ __forceinline__ R32 ld256(const R32* p) {
  R32 r;
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st256(R32* p, const R32& r) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.a.x), "r"(r.a.y), "r"(r.a.z), "r"(r.a.w),
               "r"(r.b.x), "r"(r.b.y), "r"(r.b.z), "r"(r.b.w) : "memory");
}

template <int MODE>   // 0 gather 2x128, 1 gather 256, 2 scatter 2x128, 3 scatter 256, 4 copy
__global__ void __launch_bounds__(256) k_move(const R32* __restrict__ in, R32* __restrict__ out, size_t n, int wbits) {
  size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  if (MODE == 0) { size_t j = dest(i, wbits); R32 r; r.a = in[j].a; r.b = in[j].b; out[i] = r; }
  if (MODE == 1) { size_t j = dest(i, wbits); st256(out + i, ld256(in + j)); }
  if (MODE == 2) { size_t j = dest(i, wbits); R32 r = in[i]; out[j].a = r.a; out[j].b = r.b; }
  if (MODE == 3) { size_t j = dest(i, wbits); st256(out + j, ld256(in + i)); }
  if (MODE == 4) { st256(out + i, ld256(in + i)); }
}
// 16-byte records scattered (half sectors)
__global__ void __launch_bounds__(256) k_scatter16(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, int wbits) {
  size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) out[dest(i, wbits)] = in[i];
}
__global__ void __launch_bounds__(256) k_gather16(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, int wbits) {
  size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) out[i] = in[dest(i, wbits)];
}
// runs of 2^rbits consecutive records keep together; the runs are permuted over the whole array.  `shift` records of
// misalignment (0: runs start on a multiple of their size)
__global__ void __launch_bounds__(256) k_runs(const R32* __restrict__ in, R32* __restrict__ out, size_t n, int rbits, int nbits, int shift) {
  size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  size_t run = i >> rbits, off = i & ((size_t(1) << rbits) - 1);
  size_t drun = dest(run, nbits - rbits);
  size_t j = ((drun << rbits) + off + shift) & (n - 1);
  st256(out + j, ld256(in + i));
}

int main(int argc, char** argv) {
  const int nbits = argc > 1 ? atoi(argv[1]) : 28;
  const size_t n = size_t(1) << nbits;
  R32 *a, *b;
  cudaMalloc(&a, n * 32);
  cudaMalloc(&b, n * 32);
  cudaMemset(a, 1, n * 32);
  cudaMemset(b, 2, n * 32);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const unsigned nb = unsigned((n + 255) / 256);
  auto report = [&](const char* name, int p, float ms, double bytes_per_rec) {
    printf("{\"test\": \"%s\", \"param\": %d, \"ms\": %.3f, \"Grec_s\": %.2f, \"alg_GBs\": %.0f}\n", name, p, ms, n / ms * 1e-6,
           n * bytes_per_rec / ms * 1e-6);
    fflush(stdout);
  };
#define TIME(name, p, bpr, launch)                        \
  do {                                                    \
    launch; launch;                                       \
    cudaEventRecord(e0);                                  \
    for (int r = 0; r < 3; ++r) { launch; }               \
    cudaEventRecord(e1);                                  \
    cudaEventSynchronize(e1);                             \
    float ms;                                             \
    cudaEventElapsedTime(&ms, e0, e1);                    \
    report(name, p, ms / 3, bpr);                         \
  } while (0)
  TIME("copy256", 0, 64.0, (k_move<4><<<nb, 256>>>(a, b, n, 0)));
  const int wins[] = {12, 13, 15, 17, 20, 22, 32};
  for (int w : wins) {
    int wb = w > nbits ? nbits : w;
    TIME("gather_2x128_window", wb, 64.0, (k_move<0><<<nb, 256>>>(a, b, n, wb)));
    TIME("gather_256_window", wb, 64.0, (k_move<1><<<nb, 256>>>(a, b, n, wb)));
    TIME("scatter_2x128_window", wb, 64.0, (k_move<2><<<nb, 256>>>(a, b, n, wb)));
    TIME("scatter_256_window", wb, 64.0, (k_move<3><<<nb, 256>>>(a, b, n, wb)));
    TIME("scatter_16B_window", wb, 32.0, (k_scatter16<<<nb, 256>>>((const uint4*)a, (uint4*)b, n, wb)));
    TIME("gather_16B_window", wb, 32.0, (k_gather16<<<nb, 256>>>((const uint4*)a, (uint4*)b, n, wb)));
  }
  for (int rb = 1; rb <= 6; ++rb) {
    TIME("runs_aligned", 1 << rb, 64.0, (k_runs<<<nb, 256>>>(a, b, n, rb, nbits, 0)));
    TIME("runs_shifted1", 1 << rb, 64.0, (k_runs<<<nb, 256>>>(a, b, n, rb, nbits, 1)));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"status\": \"%s\"}\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
