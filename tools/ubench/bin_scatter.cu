// Ablation microbenchmark of the bucket pass (csrc/nn_grid.cu k_bin_scatter): the same kernel on synthetic uniform particles
// with pieces switched off, to see which piece the 25 ms at 2^30 particles belong to.  f32 input with payload, 1024 buckets,
// 4096-particle tiles, ~1 particle per cell -- the statistics of cfg4 at a quarter of its size by default.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin_scatter bin_scatter.cu ; run on one B200:  ./bin_scatter [log2 np]
// Flags (template F): 1 VEC loads | 2 no payload loads | 4 no record stores | 8 ranking atomics without return value |
//                     16 L2 prefetch of the payload during phase 1 | 32 all payload loads before the first store |
//                     64 no global cursor atomics (phase 2 reads the cursor) | 128 record staged through shared memory (runs)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct Grid { double ox, oy, oz, ihx, ihy, ihz; int gx, gy, gz; int bshift; uint32_t nb; };
constexpr int kFixBits = 21;
constexpr uint32_t kFixMax = (1u << kFixBits) - 1u;
constexpr int kTile = 4096, kSub = 8, kClu = 8;
struct __align__(32) Rec32 { uint32_t w[8]; };

__device__ __forceinline__ int cell_fix(double x, double o, double ih, int g, uint32_t& fix, bool& far) {
  const double f = __dmul_rn(__dsub_rn(x, o), ih);
  int c;
  if (!(f > 0.0)) c = 0;
  else if (f >= double(g)) c = g - 1;
  else c = int(f);
  double u = __dsub_rn(f, double(c));
  if (!(u >= 0.0)) { far = far || (u < 0.0) || (u != u); u = 0.0; }
  if (u >= 1.0) { far = true; u = 1.0; }
  uint32_t q = uint32_t(__dmul_rn(u, 2097152.0));
  fix = q > kFixMax ? kFixMax : q;
  return c;
}
__device__ __forceinline__ void st256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7) : "memory");
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void k_fill(float* pos, float* vel, float* rho, int64_t np) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= np) return;
  for (int c = 0; c < 3; ++c) {
    pos[3 * i + c] = (hash32(uint32_t(i) * 3u + c + 0x9e3779b9u) >> 8) * (1.0f / 16777216.0f);
    vel[3 * i + c] = (hash32(uint32_t(i) * 3u + c + 0x12345u) >> 8) * (1.0f / 16777216.0f) - 0.5f;
  }
  rho[i] = 0.5f + (hash32(uint32_t(i) + 0xabcdefu) >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ uint32_t lin_of(float xf, float yf, float zf, const Grid& g, uint32_t& fx, uint32_t& fy, uint32_t& fz, bool& far) {
  const int cx = cell_fix(xf, g.ox, g.ihx, g.gx, fx, far), cy = cell_fix(yf, g.oy, g.ihy, g.gy, fy, far), cz = cell_fix(zf, g.oz, g.ihz, g.gz, fz, far);
  return (uint32_t(cx) * uint32_t(g.gy) + uint32_t(cy)) * uint32_t(g.gz) + uint32_t(cz);
}
__global__ void k_hist(const float* pos, int64_t np, Grid g, uint32_t* hist) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= np) return;
  uint32_t fx, fy, fz; bool far = false;
  const uint32_t lin = lin_of(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], g, fx, fy, fz, far);
  const uint32_t sub = uint32_t(((i / kTile) / kClu) % kSub);
  atomicAdd(hist + (lin >> g.bshift) * kSub + sub, 1u);
}

template <int F>
__global__ void __launch_bounds__(1024, 2) k_scatter(const float* __restrict__ pos, const float* __restrict__ vel, const float* __restrict__ rho,
                                                    float lcell3, int64_t np, Grid g, uint32_t* __restrict__ cursor, Rec32* __restrict__ rec1) {
  constexpr bool VEC = F & 1, NOPAY = F & 2, NOST = F & 4, NORANK = F & 8, PREF = F & 16, HOIST = F & 32, NOCUR = F & 64;
  extern __shared__ uint32_t sh_cnt[];
  for (uint32_t b = threadIdx.x; b < g.nb; b += 1024) sh_cnt[b] = 0u;
  __syncthreads();
  const uint32_t sub = (uint32_t(blockIdx.x) / kClu) % kSub;
  const int64_t base = int64_t(blockIdx.x) * kTile;
  uint32_t ra[4][3], slot[4];
  float pv[12];
  if (VEC) {
    const uint4* q = reinterpret_cast<const uint4*>(pos + 3 * (base + 4 * int64_t(threadIdx.x)));
    uint4 t0 = q[0], t1 = q[1], t2 = q[2];
    memcpy(pv, &t0, 16); memcpy(pv + 4, &t1, 16); memcpy(pv + 8, &t2, 16);
  }
  if (PREF) {
    const int64_t i4 = base + 4 * int64_t(threadIdx.x);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(vel + 3 * i4));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(vel + 3 * i4 + 11));
    if ((threadIdx.x & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(rho + i4));
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t i = VEC ? base + 4 * int64_t(threadIdx.x) + r : base + r * 1024 + threadIdx.x;
    float x, y, z;
    if (VEC) { x = pv[3 * r]; y = pv[3 * r + 1]; z = pv[3 * r + 2]; }
    else { x = pos[3 * i]; y = pos[3 * i + 1]; z = pos[3 * i + 2]; }
    uint32_t fx, fy, fz; bool far = false;
    const uint32_t lin = lin_of(x, y, z, g, fx, fy, fz, far);
    if (NORANK) { atomicAdd(&sh_cnt[lin >> g.bshift], 1u); slot[r] = 0u; }
    else slot[r] = atomicAdd(&sh_cnt[lin >> g.bshift], 1u);
    const unsigned long long w = (unsigned long long)fx | ((unsigned long long)fy << kFixBits) | ((unsigned long long)fz << (2 * kFixBits));
    ra[r][0] = uint32_t(w); ra[r][1] = uint32_t(w >> 32); ra[r][2] = lin;
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < g.nb; b += 1024) {
    const uint32_t c = sh_cnt[b];
    if (c) sh_cnt[b] = NOCUR ? cursor[b * kSub + sub] : atomicAdd(cursor + b * kSub + sub, c);
  }
  __syncthreads();
  float pay[HOIST ? 16 : 1];
  if (HOIST && !NOPAY) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t i = VEC ? base + 4 * int64_t(threadIdx.x) + r : base + r * 1024 + threadIdx.x;
      pay[4 * r] = vel[3 * i]; pay[4 * r + 1] = vel[3 * i + 1]; pay[4 * r + 2] = vel[3 * i + 2]; pay[4 * r + 3] = rho[i];
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t i = VEC ? base + 4 * int64_t(threadIdx.x) + r : base + r * 1024 + threadIdx.x;
    const uint32_t dst = sh_cnt[ra[r][2] >> g.bshift] + slot[r];
    float vx = 1.f, vy = 2.f, vz = 3.f, rr = 1.5f;
    if (!NOPAY) {
      if (HOIST) { vx = pay[4 * r]; vy = pay[4 * r + 1]; vz = pay[4 * r + 2]; rr = pay[4 * r + 3]; }
      else { vx = vel[3 * i]; vy = vel[3 * i + 1]; vz = vel[3 * i + 2]; rr = rho[i]; }
    }
    vx = (vx * rr) / rr; vy = (vy * rr) / rr; vz = (vz * rr) / rr;
    const float m = rr * lcell3;
    if (!NOST || dst == 0xfffffff7u)
      st256(rec1 + (NOST ? 0 : dst), ra[r][0], ra[r][1], ra[r][2], uint32_t(i), __float_as_uint(vx), __float_as_uint(vy), __float_as_uint(vz), __float_as_uint(m));
  }
}

// Write-combining variant: the tile's records are staged in shared memory in bucket order (position = exclusive prefix of the
// tile's bucket counts + rank), half a tile at a time, and written out by consecutive lanes -- the lanes of a warp then cover
// ~8 runs of consecutive destinations instead of 32 unrelated sectors.
// flags: 1 VEC (always), 256 = 2048-record window per round (two rounds), else one round of 4096 records (1 CTA/SM)
template <int WIN>
__global__ void __launch_bounds__(1024, WIN == 2048 ? 2 : 1) k_scatter_wc(const float* __restrict__ pos, const float* __restrict__ vel,
                                                                         const float* __restrict__ rho, float lcell3, int64_t np, Grid g,
                                                                         uint32_t* __restrict__ cursor, Rec32* __restrict__ rec1) {
  extern __shared__ __align__(16) unsigned char wsm[];
  uint4* sA = reinterpret_cast<uint4*>(wsm);                 // [WIN] search half
  uint4* sB = sA + WIN;                                      // [WIN] payload half
  uint32_t* sh_cnt = reinterpret_cast<uint32_t*>(sB + WIN);  // [1024] counts, then tile-local exclusive prefix
  uint32_t* sh_delta = sh_cnt + 1024;                        // [1024] global run start - tile-local prefix
  uint32_t* sh_w = sh_delta + 1024;                          // [33] warp totals
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  sh_cnt[tid] = 0u;
  __syncthreads();
  const uint32_t sub = (uint32_t(blockIdx.x) / kClu) % kSub;
  const int64_t base = int64_t(blockIdx.x) * kTile;
  uint32_t ra[4][3], slot[4];
  {
    float pv[12];
    const uint4* q = reinterpret_cast<const uint4*>(pos + 3 * (base + 4 * int64_t(tid)));
    uint4 t0 = q[0], t1 = q[1], t2 = q[2];
    memcpy(pv, &t0, 16); memcpy(pv + 4, &t1, 16); memcpy(pv + 8, &t2, 16);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      uint32_t fx, fy, fz; bool far = false;
      const uint32_t lin = lin_of(pv[3 * r], pv[3 * r + 1], pv[3 * r + 2], g, fx, fy, fz, far);
      slot[r] = atomicAdd(&sh_cnt[lin >> g.bshift], 1u);
      const unsigned long long w = (unsigned long long)fx | ((unsigned long long)fy << kFixBits) | ((unsigned long long)fz << (2 * kFixBits));
      ra[r][0] = uint32_t(w); ra[r][1] = uint32_t(w >> 32); ra[r][2] = lin;
    }
  }
  __syncthreads();
  // exclusive scan of the 1024 bucket counts (thread b owns bucket b), run claim
  const uint32_t c = sh_cnt[tid];
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) sh_w[wp] = incl;
  __syncthreads();
  if (wp == 0) {
    const uint32_t v = sh_w[lane];
    uint32_t is = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, is, o); if (lane >= o) is += t; }
    sh_w[lane] = is - v;
    if (lane == 31) sh_w[32] = is;
  }
  __syncthreads();
  const uint32_t tp = sh_w[wp] + incl - c;
  const uint32_t gb = c ? atomicAdd(cursor + tid * kSub + sub, c) : 0u;
  sh_cnt[tid] = tp;
  sh_delta[tid] = gb - tp;
  const uint32_t nv = sh_w[32];
  __syncthreads();
  for (uint32_t w0 = 0; w0 < nv; w0 += WIN) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t p = sh_cnt[ra[r][2] >> g.bshift] + slot[r] - w0;
      if (p < uint32_t(WIN)) {
        const int64_t i = base + 4 * int64_t(tid) + r;
        float vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
        const float rr = rho[i];
        vx = (vx * rr) / rr; vy = (vy * rr) / rr; vz = (vz * rr) / rr;
        sA[p] = make_uint4(ra[r][0], ra[r][1], ra[r][2], uint32_t(i));
        sB[p] = make_uint4(__float_as_uint(vx), __float_as_uint(vy), __float_as_uint(vz), __float_as_uint(rr * lcell3));
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WIN / 1024; ++k) {
      const uint32_t q = uint32_t(tid) + k * 1024;
      if (w0 + q < nv) {
        const uint4 a = sA[q], b = sB[q];
        const uint32_t dst = sh_delta[a.z >> g.bshift] + w0 + q;
        st256(rec1 + dst, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w);
      }
    }
    if (w0 + WIN < nv) __syncthreads();
  }
}

// every slot of the output holds a record whose bucket owns the slot; the indices sum up to np(np-1)/2
__global__ void k_check(const Rec32* rec, int64_t np, Grid g, const uint32_t* bstart, unsigned long long* out) {
  int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= np) return;
  const uint32_t b = rec[j].w[2] >> g.bshift;
  if (b >= g.nb || j < bstart[b] || j >= bstart[b + 1]) atomicAdd(out, 1ull);
  atomicAdd(out + 1, (unsigned long long)rec[j].w[3]);
}

int main(int argc, char** argv) {
  const int nbits = argc > 1 ? atoi(argv[1]) : 28;
  const int64_t np = int64_t(1) << nbits;
  Grid g;
  g.ox = g.oy = g.oz = 0.0;
  const int bx = (nbits + 2) / 3, by = (nbits + 1) / 3, bz = nbits / 3;
  g.gx = 1 << bx; g.gy = 1 << by; g.gz = 1 << bz;
  g.ihx = g.gx; g.ihy = g.gy; g.ihz = g.gz;
  g.nb = 1024; g.bshift = nbits - 10;
  float *pos, *vel, *rho; Rec32* rec; uint32_t *hist, *cursor, *bstart; unsigned long long* chk;
  cudaMalloc(&pos, np * 12); cudaMalloc(&vel, np * 12); cudaMalloc(&rho, np * 4); cudaMalloc(&rec, np * 32);
  cudaMalloc(&hist, g.nb * kSub * 4); cudaMalloc(&cursor, g.nb * kSub * 4); cudaMalloc(&bstart, (g.nb + 1) * 4); cudaMalloc(&chk, 16);
  k_fill<<<unsigned((np + 255) / 256), 256>>>(pos, vel, rho, np);
  cudaMemset(hist, 0, g.nb * kSub * 4);
  k_hist<<<unsigned((np + 255) / 256), 256>>>(pos, np, g, hist);
  std::vector<uint32_t> h(g.nb * kSub), c0(g.nb * kSub), bs(g.nb + 1);
  cudaMemcpy(h.data(), hist, h.size() * 4, cudaMemcpyDeviceToHost);
  uint32_t acc = 0;
  for (size_t i = 0; i < h.size(); ++i) { if (i % kSub == 0) bs[i / kSub] = acc; c0[i] = acc; acc += h[i]; }
  bs[g.nb] = acc;
  cudaMemcpy(bstart, bs.data(), bs.size() * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const unsigned ntiles = unsigned(np / kTile);
  auto run = [&](auto kern, int flags, size_t smem = 4096) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    float best = 1e9f, sum = 0.f;
    for (int it = 0; it < 4; ++it) {
      cudaMemcpy(cursor, c0.data(), c0.size() * 4, cudaMemcpyHostToDevice);
      cudaEventRecord(e0);
      kern<<<ntiles, 1024, smem>>>(pos, vel, rho, 1e-3f, np, g, cursor, rec);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it) { sum += ms; best = ms < best ? ms : best; }
    }
    unsigned long long res[2] = {0, 0};
    if (!(flags & (4 | 8 | 64))) {
      cudaMemset(chk, 0, 16);
      k_check<<<unsigned((np + 255) / 256), 256>>>(rec, np, g, bstart, chk);
      cudaMemcpy(res, chk, 16, cudaMemcpyDeviceToHost);
    }
    const bool ok = (flags & (4 | 8 | 64)) || (res[0] == 0 && res[1] == (unsigned long long)np * (np - 1) / 2);
    printf("{\"flags\": %d, \"ms\": %.3f, \"best_ms\": %.3f, \"ms_at_2^30\": %.2f, \"check\": \"%s\", \"err\": \"%s\"}\n", flags, sum / 3, best,
           sum / 3 * double(1ull << 30) / np, (flags & (4 | 8 | 64)) ? "n/a" : (ok ? "ok" : "BAD"), cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
  };
#define RUN(F) run(k_scatter<F>, F)
  run(k_scatter_wc<2048>, 128 | 256 | 1, 2048 * 32 + 8192 + 256);
  run(k_scatter_wc<4096>, 128 | 1, 4096 * 32 + 8192 + 256);
  RUN(0); RUN(1); RUN(1 | 2); RUN(1 | 4); RUN(1 | 8 | 4); RUN(1 | 2 | 4); RUN(1 | 2 | 4 | 8); RUN(1 | 2 | 4 | 8 | 64); RUN(1 | 16); RUN(1 | 32); RUN(1 | 64 | 4); RUN(1 | 2 | 8 | 64);
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"status\": \"%s\"}\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
