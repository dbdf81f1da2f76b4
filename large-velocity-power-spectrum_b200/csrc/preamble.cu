// Snapshot preamble on the device: what GasParticles.shift_to_origin / remove_bulk_velocity
// (vpower/interp.py:169-182) and the MPI script's preamble (scripts/parallel_optimized.py:278-288) do on the host with
// numpy -- two O(Np) reductions and two O(Np) updates, in place on the device arrays so that a snapshot that has been
// uploaded once never goes back to the host.
//   shift_to_origin      : pos[:, c] -= min(pos[:, c])                      (exact: a minimum and one subtraction in dtype T)
//   remove_bulk_velocity : v[:, c]  -= sum(m * v[:, c]) / sum(m)            (products in dtype T as numpy forms them; the
//                           sums are accumulated in f64 -- numpy's pairwise f32 sum differs from this by ~1e-7 relative)
#include <math.h>

#include "common.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) k_min3(const T* __restrict__ pos, int64_t np, double* __restrict__ out /*[3], preset to +inf*/) {
  double m[3] = {INFINITY, INFINITY, INFINITY};
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < np; i += int64_t(gridDim.x) * 256) {
#pragma unroll
    for (int c = 0; c < 3; ++c) m[c] = fmin(m[c], double(pos[3 * i + c]));
  }
  __shared__ double sm[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o; o >>= 1) m[c] = fmin(m[c], __shfl_xor_sync(0xffffffffu, m[c], o));
    if ((threadIdx.x & 31) == 0) sm[c][threadIdx.x >> 5] = m[c];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = sm[threadIdx.x][0];
    for (int w = 1; w < 8; ++w) v = fmin(v, sm[threadIdx.x][w]);
    // atomic min on the ordered bit pattern: doubles of one sign order like their integer images
    unsigned long long* a = reinterpret_cast<unsigned long long*>(out + threadIdx.x);
    unsigned long long old = *a;
    while (v < __longlong_as_double((long long)old)) {
      const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
      if (seen == old) break;
      old = seen;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_sub3(T* __restrict__ a, int64_t np, const double* __restrict__ d /*[3]*/) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= 3 * np) return;
  a[i] = a[i] - T(d[i % 3]);
}

template <typename T>
__global__ void __launch_bounds__(256) k_mv_sums(const T* __restrict__ vel, const T* __restrict__ mass, int64_t np,
                                                 double* __restrict__ out /*[4]: sum m*vx, m*vy, m*vz, m*/) {
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < np; i += int64_t(gridDim.x) * 256) {
    const T m = mass[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) s[c] += double(T(m * vel[3 * i + c]));
    s[3] += double(m);
  }
  __shared__ double sm[4][8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int o = 16; o; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
    if ((threadIdx.x & 31) == 0) sm[c][threadIdx.x >> 5] = s[c];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sm[threadIdx.x][w];
    atomicAdd(out + threadIdx.x, v);
  }
}

__global__ void k_bulk_from_sums(double* s /*[4] -> [3] bulk velocity*/) {
  if (threadIdx.x < 3 && blockIdx.x == 0) {
    const double M = s[3];
    s[threadIdx.x] = s[threadIdx.x] / M;
  }
}

template <typename T>
int preamble_typed(vp_ctx* ctx, T* pos, T* vel, const T* mass, int64_t np, int do_shift, int do_bulk, double* min_h, double* bulk_h,
                   cudaStream_t st) {
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, 1024));
  double* d = static_cast<double*>(vp_arena_alloc(ctx, 256));
  VP_REQUIRE(d, "vp_snapshot_preamble: arena carve failed");
  const double init[8] = {INFINITY, INFINITY, INFINITY, 0.0, 0.0, 0.0, 0.0, 0.0};
  VP_CUDA(cudaMemcpyAsync(d, init, sizeof init, cudaMemcpyHostToDevice, st));
  const unsigned grid = unsigned(ctx->sm_count * 8);
  const unsigned nb3 = unsigned((3 * np + 255) / 256);
  if (np > 0 && do_bulk) {
    VP_REQUIRE(vel && mass, "vp_snapshot_preamble: bulk-velocity removal needs velocities and masses");
    vp_stage stage(ctx, "k0_bulk_velocity", st, 3, double(np) * sizeof(T) * (4.0 + 6.0));
    k_mv_sums<T><<<grid, 256, 0, st>>>(vel, mass, np, d + 4);
    k_bulk_from_sums<<<1, 32, 0, st>>>(d + 4);
    k_sub3<T><<<nb3, 256, 0, st>>>(vel, np, d + 4);
  }
  if (np > 0 && do_shift) {
    VP_REQUIRE(pos, "vp_snapshot_preamble: shift needs positions");
    vp_stage stage(ctx, "k0_shift_to_origin", st, 2, double(np) * sizeof(T) * (3.0 + 6.0));
    k_min3<T><<<grid, 256, 0, st>>>(pos, np, d);
    k_sub3<T><<<nb3, 256, 0, st>>>(pos, np, d);
  }
  VP_CHECK_LAUNCH();
  double h[8];
  VP_CUDA(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaStreamSynchronize(st));
  if (min_h) for (int c = 0; c < 3; ++c) min_h[c] = do_shift ? h[c] : 0.0;
  if (bulk_h) for (int c = 0; c < 3; ++c) bulk_h[c] = do_bulk ? h[4 + c] : 0.0;
  return VP_OK;
}

}  // namespace

extern "C" int vp_snapshot_preamble(vp_ctx* ctx, void* pos_d, void* vel_d, const void* mass_d, int dtype, int64_t np, int do_shift,
                                    int do_bulk, double* min_h, double* bulk_h, void* stream) {
  VP_REQUIRE(ctx && np >= 0, "vp_snapshot_preamble: bad argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == VP_F32)
    return preamble_typed<float>(ctx, static_cast<float*>(pos_d), static_cast<float*>(vel_d), static_cast<const float*>(mass_d), np,
                                 do_shift, do_bulk, min_h, bulk_h, st);
  if (dtype == VP_F64)
    return preamble_typed<double>(ctx, static_cast<double*>(pos_d), static_cast<double*>(vel_d), static_cast<const double*>(mass_d), np,
                                  do_shift, do_bulk, min_h, bulk_h, st);
  vp_set_error("vp_snapshot_preamble: unknown dtype %d", dtype);
  return VP_ERR_ARG;
}
