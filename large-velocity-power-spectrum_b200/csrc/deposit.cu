// K2: nearest-grid-point deposit, the device form of deposit_to_grid (vpower/interp.py:996-1015).
//   index = int((pos // Lcell) % Nsize)   with numpy's floor-division / remainder semantics, evaluated in
//   the dtype of `pos` (numpy demotes the Python-float Lcell to float32 for a float32 array);
//   np.add.at(grid, index, f)  -> f64 accumulation (order differs: sums agree to rounding, integer-valued
//   weights agree exactly).
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ double fmod_t(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ float fmod_t(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double floor_t(double a) { return floor(a); }
__device__ __forceinline__ float floor_t(float a) { return floorf(a); }
__device__ __forceinline__ double copysign_t(double a, double b) { return copysign(a, b); }
__device__ __forceinline__ float copysign_t(float a, float b) { return copysignf(a, b); }

// numpy npy_divmod (numpy/_core/src/npymath/npy_math_internal.h.src): floor division built on an exact fmod
template <typename T>
__device__ __forceinline__ T npy_floor_divide(T a, T b, T* modulus) {
  T mod = fmod_t(a, b);
  if (!b) { *modulus = mod; return a / b; }
  T div = (a - mod) / b;
  if (mod) {
    if ((b < 0) != (mod < 0)) { mod += b; div -= T(1); }
  } else {
    mod = copysign_t(T(0), b);
  }
  T fd;
  if (div) {
    fd = floor_t(div);
    if (div - fd > T(0.5)) fd += T(1);
  } else {
    fd = copysign_t(T(0), a / b);
  }
  *modulus = mod;
  return fd;
}

template <typename T>
__device__ __forceinline__ int cell_index(T x, T lcell, T n) {
  T m;
  T q = npy_floor_divide<T>(x, lcell, &m);
  npy_floor_divide<T>(q, n, &m);  // remainder of q by N
  return int(m);
}

template <typename T>
__global__ void __launch_bounds__(256) k_deposit(const T* __restrict__ pos, int64_t np, const double* __restrict__ w, int C, int N,
                                                  T lcell, double* __restrict__ grid) {
  int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= np) return;
  int i = cell_index<T>(pos[3 * p], lcell, T(N));
  int j = cell_index<T>(pos[3 * p + 1], lcell, T(N));
  int k = cell_index<T>(pos[3 * p + 2], lcell, T(N));
  if (unsigned(i) >= unsigned(N) || unsigned(j) >= unsigned(N) || unsigned(k) >= unsigned(N)) return;  // NaN / inf positions
  size_t cell = (size_t(i) * N + j) * N + k;
  for (int c = 0; c < C; ++c) atomicAdd(grid + cell * C + c, w[size_t(p) * C + c]);
}

}  // namespace

extern "C" int vp_deposit_ngp(vp_ctx* ctx, const void* pos_d, int pos_dtype, int64_t np, const double* w_d, int ncomp, int N,
                              double Lbox, double* grid_d, void* stream) {
  VP_REQUIRE(ctx && pos_d && w_d && grid_d, "vp_deposit_ngp: null argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_REQUIRE(np >= 0 && ncomp >= 1 && N >= 1, "vp_deposit_ngp: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VP_CUDA(cudaMemsetAsync(grid_d, 0, sizeof(double) * size_t(N) * N * N * ncomp, st));
  if (np == 0) return VP_OK;
  const double lcell = Lbox / double(N);
  unsigned nb = unsigned((np + 255) / 256);
  vp_stage stage(ctx, "k2_deposit_ngp", st, 1, double(np) * ((pos_dtype == VP_F64 ? 24.0 : 12.0) + 16.0 * ncomp));
  if (pos_dtype == VP_F32)
    k_deposit<float><<<nb, 256, 0, st>>>(static_cast<const float*>(pos_d), np, w_d, ncomp, N, float(lcell), grid_d);
  else if (pos_dtype == VP_F64)
    k_deposit<double><<<nb, 256, 0, st>>>(static_cast<const double*>(pos_d), np, w_d, ncomp, N, lcell, grid_d);
  else {
    vp_set_error("vp_deposit_ngp: unknown dtype %d", pos_dtype);
    return VP_ERR_ARG;
  }
  VP_CHECK_LAUNCH();
  return VP_OK;
}
