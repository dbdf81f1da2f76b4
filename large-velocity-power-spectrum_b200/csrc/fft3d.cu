// K4 + K5: in-place 3-D r2c FFT (z, y, x line passes) with |F|^2 and k-shell binning fused into the x pass.
// See include/vpower_b200.h (vp_pk_plan_create / vp_pk_fields) and DESIGN.md for the data layout.
//
// Half-spectrum layout ("packed"): the real cube [N][N][N] f32 is overwritten by [N][N][N/2] complex64.
// After the z pass entry kz=0 of every line holds (Re X[0], Re X[N/2]) -- both are real for a real line --
// so the whole transform is in place with a power-of-two pitch.  The kz=0 column therefore carries two real
// planes a+ib through the y and x passes; they are separated at the end by Hermitian symmetry
// (k_plane_bin), all other columns are ordinary half-spectrum modes with weight 2.
// The x pass reads the (blocked) half spectrum through the TMA engine: k_fft_x_pow_tma, tensor-map boxes into a shared-memory
// ring with mbarrier completion; k_fft_x_pow (per-thread loads) serves the row-major layout, N = 250 and N = 2048.
#include <cuda.h>   // CUtensorMap and its enums (types only: the encoder is looked up through the runtime, no -lcuda)
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "fft_core.cuh"

struct vp_pk_plan {
  vp_ctx* ctx = nullptr;
  int N = 0, nbins = 0;
  bool pow2 = false;
  float2* tw_full = nullptr;  // W_N^m
  float2* tw_half = nullptr;  // W_{N/2}^m
  double* kk2 = nullptr;      // k[i]^2, f64
  double* thr = nullptr;      // nbins+1 thresholds on the squared magnitude
  float2* plane0 = nullptr;   // [3][N][N] x-pass output of the kz=0 column
  float inv_kf = 0.f;
  double e0 = 0.0, inv_de = 1.0;                 // linear estimate of the shell of |k|: (|k| - e0) * inv_de
  unsigned long long* ns_tiles = nullptr;        // mode counts of the x-pass tiles (geometry only; filled on first use)
  bool ns_valid = false;
  int ns_kz_offset = 0, ns_NZ = 0;
  // slab decomposition (one process per GPU): this rank owns x planes [rank*N/nranks, ...) before the exchange
  // and half-spectrum columns kz in [rank*kzc, (rank+1)*kzc) after it
  int nranks = 1, rank = 0;
  int kzc = 0;                // N/2/nranks
  // peer-to-peer transpose: receive buffers [N][N][kzc] complex64 per component, allocated here (cudaMalloc, exported
  // through CUDA IPC) and the same buffers of every other rank mapped into this process
  int p2p_ncomp = 0;
  float2* recv[3] = {nullptr, nullptr, nullptr};
  float2* peer[3][16];        // peer[c][d]: rank d's recv[c] (own pointer for d == rank)
  bool peer_open = false;
};

namespace {

// ------------------------------------------------------------------ z pass: N reals -> N/2 packed complex
template <int R1, int R2, int R3>
__global__ void __launch_bounds__(256) k_fft_z(float* __restrict__ data, const float2* __restrict__ tw_half,
                                               const float2* __restrict__ tw_full) {
  using F = LineFFT<R1, R2, R3, 1>;
  constexpr int L = F::L, T = F::T, LINES = 256 / T, XS = xsize<L, 1>();
  extern __shared__ float2 sm[];
  const int tid = threadIdx.x, ll = tid / T, t = tid % T;
  float2* g = reinterpret_cast<float2*>(data) + (size_t(blockIdx.x) * LINES + ll) * L;
  float2* s = sm + ll * XS;
  float2 v[F::P];
#pragma unroll
  for (int j = 0; j < F::P; ++j) v[j] = g[j * T + t];
  F::run(v, t, s, tw_half);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < F::P; ++j) s[F::kout(j, t)] = v[j];
  __syncthreads();
  for (int k = t; k <= L / 2; k += T) {
    if (k == 0) {
      float2 z0 = s[0];
      g[0] = make_float2(z0.x + z0.y, z0.x - z0.y);  // (X[0], X[N/2]) packed
    } else if (L % 2 == 0 && k == L / 2) {
      g[k] = cconj(s[k]);
    } else {
      float2 zk = s[k], zc = cconj(s[L - k]);
      float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
      float2 d = cmul_mi(csub(zk, zc));
      float2 o = make_float2(0.5f * d.x, 0.5f * d.y);
      float2 wo = cmul(__ldg(tw_full + k), o);
      g[k] = cadd(e, wo);
      g[L - k] = cconj(csub(e, wo));
    }
  }
}

// ------------------------------------------------------------------ y pass: lines strided by NZ, C columns per CTA
// Destination of the y pass, two layouts of the half spectrum [x][ky][kz]:
//   row-major (blocked = 0):  base[d] + (xoff + x_local)*N*kzc + ky*kzc + (kz - d*kzc)
//   blocked   (blocked = 1):  base[d] + ((((ky/16)*nxtot + xoff + x_local)*(kzc/C) + zt_local)*16 + ky%16)*C + c
//       -- 16 consecutive ky of one (x, kz tile) form ONE contiguous block of 16*C*8 bytes (1 KB at C = 8), and for a fixed
//       (ky, kz tile) consecutive x are only (kzc/C)*16*C*8 bytes apart (64 KB at N = 1024 on one GPU instead of the 4 MB
//       of the row-major layout).  Why: a line FFT along x reads 64-byte pieces one x-stride apart, and
//       profiles/r2_ubench_strided.jsonl shows what that costs on B200 -- 7.5 TB/s at 4 KB, 6.1 at 64 KB, 3.0 at 4 MB.
//       The y pass stores whole 1 KB blocks (a warp store = 256 contiguous bytes), which is also what NVLink wants.
//   single GPU        : base[0] = a second buffer (the y pass cannot be in place with this layout), blocked
//   NCCL exchange     : base[d] = send buffer block d ([dest][x_local][ky][kzc]), xoff = 0, row-major
//   peer-to-peer      : base[d] = rank d's receive buffer mapped through CUDA IPC, xoff = rank*N/nranks, blocked --
//                       the blocks are stored over NVLink where the x pass of rank d will read them
//   diagnostics       : base[0] = the field itself, row-major, in place (a CTA writes where it read)
template <int L>
__host__ __device__ constexpr int yb_of() { return L % 16 == 0 ? 16 : 10; }   // ky per block of the blocked layout (divides L)
struct YDest {
  float2* base[16];
  int xoff;
  int blocked;
  int nxtot;    // x planes of the destination array (N)
};

template <int R1, int R2, int R3, int C>
__global__ void __launch_bounds__(R2* R3* C, (R2 * R3 * C <= 512 ? 2 : 1)) k_fft_y(const float2* __restrict__ data, YDest dst, int NZ, int kzc,
                                                                                   const float2* __restrict__ tw) {
  using F = LineFFT<R1, R2, R3, C>;
  constexpr int L = F::L, T = F::T, kYB = yb_of<L>();
  extern __shared__ float2 sm[];
  const int tid = threadIdx.x, c = tid % C, t = tid / C;
  const int tiles = NZ / C;
  const int x = blockIdx.x / tiles, zt = blockIdx.x % tiles;
  const float2* base = data + size_t(x) * L * NZ + zt * C + c;
  float2 v[F::P];
#pragma unroll
  for (int j = 0; j < F::P; ++j) v[j] = base[size_t(j * T + t) * NZ];
  F::run(v, t, sm + c, tw);
  const int kz0 = zt * C, d = kz0 / kzc;
  if (dst.blocked) {
    const int tiles_r = kzc / C;
    float2* ob = dst.base[d] + ((size_t(dst.xoff) + x) * tiles_r + (kz0 - d * kzc) / C) * (kYB * C) + c;
    const size_t blk = size_t(dst.nxtot) * tiles_r * (kYB * C);      // one ky block of all x
#pragma unroll
    for (int j = 0; j < F::P; ++j) {
      const int ky = F::kout(j, t);
      ob[size_t(ky / kYB) * blk + (ky % kYB) * C] = v[j];
    }
  } else {
    float2* ob = dst.base[d] + (size_t(dst.xoff) + x) * L * kzc + (kz0 - d * kzc) + c;
#pragma unroll
    for (int j = 0; j < F::P; ++j) ob[size_t(F::kout(j, t)) * kzc] = v[j];
  }
}

// ------------------------------------------------------------------ x pass with |F|^2, then shell binning
struct FieldSet {
  float2* f[3];
  int n;
  int blocked;      // layout of the half spectrum the x pass reads (see YDest)
};

// x pass of 1..3 components of one (ky, kz-tile): the sum over components of |F|^2 is collected in a shared-memory tile
// [C columns][L rows] and written out with rows kx and -kx folded together, tile-major and column-major inside the tile:
//   P[((ky * tiles_z + zt) * C + c) * NRP + r],  r = 0 .. L/2  (NRP = L/2+1 rounded up to a multiple of 32; padding unwritten)
// so that the binning kernel reads 32 consecutive rows of ONE column per warp -- along a column the shell index never
// decreases, which is what its segmented reduction relies on.  The transform itself is never written back.
template <int L>
__host__ __device__ constexpr int nrp_of() { return ((L / 2 + 1) + 31) / 32 * 32; }

template <int R1, int R2, int R3, int C>
__global__ void __launch_bounds__(R2* R3* C, (R2 * R3 * C <= 512 ? 2 : 1)) k_fft_x_pow(FieldSet fs, int NZ, const float2* __restrict__ tw,
                                                                                       float2* __restrict__ plane0, int kz_offset,
                                                                                       float* __restrict__ P) {
  using F = LineFFT<R1, R2, R3, C>;
  constexpr int L = F::L, T = F::T, NT = T * C, XS = xsize<L, C>(), PP = L + 4, NR = L / 2 + 1, NRP = nrp_of<L>(), kYB = yb_of<L>();
  extern __shared__ float2 sm[];                                   // exchange area
  float* pt = reinterpret_cast<float*>(sm + XS);                   // [C][PP] sum over components of |F|^2
  const int tid = threadIdx.x, c = tid % C, t = tid / C;
  const int tiles_z = NZ / C;
  // (launch order: tried with the kYB tiles that share the 1 KB blocks of the blocked layout as neighbours, so that their
  // 64-byte pieces of one block are requested together -- no difference, 13.7 vs 13.5 ms at cfg4)
  const int ky = blockIdx.x / tiles_z, zt = blockIdx.x % tiles_z;
  // element (x, ky, zt, c): row-major  x*L*NZ + ky*NZ + zt*C + c;  blocked  ((((ky/16)*L + x)*tiles_z + zt)*16 + ky%16)*C + c
  const size_t xstride = fs.blocked ? size_t(tiles_z) * (kYB * C) : size_t(L) * NZ;
  const size_t tile_off = fs.blocked ? (size_t(ky / kYB) * L * tiles_z + zt) * (kYB * C) + (ky % kYB) * C : size_t(ky) * NZ + size_t(zt) * C;
  for (int comp = 0; comp < fs.n; ++comp) {
    const float2* base = fs.f[comp] + tile_off + c;
    float2 v[F::P];
#pragma unroll
    for (int j = 0; j < F::P; ++j) v[j] = base[size_t(j * T + t) * xstride];
    if (comp) __syncthreads();  // previous component's last exchange reads are done
    F::run(v, t, sm + c, tw);
    if (kz_offset + zt * C + c == 0) {
      float2* pl = plane0 + (size_t(comp) * L + ky) * L;
#pragma unroll
      for (int j = 0; j < F::P; ++j) pl[F::kout(j, t)] = v[j];
    }
    float* col = pt + c * PP;
    if (comp == 0) {
#pragma unroll
      for (int j = 0; j < F::P; ++j) col[F::kout(j, t)] = v[j].x * v[j].x + v[j].y * v[j].y;
    } else {
#pragma unroll
      for (int j = 0; j < F::P; ++j) col[F::kout(j, t)] += v[j].x * v[j].x + v[j].y * v[j].y;   // same thread, same slots
    }
  }
  __syncthreads();
  float* out = P + size_t(blockIdx.x) * (C * NRP);
  for (int i = tid; i < C * NRP; i += NT) {
    const int cc = i / NRP, r = i - cc * NRP;
    if (r < NR) {
      float q = pt[cc * PP + r];
      if (r > 0 && r < L / 2) q += pt[cc * PP + L - r];
      out[i] = q;
    }
  }
}

// ---- the same x pass fed by the TMA engine (blocked layout only) ------------------------------------------------------------
// One persistent CTA per shared-memory slot of the SM walks its tiles; the components of its tiles form one stream of "items".
// An item (one component of one (ky, kz-tile)) is L x C complex values that lie in the blocked half spectrum as L pieces of
// C*8 bytes, one per x plane: as a 5-D tensor  [ky block][x][kz tile][ky in block][c]  that is NBOX boxes of  BX x 1 x 1 x 1 x C
// elements, which ONE thread requests with cp.async.bulk.tensor (SASS UTMALDG) into a ring of NST item buffers; the
// boxes of an item complete on the item's mbarrier (complete_tx).  While the CTA transforms item m out of its registers, the
// TMA engine fills the buffers of items m+1 .. m+NST -- the loads the old kernel issued per thread (sixteen strided 8-byte
// LDGs, nothing in flight during the butterflies) are off the instruction stream altogether.  The sum over components of
// |F|^2 stays in registers (same thread, same slots for every component) and goes through the exchange area once per tile.
struct XMaps {
  CUtensorMap m[3];
};
template <int L>
__host__ __device__ constexpr int xbox_of() { return L <= 256 ? L : (L % 256 == 0 ? 256 : 250); }   // x planes per box (<= 256, divides L)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_box_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int R1, int R2, int R3, int C>
struct XTma {
  using F = LineFFT<R1, R2, R3, C>;
  static constexpr int L = F::L, T = F::T, NT = T * C, BX = xbox_of<L>(), NBOX = L / BX;
  static constexpr int BOXB = BX * C * 8, BOXP = (BOXB + 127) / 128 * 128;        // bytes of a box; its slot (TMA wants 128-byte aligned destinations)
  static constexpr int ITEMB = NBOX * BOXP;
  static constexpr int XSB = (xsize<L, C>() * 8 + 127) / 128 * 128;                // exchange area (also holds the [C][L+4] power tile)
  static constexpr int NST = (2 * ITEMB + XSB + 64 <= 227 * 1024) ? 2 : 1;         // item buffers in the ring
  static constexpr int SMEM = NST * ITEMB + XSB + 64;
  static constexpr bool ok = (C * 8) % 16 == 0 && (ITEMB + XSB + 64 <= 227 * 1024) && NT <= 1024;
};

// Tensor map of the blocked half spectrum [N (x)][N (ky)][NZ (kz)] of one rank (see YDest), as the x pass requests it: f32
// elements, innermost dimension first.  Host arithmetic only; vp_fft_x_layout() reports it, launch_x_pow() encodes it.
struct XGeom {
  unsigned long long gdim[5];       // [2C][kYB][NZ/C][N][N/kYB]
  unsigned long long gstride[4];    // bytes, dimensions 1..4
  unsigned box[5];                  // 2C x 1 x 1 x BX x 1
};
template <int L, int C>
XGeom x_geometry(int NZ) {
  constexpr int kYB = yb_of<L>();
  const unsigned long long tiles_z = (unsigned long long)(NZ / C), row = (unsigned long long)(C) * 8;
  XGeom g;
  g.gdim[0] = 2 * C; g.gdim[1] = kYB; g.gdim[2] = tiles_z; g.gdim[3] = L; g.gdim[4] = L / kYB;
  g.gstride[0] = row; g.gstride[1] = row * kYB; g.gstride[2] = row * kYB * tiles_z; g.gstride[3] = row * kYB * tiles_z * L;
  g.box[0] = 2 * C; g.box[1] = 1; g.box[2] = 1; g.box[3] = xbox_of<L>(); g.box[4] = 1;
  return g;
}

template <int R1, int R2, int R3, int C>
__global__ void __launch_bounds__(R2* R3* C) k_fft_x_pow_tma(const __grid_constant__ XMaps maps, int ncomp, int NZ,
                                                              const float2* __restrict__ tw, float2* __restrict__ plane0, int kz_offset,
                                                              float* __restrict__ P, int ntiles) {
  using X = XTma<R1, R2, R3, C>;
  using F = typename X::F;
  constexpr int L = F::L, T = F::T, NT = X::NT, PP = L + 4, NR = L / 2 + 1, NRP = nrp_of<L>(), kYB = yb_of<L>(), BX = X::BX, NST = X::NST;
  extern __shared__ __align__(128) unsigned char xsm[];
  unsigned char* ring = xsm;                                                  // [NST][NBOX][BOXP]
  float2* sm = reinterpret_cast<float2*>(xsm + NST * X::ITEMB);                // exchange area
  float* pt = reinterpret_cast<float*>(sm);                                    // power tile: the same memory, between two barriers
  const uint32_t bar0 = smem_u32(xsm + NST * X::ITEMB + X::XSB);               // NST mbarriers
  const int tid = threadIdx.x, c = tid % C, t = tid / C;
  const int tiles_z = NZ / C;
  const int my_tiles = (ntiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int nitems = my_tiles * ncomp;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // item m of this CTA = component m % ncomp of tile blockIdx.x + (m / ncomp) * gridDim.x; requested by thread 0 only
  auto request = [&](int tile, int comp, int slot) {
    const int ky = tile / tiles_z, zt = tile - ky * tiles_z;
    const uint32_t bar = bar0 + 8 * slot, dst = smem_u32(ring + slot * X::ITEMB);
    const CUtensorMap* mp = comp == 0 ? &maps.m[0] : (comp == 1 ? &maps.m[1] : &maps.m[2]);
    mbar_expect_tx(bar, uint32_t(X::NBOX * X::BOXB));
#pragma unroll
    for (int q = 0; q < X::NBOX; ++q) tma_box_5d(dst + q * X::BOXP, mp, bar, 0, ky % kYB, zt, q * BX, ky / kYB);
  };
  int rq_tile = blockIdx.x, rq_comp = 0, rq_m = 0;                             // next item to request (thread 0)
  if (tid == 0) {
    for (; rq_m < NST && rq_m < nitems; ++rq_m) {
      request(rq_tile, rq_comp, rq_m % NST);
      if (++rq_comp == ncomp) { rq_comp = 0; rq_tile += gridDim.x; }
    }
  }
  float acc[F::P];
  int tile = blockIdx.x, comp = 0;
  for (int m = 0; m < nitems; ++m) {
    const int slot = m % NST;
    mbar_wait(bar0 + 8 * slot, uint32_t(m / NST) & 1u);
    const unsigned char* it = ring + slot * X::ITEMB;
    float2 v[F::P];
#pragma unroll
    for (int j = 0; j < F::P; ++j) {
      const int x = j * T + t;
      v[j] = *reinterpret_cast<const float2*>(it + (x / BX) * X::BOXP + ((x % BX) * C + c) * 8);
    }
    __syncthreads();   // the item buffer is free again; the previous item's last exchange reads / power-tile reads are done
    if (tid == 0 && rq_m < nitems) {
      request(rq_tile, rq_comp, slot);
      ++rq_m;
      if (++rq_comp == ncomp) { rq_comp = 0; rq_tile += gridDim.x; }
    }
    F::run(v, t, sm + c, tw);
    const int ky = tile / tiles_z, zt = tile - ky * tiles_z;
    if (kz_offset + zt * C + c == 0) {
      float2* pl = plane0 + (size_t(comp) * L + ky) * L;
#pragma unroll
      for (int j = 0; j < F::P; ++j) pl[F::kout(j, t)] = v[j];
    }
    if (comp == 0) {
#pragma unroll
      for (int j = 0; j < F::P; ++j) acc[j] = v[j].x * v[j].x + v[j].y * v[j].y;
    } else {
#pragma unroll
      for (int j = 0; j < F::P; ++j) acc[j] += v[j].x * v[j].x + v[j].y * v[j].y;
    }
    if (++comp == ncomp) {
      __syncthreads();   // the last exchange reads are done: the area becomes the power tile [C][PP]
      float* col = pt + c * PP;
#pragma unroll
      for (int j = 0; j < F::P; ++j) col[F::kout(j, t)] = acc[j];
      __syncthreads();
      float* out = P + size_t(tile) * (C * NRP);
      for (int i = tid; i < C * NRP; i += NT) {
        const int cc = i / NRP, r = i - cc * NRP;
        if (r < NR) {
          float q = pt[cc * PP + r];
          if (r > 0 && r < L / 2) q += pt[cc * PP + L - r];
          out[i] = q;
        }
      }
      comp = 0;
      tile += gridDim.x;
    }
  }
}

// Shell binning of the folded power tiles.  A warp takes 32 consecutive rows of one column: |k|^2 = (kx^2 + ky^2) + kz^2 in
// f64 with numpy's association, the shell from a float estimate corrected against the f64 thresholds (identical decisions
// to numpy.histogram on sqrt(s), see sq_threshold), then a segmented warp reduction over runs of equal shell and ONE
// read-modify-write per run into the warp's private f64 accumulators in shared memory -- no atomics until the final flush.
// COUNT: accumulate the number of modes (1 or 2 per folded row) instead of the power -- geometry only, done once per plan.
template <bool COUNT>
__global__ void __launch_bounds__(256) k_bin_tiles(const float* __restrict__ P, int L, int C, int NRP, int tiles_z, int ntiles,
                                                   const double* __restrict__ kk2, const double* __restrict__ thr_g, int nbins,
                                                   double e0, double inv_de, int kz_offset, double* __restrict__ psum_g,
                                                   unsigned long long* __restrict__ cnt_g) {
  extern __shared__ double bsm[];
  double* thr = bsm;                                   // [nbins + 1]
  const int nwarps = blockDim.x >> 5;
  double* acc = bsm + (nbins + 2);                     // [nwarps][nbins]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i <= nbins; i += blockDim.x) thr[i] = thr_g[i];
  for (int i = tid; i < nwarps * nbins; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  double* my = acc + size_t(w) * nbins;
  const int NR = L / 2 + 1;
  const int chunks_per_col = NRP / 32;
  const uint32_t ncols = uint32_t(ntiles) * uint32_t(C);
  // a warp takes whole columns (tile, c): the column geometry is decoded once, then its rows in chunks of 32
  for (uint32_t colid = blockIdx.x * nwarps + w; colid < ncols; colid += gridDim.x * nwarps) {
    const uint32_t tile = colid / uint32_t(C), c = colid - tile * uint32_t(C);
    const uint32_t ky = tile / uint32_t(tiles_z), zt = tile - ky * uint32_t(tiles_z);
    const int kz = kz_offset + int(zt) * C + int(c);
    if (kz == 0) continue;                               // the packed kz = 0 column is binned by k_plane_bin
    const double kyz2a = __ldg(kk2 + ky), kz2 = __ldg(kk2 + kz);
    const float* col = COUNT ? nullptr : P + size_t(colid) * NRP;
    for (int rc = 0; rc < chunks_per_col; ++rc) {
      const int r = rc * 32 + lane;
      int b = -1;
      float v = 0.f;
      if (r < NR) {
        const double s = __dadd_rn(__dadd_rn(__ldg(kk2 + r), kyz2a), kz2);
        int e = int((sqrtf(float(s)) - float(e0)) * float(inv_de));   // estimate of the shell; corrected below against the exact thresholds
        e = e < 0 ? 0 : (e > nbins ? nbins : e);
        while (e < nbins + 1 && thr[e] <= s) ++e;        // e = number of thresholds <= s ...
        while (e > 0 && thr[e - 1] > s) --e;
        b = e - 1;                                        // ... minus one: -1 below the first edge, nbins beyond the last
        if (b >= nbins) b = -1;
        if (b >= 0) v = COUNT ? ((r > 0 && r < L / 2) ? 2.f : 1.f) : col[r];
      }
      // runs of equal shell (never decreasing along a column): inclusive segmented scan, the last lane of a run holds its sum
      const int prev = __shfl_up_sync(0xffffffffu, b, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != b);
      const int seg0 = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float tv = __shfl_up_sync(0xffffffffu, v, d);
        if (lane - d >= seg0) v += tv;
      }
      const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
      if (tail && b >= 0) my[b] += double(v);
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < nbins; i += blockDim.x) {
    double a = 0.0;
    for (int q = 0; q < nwarps; ++q) a += acc[size_t(q) * nbins + i];
    if (a != 0.0) {
      // Hermitian partner of every kz in [1, N/2-1]
      if (COUNT) atomicAdd(cnt_g + i, 2ull * (unsigned long long)(a + 0.5));
      else atomicAdd(psum_g + i, 2.0 * a);
    }
  }
}

__device__ __forceinline__ int bin_of(double s, const double* __restrict__ thr, int nbins) {
  // number of thresholds <= s, minus one
  int lo = 0, hi = nbins + 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (thr[mid] <= s) lo = mid + 1; else hi = mid;
  }
  return lo - 1;  // -1: below first edge; nbins: beyond last
}

// kz = 0 and kz = N/2 planes from the packed column:  A = (Z(k) + conj Z(-k))/2,  B = (Z(k) - conj Z(-k))/(2i)
__global__ void __launch_bounds__(256) k_plane_bin(const float2* __restrict__ plane0, int ncomp, int N,
                                                   const double* __restrict__ kk2, const double* __restrict__ thr, int nbins,
                                                   double* __restrict__ psum_g, unsigned long long* __restrict__ cnt_g) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * N) return;
  int ky = i / N, kx = i % N;
  int my = (N - ky) % N, mx = (N - kx) % N;
  float pa = 0.f, pb = 0.f;
  for (int comp = 0; comp < ncomp; ++comp) {
    const float2* pl = plane0 + size_t(comp) * N * N;
    float2 zp = pl[size_t(ky) * N + kx], zm = cconj(pl[size_t(my) * N + mx]);
    float2 a = make_float2(0.5f * (zp.x + zm.x), 0.5f * (zp.y + zm.y));
    float2 d = csub(zp, zm);
    pa += a.x * a.x + a.y * a.y;
    pb += 0.25f * (d.x * d.x + d.y * d.y);
  }
  double w = __dadd_rn(kk2[kx], kk2[ky]);
  int ba = bin_of(__dadd_rn(w, kk2[0]), thr, nbins), bb = bin_of(__dadd_rn(w, kk2[N / 2]), thr, nbins);
  if (ba >= 0 && ba < nbins) { atomicAdd(psum_g + ba, double(pa)); atomicAdd(cnt_g + ba, 1ull); }
  if (bb >= 0 && bb < nbins) { atomicAdd(psum_g + bb, double(pb)); atomicAdd(cnt_g + bb, 1ull); }
}

// ------------------------------------------------------------------ standalone binning of a full power cube
__global__ void __launch_bounds__(256) k_bin_full(const double* __restrict__ P, int N, const double* __restrict__ kk2,
                                                  const double* __restrict__ thr, int nbins, double* __restrict__ psum_g,
                                                  unsigned long long* __restrict__ cnt_g) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(N) * N * N) return;
  int kz = int(i % N);
  size_t t = i / N;
  int ky = int(t % N), kx = int(t / N);
  double s = __dadd_rn(__dadd_rn(kk2[kx], kk2[ky]), kk2[kz]);
  int b = bin_of(s, thr, nbins);
  if (b >= 0 && b < nbins) { atomicAdd(psum_g + b, P[i]); atomicAdd(cnt_g + b, 1ull); }
}

// ------------------------------------------------------------------ any-N fallback (direct DFT per axis)
__global__ void k_real_to_complex(const float* __restrict__ f, float2* __restrict__ z, size_t n) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) z[i] = make_float2(f[i], 0.f);
}
// one CTA per line; line element n at base + n*stride
__global__ void __launch_bounds__(128) k_dft_axis(float2* __restrict__ z, int N, size_t stride, int n_inner, size_t outer_stride,
                                                  size_t inner_stride, const double2* __restrict__ tw) {
  extern __shared__ float2 line[];
  size_t base = size_t(blockIdx.x / n_inner) * outer_stride + size_t(blockIdx.x % n_inner) * inner_stride;
  for (int n = threadIdx.x; n < N; n += blockDim.x) line[n] = z[base + n * stride];
  __syncthreads();
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    double re = 0, im = 0;
    int m = 0;
    for (int n = 0; n < N; ++n) {
      double2 w = tw[m];
      re += double(line[n].x) * w.x - double(line[n].y) * w.y;
      im += double(line[n].x) * w.y + double(line[n].y) * w.x;
      m += k;
      if (m >= N) m -= N;
    }
    z[base + k * stride] = make_float2(float(re), float(im));
  }
}
__global__ void k_accum_power(const float2* __restrict__ z, double* __restrict__ P, size_t n, int first) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = double(z[i].x) * z[i].x + double(z[i].y) * z[i].y;
  P[i] = first ? v : P[i] + v;
}

__global__ void k_unpack_half(const float2* __restrict__ packed, int N, float2* __restrict__ half) {
  // half[x][y][kz], kz in [0, N/2]
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int NZ = N / 2, NH = NZ + 1;
  if (i >= size_t(N) * N * NH) return;
  int kz = int(i % NH);
  size_t t = i / NH;
  int y = int(t % N), x = int(t / N);
  if (kz > 0 && kz < NZ) { half[i] = packed[(size_t(x) * N + y) * NZ + kz]; return; }
  int my = (N - y) % N, mx = (N - x) % N;
  float2 zp = packed[(size_t(x) * N + y) * NZ], zm = cconj(packed[(size_t(mx) * N + my) * NZ]);
  if (kz == 0) half[i] = make_float2(0.5f * (zp.x + zm.x), 0.5f * (zp.y + zm.y));
  else {
    float2 d = cmul_mi(csub(zp, zm));
    half[i] = make_float2(0.5f * d.x, 0.5f * d.y);
  }
}

// full-cube power from the packed half spectrum: every (x,y,kz<=N/2) writes itself and its Hermitian partner
__global__ void k_expand_power(const float2* __restrict__ packed, int N, double* __restrict__ P, int first) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int NZ = N / 2, NH = NZ + 1;
  if (i >= size_t(N) * N * NH) return;
  int kz = int(i % NH);
  size_t t = i / NH;
  int y = int(t % N), x = int(t / N);
  int my = (N - y) % N, mx = (N - x) % N;
  float2 f;
  if (kz > 0 && kz < NZ) f = packed[(size_t(x) * N + y) * NZ + kz];
  else {
    float2 zp = packed[(size_t(x) * N + y) * NZ], zm = cconj(packed[(size_t(mx) * N + my) * NZ]);
    if (kz == 0) f = make_float2(0.5f * (zp.x + zm.x), 0.5f * (zp.y + zm.y));
    else { float2 d = cmul_mi(csub(zp, zm)); f = make_float2(0.5f * d.x, 0.5f * d.y); }
  }
  double p = double(f.x) * f.x + double(f.y) * f.y;
  size_t o = (size_t(x) * N + y) * N + kz;
  P[o] = first ? p : P[o] + p;
  if (kz > 0 && kz < NZ) {
    size_t o2 = (size_t(mx) * N + my) * N + (N - kz);
    P[o2] = first ? p : P[o2] + p;
  }
}

// ------------------------------------------------------------------ launch helpers
// cuTensorMapEncodeTiled through the runtime's driver entry point table (the library does not link libcuda)
typedef CUresult (*vp_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int encode_tensor_map(CUtensorMap* out, const void* base, int rank, const cuuint64_t* gdim, const cuuint64_t* gstride_bytes,
                      const cuuint32_t* box) {
  static vp_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    VP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
    VP_REQUIRE(p && qr == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
    fn = reinterpret_cast<vp_encode_tiled_fn>(p);
  }
  static const int promo = getenv("VP_X_L2PROMO") ? atoi(getenv("VP_X_L2PROMO")) : 1;   // 0 none, 1 128 B, 2 256 B
  const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, cuuint32_t(rank), const_cast<void*>(base), gdim, gstride_bytes, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B),
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vp_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, base %p)", int(r), rank, base);
    return VP_ERR_CUDA;
  }
  return VP_OK;
}

template <int R1, int R2, int R3>
int launch_z(float* data, int N, int nx, const vp_pk_plan* pl, cudaStream_t st) {
  using F = LineFFT<R1, R2, R3, 1>;
  constexpr int LINES = 256 / F::T;
  VP_REQUIRE((size_t(nx) * N) % LINES == 0, "fft z pass: %d x %d lines is not a multiple of %d", nx, N, LINES);
  size_t smem = size_t(LINES) * xsize<F::L, 1>() * sizeof(float2);
  static bool attr = false;
  if (!attr) { VP_CUDA(cudaFuncSetAttribute(k_fft_z<R1, R2, R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); attr = true; }
  size_t nlines = size_t(nx) * N;
  vp_stage stage(pl->ctx, "k4a_fft_z", st, 1, 8.0 * double(nx) * N * N);   // 4 B/real read + 4 B/real written in place
  k_fft_z<R1, R2, R3><<<unsigned(nlines / LINES), LINES * F::T, smem, st>>>(data, pl->tw_half, pl->tw_full);
  VP_CHECK_LAUNCH();
  return VP_OK;
}

template <int R1, int R2, int R3, int C>
int launch_y(const float2* data, const YDest& dst, int N, int nx, int kzc, const vp_pk_plan* pl, cudaStream_t st) {
  using F = LineFFT<R1, R2, R3, C>;
  size_t smem = size_t(xsize<F::L, C>()) * sizeof(float2);
  static bool attr = false;
  if (!attr) { VP_CUDA(cudaFuncSetAttribute(k_fft_y<R1, R2, R3, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); attr = true; }
  const int NZ = N / 2;
  VP_REQUIRE(kzc % C == 0, "fft y pass: %d columns per rank is not a multiple of the tile width %d", kzc, C);
  vp_stage stage(pl->ctx, "k4b_fft_y", st, 1, 8.0 * double(nx) * N * N);   // 8 B/mode read + written, nx*N*N/2 modes
  k_fft_y<R1, R2, R3, C><<<unsigned(nx * (NZ / C)), F::T * C, smem, st>>>(data, dst, NZ, kzc, pl->tw_full);
  VP_CHECK_LAUNCH();
  return VP_OK;
}

template <int R1, int R2, int R3, int C>
int launch_x_pow(FieldSet fs, int N, int NZ, int kz_offset, vp_pk_plan* pl, double* psum, unsigned long long* cnt, cudaStream_t st) {
  using F = LineFFT<R1, R2, R3, C>;
  constexpr int NT = F::T * C, NRP = nrp_of<F::L>();
  vp_ctx* ctx = pl->ctx;
  VP_REQUIRE(NZ % C == 0, "fft x pass: %d columns is not a multiple of the tile width %d", NZ, C);
  const int tiles_z = NZ / C, ntiles = N * tiles_z;
  const size_t pbytes = size_t(ntiles) * C * NRP * sizeof(float);
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, pbytes + 1024));
  float* P = static_cast<float*>(vp_arena_alloc(ctx, pbytes));
  VP_REQUIRE(P, "fft x pass: arena carve failed");
  const size_t smem = size_t(xsize<F::L, C>()) * sizeof(float2) + size_t(C) * (F::L + 4) * sizeof(float);
  static bool attr = false;
  if (!attr) { VP_CUDA(cudaFuncSetAttribute(k_fft_x_pow<R1, R2, R3, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); attr = true; }
  using X = XTma<R1, R2, R3, C>;
  static const bool tma_on = !(getenv("VP_X_TMA") && atoi(getenv("VP_X_TMA")) == 0);
  const bool use_tma = X::ok && fs.blocked && tma_on;
  if constexpr (X::ok) if (use_tma) {
    // blocked half spectrum as a 5-D tensor of f32 (innermost first): [2C][kYB][tiles_z][N (x)][N/kYB]; one box = BX x planes of
    // one (ky, kz tile): 2C x 1 x 1 x BX x 1
    const XGeom xg = x_geometry<F::L, C>(NZ);
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5];
    for (int i = 0; i < 5; ++i) { gdim[i] = xg.gdim[i]; box[i] = xg.box[i]; }
    for (int i = 0; i < 4; ++i) gstr[i] = xg.gstride[i];
    XMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int c = 0; c < fs.n; ++c) VP_TRY(encode_tensor_map(&maps.m[c], fs.f[c], 5, gdim, gstr, box));
    static int occ = 0;
    if (!occ) {
      VP_CUDA(cudaFuncSetAttribute(k_fft_x_pow_tma<R1, R2, R3, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, X::SMEM));
      VP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fft_x_pow_tma<R1, R2, R3, C>, NT, X::SMEM));
      VP_REQUIRE(occ >= 1, "fft x pass: the TMA kernel for N=%d does not fit an SM", N);
    }
    const int grid = ntiles < ctx->sm_count * occ ? ntiles : ctx->sm_count * occ;
    vp_stage stage(ctx, "k4c_fft_x_pow", st, 1, 8.0 * double(N) * N * NZ * fs.n + 4.0 * double(N) * NZ * (N / 2 + 1));
    k_fft_x_pow_tma<R1, R2, R3, C><<<grid, NT, X::SMEM, st>>>(maps, fs.n, NZ, pl->tw_full, pl->plane0, kz_offset, P, ntiles);
    VP_CHECK_LAUNCH();
  }
  if (!use_tma) {
    // row-major input (NCCL exchange), N = 250 (40-byte pieces) and N = 2048 (an item does not fit beside its exchange area):
    // per-thread loads
    // 8 B/mode read per component, the folded power tile written (4 B per pair of modes)
    vp_stage stage(ctx, "k4c_fft_x_pow", st, 1, 8.0 * double(N) * N * NZ * fs.n + 4.0 * double(N) * NZ * (N / 2 + 1));
    k_fft_x_pow<R1, R2, R3, C><<<ntiles, NT, smem, st>>>(fs, NZ, pl->tw_full, pl->plane0, kz_offset, P);
    VP_CHECK_LAUNCH();
  }
  // binning: per-warp f64 accumulators in shared memory
  const int nbins = pl->nbins;
  const int nwarps = nbins <= 2048 ? 8 : 4;
  const size_t bsmem = (size_t(nbins) + 2 + size_t(nwarps) * nbins) * sizeof(double);
  VP_REQUIRE(bsmem <= size_t(200) * 1024, "vp_pk_fields: nbins=%d is more than the binning kernel holds in shared memory", nbins);
  static bool battr = false;
  if (!battr) {
    VP_CUDA(cudaFuncSetAttribute(k_bin_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VP_CUDA(cudaFuncSetAttribute(k_bin_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    battr = true;
  }
  int per_sm = int((size_t(220) * 1024) / (bsmem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const long long ncols = (long long)ntiles * C;
  long long grid = (long long)ctx->sm_count * per_sm;
  if (grid > (ncols + nwarps - 1) / nwarps) grid = (ncols + nwarps - 1) / nwarps;
  if (!pl->ns_valid || pl->ns_kz_offset != kz_offset || pl->ns_NZ != NZ) {
    // mode counts of the tiles: geometry only -- once per plan
    if (!pl->ns_tiles) VP_CUDA(cudaMalloc(reinterpret_cast<void**>(&pl->ns_tiles), sizeof(unsigned long long) * nbins));
    VP_CUDA(cudaMemsetAsync(pl->ns_tiles, 0, sizeof(unsigned long long) * nbins, st));
    vp_stage stage(ctx, "k5_bin_counts", st, 1);
    k_bin_tiles<true><<<unsigned(grid), nwarps * 32, bsmem, st>>>(nullptr, N, C, NRP, tiles_z, ntiles, pl->kk2, pl->thr, nbins, pl->e0,
                                                                pl->inv_de, kz_offset, nullptr, pl->ns_tiles);
    VP_CHECK_LAUNCH();
    pl->ns_valid = true;
    pl->ns_kz_offset = kz_offset;
    pl->ns_NZ = NZ;
  }
  VP_CUDA(cudaMemcpyAsync(cnt, pl->ns_tiles, sizeof(unsigned long long) * nbins, cudaMemcpyDeviceToDevice, st));
  {
    vp_stage stage(ctx, "k5_bin_tiles", st, 1, 4.0 * double(N) * NZ * (N / 2 + 1));
    k_bin_tiles<false><<<unsigned(grid), nwarps * 32, bsmem, st>>>(P, N, C, NRP, tiles_z, ntiles, pl->kk2, pl->thr, nbins, pl->e0,
                                                                 pl->inv_de, kz_offset, psum, nullptr);
    VP_CHECK_LAUNCH();
  }
  return VP_OK;
}

int run_z(float* d, int N, int nx, const vp_pk_plan* pl, cudaStream_t st) {
  switch (N) {
    case 64: return launch_z<16, 2, 1>(d, N, nx, pl, st);
    case 128: return launch_z<16, 4, 1>(d, N, nx, pl, st);
    case 256: return launch_z<16, 8, 1>(d, N, nx, pl, st);
    case 512: return launch_z<16, 16, 1>(d, N, nx, pl, st);
    case 1024: return launch_z<16, 16, 2>(d, N, nx, pl, st);
    case 2048: return launch_z<16, 16, 4>(d, N, nx, pl, st);
    case 250: return launch_z<5, 5, 5>(d, N, nx, pl, st);      // half length 125
    case 500: return launch_z<10, 5, 5>(d, N, nx, pl, st);     // 250
    case 1000: return launch_z<10, 10, 5>(d, N, nx, pl, st);   // 500
  }
  vp_set_error("fft z pass: unsupported N=%d", N);
  return VP_ERR_UNSUPPORTED;
}
int run_y(const float2* d, const YDest& dst, int N, int nx, int kzc, const vp_pk_plan* pl, cudaStream_t st) {
  switch (N) {
    case 64: return launch_y<16, 4, 1, 32>(d, dst, N, nx, kzc, pl, st);
    case 128: return launch_y<16, 8, 1, 32>(d, dst, N, nx, kzc, pl, st);
    case 256: return launch_y<16, 16, 1, 16>(d, dst, N, nx, kzc, pl, st);
    case 512: return launch_y<16, 16, 2, 8>(d, dst, N, nx, kzc, pl, st);
    case 1024: return launch_y<16, 16, 4, 8>(d, dst, N, nx, kzc, pl, st);
    case 2048: return launch_y<16, 16, 8, 8>(d, dst, N, nx, kzc, pl, st);
    case 250: return launch_y<10, 5, 5, 5>(d, dst, N, nx, kzc, pl, st);
    case 500: return launch_y<10, 10, 5, 10>(d, dst, N, nx, kzc, pl, st);
    case 1000: return launch_y<10, 10, 10, 10>(d, dst, N, nx, kzc, pl, st);
  }
  vp_set_error("fft y pass: unsupported N=%d", N);
  return VP_ERR_UNSUPPORTED;
}
// destinations for the two single-buffer forms
YDest ydest_blocks(float2* out, int nranks, int nx, int N, int kzc) {
  YDest d;
  for (int r = 0; r < 16; ++r) d.base[r] = r < nranks ? out + size_t(r) * nx * N * kzc : nullptr;
  d.xoff = 0;
  d.blocked = 0;
  d.nxtot = nx;
  return d;
}
// x pass + |F|^2 (k_fft_x_pow) and shell binning (k_bin_tiles).  psum is accumulated into (the caller zeroes it); cnt is
// overwritten with the mode counts of this rank's tiles.
int run_x_bin(FieldSet fs, int N, int NZ, int kz_offset, vp_pk_plan* pl, double* psum, unsigned long long* cnt, cudaStream_t st) {
  switch (N) {
    case 64: return launch_x_pow<16, 4, 1, 32>(fs, N, NZ, kz_offset, pl, psum, cnt, st);
    case 128: return launch_x_pow<16, 8, 1, 32>(fs, N, NZ, kz_offset, pl, psum, cnt, st);
    case 256: return launch_x_pow<16, 16, 1, 16>(fs, N, NZ, kz_offset, pl, psum, cnt, st);
    case 512: return launch_x_pow<16, 16, 2, 8>(fs, N, NZ, kz_offset, pl, psum, cnt, st);
    case 1024: return launch_x_pow<16, 16, 4, 8>(fs, N, NZ, kz_offset, pl, psum, cnt, st);
    case 2048: return launch_x_pow<16, 16, 8, 8>(fs, N, NZ, kz_offset, pl, psum, cnt, st);
    case 250: return launch_x_pow<10, 5, 5, 5>(fs, N, NZ, kz_offset, pl, psum, cnt, st);
    case 500: return launch_x_pow<10, 10, 5, 10>(fs, N, NZ, kz_offset, pl, psum, cnt, st);
    case 1000: return launch_x_pow<10, 10, 10, 10>(fs, N, NZ, kz_offset, pl, psum, cnt, st);
  }
  vp_set_error("fft x pass: unsupported N=%d", N);
  return VP_ERR_UNSUPPORTED;
}

// the x pass without binning (diagnostic transform): reuse the y kernel on a transposed view is not possible
// in place, so the diagnostic transform runs the x lines with the generic strided kernel below.
template <int R1, int R2, int R3, int C>
__global__ void __launch_bounds__(R2* R3* C, (R2 * R3 * C <= 512 ? 2 : 1)) k_fft_x(float2* __restrict__ data, int NZ, const float2* __restrict__ tw) {
  using F = LineFFT<R1, R2, R3, C>;
  constexpr int L = F::L, T = F::T;
  extern __shared__ float2 sm[];
  const int tid = threadIdx.x, c = tid % C, t = tid / C;
  const int tiles = NZ / C;
  const int y = blockIdx.x / tiles, zt = blockIdx.x % tiles;
  const size_t xstride = size_t(L) * NZ;
  float2* base = data + size_t(y) * NZ + zt * C + c;
  float2 v[F::P];
#pragma unroll
  for (int j = 0; j < F::P; ++j) v[j] = base[size_t(j * T + t) * xstride];
  F::run(v, t, sm + c, tw);
#pragma unroll
  for (int j = 0; j < F::P; ++j) base[size_t(F::kout(j, t)) * xstride] = v[j];
}
template <int R1, int R2, int R3, int C>
int launch_x(float2* data, int N, const vp_pk_plan* pl, cudaStream_t st) {
  using F = LineFFT<R1, R2, R3, C>;
  size_t smem = size_t(xsize<F::L, C>()) * sizeof(float2);
  static bool attr = false;
  if (!attr) { VP_CUDA(cudaFuncSetAttribute(k_fft_x<R1, R2, R3, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); attr = true; }
  const int NZ = N / 2;
  vp_stage stage(pl->ctx, "k4c_fft_x", st, 1, 8.0 * double(N) * N * N);
  k_fft_x<R1, R2, R3, C><<<unsigned(N * (NZ / C)), F::T * C, smem, st>>>(data, NZ, pl->tw_full);
  VP_CHECK_LAUNCH();
  return VP_OK;
}
int run_x(float2* d, int N, const vp_pk_plan* pl, cudaStream_t st) {
  switch (N) {
    case 64: return launch_x<16, 4, 1, 32>(d, N, pl, st);
    case 128: return launch_x<16, 8, 1, 32>(d, N, pl, st);
    case 256: return launch_x<16, 16, 1, 16>(d, N, pl, st);
    case 512: return launch_x<16, 16, 2, 8>(d, N, pl, st);
    case 1024: return launch_x<16, 16, 4, 8>(d, N, pl, st);
    case 2048: return launch_x<16, 16, 8, 8>(d, N, pl, st);
    case 250: return launch_x<10, 5, 5, 5>(d, N, pl, st);
    case 500: return launch_x<10, 10, 5, 10>(d, N, pl, st);
    case 1000: return launch_x<10, 10, 10, 10>(d, N, pl, st);
  }
  vp_set_error("fft x pass: unsupported N=%d", N);
  return VP_ERR_UNSUPPORTED;
}

// threshold on s = |k|^2 equivalent to the comparison  sqrt(s) >= e  (strict = false) or  sqrt(s) > e  (strict = true)
double sq_threshold(double e, bool strict) {
  if (!(e > 0.0)) {
    if (!strict || e < 0.0) return 0.0;       // every s >= 0 qualifies
    return nextafter(0.0, 1.0);               // sqrt(s) > 0
  }
  double s = e * e;
  auto ok = [&](double v) { double r = sqrt(v); return strict ? r > e : r >= e; };
  while (ok(s)) { double p = nextafter(s, -INFINITY); if (p < 0.0) break; if (!ok(p)) break; s = p; }
  while (!ok(s)) s = nextafter(s, INFINITY);
  return s;
}

}  // namespace

namespace {
template <int R1, int R2, int R3, int C>
int x_layout_report(int NZ, int64_t* o) {
  using X = XTma<R1, R2, R3, C>;
  constexpr int L = R1 * R2 * R3;
  if (NZ <= 0 || NZ % C != 0) { vp_set_error("vp_fft_x_layout: %d columns is not a multiple of the tile width %d", NZ, C); return VP_ERR_ARG; }
  const XGeom g = x_geometry<L, C>(NZ);
  o[0] = C; o[1] = yb_of<L>(); o[2] = NZ / C; o[3] = X::BX; o[4] = X::NBOX; o[5] = X::ok ? 1 : 0;
  for (int i = 0; i < 5; ++i) o[6 + i] = int64_t(g.gdim[i]);
  for (int i = 0; i < 4; ++i) o[11 + i] = int64_t(g.gstride[i]);
  for (int i = 0; i < 5; ++i) o[15 + i] = int64_t(g.box[i]);
  o[20] = X::BOXP; o[21] = X::NST; o[22] = X::SMEM; o[23] = X::NT;
  return VP_OK;
}
}  // namespace

extern "C" int vp_fft_x_layout(int N, int kz_columns, int64_t* info_out) {
  VP_REQUIRE(info_out, "vp_fft_x_layout: null argument");
  switch (N) {
    case 64: return x_layout_report<16, 4, 1, 32>(kz_columns, info_out);
    case 128: return x_layout_report<16, 8, 1, 32>(kz_columns, info_out);
    case 256: return x_layout_report<16, 16, 1, 16>(kz_columns, info_out);
    case 512: return x_layout_report<16, 16, 2, 8>(kz_columns, info_out);
    case 1024: return x_layout_report<16, 16, 4, 8>(kz_columns, info_out);
    case 2048: return x_layout_report<16, 16, 8, 8>(kz_columns, info_out);
    case 250: return x_layout_report<10, 5, 5, 5>(kz_columns, info_out);
    case 500: return x_layout_report<10, 10, 5, 10>(kz_columns, info_out);
    case 1000: return x_layout_report<10, 10, 10, 10>(kz_columns, info_out);
  }
  vp_set_error("vp_fft_x_layout: N=%d has no line-FFT path", N);
  return VP_ERR_UNSUPPORTED;
}

extern "C" int vp_pk_plan_create(vp_ctx* ctx, int N, const double* k_h, const double* edges_h, int nbins, vp_pk_plan** out) {
  VP_REQUIRE(ctx && k_h && edges_h && out, "vp_pk_plan_create: null argument");
  VP_REQUIRE(N >= 2 && nbins >= 1, "vp_pk_plan_create: bad sizes");
  for (int j = 0; j < nbins; ++j)
    VP_REQUIRE(edges_h[j + 1] > edges_h[j], "vp_pk_plan_create: edges must increase");
  VP_CUDA(cudaSetDevice(ctx->device));
  vp_pk_plan* p = new vp_pk_plan();
  p->ctx = ctx;
  p->N = N;
  p->nbins = nbins;
  // lines with a register/shared-memory transform: powers of two 64..2048, and the 2^a 5^b sizes the reference is run at
  // (scripts/parallel_optimized.py:30 NTOT = 1000; scripts/buffer_test.sh -N 500; 250 as their half)
  bool is_pow2 = ((N & (N - 1)) == 0 && N >= 64 && N <= 2048) || N == 250 || N == 500 || N == 1000;
  // the fused binning folds kx <-> -kx: needs a symmetric k table (true for fftfreq; false with a fold shift)
  bool symmetric = true;
  for (int r = 1; r < N / 2; ++r)
    if (k_h[r] * k_h[r] != k_h[N - r] * k_h[N - r]) symmetric = false;
  for (int r = 1; r <= N / 2; ++r)
    if (!(k_h[r] * k_h[r] >= k_h[r - 1] * k_h[r - 1])) symmetric = false;  // rows must be ordered in |kx|
  p->pow2 = is_pow2 && symmetric && (N % 2 == 0);
  std::vector<double> kk2(N), thr(nbins + 1);
  for (int i = 0; i < N; ++i) kk2[i] = k_h[i] * k_h[i];
  for (int j = 0; j < nbins; ++j) thr[j] = sq_threshold(edges_h[j], false);
  thr[nbins] = sq_threshold(edges_h[nbins], true);
  p->inv_kf = (N > 1 && fabs(k_h[1]) > 0) ? float(1.0 / fabs(k_h[1])) : 1.f;
  p->e0 = edges_h[0];
  p->inv_de = double(nbins) / (edges_h[nbins] - edges_h[0]);
  std::vector<float2> twf(N), twh(N / 2 > 0 ? N / 2 : 1);
  for (int m = 0; m < N; ++m) twf[m] = make_float2(float(cos(2.0 * M_PI * m / N)), float(-sin(2.0 * M_PI * m / N)));
  for (int m = 0; m < N / 2; ++m)
    twh[m] = make_float2(float(cos(2.0 * M_PI * m / (N / 2))), float(-sin(2.0 * M_PI * m / (N / 2))));
  auto fail = [&](const char* what) { vp_set_error("vp_pk_plan_create: %s", what); vp_pk_plan_destroy(p); return VP_ERR_NOMEM; };
  if (cudaMalloc(&p->tw_full, sizeof(float2) * N) != cudaSuccess) return fail("cudaMalloc tw_full");
  if (cudaMalloc(&p->tw_half, sizeof(float2) * twh.size()) != cudaSuccess) return fail("cudaMalloc tw_half");
  if (cudaMalloc(&p->kk2, sizeof(double) * N) != cudaSuccess) return fail("cudaMalloc kk2");
  if (cudaMalloc(&p->thr, sizeof(double) * (nbins + 1)) != cudaSuccess) return fail("cudaMalloc thr");
  if (p->pow2 && cudaMalloc(&p->plane0, sizeof(float2) * 3 * size_t(N) * N) != cudaSuccess) return fail("cudaMalloc plane0");
  VP_CUDA(cudaMemcpy(p->tw_full, twf.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
  VP_CUDA(cudaMemcpy(p->tw_half, twh.data(), sizeof(float2) * twh.size(), cudaMemcpyHostToDevice));
  VP_CUDA(cudaMemcpy(p->kk2, kk2.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
  VP_CUDA(cudaMemcpy(p->thr, thr.data(), sizeof(double) * (nbins + 1), cudaMemcpyHostToDevice));
  *out = p;
  return VP_OK;
}

extern "C" int vp_pk_plan_destroy(vp_pk_plan* p) {
  if (!p) return VP_OK;
  cudaSetDevice(p->ctx->device);
  cudaDeviceSynchronize();
  if (p->tw_full) cudaFree(p->tw_full);
  if (p->tw_half) cudaFree(p->tw_half);
  if (p->kk2) cudaFree(p->kk2);
  if (p->thr) cudaFree(p->thr);
  if (p->plane0) cudaFree(p->plane0);
  if (p->ns_tiles) cudaFree(p->ns_tiles);
  for (int c = 0; c < 3; ++c) {
    if (p->peer_open)
      for (int d = 0; d < p->nranks; ++d)
        if (d != p->rank && p->peer[c][d]) cudaIpcCloseMemHandle(p->peer[c][d]);
    if (p->recv[c]) cudaFree(p->recv[c]);
  }
  delete p;
  return VP_OK;
}

size_t vp_pk_fields_scratch_bytes(const vp_pk_plan* pl) {
  if (pl->pow2)   // folded power tiles + the second cube of the out-of-place y pass
    return vp_align256(size_t(pl->N) * (pl->N / 2) * (((pl->N / 2 + 1) + 31) / 32 * 32) * sizeof(float)) +
           vp_align256(size_t(pl->N) * pl->N * (pl->N / 2) * sizeof(float2)) + 16384;
  const size_t n3 = size_t(pl->N) * pl->N * pl->N;
  return vp_align256(n3 * sizeof(float2)) + vp_align256(n3 * sizeof(double)) + vp_align256(sizeof(double2) * pl->N) + 16384;
}

// any-N path: full complex transform by direct DFT along each axis -> P = sum_c |F_c|^2 (f64 cube)
static int power_cube_generic(vp_pk_plan* pl, float* const* field_d, int ncomp, double* P, cudaStream_t st) {
  const int N = pl->N;
  const size_t n3 = size_t(N) * N * N;
  vp_ctx* ctx = pl->ctx;
  size_t need = vp_align256(n3 * sizeof(float2)) + vp_align256(sizeof(double2) * N);
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, need));
  float2* z = static_cast<float2*>(vp_arena_alloc(ctx, n3 * sizeof(float2)));
  double2* tw = static_cast<double2*>(vp_arena_alloc(ctx, sizeof(double2) * N));
  VP_REQUIRE(z && tw, "power_cube_generic: arena carve failed");
  std::vector<double2> twh(N);
  for (int m = 0; m < N; ++m) twh[m] = make_double2(cos(2.0 * M_PI * m / N), -sin(2.0 * M_PI * m / N));
  VP_CUDA(cudaMemcpyAsync(tw, twh.data(), sizeof(double2) * N, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaStreamSynchronize(st));
  const unsigned nb = unsigned((n3 + 255) / 256);
  const size_t smem = sizeof(float2) * N;
  vp_stage stage(ctx, "generic_dft", st, 5 * ncomp);
  for (int c = 0; c < ncomp; ++c) {
    k_real_to_complex<<<nb, 256, 0, st>>>(field_d[c], z, n3);
    k_dft_axis<<<unsigned(N * N), 128, smem, st>>>(z, N, 1, N, size_t(N) * N, size_t(N), tw);   // z lines
    k_dft_axis<<<unsigned(N * N), 128, smem, st>>>(z, N, size_t(N), N, size_t(N) * N, 1, tw);   // y lines
    k_dft_axis<<<unsigned(N * N), 128, smem, st>>>(z, N, size_t(N) * N, N, size_t(N), 1, tw);   // x lines
    k_accum_power<<<nb, 256, 0, st>>>(z, P, n3, c == 0);
    VP_CHECK_LAUNCH();
  }
  return VP_OK;
}

static int pk_fields_generic(vp_pk_plan* pl, float* const* field_d, int ncomp, double* psum_d, uint64_t* nsample_d, cudaStream_t st) {
  const size_t n3 = size_t(pl->N) * pl->N * pl->N;
  vp_arena_scope scope(pl->ctx);
  VP_TRY(vp_arena_reserve(pl->ctx, vp_pk_fields_scratch_bytes(pl)));
  double* P = static_cast<double*>(vp_arena_alloc(pl->ctx, n3 * sizeof(double)));
  VP_REQUIRE(P, "pk_fields_generic: arena carve failed");
  VP_TRY(power_cube_generic(pl, field_d, ncomp, P, st));
  return vp_power_bin_full(pl, P, psum_d, nsample_d, st);
}

extern "C" int vp_pk_fields(vp_pk_plan* pl, float* const* field_d, int ncomp, double* psum_d, uint64_t* nsample_d, void* stream) {
  VP_REQUIRE(pl && field_d && psum_d && nsample_d, "vp_pk_fields: null argument");
  vp_call_guard guard(pl->ctx, static_cast<cudaStream_t>(stream));
  VP_REQUIRE(ncomp >= 1 && ncomp <= 3, "vp_pk_fields: ncomp must be 1..3");
  VP_REQUIRE(pl->nranks == 1, "vp_pk_fields: plan is distributed over %d ranks; use vp_pk_dist_local/final", pl->nranks);
  VP_CUDA(cudaSetDevice(pl->ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!pl->pow2) return pk_fields_generic(pl, field_d, ncomp, psum_d, nsample_d, st);
  const int N = pl->N;
  VP_CUDA(cudaMemsetAsync(psum_d, 0, sizeof(double) * pl->nbins, st));
  VP_CUDA(cudaMemsetAsync(nsample_d, 0, sizeof(uint64_t) * pl->nbins, st));
  // The y pass writes the blocked layout out of place: component 0 into a scratch cube, component c into the (dead) cube of
  // component c - 1 -- one extra cube whatever the number of components.
  vp_arena_scope scope(pl->ctx);
  const size_t cube = vp_align256(size_t(N) * N * (N / 2) * sizeof(float2));
  VP_TRY(vp_arena_reserve(pl->ctx, vp_pk_fields_scratch_bytes(pl)));
  float2* extra = static_cast<float2*>(vp_arena_alloc(pl->ctx, cube));
  VP_REQUIRE(extra, "vp_pk_fields: arena carve failed");
  FieldSet fs;
  fs.n = ncomp;
  fs.blocked = 1;
  for (int c = 0; c < 3; ++c) fs.f[c] = nullptr;
  // (Alternating the z and y passes over groups of 8-32 x planes, so that a group's z-pass output is still in L2 when its y
  // pass reads it, was measured and lost: 26.9 / 23.7 / 33.0 ms for groups of 16 / 32 / 8 planes against 19.1 ms for whole-cube
  // passes at cfg4 -- the short launches do not fill the machine.)
  for (int c = 0; c < ncomp; ++c) {
    VP_REQUIRE(field_d[c], "vp_pk_fields: null field %d", c);
    VP_TRY(run_z(field_d[c], N, N, pl, st));
    YDest dst;
    for (int r = 0; r < 16; ++r) dst.base[r] = nullptr;
    dst.base[0] = c == 0 ? extra : reinterpret_cast<float2*>(field_d[c - 1]);
    dst.xoff = 0;
    dst.blocked = 1;
    dst.nxtot = N;
    VP_TRY(run_y(reinterpret_cast<float2*>(field_d[c]), dst, N, N, N / 2, pl, st));
    fs.f[c] = dst.base[0];
  }
  VP_TRY(run_x_bin(fs, N, N / 2, 0, pl, psum_d, reinterpret_cast<unsigned long long*>(nsample_d), st));
  vp_stage stage(pl->ctx, "k5_plane_bin", st, 1, 8.0 * double(N) * N * ncomp);
  k_plane_bin<<<unsigned((size_t(N) * N + 255) / 256), 256, 0, st>>>(pl->plane0, ncomp, N, pl->kk2, pl->thr, pl->nbins, psum_d,
                                                                  reinterpret_cast<unsigned long long*>(nsample_d));
  VP_CHECK_LAUNCH();
  return VP_OK;
}

extern "C" int vp_pk_plan_create_dist(vp_ctx* ctx, int N, int nranks, int rank, const double* k_h, const double* edges_h, int nbins,
                                      vp_pk_plan** out) {
  VP_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "vp_pk_plan_create_dist: bad rank %d of %d", rank, nranks);
  VP_TRY(vp_pk_plan_create(ctx, N, k_h, edges_h, nbins, out));
  vp_pk_plan* p = *out;
  if (nranks > 1) {
    bool ok = p->pow2 && (N % nranks == 0) && ((N / 2) % nranks == 0);
    if (!ok) {
      vp_pk_plan_destroy(p);
      *out = nullptr;
      vp_set_error("vp_pk_plan_create_dist: N=%d over %d ranks needs the power-of-two path and N/2 divisible by the rank count", N, nranks);
      return VP_ERR_UNSUPPORTED;
    }
  }
  p->nranks = nranks;
  p->rank = rank;
  p->kzc = (N / 2) / nranks;
  return VP_OK;
}

extern "C" int vp_pk_dist_local(vp_pk_plan* pl, float* const* field_d, int ncomp, float* const* send_d, void* stream) {
  VP_REQUIRE(pl && field_d && send_d && ncomp >= 1 && ncomp <= 3, "vp_pk_dist_local: bad argument");
  vp_call_guard guard(pl->ctx, static_cast<cudaStream_t>(stream));
  VP_REQUIRE(pl->pow2, "vp_pk_dist_local: N=%d has no slab path", pl->N);
  VP_CUDA(cudaSetDevice(pl->ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = pl->N, nx = N / pl->nranks;
  for (int c = 0; c < ncomp; ++c) {
    VP_REQUIRE(field_d[c] && send_d[c], "vp_pk_dist_local: null buffer %d", c);
    VP_TRY(run_z(field_d[c], N, nx, pl, st));
    VP_TRY(run_y(reinterpret_cast<const float2*>(field_d[c]), ydest_blocks(reinterpret_cast<float2*>(send_d[c]), pl->nranks, nx, N, pl->kzc), N, nx,
                 pl->kzc, pl, st));
  }
  return VP_OK;
}

static int dist_final_impl(vp_pk_plan* pl, float* const* recv_d, int ncomp, int blocked, double* psum_d, uint64_t* nsample_d, void* stream);
extern "C" int vp_pk_dist_final(vp_pk_plan* pl, float* const* recv_d, int ncomp, double* psum_d, uint64_t* nsample_d, void* stream) {
  return dist_final_impl(pl, recv_d, ncomp, 0, psum_d, nsample_d, stream);
}
static int dist_final_impl(vp_pk_plan* pl, float* const* recv_d, int ncomp, int blocked, double* psum_d, uint64_t* nsample_d, void* stream) {
  VP_REQUIRE(pl && recv_d && psum_d && nsample_d && ncomp >= 1 && ncomp <= 3, "vp_pk_dist_final: bad argument");
  vp_call_guard guard(pl->ctx, static_cast<cudaStream_t>(stream));
  VP_REQUIRE(pl->pow2, "vp_pk_dist_final: N=%d has no slab path", pl->N);
  VP_CUDA(cudaSetDevice(pl->ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = pl->N;
  VP_CUDA(cudaMemsetAsync(psum_d, 0, sizeof(double) * pl->nbins, st));
  VP_CUDA(cudaMemsetAsync(nsample_d, 0, sizeof(uint64_t) * pl->nbins, st));
  FieldSet fs;
  fs.n = ncomp;
  fs.blocked = blocked;
  for (int c = 0; c < 3; ++c) fs.f[c] = c < ncomp ? reinterpret_cast<float2*>(recv_d[c]) : nullptr;
  VP_TRY(run_x_bin(fs, N, pl->kzc, pl->rank * pl->kzc, pl, psum_d, reinterpret_cast<unsigned long long*>(nsample_d), st));
  if (pl->rank == 0) {   // the packed kz=0 column (planes kz=0 and kz=N/2) lives on rank 0
    vp_stage stage(pl->ctx, "k5_plane_bin", st, 1, 8.0 * double(N) * N * ncomp);
    k_plane_bin<<<unsigned((size_t(N) * N + 255) / 256), 256, 0, st>>>(pl->plane0, ncomp, N, pl->kk2, pl->thr, pl->nbins, psum_d,
                                                                    reinterpret_cast<unsigned long long*>(nsample_d));
    VP_CHECK_LAUNCH();
  }
  return VP_OK;
}

// ---- peer-to-peer transpose (the exchange fused into the y pass)
extern "C" int vp_pk_dist_p2p_alloc(vp_pk_plan* pl, int ncomp_max, unsigned char* handles_out) {
  VP_REQUIRE(pl && handles_out && ncomp_max >= 1 && ncomp_max <= 3, "vp_pk_dist_p2p_alloc: bad argument");
  VP_REQUIRE(pl->pow2 && pl->nranks <= 16, "vp_pk_dist_p2p_alloc: needs the power-of-two path and <= 16 ranks");
  VP_CUDA(cudaSetDevice(pl->ctx->device));
  const size_t bytes = sizeof(float2) * size_t(pl->N) * pl->N * pl->kzc;
  for (int c = 0; c < 3; ++c)
    for (int d = 0; d < 16; ++d) pl->peer[c][d] = nullptr;
  for (int c = 0; c < ncomp_max; ++c) {
    if (!pl->recv[c]) VP_CUDA(cudaMalloc(reinterpret_cast<void**>(&pl->recv[c]), bytes));
    cudaIpcMemHandle_t h;
    VP_CUDA(cudaIpcGetMemHandle(&h, pl->recv[c]));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handles_out + size_t(c) * 64, &h, 64);
    pl->peer[c][pl->rank] = pl->recv[c];
  }
  pl->p2p_ncomp = ncomp_max;
  return VP_OK;
}

// Unmap the other ranks' receive buffers.  Teardown order: every rank calls this, the caller runs a barrier across the ranks,
// and only then vp_pk_plan_destroy() frees this rank's exported buffers.
extern "C" int vp_pk_dist_p2p_close(vp_pk_plan* pl) {
  VP_REQUIRE(pl, "vp_pk_dist_p2p_close: null plan");
  VP_CUDA(cudaSetDevice(pl->ctx->device));
  VP_CUDA(cudaDeviceSynchronize());
  if (pl->peer_open) {
    for (int c = 0; c < 3; ++c)
      for (int d = 0; d < pl->nranks; ++d)
        if (d != pl->rank && pl->peer[c][d]) { cudaIpcCloseMemHandle(pl->peer[c][d]); pl->peer[c][d] = nullptr; }
    pl->peer_open = false;
  }
  return VP_OK;
}

extern "C" int vp_pk_dist_p2p_open(vp_pk_plan* pl, const unsigned char* all_handles) {
  VP_REQUIRE(pl && all_handles && pl->p2p_ncomp > 0, "vp_pk_dist_p2p_open: call vp_pk_dist_p2p_alloc first");
  VP_CUDA(cudaSetDevice(pl->ctx->device));
  for (int d = 0; d < pl->nranks; ++d) {
    if (d == pl->rank) continue;
    for (int c = 0; c < pl->p2p_ncomp; ++c) {
      cudaIpcMemHandle_t h;
      memcpy(&h, all_handles + (size_t(d) * pl->p2p_ncomp + c) * 64, 64);
      void* p = nullptr;
      VP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      pl->peer[c][d] = static_cast<float2*>(p);
    }
  }
  pl->peer_open = true;
  return VP_OK;
}

extern "C" int vp_pk_dist_local_p2p(vp_pk_plan* pl, float* const* field_d, int ncomp, void* stream) {
  VP_REQUIRE(pl && field_d && ncomp >= 1 && ncomp <= pl->p2p_ncomp, "vp_pk_dist_local_p2p: bad argument");
  vp_call_guard guard(pl->ctx, static_cast<cudaStream_t>(stream));
  VP_REQUIRE(pl->nranks == 1 || pl->peer_open, "vp_pk_dist_local_p2p: peer buffers not opened");
  VP_CUDA(cudaSetDevice(pl->ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = pl->N, nx = N / pl->nranks;
  for (int c = 0; c < ncomp; ++c) {
    VP_REQUIRE(field_d[c], "vp_pk_dist_local_p2p: null field %d", c);
    VP_TRY(run_z(field_d[c], N, nx, pl, st));
    YDest dst;
    for (int d = 0; d < 16; ++d) dst.base[d] = d < pl->nranks ? pl->peer[c][d] : nullptr;
    dst.xoff = pl->rank * nx;
    dst.blocked = 1;
    dst.nxtot = N;
    VP_TRY(run_y(reinterpret_cast<const float2*>(field_d[c]), dst, N, nx, pl->kzc, pl, st));
  }
  return VP_OK;
}

extern "C" int vp_pk_dist_final_p2p(vp_pk_plan* pl, int ncomp, double* psum_d, uint64_t* nsample_d, void* stream) {
  VP_REQUIRE(pl && ncomp >= 1 && ncomp <= pl->p2p_ncomp, "vp_pk_dist_final_p2p: bad argument");
  float* r[3] = {reinterpret_cast<float*>(pl->recv[0]), reinterpret_cast<float*>(pl->recv[1]), reinterpret_cast<float*>(pl->recv[2])};
  return dist_final_impl(pl, r, ncomp, 1, psum_d, nsample_d, stream);   // the peer stores arrive in the blocked layout (see YDest)
}

extern "C" int vp_fft_r2c_inplace(vp_pk_plan* pl, float* field_d, void* stream) {
  VP_REQUIRE(pl && field_d, "vp_fft_r2c_inplace: null argument");
  vp_call_guard guard(pl->ctx, static_cast<cudaStream_t>(stream));
  if (!pl->pow2) { vp_set_error("vp_fft_r2c_inplace: N=%d has no packed fast path", pl->N); return VP_ERR_UNSUPPORTED; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VP_TRY(run_z(field_d, pl->N, pl->N, pl, st));
  VP_TRY(run_y(reinterpret_cast<float2*>(field_d), ydest_blocks(reinterpret_cast<float2*>(field_d), 1, pl->N, pl->N, pl->N / 2), pl->N, pl->N,
               pl->N / 2, pl, st));
  return run_x(reinterpret_cast<float2*>(field_d), pl->N, pl, st);
}

extern "C" int vp_power_cube(vp_pk_plan* pl, float* const* field_d, int ncomp, double* P_d, void* stream) {
  VP_REQUIRE(pl && field_d && P_d && ncomp >= 1 && ncomp <= 3, "vp_power_cube: bad argument");
  vp_call_guard guard(pl->ctx, static_cast<cudaStream_t>(stream));
  VP_CUDA(cudaSetDevice(pl->ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!pl->pow2) return power_cube_generic(pl, field_d, ncomp, P_d, st);
  const int N = pl->N;
  size_t n = size_t(N) * N * (N / 2 + 1);
  for (int c = 0; c < ncomp; ++c) {
    VP_TRY(vp_fft_r2c_inplace(pl, field_d[c], stream));
    vp_stage stage(pl->ctx, "expand_power", st, 1);
    k_expand_power<<<unsigned((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float2*>(field_d[c]), N, P_d, c == 0);
    VP_CHECK_LAUNCH();
  }
  return VP_OK;
}

extern "C" int vp_fft_unpack_half(vp_pk_plan* pl, const float* packed_d, float* half_d, void* stream) {
  VP_REQUIRE(pl && packed_d && half_d, "vp_fft_unpack_half: null argument");
  vp_call_guard guard(pl->ctx, static_cast<cudaStream_t>(stream));
  const int N = pl->N;
  size_t n = size_t(N) * N * (N / 2 + 1);
  vp_stage stage(pl->ctx, "unpack_half", static_cast<cudaStream_t>(stream), 1);
  k_unpack_half<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(packed_d), N, reinterpret_cast<float2*>(half_d));
  VP_CHECK_LAUNCH();
  return VP_OK;
}

extern "C" int vp_power_bin_full(vp_pk_plan* pl, const double* P_d, double* psum_d, uint64_t* nsample_d, void* stream) {
  VP_REQUIRE(pl && P_d && psum_d && nsample_d, "vp_power_bin_full: null argument");
  vp_call_guard guard(pl->ctx, static_cast<cudaStream_t>(stream));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = pl->N;
  VP_CUDA(cudaMemsetAsync(psum_d, 0, sizeof(double) * pl->nbins, st));
  VP_CUDA(cudaMemsetAsync(nsample_d, 0, sizeof(uint64_t) * pl->nbins, st));
  size_t n3 = size_t(N) * N * N;
  vp_stage stage(pl->ctx, "k5_bin_full", st, 1, 8.0 * double(n3));
  k_bin_full<<<unsigned((n3 + 255) / 256), 256, 0, st>>>(P_d, N, pl->kk2, pl->thr, pl->nbins, psum_d,
                                                       reinterpret_cast<unsigned long long*>(nsample_d));
  VP_CHECK_LAUNCH();
  return VP_OK;
}
