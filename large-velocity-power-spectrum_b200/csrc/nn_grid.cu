// K1: exact nearest-particle gridding on a cell list.  See include/vpower_b200.h (vp_nn_grid).
//
// Cell-list build (no sort; every record moves twice, both times through L2-resident windows -- the measurements behind
// this are in profiles/r2_ubench_scatter_gather.jsonl: a random scatter of whole 32-byte sectors inside a <= 32 MB
// window runs at the copy rate, a random gather over the whole array at a quarter of it):
//   1. k_bin_hist     particles per BUCKET (a bucket = 2^bshift consecutive cells of the linear cell order)
//   2. k_bin_scatter_wc  every particle's packed 32-byte record is appended to its bucket: a tile of 4096 particles claims one
//                        run per touched bucket, is put in bucket order in shared memory and written out by consecutive lanes
//                        (the append frontiers of all buckets together are a few hundred KB, so L2 merges the sectors);
//                        k_bin_scatter (records stored straight from registers) serves ragged tiles and strided input
//   3. k_cell_count      records are now grouped by bucket: per-cell counts with L2-resident atomics
//   4. exclusive scan    -> the cell-start table
//   5. k_cell_place      counting-sort placement inside the bucket's window; the record is rewritten in search format
// Search: a shared-memory staged brick kernel (one CTA = 8 x 8 x 32 lattice nodes, the particles of the covering
// 9 x 9 x 33 cells staged once, coordinates re-based to the brick so that f32 carries ~2^-16 of a cell), a plus-shaped
// wider stage for the nodes it cannot prove, and an exact f64 stage (warp per node, growing block) for near ties.
//
// Exactness: a candidate is accepted only if its distance is strictly below the distance from the
// node to every face of the searched cell block behind which unexamined particles can exist.  All
// deciding arithmetic is f64 with the reference's association ((dx*dx+dy*dy)+dz*dz) and no FMA
// contraction (explicit __dmul_rn/__dadd_rn); ties go to the lowest particle index.
#include <math.h>
#include <algorithm>
#include <stdlib.h>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

struct Grid {
  double ox, oy, oz;     // origin of cell (0,0,0)
  double hx, hy, hz;     // cell size
  double ihx, ihy, ihz;  // 1/cell size
  int gx, gy, gz;
  int use_keep;
  double keep_lo, keep_hi;
  int closed_xlo, closed_xhi;  // 1: particles beyond that x face were dropped (face constrains the proof)
  int ps, vs, rs;              // element strides between consecutive particles in pos / vel / rho (3,3,1 when compact)
  int bshift;                  // bucket of a cell = linear cell index >> bshift,  linear = (cx*gy + cy)*gz + cz
  uint32_t nb;                 // number of buckets
  float hxf, hyf, hzf;         // cell size in f32 (operands straight from the constant bank in the search loops)
};

// ---- record formats
// Bucketed record (between k_bin_scatter and k_cell_place): the position is kept as the cell index plus a 21-bit
// fixed-point offset inside the cell per axis -- exact cell assignment (f64) travels with the record, and the offset
// is good to 2^-22 of a cell, far better than an origin-relative f32.
constexpr int kFixBits = 21;
constexpr uint32_t kFixMax = (1u << kFixBits) - 1u;
constexpr uint32_t kFarBit = 0x80000000u;   // idx bit 31: the particle lies outside the cell grid (clamped into an end cell)
struct __align__(16) RecA { uint32_t u0, u1, cell, idx; };
struct __align__(32) Rec32 { RecA a; float4 b; };
// Sorted record (search format), one per particle in cell order: float4 a = (ux, uy, uz32, idx bits) [+ float4 b = (v'x, v'y,
// v'z, m) with a payload, the two halves in one 32-byte record]: ux, uy = offset from the low corner of the particle's cell; uz32 = offset from the low corner of
// the aligned 32-cell z block the cell lies in (cz & ~31) -- a brick of the search kernel spans exactly one such block
// plus one cell, so staging re-bases z with one add.  Every consumer knows the cell (it walks the cell table).
typedef float4 rec_t;

// (explicit round-to-nearest operations: with FMA contraction the compiler may form f - c from the UNROUNDED product, which
// comes out at -1e-17 for a particle whose f rounds up to an integer -- and every kernel must assign the same cell)
__device__ __forceinline__ int cell_fix(double x, double o, double ih, int g, uint32_t& fix, bool& far) {
  const double f = __dmul_rn(__dsub_rn(x, o), ih);
  int c;
  if (!(f > 0.0)) c = 0;
  else if (f >= double(g)) c = g - 1;
  else c = int(f);
  double u = __dsub_rn(f, double(c));            // in [0,1) unless the particle was clamped into an end cell (or is NaN)
  if (!(u >= 0.0)) { far = far || (u < 0.0) || (u != u); u = 0.0; }
  if (u >= 1.0) { far = true; u = 1.0; }
  uint32_t q = uint32_t(__dmul_rn(u, 2097152.0));
  fix = q > kFixMax ? kFixMax : q;
  return c;
}

__device__ __forceinline__ void st256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                      uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5),
               "r"(a6), "r"(a7) : "memory");
}
__device__ __forceinline__ void ld256(const void* p, uint32_t (&a)[8]) {
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "l"(p));
}

// true: the particle takes part (inside the kept x range of a slab, or no filter)
template <typename T>
__device__ __forceinline__ bool load_pos(const T* __restrict__ pos, const Grid& g, int64_t i, double& x, double& y, double& z) {
  x = pos[size_t(g.ps) * i];
  y = pos[size_t(g.ps) * i + 1];
  z = pos[size_t(g.ps) * i + 2];
  // slab filter: a particle beyond a CLOSED face is dropped; beyond an open (domain-edge) face it is kept and clamped
  return !(g.use_keep && ((g.closed_xlo && !(x >= g.keep_lo)) || (g.closed_xhi && !(x <= g.keep_hi))));
}

// N consecutive values (N * sizeof(T) bytes, W-byte aligned) with W-byte loads: a thread that owns FOUR CONSECUTIVE particles
// fetches their 12 position components with three 16-byte loads instead of twelve 4-byte loads 12 bytes apart (the bucket
// kernels are bound by the load/store and shared-memory pipe, not by DRAM: ncu MIO-throttle + short-scoreboard 25 %)
template <int W, typename T, int N>
__device__ __forceinline__ void load_run(const T* __restrict__ p, T (&out)[N]) {
  static_assert((N * sizeof(T)) % W == 0 && (W == 8 || W == 16), "load_run: whole vectors only");
  constexpr int NV = int(N * sizeof(T)) / W;
  if (W == 16) {
    uint4 tmp[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) tmp[k] = reinterpret_cast<const uint4*>(p)[k];
    memcpy(out, tmp, sizeof(out));
  } else {
    uint2 tmp[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) tmp[k] = reinterpret_cast<const uint2*>(p)[k];
    memcpy(out, tmp, sizeof(out));
  }
}
__device__ __forceinline__ bool keep_x(const Grid& g, double x) {
  return !(g.use_keep && ((g.closed_xlo && !(x >= g.keep_lo)) || (g.closed_xhi && !(x <= g.keep_hi))));
}

// Bucket pass.  Particles are taken in TILES of kBinTile consecutive particles; tile t appends to SUB-STREAM t % kSub of each
// bucket (kSub independent append cursors per bucket keep the per-address atomic rate low; the sub-streams of a bucket are
// laid out one after the other, so downstream a bucket is still one contiguous run of records).
// Part 1 (k_bin_hist): particles per (bucket, sub-stream) -- persistent CTAs, shared-memory counters, one flush per CTA.
constexpr int kBinTile = 4096;
constexpr int kSub = 8;
constexpr int kClu = 8;      // tiles per thread-block cluster of the scatter kernel: they claim their runs together
template <typename T, int VEC>   // 0: element by element; 1: compact [np,3], 16-byte loads; 2: 32-byte rows (x y z vx | vy vz rho -), f32
__global__ void __launch_bounds__(256) k_bin_hist(const T* __restrict__ pos, int64_t np, Grid g, uint32_t* __restrict__ hist_g,
                                                   uint32_t tile0) {
  extern __shared__ uint32_t sh_hist[];         // [nb][kSub]
  const uint32_t nh = g.nb * kSub;
  for (uint32_t b = threadIdx.x; b < nh; b += 256) sh_hist[b] = 0u;
  __syncthreads();
  const int64_t ntiles = (np + kBinTile - 1) / kBinTile;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint32_t sub = uint32_t(((tile0 + tile) / kClu) % kSub);   // all tiles of a cluster append to the same sub-stream
    const int64_t base = tile * kBinTile;
    if constexpr (VEC != 0) if (base + kBinTile <= np) {
      // compact, 16-byte aligned positions: four consecutive particles per thread and step
#pragma unroll 2
      for (int r = 0; r < kBinTile / 1024; ++r) {
        T pv[12];
        if constexpr (VEC == 2) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            // (consecutive lanes take consecutive rows: a warp load covers 1 KB, not 32 separate lines)
            const uint4 q = *reinterpret_cast<const uint4*>(pos + 8 * (base + (r * 4 + u) * 256 + threadIdx.x));
            pv[3 * u] = T(__uint_as_float(q.x)); pv[3 * u + 1] = T(__uint_as_float(q.y)); pv[3 * u + 2] = T(__uint_as_float(q.z));
          }
        } else {
          load_run<16>(pos + 3 * (base + (r * 256 + threadIdx.x) * 4), pv);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double x = pv[3 * u], y = pv[3 * u + 1], z = pv[3 * u + 2];
          if (!keep_x(g, x)) continue;
          uint32_t fx, fy, fz;
          bool far = false;
          const int cx = cell_fix(x, g.ox, g.ihx, g.gx, fx, far), cy = cell_fix(y, g.oy, g.ihy, g.gy, fy, far),
                    cz = cell_fix(z, g.oz, g.ihz, g.gz, fz, far);
          const uint32_t lin = (uint32_t(cx) * uint32_t(g.gy) + uint32_t(cy)) * uint32_t(g.gz) + uint32_t(cz);
          atomicAdd(&sh_hist[(lin >> g.bshift) * kSub + sub], 1u);
        }
      }
      continue;
    }
#pragma unroll 4
    for (int r = 0; r < kBinTile / 256; ++r) {
      const int64_t i = base + r * 256 + threadIdx.x;
      if (i >= np) break;
      double x, y, z;
      if (!load_pos(pos, g, i, x, y, z)) continue;
      uint32_t fx, fy, fz;
      bool far = false;
      const int cx = cell_fix(x, g.ox, g.ihx, g.gx, fx, far), cy = cell_fix(y, g.oy, g.ihy, g.gy, fy, far),
                cz = cell_fix(z, g.oz, g.ihz, g.gz, fz, far);
      const uint32_t lin = (uint32_t(cx) * uint32_t(g.gy) + uint32_t(cy)) * uint32_t(g.gz) + uint32_t(cz);
      atomicAdd(&sh_hist[(lin >> g.bshift) * kSub + sub], 1u);
    }
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < nh; b += 256) {
    const uint32_t c = sh_hist[b];
    if (c) atomicAdd(hist_g + b, c);
  }
}

// exclusive prefix of the (bucket, sub-stream) histogram -> append cursors; the total is the number of kept particles.  One CTA.
__global__ void __launch_bounds__(1024) k_bin_offsets(const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor, uint32_t n,
                                                       vp_nn_stats_dev* __restrict__ stats) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_s = 0u;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t b = base + tid;
    const uint32_t v = b < n ? hist[b] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
      const uint32_t sv = wsum[lane];
      uint32_t is = sv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, is, o);
        if (lane >= o) is += t;
      }
      wsum[lane] = is - sv;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    if (b < n) cursor[b] = carry + wsum[w] + incl - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[31] + incl;
    __syncthreads();
  }
  if (tid == 0) stats->n_kept = carry_s;
}

template <typename T>
struct PayloadIn {
  const T* vel;   // [np,3] or null
  const T* rho;   // [np] or null (rho = 1)
  T lcell3;
};

// One thread per particle: cell + in-cell offset in f64, the record appended to the particle's bucket.
//   a = (21-bit offsets x|y|z, linear cell, particle index [| far flag])     b = (vx', vy', vz', m)  [only with a payload]
// with v' = (rho*v)/rho and m = rho*Lcell^3 evaluated in the input dtype (interp.py:199-213,272-273).
// Part 2 (k_bin_scatter): one CTA of 1024 threads per tile, four particles per thread held in registers.  The tile's
// particles are ranked per bucket with shared-memory atomics, every touched bucket claims ONE contiguous run of slots from
// its sub-stream cursor, and the records go straight from registers to their slots: all records of a run are written
// within the same few microseconds, and temporally adjacent tiles append adjacent runs, so L2 assembles whole lines.
template <typename T, bool PAY, int THREADS, bool VEC>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS) k_bin_scatter(const T* __restrict__ pos, PayloadIn<T> pin, int64_t np, int64_t i0, Grid g,
                                                                        uint32_t* __restrict__ cursor, uint32_t tile0, void* __restrict__ rec1) {
  // pos / pin.vel / pin.rho point at particle i0 (a chunk); the stored index is global (i0 + local).
  // 32 registers per thread = two CTAs per SM, so that one tile's loads overlap the other's stores: only the search half
  // of the record is carried across the two barriers, the payload is loaded right before the store.  (Persistent CTAs with
  // all position loads hoisted in front were slower: 31.0 vs 25.4 ms at 2^30 particles.)
  extern __shared__ uint32_t sh_cnt[];          // [nb] particles of this tile per bucket, then the first slot of its run
  for (uint32_t b = threadIdx.x; b < g.nb; b += THREADS) sh_cnt[b] = 0u;
  __syncthreads();
  constexpr int kItems = 4, kTile = THREADS * kItems;     // (kTile divides kBinTile: the sub-stream is that of the histogram's tile)
  const uint32_t sub = ((tile0 + uint32_t((int64_t(blockIdx.x) * kTile) / kBinTile)) / kClu) % kSub;
  const int64_t base = int64_t(blockIdx.x) * kTile;
  uint32_t ra[kItems][3], slot[kItems];          // fixed-point offsets (2 words), linear cell; bucket-local rank, later the slot
  bool farv[kItems];
  // VEC (compact 16-byte aligned arrays; the host launches it on whole tiles only): thread t owns particles 4t .. 4t+3 of the
  // tile and fetches them with 16-byte loads; otherwise particle r*THREADS + t, element by element
  constexpr bool vec = VEC;
  static_assert(!VEC || kItems == 4, "the vector path takes four consecutive particles per thread");
  T pv[VEC ? 12 : 1];
  if constexpr (VEC) { if (vec) load_run<16>(pos + 3 * (base + 4 * int64_t(threadIdx.x)), pv); }
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int64_t i = vec ? base + 4 * int64_t(threadIdx.x) + r : base + r * THREADS + threadIdx.x;
    slot[r] = 0xffffffffu;
    farv[r] = false;
    double x, y, z;
    if (VEC && vec) {
      x = pv[3 * r]; y = pv[3 * r + 1]; z = pv[3 * r + 2];
      if (!keep_x(g, x)) continue;
    } else if (i >= np || !load_pos(pos, g, i, x, y, z)) continue;
    uint32_t fx, fy, fz;
    bool far = false;
    const int cx = cell_fix(x, g.ox, g.ihx, g.gx, fx, far), cy = cell_fix(y, g.oy, g.ihy, g.gy, fy, far),
              cz = cell_fix(z, g.oz, g.ihz, g.gz, fz, far);
    const uint32_t lin = (uint32_t(cx) * uint32_t(g.gy) + uint32_t(cy)) * uint32_t(g.gz) + uint32_t(cz);
    slot[r] = atomicAdd(&sh_cnt[lin >> g.bshift], 1u);
    const unsigned long long w = (unsigned long long)fx | ((unsigned long long)fy << kFixBits) | ((unsigned long long)fz << (2 * kFixBits));
    ra[r][0] = uint32_t(w); ra[r][1] = uint32_t(w >> 32); ra[r][2] = lin;
    farv[r] = far;
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < g.nb; b += THREADS) {
    const uint32_t c = sh_cnt[b];
    if (c) sh_cnt[b] = atomicAdd(cursor + b * kSub + sub, c);
  }
  __syncthreads();
  T vpair[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, rpair[2] = {T(1), T(1)};
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    // payload of two consecutive particles at a time on the vector path (8-byte loads for f32, 16-byte loads for f64)
    T vv[(VEC && PAY) ? 6 : 1], rv[(VEC && PAY) ? 2 : 1];
    if constexpr (VEC && PAY) if (vec && (r & 1) == 0) {
      const int64_t i2 = base + 4 * int64_t(threadIdx.x) + r;
      load_run<sizeof(T) == 4 ? 8 : 16>(pin.vel + 3 * i2, vv);
      if (pin.rho) load_run<sizeof(T) == 4 ? 8 : 16>(pin.rho + i2, rv);
      vpair[0] = vv[0]; vpair[1] = vv[1]; vpair[2] = vv[2]; vpair[3] = vv[3]; vpair[4] = vv[4]; vpair[5] = vv[5];
      rpair[0] = rv[0]; rpair[1] = rv[1];
    }
    if (slot[r] == 0xffffffffu) continue;
    const int64_t i = vec ? base + 4 * int64_t(threadIdx.x) + r : base + r * THREADS + threadIdx.x;
    const uint32_t dst = sh_cnt[ra[r][2] >> g.bshift] + slot[r];
    const uint32_t idx = uint32_t(i0 + i) | (farv[r] ? kFarBit : 0u);
    if (PAY) {
      T vx, vy, vz;
      if (VEC && vec) { vx = vpair[3 * (r & 1)]; vy = vpair[3 * (r & 1) + 1]; vz = vpair[3 * (r & 1) + 2]; }
      else { vx = pin.vel[size_t(g.vs) * i]; vy = pin.vel[size_t(g.vs) * i + 1]; vz = pin.vel[size_t(g.vs) * i + 2]; }
      T m = pin.lcell3;
      if (pin.rho) {
        const T rr = (VEC && vec) ? rpair[r & 1] : pin.rho[size_t(g.rs) * i];
        vx = (vx * rr) / rr;
        vy = (vy * rr) / rr;
        vz = (vz * rr) / rr;
        m = rr * pin.lcell3;
      }
      st256(static_cast<Rec32*>(rec1) + dst, ra[r][0], ra[r][1], ra[r][2], idx, __float_as_uint(float(vx)), __float_as_uint(float(vy)),
            __float_as_uint(float(vz)), __float_as_uint(float(m)));
    } else {
      static_cast<uint4*>(rec1)[dst] = make_uint4(ra[r][0], ra[r][1], ra[r][2], idx);
    }
  }
}

// Part 2, write-combining form (whole tiles of compact, 16-byte aligned arrays -- the usual case; everything else takes
// k_bin_scatter above).  What the plain kernel pays for is not DRAM but the store instruction itself: a warp whose 32 lanes write
// 32 unrelated sectors is served one sector at a time by the load/store unit (tools/ubench/bin_scatter.cu: 17 of its 25 ms at
// 2^30 particles are the stores; with the stores removed the kernel takes 8 ms).  Here the tile's records are first put in BUCKET
// ORDER in shared memory -- position = exclusive prefix of the tile's bucket counts + rank inside the bucket, half a tile (kWin
// records, 64 KB) at a time so that two CTAs fit an SM -- and then written out by consecutive lanes: a warp covers ~8 runs of
// consecutive destinations instead of 32 sectors.  Same records, same runs, same cursors as the plain kernel.
constexpr int kWin = 2048;
constexpr uint32_t kWcBuckets = 2048;      // = kMaxBuckets: fixed shared-memory layout (compile-time offsets, two registers less)
template <typename T, bool PAY, bool ROWS>   // ROWS: pos/vel/rho are columns 0-2 / 3-5 / 6 of one [np,8] f32 row array (the slab exchange's rows)
__global__ void __launch_bounds__(1024, 2) k_bin_scatter_wc(const T* __restrict__ pos, PayloadIn<T> pin, int64_t i0, Grid g,
                                                           uint32_t* __restrict__ cursor, uint32_t tile0, void* __restrict__ rec1) {
  static_assert(kBinTile == 4096, "four consecutive particles per thread, 1024 threads");
  extern __shared__ __align__(16) unsigned char wsm[];
  uint4* sA = reinterpret_cast<uint4*>(wsm);                               // [kWin] search half of the staged records
  uint4* sB = sA + kWin;                                                   // [kWin] payload half (PAY only)
  uint32_t* sh_cnt = reinterpret_cast<uint32_t*>(sA + (PAY ? 2 : 1) * kWin);   // [nb] counts, then the tile-local exclusive prefix
  uint32_t* sh_delta = sh_cnt + kWcBuckets;                                // [nb] first slot of the claimed run - tile-local prefix
  uint32_t* sh_w = sh_delta + kWcBuckets;                                      // [33] warp totals of the scan
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  for (uint32_t b = tid; b < g.nb; b += 1024) sh_cnt[b] = 0u;
  __syncthreads();
  const uint32_t sub = ((tile0 + blockIdx.x) / kClu) % kSub;
  // this thread's four particles: base + r*kStep -- four consecutive ones of compact arrays (one batch of 16-byte loads), or
  // particles 1024 apart of a row array (consecutive lanes then read consecutive 32-byte rows)
  constexpr int kStep = ROWS ? 1024 : 1;
  const int64_t base = int64_t(blockIdx.x) * kBinTile + (ROWS ? 1 : 4) * int64_t(tid);
  const uint32_t idx0 = uint32_t(i0 + base);
  const T* __restrict__ velp = (PAY && !ROWS) ? pin.vel + 3 * base : nullptr;      // this thread's four particles
  const T* __restrict__ rhop = (PAY && pin.rho) ? (ROWS ? pin.rho : pin.rho + base) : nullptr;   // (ROWS: only its presence matters)
  uint32_t ra[4][3], slot[4];       // slot: rank inside the tile's bucket (< 4096) | far flag in bit 30; all ones = not taking part
  constexpr int kHalves = sizeof(T) == 4 ? 1 : 2, kPer = 4 / kHalves;      // f64: two particles (48 bytes) per batch of loads
#pragma unroll
  for (int h = 0; h < kHalves; ++h) {
    T pv[3 * kPer];
    if constexpr (ROWS) {
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const uint4 q = *reinterpret_cast<const uint4*>(pos + 8 * (base + (h * kPer + u) * kStep));
        pv[3 * u] = T(__uint_as_float(q.x)); pv[3 * u + 1] = T(__uint_as_float(q.y)); pv[3 * u + 2] = T(__uint_as_float(q.z));
      }
    } else {
      load_run<16>(pos + 3 * (base + h * kPer), pv);
    }
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int r = h * kPer + u;
      slot[r] = 0xffffffffu;
      const double x = pv[3 * u], y = pv[3 * u + 1], z = pv[3 * u + 2];
      if (!keep_x(g, x)) continue;
      uint32_t fx, fy, fz;
      bool far = false;
      const int cx = cell_fix(x, g.ox, g.ihx, g.gx, fx, far), cy = cell_fix(y, g.oy, g.ihy, g.gy, fy, far),
                cz = cell_fix(z, g.oz, g.ihz, g.gz, fz, far);
      const uint32_t lin = (uint32_t(cx) * uint32_t(g.gy) + uint32_t(cy)) * uint32_t(g.gz) + uint32_t(cz);
      const uint32_t rk = atomicAdd(&sh_cnt[lin >> g.bshift], 1u);
      const unsigned long long w = (unsigned long long)fx | ((unsigned long long)fy << kFixBits) | ((unsigned long long)fz << (2 * kFixBits));
      ra[r][0] = uint32_t(w); ra[r][1] = uint32_t(w >> 32); ra[r][2] = lin;
      slot[r] = rk | (far ? 0x40000000u : 0u);
    }
  }
  __syncthreads();
  // exclusive scan of the bucket counts in bucket order (thread t owns buckets 2t, 2t+1: nb <= 2048), one run claimed per bucket
  const uint32_t b0 = 2u * tid, b1 = b0 + 1u;
  const uint32_t c0 = b0 < g.nb ? sh_cnt[b0] : 0u, c1 = b1 < g.nb ? sh_cnt[b1] : 0u;
  uint32_t incl = c0 + c1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sh_w[wp] = incl;
  __syncthreads();
  if (wp == 0) {
    const uint32_t v = sh_w[lane];
    uint32_t is = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, is, o);
      if (lane >= o) is += t;
    }
    sh_w[lane] = is - v;
    if (lane == 31) sh_w[32] = is;
  }
  __syncthreads();
  const uint32_t tp0 = sh_w[wp] + incl - (c0 + c1), tp1 = tp0 + c0;
  const uint32_t gb0 = c0 ? atomicAdd(cursor + b0 * kSub + sub, c0) : 0u;     // (both claims in flight before either is used)
  const uint32_t gb1 = c1 ? atomicAdd(cursor + b1 * kSub + sub, c1) : 0u;
  if (b0 < g.nb) { sh_cnt[b0] = tp0; sh_delta[b0] = gb0 - tp0; }
  if (b1 < g.nb) { sh_cnt[b1] = tp1; sh_delta[b1] = gb1 - tp1; }
  const uint32_t nv = sh_w[32];      // particles of the tile that take part
  __syncthreads();
  for (uint32_t w0 = 0; w0 < nv; w0 += kWin) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (slot[r] == 0xffffffffu) continue;
      const uint32_t p = sh_cnt[ra[r][2] >> g.bshift] + (slot[r] & 0xffffu) - w0;
      if (p >= uint32_t(kWin)) continue;
      sA[p] = make_uint4(ra[r][0], ra[r][1], ra[r][2], (idx0 + r * kStep) | ((slot[r] & 0x40000000u) ? kFarBit : 0u));
      if (PAY) {
        T vx, vy, vz, rrow = T(1);
        if constexpr (ROWS) {
          const T* row = pos + 8 * (base + r * kStep);
          const uint4 hi = *reinterpret_cast<const uint4*>(row + 4);
          vx = row[3]; vy = T(__uint_as_float(hi.x)); vz = T(__uint_as_float(hi.y)); rrow = T(__uint_as_float(hi.z));
        } else {
          vx = velp[3 * r]; vy = velp[3 * r + 1]; vz = velp[3 * r + 2];
        }
        T m = pin.lcell3;
        if (rhop) {
          const T rr = ROWS ? rrow : rhop[r];
          vx = (vx * rr) / rr;
          vy = (vy * rr) / rr;
          vz = (vz * rr) / rr;
          m = rr * pin.lcell3;
        }
        sB[p] = make_uint4(__float_as_uint(float(vx)), __float_as_uint(float(vy)), __float_as_uint(float(vz)), __float_as_uint(float(m)));
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kWin / 1024; ++k) {
      const uint32_t q = uint32_t(tid) + k * 1024;
      if (w0 + q < nv) {
        const uint4 a = sA[q];
        const uint32_t dst = sh_delta[a.z >> g.bshift] + w0 + q;
        if (PAY) {
          const uint4 b = sB[q];
          st256(static_cast<Rec32*>(rec1) + dst, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w);
        } else {
          static_cast<uint4*>(rec1)[dst] = a;
        }
      }
    }
    if (w0 + kWin < nv) __syncthreads();
  }
}

// The same with thread-block clusters (sm_90+/sm_100): the kClu CTAs of a cluster rank their tiles independently, add up
// their per-bucket counts through distributed shared memory, claim ONE run per bucket for the whole cluster (kClu times
// longer: ~1 KB instead of ~128 B at 1024 buckets) and write into it side by side -- DRAM sees the long runs it needs
// (profiles/r2_ubench_scatter_gather.jsonl: runs of 32 records move at the copy rate, runs of 4 at ~55 %).
template <typename T, bool PAY>
__global__ void __cluster_dims__(kClu, 1, 1) __launch_bounds__(1024)
k_bin_scatter_clu(const T* __restrict__ pos, PayloadIn<T> pin, int64_t np, int64_t i0, Grid g, uint32_t* __restrict__ cursor,
                  uint32_t tile0, void* __restrict__ rec1) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ uint32_t sh_cnt[];          // [nb] counts of this tile, later the first slot of this tile's part of the run
  uint32_t* sh_base = sh_cnt + g.nb;            // [nb] (cluster rank 0 only) first slot of the cluster's run
  const unsigned crank = cluster.block_rank();
  for (uint32_t b = threadIdx.x; b < g.nb; b += 1024) sh_cnt[b] = 0u;
  __syncthreads();
  constexpr int kItems = kBinTile / 1024;
  const uint32_t sub = ((tile0 + blockIdx.x) / kClu) % kSub;
  const int64_t base = int64_t(blockIdx.x) * kBinTile;
  uint32_t ra[kItems][4], rb[kItems][4], bkt[kItems], rank[kItems];
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int64_t i = base + r * 1024 + threadIdx.x;
    bkt[r] = 0xffffffffu;
    double x, y, z;
    if (i >= np || !load_pos(pos, g, i, x, y, z)) continue;
    uint32_t fx, fy, fz;
    bool far = false;
    const int cx = cell_fix(x, g.ox, g.ihx, g.gx, fx, far), cy = cell_fix(y, g.oy, g.ihy, g.gy, fy, far),
              cz = cell_fix(z, g.oz, g.ihz, g.gz, fz, far);
    const uint32_t lin = (uint32_t(cx) * uint32_t(g.gy) + uint32_t(cy)) * uint32_t(g.gz) + uint32_t(cz);
    bkt[r] = lin >> g.bshift;
    rank[r] = atomicAdd(&sh_cnt[bkt[r]], 1u);
    const unsigned long long w = (unsigned long long)fx | ((unsigned long long)fy << kFixBits) | ((unsigned long long)fz << (2 * kFixBits));
    ra[r][0] = uint32_t(w); ra[r][1] = uint32_t(w >> 32); ra[r][2] = lin; ra[r][3] = uint32_t(i0 + i) | (far ? kFarBit : 0u);
    if (PAY) {
      T vx = pin.vel[size_t(g.vs) * i], vy = pin.vel[size_t(g.vs) * i + 1], vz = pin.vel[size_t(g.vs) * i + 2];
      T m = pin.lcell3;
      if (pin.rho) {
        const T rr = pin.rho[size_t(g.rs) * i];
        vx = (vx * rr) / rr;
        vy = (vy * rr) / rr;
        vz = (vz * rr) / rr;
        m = rr * pin.lcell3;
      }
      rb[r][0] = __float_as_uint(float(vx)); rb[r][1] = __float_as_uint(float(vy)); rb[r][2] = __float_as_uint(float(vz));
      rb[r][3] = __float_as_uint(float(m));
    }
  }
  cluster.sync();                               // every tile of the cluster has its counts
  // per bucket (at most two per thread, nb <= 2048): slots taken by the lower-ranked tiles, and the cluster's total
  uint32_t off[2] = {0u, 0u}, tot[2] = {0u, 0u};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint32_t b = threadIdx.x + q * 1024;
    if (b < g.nb) {
      for (unsigned r = 0; r < unsigned(kClu); ++r) {
        const uint32_t c = *cluster.map_shared_rank(sh_cnt + b, r);
        if (r < crank) off[q] += c;
        tot[q] += c;
      }
    }
  }
  cluster.sync();                               // all remote reads of the counts are done: they may be overwritten now
  if (crank == 0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint32_t b = threadIdx.x + q * 1024;
      if (b < g.nb) sh_base[b] = tot[q] ? atomicAdd(cursor + b * kSub + sub, tot[q]) : 0u;
    }
  }
  cluster.sync();                               // the claims of rank 0 are visible
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint32_t b = threadIdx.x + q * 1024;
    if (b < g.nb) sh_cnt[b] = *cluster.map_shared_rank(sh_base + b, 0) + off[q];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    if (bkt[r] == 0xffffffffu) continue;
    const uint32_t dst = sh_cnt[bkt[r]] + rank[r];
    if (PAY) st256(static_cast<Rec32*>(rec1) + dst, ra[r][0], ra[r][1], ra[r][2], ra[r][3], rb[r][0], rb[r][1], rb[r][2], rb[r][3]);
    else static_cast<uint4*>(rec1)[dst] = make_uint4(ra[r][0], ra[r][1], ra[r][2], ra[r][3]);
  }
  cluster.sync();                               // rank 0's shared memory stays alive until every tile has read its claims
}

// per-cell counts: T[cell] += 1 (the records of one bucket are contiguous, so the counters in flight are L2 resident).
// kIlp records per thread, block-strided, so that a thread has several independent loads / atomics in flight (one record per
// thread left the kernels latency bound: ncu 93 % long-scoreboard stalls at 80 % occupancy).
constexpr int kIlp = 4;
template <bool PAY>
__global__ void __launch_bounds__(256) k_cell_count(const void* __restrict__ rec1, const vp_nn_stats_dev* __restrict__ stats,
                                                     uint32_t* __restrict__ tab) {
  const uint32_t n = uint32_t(stats->n_kept);
  const uint32_t i0 = blockIdx.x * (256u * kIlp) + threadIdx.x;
  if (i0 >= n) return;
  uint32_t cell[kIlp];
#pragma unroll
  for (int u = 0; u < kIlp; ++u) {
    const uint32_t i = i0 + u * 256u;
    cell[u] = 0xffffffffu;
    if (i < n) cell[u] = PAY ? static_cast<const Rec32*>(rec1)[i].a.cell : static_cast<const RecA*>(rec1)[i].cell;
  }
#pragma unroll
  for (int u = 0; u < kIlp; ++u)
    if (cell[u] != 0xffffffffu) atomicAdd(tab + cell[u], 1u);
}

// counting-sort placement: the record goes to (start of its cell) + (a slot handed out by the cell's cursor), rewritten in
// search format.  The order of the particles INSIDE a cell is whatever the cursors hand out; no result depends on it.
struct PlaceGeom {
  float sx, sy, sz;   // cell size * 2^-21 per axis
  float hz;
  uint32_t gz;
};
template <bool PAY>
__global__ void __launch_bounds__(256) k_cell_place(const void* __restrict__ rec1, vp_nn_stats_dev* __restrict__ stats,
                                                     uint32_t* __restrict__ tab, PlaceGeom pg, void* __restrict__ srec) {
  const uint32_t n = uint32_t(stats->n_kept);
  const uint32_t i0 = blockIdx.x * (256u * kIlp) + threadIdx.x;
  if (i0 >= n) return;
  uint32_t r[kIlp][8], dst[kIlp];
  bool on[kIlp];
#pragma unroll
  for (int u = 0; u < kIlp; ++u) {
    const uint32_t i = i0 + u * 256u;
    on[u] = i < n;
    if (on[u]) {
      if (PAY) {
        ld256(static_cast<const Rec32*>(rec1) + i, r[u]);
      } else {
        const uint4 v = static_cast<const uint4*>(rec1)[i];
        r[u][0] = v.x; r[u][1] = v.y; r[u][2] = v.z; r[u][3] = v.w;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kIlp; ++u)
    if (on[u]) dst[u] = atomicAdd(tab + r[u][2], 1u);
  unsigned nfar = 0;
#pragma unroll
  for (int u = 0; u < kIlp; ++u) {
    if (!on[u]) continue;
    const unsigned long long w = (unsigned long long)r[u][0] | ((unsigned long long)r[u][1] << 32);
    const uint32_t fx = uint32_t(w) & kFixMax, fy = uint32_t(w >> kFixBits) & kFixMax, fz = uint32_t(w >> (2 * kFixBits)) & kFixMax;
    const uint32_t cz = r[u][2] % pg.gz;
    const float ux = (float(fx) + 0.5f) * pg.sx, uy = (float(fy) + 0.5f) * pg.sy;
    const float uz = fmaf(float(cz & 31u), pg.hz, (float(fz) + 0.5f) * pg.sz);
    nfar += (r[u][3] & kFarBit) ? 1u : 0u;
    // ONE 32-byte store per record: a whole sector.  (Writing the search half and the payload half to two arrays -- two
    // scattered 16-byte partial-sector stores -- took 43.5 ms instead of 18.0 ms at 2^30 particles.)
    if (PAY)
      st256(static_cast<Rec32*>(srec) + dst[u], __float_as_uint(ux), __float_as_uint(uy), __float_as_uint(uz), r[u][3], r[u][4], r[u][5],
            r[u][6], r[u][7]);
    else
      static_cast<float4*>(srec)[dst[u]] = make_float4(ux, uy, uz, __uint_as_float(r[u][3]));
  }
  if (nfar) atomicAdd(&stats->n_far, (unsigned long long)nfar);
}

struct Best {
  double d2;
  int idx;
};

__device__ __forceinline__ void consider(Best& b, double qx, double qy, double qz, double x, double y, double z, int i) {
  double dx = qx - x, dy = qy - y, dz = qz - z;
  double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  if (d2 < b.d2 || (d2 == b.d2 && i < b.idx)) { b.d2 = d2; b.idx = i; }
}

// Distance from coordinate q to the nearest face of the cell block [c0,c1] (inclusive) along one axis
// behind which unexamined particles may exist; +inf when both sides are open ends of the grid.
__device__ __forceinline__ double axis_margin(double q, double o, double h, int c0, int c1, int g, bool closed_lo,
                                              bool closed_hi) {
  double m = INFINITY;
  if (c0 > 0) m = fmin(m, q - (o + double(c0) * h));
  else if (closed_lo) m = fmin(m, q - o);
  if (c1 < g - 1) m = fmin(m, (o + double(c1 + 1) * h) - q);
  else if (closed_hi) m = fmin(m, (o + double(g) * h) - q);
  return m;
}

__device__ __forceinline__ bool proven(const Best& b, double margin) {
  if (margin == INFINITY) return true;
  if (!(margin > 0.0)) return false;
  double ms = margin * (1.0 - 1.0 / 1048576.0);  // slack for the rounding of the cell assignment
  return b.d2 < ms * ms;
}

struct Lattice {
  const double *qx, *qy, *qz;  // node coordinates
  // first cell of the node's 2-cell window per axis: [w, w+1] is the pair of cells whose common face is nearest
  // to the node (with the default grid the nodes sit exactly on cell corners, so the 2x2x2 block proves radius h)
  const int *wx, *wy, *wz;
  // node coordinate minus the low corner of cell w, evaluated in f64 on the host and rounded to f32: every distance
  // of the f32 prefilter is formed from numbers of the size of a cell
  const float *rx, *ry, *rz;
  // per axis, the distance from the node to the nearest face of its 2-cell window behind which unexamined particles may
  // exist (rounded down, with the cell-assignment slack already taken off; +inf when the window reaches an open end)
  const float *mx, *my, *mz;
  int nx, ny, nz;
};

// f32 prefilter.  Distances are evaluated in f32 on window-relative coordinates and the two smallest are tracked.  A node
// is settled by the prefilter only if (a) the runner-up is farther than the winner by more than a rigorous bound on the
// f32 error of both and (b) the winner (plus that bound) is strictly inside the proof margin of the examined block.
//
// Error bound: the stored offset is good to 2^-22 h (fixed point) and every f32 operation that forms a coordinate
// difference works on numbers below 34 h (brick-relative z) -- at most seven roundings of 2^-24 * 34 h: per axis
//   |d~x - dx| <= 2e-5 h =: eps,      |d~^2 - d^2| <= 2 sqrt(3) d eps + 3 eps^2 + 2^-22 d^2
struct Cand {
  float b1, b2;
  int bi;
};
__device__ __forceinline__ void cand_update(Cand& c, float d, int p) {
  const bool lt = d < c.b1;
  c.b2 = fminf(c.b2, lt ? c.b1 : d);
  c.bi = lt ? p : c.bi;
  c.b1 = lt ? d : c.b1;
}
// records [s, e) of the sorted array (element stride `rs` float4), query relative to the frame the records are stored in
__device__ __forceinline__ void scan_range(const rec_t* __restrict__ part, int rs, uint32_t s, uint32_t e, float qx, float qy, float qz,
                                           Cand& c) {
#pragma unroll 1
  for (uint32_t p = s; p < e; ++p) {
    const float4 q = __ldg(part + size_t(p) * rs);
    const float dx = qx - q.x, dy = qy - q.y, dz = qz - q.z;
    cand_update(c, fmaf(dz, dz, fmaf(dy, dy, dx * dx)), int(p));
  }
}
// 0: settled, 1: proven-or-not but ambiguous / empty -> exact, 2: unambiguous but unproven -> wider block
__device__ __forceinline__ int judge(const Cand& c, float eps, float margin) {
  if (c.bi < 0) return 2;
  const float rb = sqrtf(c.b2 < INFINITY ? c.b2 : c.b1);
  const float tol = 8.f * rb * eps + 8.f * eps * eps + 1e-6f * rb * rb;   // >= err(b1) + err(b2)
  const bool proven = (margin == INFINITY) || (margin > 0.f && c.b1 + tol < margin * margin);
  if (!proven) return 2;
  return (c.b2 - c.b1 > tol) ? 0 : 1;
}

// Field planes written straight from the search (K3 fused into K1): the stage that settles a node reads the payload half
// of the winner's record -- the same 32-byte sector the search has just read -- and writes the node's value of every
// requested plane (interp.py:272-273, 501-557).  Null pointers = plane not wanted; all null = no fusion.
struct FieldOut {
  float *vx, *vy, *vz, *px, *py, *pz, *e, *m;
  int on;
};
__device__ __forceinline__ void write_fields(const FieldOut& f, const rec_t* __restrict__ part, int rs, size_t node, int pos) {
  const float4 w = __ldg(part + size_t(pos) * rs + 1);
  if (f.vx) f.vx[node] = w.x;
  if (f.vy) f.vy[node] = w.y;
  if (f.vz) f.vz[node] = w.z;
  if (f.px) f.px[node] = w.x * w.w;
  if (f.py) f.py[node] = w.y * w.w;
  if (f.pz) f.pz[node] = w.z * w.w;
  if (f.e) f.e[node] = w.w * (w.x * w.x + w.y * w.y + w.z * w.z);   // interp.py:546 (no 1/2)
  if (f.m) f.m[node] = w.w;
}

struct SearchOut {
  int32_t* nn;        // particle index per node (may be null)
  int32_t* nn_pos;    // sorted position per node (may be null)
  uint32_t* list_b;   // nodes for the wider stage
  uint32_t* list_c;   // nodes for the exact stage
  uint32_t* crowded;  // bricks for k_search_crowded
  vp_nn_stats_dev* stats;
  FieldOut f;
};
__device__ __forceinline__ void settle(const SearchOut& o, const rec_t* __restrict__ part, int rs, size_t node, int pos) {
  if (o.nn) o.nn[node] = int(__float_as_uint(__ldg(&part[size_t(pos) * rs].w)) & ~kFarBit);
  if (o.nn_pos) o.nn_pos[node] = pos;
  if (o.f.on) write_fields(o.f, part, rs, node, pos);
}
__device__ __forceinline__ void emit(const SearchOut& o, const rec_t* __restrict__ part, int rs, size_t node, int verdict, int pos) {
  if (verdict == 0) {
    settle(o, part, rs, node, pos);
  } else if (verdict == 1) {
    o.list_c[atomicAdd(&o.stats->n_wide, 1ull)] = uint32_t(node);
  } else {
    o.list_b[atomicAdd(&o.stats->n_b, 1ull)] = uint32_t(node);
  }
}

// true: with particles outside the cell grid (clamped into end cells, their stored offsets saturated) every node whose
// examined block touches an end cell is decided by the exact stage, which reads the caller's coordinates
__device__ __forceinline__ bool touches_end(int c0, int c1, int g) { return c0 <= 0 || c1 >= g - 1; }

// Records of cells [za, zb] (zb - za <= 31) of one cell row.  Their z offsets are relative to the 32-cell block of their own
// cell, so the range is cut where it enters the next block: each piece is one contiguous run of records with a single
// query offset  rz + (wz - block base) * hz  (rz = node z relative to the low corner of cell wz).
__device__ __forceinline__ void scan_zcells(const rec_t* __restrict__ part, int rs, const uint32_t* __restrict__ start, size_t row,
                                            int za, int zb, int wz, float qx, float qy, float rz, float hz, Cand& c) {
  const int zs = min(zb, za | 31);
  const uint32_t s0 = __ldg(start + row + za), s1 = __ldg(start + row + zs + 1);
  scan_range(part, rs, s0, s1, qx, qy, rz + float(wz - (za & ~31)) * hz, c);
  if (zs < zb) scan_range(part, rs, s1, __ldg(start + row + zb + 1), qx, qy, rz + float(wz - ((zs + 1) & ~31)) * hz, c);
}

// Stage A (any lattice / any cell size): one thread per node (consecutive lanes = consecutive z nodes), the 2x2x2 window as
// 4 cell rows x 2 contiguous cells.  The z offsets of the records are relative to the 32-cell block of their own cell: when
// cell wz + 1 opens a new block (one lane in 32) a row is two runs with two query offsets -- those lanes take the second
// cell in a separate, rare tail.  Register budget: 32 (8 CTAs of 256 threads per SM) -- the loop is latency bound, and a
// version carrying the wider stage inline needed 64 registers and ran 2.2x slower.
template <int RS, int MINB>
__global__ void __launch_bounds__(256, MINB) k_search_rows(const rec_t* __restrict__ part, const uint32_t* __restrict__ start, Grid g,
                                                            Lattice L, float eps, SearchOut out) {
  // block = (z nodes, y rows); grid = (z chunks, y chunks, x).  32-bit cell and node arithmetic (both counts are < 2^32).
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (k >= L.nz || j >= L.ny) return;
  const uint32_t node = (uint32_t(blockIdx.z) * uint32_t(L.ny) + uint32_t(j)) * uint32_t(L.nz) + uint32_t(k);
  const int wx = __ldg(L.wx + blockIdx.z), wy = __ldg(L.wy + j), wz = __ldg(L.wz + k);
  if (out.stats->n_far && (touches_end(wx, wx + 1, g.gx) || touches_end(wy, wy + 1, g.gy) || touches_end(wz, wz + 1, g.gz))) {
    out.list_c[atomicAdd(&out.stats->n_wide, 1ull)] = node;
    return;
  }
  const bool cross = (wz & 31) == 31 && wz + 1 < g.gz;
  uint32_t rs0[4], re0[4];
  {
    const uint32_t zl = uint32_t(cross ? wz : min(wz + 1, g.gz - 1)) + 1u;   // one past the last cell of the first run
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int X = wx + (r >> 1), Y = wy + (r & 1);
      const bool in = X < g.gx && Y < g.gy;
      const uint32_t row = in ? (uint32_t(X) * uint32_t(g.gy) + uint32_t(Y)) * uint32_t(g.gz) : 0u;
      rs0[r] = in ? __ldg(start + row + uint32_t(wz)) : 0u;
      re0[r] = in ? __ldg(start + row + zl) : 0u;
    }
  }
  const float rx = __ldg(L.rx + blockIdx.z), ry = __ldg(L.ry + j), rz = __ldg(L.rz + k);
  Cand c;
  c.b1 = INFINITY; c.b2 = INFINITY; c.bi = -1;
  {
    const float qz0 = rz + float(wz & 31) * g.hzf;
#pragma unroll
    for (int r = 0; r < 4; ++r) scan_range(part, RS, rs0[r], re0[r], rx - float(r >> 1) * g.hxf, ry - float(r & 1) * g.hyf, qz0, c);
  }
  if (cross) {   // (one lane in 32) cell wz + 1 opens a new 32-cell block: its records have another z reference
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
      const int X = wx + (r >> 1), Y = wy + (r & 1);
      if (X >= g.gx || Y >= g.gy) continue;
      const uint32_t row = (uint32_t(X) * uint32_t(g.gy) + uint32_t(Y)) * uint32_t(g.gz) + uint32_t(wz);
      scan_range(part, RS, __ldg(start + row + 1), __ldg(start + row + 2), rx - float(r >> 1) * g.hxf, ry - float(r & 1) * g.hyf, rz - g.hzf, c);
    }
  }
  const float m = fminf(__ldg(L.mx + blockIdx.z), fminf(__ldg(L.my + j), __ldg(L.mz + k)));
  emit(out, part, RS, node, judge(c, eps, m), c.bi);
}

// Stage A, brick form (particle-centric): the lattice windows are consecutive cells along every axis (node i <-> cells
// [w0+i, w0+i+1]) and the z windows start on a multiple of 32.  One CTA = kBX x kBY x 32 nodes and the particles of the covering
// (kBX+1) x (kBY+1) cell rows x 33 cells -- 81 contiguous runs of the sorted array.  The roles are turned round: instead of every
// node walking its 8 cells (four Poisson-length loops per lane: a warp runs to its longest lane, ~3x the mean trip count, and
// the kernel is issue bound), every PARTICLE visits the 8 nodes whose window contains its cell -- one lane per particle,
// exactly 8 pairs each, no divergence, and the 8 squared distances cost 6 squares + 12 adds together.
//
// One sweep.  Per node one 32-bit word in shared memory:  key = (n << 13) | (slot << 1) | flag.  n is the squared distance in
// units of 1/S, formed in INTEGERS: each of the six per-axis squares (dx^2 * S, clamped to 2^17) is rounded once (adding 2^23
// leaves the integer in the low mantissa bits), and a corner's n is one three-input add -- |d^2 S - n| <= 1.6.  S puts the
// clamp just above the largest proof margin: a candidate with n >= 2^17 can never be proven, nor can it threaten a proven
// winner.  slot = the particle's number inside the brick.  A particle tests its key against the word and only then issues
// atomicMin (a shared-memory atomic costs ~1-2 cycles per lane on B200, a load + compare almost nothing; of the ~8
// candidates of a node only the running minima, H_8 = 2.7 on average, get through).  Ambiguity is settled in the
// same sweep.  The value a candidate c meets -- the word it read when it loses, the word the atomic returned when it goes
// through -- is the current holder h; with M = n_h / S, q = 1.6 / S (quantisation) and tol = the rigorous f32 error bound
// of judge():
//     d_c < M - 2 tol - 2q   c replaces h, and its key carries flag 0: it beats h and every rival of h by more than tol
//     d_c > M + 2 tol + 2q   c is clearly worse than the holder (whose key can only decrease from here): nothing to do
//     otherwise              the flag bit of the word is set (atomicOr) -- the node's holder has a rival within the tolerance
// Invariant: every candidate that has arrived is no better than the holder (up to q), and if the flag is clear the holder
// beats each of them by more than tol; so an unflagged final holder is the true nearest particle whatever the order of
// arrival -- the guarantee of judge()'s runner-up test (its flags are a subset of these).  The cell along z comes from the
// coordinate itself: a particle within a rounding error (4e-6 h) of a cell face may be booked one cell off, but it is then
// within that distance of the window face of every node it is wrongly offered to (or withheld from), and such a node cannot
// pass the proof (tol >= 1.6e-4 * margin * h).
// Pass 3, one thread per node: proven iff n <= the node's integer threshold (the largest n with b + tol(b) < margin^2 for
// b = (n + 0.6)/S, tabulated per axis on the host; min over the axes) -> settle (index + field planes) or hand on (lists B / C).
// The node arrays are padded by one node on every side, so the 8 updates of a particle are 8 immediate offsets from one base
// and need no bounds tests; the padding nodes are never read.  A brick with more than 4095 particles (1.5 per cell) goes to the
// node-centric kernel k_search_crowded.
constexpr int kBX = 8, kBY = 8, kBZ = 32;
constexpr int kBRows = (kBX + 1) * (kBY + 1);
constexpr int kBThreads = 256;
constexpr int kNX = kBX + 2, kNY = kBY + 2, kNZ = kBZ + 2;   // padded node box
constexpr int kNodes = kNX * kNY * kNZ;
constexpr uint32_t kSlotBits = 12, kSlotMax = (1u << kSlotBits) - 1u;      // particles per sweep
constexpr uint32_t kQ = 1u << 17;                                            // clamp of one quantised axis term; n >= kQ: beyond the clamp
struct BrickQ {
  float S, invS;       // quantisation scale of d^2 and its inverse
  float q2;            // 2 * 1.6 / S
  uint32_t nt;         // coarse filter: |n_c - n_h| <= nt  covers  2 tol + 2q  for every distance below the clamp
  const int *tx, *ty, *tz;   // per-axis proof thresholds on n (device)
};
constexpr size_t kBrickOffRow = size_t(kNodes) * 4;
constexpr size_t kBrickOffSeg = kBrickOffRow + size_t(kSlotMax + 1);
constexpr size_t kBrickOffNq = kBrickOffSeg + (3 * size_t(kBRows) + 4) * 4;
constexpr size_t kBrickSmem = kBrickOffNq + size_t(kNX + kNY + kNZ + 2) * 4 + size_t(kBX) * 4;

__host__ __device__ constexpr int brick_off(int q) { return (q >> 2) * (kNY * kNZ) + ((q >> 1) & 1) * kNZ + (q & 1); }
// h = word at [base + OFF words] (shared-memory address); if key < h the atomicMin is issued and h becomes what it returned.
// Predicated, so that the common path carries no branch bookkeeping (BSSY / BRA / BSYNC around an `if`).
template <int OFF>
__device__ __forceinline__ uint32_t brick_try_min(uint32_t base, uint32_t key) {
  uint32_t h;
  asm volatile(
      "{\n .reg .pred p;\n ld.shared.u32 %0, [%1+%3];\n setp.lt.u32 p, %2, %0;\n @p atom.shared.min.u32 %0, [%1+%3], %2;\n}"
      : "=&r"(h)
      : "r"(base), "r"(key), "n"(OFF * 4)
      : "memory");
  return h;
}

__global__ void __launch_bounds__(kBThreads, 5) k_search_brick(const rec_t* __restrict__ part, int rs, const uint32_t* __restrict__ start,
                                                                Grid g, Lattice L, float eps, BrickQ bq, SearchOut out) {
  extern __shared__ __align__(16) unsigned char brick_smem[];
  uint32_t* nkey = reinterpret_cast<uint32_t*>(brick_smem);                                 // [kNodes] (n << 13) | (slot << 1) | flag
  unsigned char* slotrow = brick_smem + kBrickOffRow;                                       // [4096] cell row of a slot
  uint32_t* seg_s = reinterpret_cast<uint32_t*>(brick_smem + kBrickOffSeg);                 // first sorted position of a row's run
  uint32_t* seg_last = seg_s + kBRows;                                                      // first position of its 33rd cell
  uint32_t* seg_off = seg_last + kBRows;                                                    // [kBRows + 1] slot of its first particle
  float* nqx = reinterpret_cast<float*>(brick_smem + kBrickOffNq);                          // brick-relative node coordinates, padded
  float* nqy = nqx + kNX;
  float* nqz = nqy + kNY;
  int* ntx = reinterpret_cast<int*>(nqz + kNZ + 2);                                         // [kBX] proof thresholds along x
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int k0 = blockIdx.x * kBZ, j0 = blockIdx.y * kBY, i0 = blockIdx.z * kBX;
  const int X0 = __ldg(L.wx + i0), Y0 = __ldg(L.wy + j0), Z0 = __ldg(L.wz + k0);   // first cell of the brick
  const int Zend = min(Z0 + kBZ, g.gz - 1);                                         // last cell of the runs along z
  const float hx = g.hxf, hy = g.hyf, hz = g.hzf;

  if (tid < kNX + kNY + kNZ) {
    float v = 0.f;
    if (tid < kNX) { const int i = tid - 1; if (i >= 0 && i < kBX && i0 + i < L.nx) v = __ldg(L.rx + i0 + i) + float(i) * hx; }
    else if (tid < kNX + kNY) { const int j = tid - kNX - 1; if (j >= 0 && j < kBY && j0 + j < L.ny) v = __ldg(L.ry + j0 + j) + float(j) * hy; }
    else { const int k = tid - kNX - kNY - 1; if (k >= 0 && k < kBZ && k0 + k < L.nz) v = __ldg(L.rz + k0 + k) + float(k) * hz; }
    nqx[tid] = v;
  } else if (tid < kNX + kNY + kNZ + kBX) {
    const int i = tid - (kNX + kNY + kNZ);
    ntx[i] = i0 + i < L.nx ? __ldg(bq.tx + i0 + i) : -1;
  }
  if (tid < kBRows) {
    const int X = X0 + tid / (kBY + 1), Y = Y0 + tid % (kBY + 1);
    uint32_t s = 0, e = 0, l = 0xffffffffu;
    if (X < g.gx && Y < g.gy) {
      const size_t row = (size_t(X) * g.gy + Y) * g.gz;
      s = __ldg(start + row + Z0);
      e = __ldg(start + row + Zend + 1);
      if (Z0 + kBZ <= Zend) l = __ldg(start + row + Z0 + kBZ);   // first record of the 33rd cell (next 32-cell z block)
    }
    seg_s[tid] = s;
    seg_last[tid] = l;
    seg_off[tid] = e - s;
  }
  __syncthreads();
  if (w == 0) {   // exclusive prefix of the kBRows run lengths
    uint32_t carry = 0;
    for (int base = 0; base < kBRows; base += 32) {
      const int r = base + lane;
      const uint32_t v = r < kBRows ? seg_off[r] : 0u;
      uint32_t incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (r < kBRows) seg_off[r] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) seg_off[kBRows] = carry;
  }
  __syncthreads();
  // (the sweep below can take a brick in slices of x planes; measured on clustered input the dense bricks are better off in the
  // node-centric kernel, so a brick either fits one sweep or is handed over)
  const int planes = seg_off[kBRows] <= kSlotMax ? kBX : 0;
  const bool far = out.stats->n_far != 0;
  const int j = j0 + w, k = k0 + lane;     // pass 3: this thread's node column
  if (planes == 0) {
    // crowded brick: handed to the node-centric kernel, k_search_crowded
    if (tid == 0) out.crowded[atomicAdd(&out.stats->n_crowded, 1ull)] = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    return;
  }
  const int nthr_yz = (j < L.ny && k < L.nz) ? min(__ldg(bq.ty + j), __ldg(bq.tz + k)) : -1;
  const float ihz = float(g.ihz);
  const uint32_t nt = bq.nt;
  for (int xa = 0; xa < kBX; xa += planes) {
    const int r0 = xa * (kBY + 1), nrows = (planes + 1) * (kBY + 1);
    const uint32_t sb = seg_off[r0], count = seg_off[r0 + nrows] - sb;
    // empty node = "a holder exactly at the clamp": candidates beyond the clamp (44 % of all pairs: the corners of the 2x2x2
    // window outside the proof sphere) then fail the key test and issue no atomic -- the sweep is bound by the shared-memory
    // atomic unit (~1.3 cycles per lane), not by instruction issue
    for (int n = tid; n < kNodes; n += kBThreads) nkey[n] = kQ << (kSlotBits + 1);
    for (int r = r0 + w; r < r0 + nrows; r += kBThreads / 32) {       // the cell row of every slot
      const uint32_t a = seg_off[r] - sb, e = seg_off[r + 1] - sb;
      for (uint32_t p = a + lane; p < e; p += 32) slotrow[p] = (unsigned char)r;
    }
    __syncthreads();
    // ---- the sweep: one lane per particle
    // (software pipelined: the record of the lane's next particle is requested before this one is worked on)
    auto locate = [&](uint32_t t, int& r, uint32_t& gpos) {
      r = slotrow[t];
      gpos = seg_s[r] + (sb + t - seg_off[r]);
    };
    int r_n = 0;
    uint32_t gpos_n = 0;
    float4 q_n = make_float4(0.f, 0.f, 0.f, 0.f);
    if (uint32_t(tid) < count) { locate(tid, r_n, gpos_n); q_n = __ldg(part + size_t(gpos_n) * rs); }
    for (uint32_t t = tid; t < count; t += kBThreads) {
      const int r = r_n;
      const uint32_t gpos = gpos_n;
      const float4 q = q_n;
      if (t + kBThreads < count) { locate(t + kBThreads, r_n, gpos_n); q_n = __ldg(part + size_t(gpos_n) * rs); }
      const uint32_t cx = uint32_t(r) / (kBY + 1), cy = uint32_t(r) - cx * (kBY + 1);
      const float pz = (gpos >= seg_last[r]) ? q.z + float(kBZ) * hz : q.z;
      int cz = int(pz * ihz);
      cz = cz < 0 ? 0 : (cz > kBZ ? kBZ : cz);
      const float px = q.x + float(cx) * hx, py = q.y + float(cy) * hy;
      // the six per-axis squares in units of 1/S as integers: bits(min(d*d*S, 2^17) + 2^23) = 0x4B000000 + round(...)
      auto quant = [&](float dd) { return __float_as_uint(fminf(dd * dd * bq.S, float(kQ)) + 8388608.f); };
      const uint32_t ux0 = quant(nqx[cx] - px), ux1 = quant(nqx[cx + 1] - px), uy0 = quant(nqy[cy] - py), uy1 = quant(nqy[cy + 1] - py);
      // (the three biases add up to 0xE1000000: taken off the z terms, so that a corner's n is one three-input add)
      const uint32_t uz0 = quant(nqz[cz] - pz) - 0xE1000000u, uz1 = quant(nqz[cz + 1] - pz) - 0xE1000000u;
      const uint32_t t2 = t << 1;
      uint32_t* m = nkey + (cx * kNY + cy) * kNZ + uint32_t(cz);
      const uint32_t msh = uint32_t(__cvta_generic_to_shared(m));
      // corner qd = 4a + 2b + c (x, y, z offsets): key, predicated atomicMin, rare ambiguity path
#define VP_BRICK_CORNER(QD)                                                                                                \
      {                                                                                                                     \
        const uint32_t n = ((QD & 4) ? ux1 : ux0) + ((QD & 2) ? uy1 : uy0) + ((QD & 1) ? uz1 : uz0);                        \
        const uint32_t key = (n << (kSlotBits + 1)) + t2;                                                                   \
        const uint32_t h = brick_try_min<brick_off(QD)>(msh, key);                                                          \
        const uint32_t nh = h >> (kSlotBits + 1);                                                                           \
        if (n - nh + nt <= 2u * nt && n < kQ - nt) { /* rare: neither clearly better nor clearly worse */                   \
          const float dc = float(n) * bq.invS, M = float(nh) * bq.invS, dm = fmaxf(dc, M);                                  \
          if (fabsf(dc - M) <= 2.f * (8.f * sqrtf(dm) * eps + 8.f * eps * eps + 1e-6f * dm) + bq.q2)                        \
            atomicOr(m + brick_off(QD), 1u);                                                                                \
        }                                                                                                                   \
      }
      VP_BRICK_CORNER(0) VP_BRICK_CORNER(1) VP_BRICK_CORNER(2) VP_BRICK_CORNER(3)
      VP_BRICK_CORNER(4) VP_BRICK_CORNER(5) VP_BRICK_CORNER(6) VP_BRICK_CORNER(7)
#undef VP_BRICK_CORNER
    }
    __syncthreads();
    // ---- pass 3: one thread per node (k = lane: plane stores are coalesced).  (Four planes at a time, with the four payload
    // records requested together, changed nothing: 22.4 vs 22.0 ms.)
    if (j < L.ny && k < L.nz) {
      const int wy = Y0 + w, wz = Z0 + lane;
      for (int ii = xa; ii < xa + planes; ++ii) {
        const int gi = i0 + ii;
        if (gi >= L.nx) break;
        const size_t node = (size_t(gi) * L.ny + j) * L.nz + k;
        const int wx = X0 + ii;
        if (far && (touches_end(wx, wx + 1, g.gx) || touches_end(wy, wy + 1, g.gy) || touches_end(wz, wz + 1, g.gz))) {
          out.list_c[atomicAdd(&out.stats->n_wide, 1ull)] = uint32_t(node);
          continue;
        }
        const int idx = ((ii + 1) * kNY + (w + 1)) * kNZ + (lane + 1);
        const uint32_t key = nkey[idx];
        int verdict = 2, pos = -1;
        if (int(key >> (kSlotBits + 1)) <= min(ntx[ii], nthr_yz)) {    // proven (an empty node holds n = 2^17 > every threshold)
          verdict = int(key & 1u);
          const uint32_t t = (key >> 1) & kSlotMax, r = slotrow[t];
          pos = int(seg_s[r] + (sb + t - seg_off[r]));
        }
        emit(out, part, rs, node, verdict, pos);
      }
    }
    __syncthreads();
  }
}

// Stage A for the bricks the sweep declined (clustered input: a few dense bricks hold most of the particles): node-centric, a
// warp = 32 consecutive z nodes, the 2x2x2 window as 4 cell rows, straight from global memory.  Persistent CTAs take
// (brick, x plane) items from a global cursor.
__global__ void __launch_bounds__(256, 6) k_search_crowded(const rec_t* __restrict__ part, int rs, const uint32_t* __restrict__ start, Grid g,
                                                            Lattice L, float eps, SearchOut out, unsigned gx_b, unsigned gy_b) {
  __shared__ unsigned long long item_s;
  const unsigned long long nitems = out.stats->n_crowded * kBX;       // work item = one x plane of a brick (256 nodes)
  const bool far = out.stats->n_far != 0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float hx = g.hxf, hy = g.hyf, hz = g.hzf;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) item_s = atomicAdd(&out.stats->crowded_cursor, 1ull);
    __syncthreads();
    const unsigned long long item = item_s;
    if (item >= nitems) break;
    const unsigned id = out.crowded[item / kBX];
    const int ii = int(item % kBX);
    const int k0 = int(id % gx_b) * kBZ, j0 = int((id / gx_b) % gy_b) * kBY, i0 = int(id / (gx_b * gy_b)) * kBX;
    const int j = j0 + w, k = k0 + lane, gi = i0 + ii;
    if (j >= L.ny || k >= L.nz || gi >= L.nx) continue;
    const int wx = __ldg(L.wx + gi), wy = __ldg(L.wy + j), wz = __ldg(L.wz + k);
    const size_t node = (size_t(gi) * L.ny + j) * L.nz + k;
    if (far && (touches_end(wx, wx + 1, g.gx) || touches_end(wy, wy + 1, g.gy) || touches_end(wz, wz + 1, g.gz))) {
      out.list_c[atomicAdd(&out.stats->n_wide, 1ull)] = uint32_t(node);
      continue;
    }
    const float rx = __ldg(L.rx + gi), ry = __ldg(L.ry + j), rz = __ldg(L.rz + k);
    Cand c;
    c.b1 = INFINITY; c.b2 = INFINITY; c.bi = -1;
    for (int a = 0; a < 2; ++a)
      for (int bb = 0; bb < 2; ++bb) {
        const int X = wx + a, Y = wy + bb;
        if (X >= g.gx || Y >= g.gy) continue;
        scan_zcells(part, rs, start, (size_t(X) * g.gy + Y) * g.gz, wz, min(wz + 1, g.gz - 1), wz, rx - float(a) * hx,
                    ry - float(bb) * hy, rz, hz, c);
      }
    emit(out, part, rs, node, judge(c, eps, fminf(__ldg(L.mx + gi), fminf(__ldg(L.my + j), __ldg(L.mz + k)))), c.bi);
  }
}

// Stage B: the nodes stage A could not prove, one thread per listed node, the 32-cell union of the three 4x2x2 bars
// through the window (proves sqrt(2) h with corner-aligned nodes), same f32 prefilter and verdict.  What is still
// unproven or ambiguous goes to the exact kernel.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_search_block4(const rec_t* __restrict__ part, int rs, const uint32_t* __restrict__ start, Grid g,
                                                        Lattice L, float eps, SearchOut out) {
  const unsigned long long nb = out.stats->n_b;
  const bool far = out.stats->n_far != 0;
  const float hx = float(g.hx), hy = float(g.hy), hz = float(g.hz);
  for (unsigned long long t = (unsigned long long)(blockIdx.x) * blockDim.x + threadIdx.x; t < nb;
       t += (unsigned long long)(gridDim.x) * blockDim.x) {
    const uint32_t node = out.list_b[t];
    const int k = int(node % uint32_t(L.nz));
    const uint32_t u = node / uint32_t(L.nz);
    const int j = int(u % uint32_t(L.ny)), i = int(u / uint32_t(L.ny));
    // examined region = union of the three 4x2x2 bars through the 2x2x2 window (32 cells instead of 64):
    //   central rows (x, y both inside the window): 4 cells along z;  rows one step outside in x OR y: 2 cells
    const int wx = L.wx[i], wy = L.wy[j], wz = L.wz[k];
    const int wx1 = min(wx + 1, g.gx - 1), wy1 = min(wy + 1, g.gy - 1), wz1 = min(wz + 1, g.gz - 1);
    const int x0 = max(wx - 1, 0), x1 = min(wx + 2, g.gx - 1);
    const int y0 = max(wy - 1, 0), y1 = min(wy + 2, g.gy - 1);
    const int z0 = max(wz - 1, 0), z1 = min(wz + 2, g.gz - 1);
    if (far && (touches_end(x0, x1, g.gx) || touches_end(y0, y1, g.gy) || touches_end(z0, z1, g.gz))) {
      out.list_c[atomicAdd(&out.stats->n_wide, 1ull)] = node;
      continue;
    }
    const float rx = L.rx[i], ry = L.ry[j], rz = L.rz[k];
    Cand c;
    c.b1 = INFINITY; c.b2 = INFINITY; c.bi = -1;
    for (int X = x0; X <= x1; ++X)
      for (int Y = y0; Y <= y1; ++Y) {
        const bool xin = X >= wx && X <= wx1, yin = Y >= wy && Y <= wy1;
        if (!xin && !yin) continue;                                  // corner rows are not part of the union
        const size_t row = (size_t(X) * g.gy + Y) * g.gz;
        const int a = (xin && yin) ? z0 : wz, b = (xin && yin) ? z1 : wz1;
        scan_zcells(part, rs, start, row, a, b, wz, rx - float(X - wx) * hx, ry - float(Y - wy) * hy, rz, hz, c);
      }
    // nearest unexamined point: beyond a 4-cell bar end along one axis, or outside the 2-cell window along two axes
    const double qxd = L.qx[i], qyd = L.qy[j], qzd = L.qz[k];
    const double m4 = fmin(axis_margin(qxd, g.ox, g.hx, x0, x1, g.gx, g.closed_xlo, g.closed_xhi),
                           fmin(axis_margin(qyd, g.oy, g.hy, y0, y1, g.gy, false, false),
                                axis_margin(qzd, g.oz, g.hz, z0, z1, g.gz, false, false)));
    double a2 = axis_margin(qxd, g.ox, g.hx, wx, wx1, g.gx, g.closed_xlo, g.closed_xhi);
    double b2 = axis_margin(qyd, g.oy, g.hy, wy, wy1, g.gy, false, false);
    double c2 = axis_margin(qzd, g.oz, g.hz, wz, wz1, g.gz, false, false);
    // two smallest of (a2, b2, c2)
    double lo1 = fmin(a2, fmin(b2, c2));
    double lo2 = (lo1 == a2) ? fmin(b2, c2) : ((lo1 == b2) ? fmin(a2, c2) : fmin(a2, b2));
    const double diag = (lo2 == INFINITY) ? INFINITY : sqrt(lo1 * lo1 + lo2 * lo2);
    const double md = fmin(m4, diag);
    float m = INFINITY;
    if (md != INFINITY) m = md > 0.0 ? __double2float_rd(md * (1.0 - 1.0 / 1048576.0)) : 0.f;
    const int verdict = judge(c, eps, m);
    if (verdict == 0) {
      settle(out, part, rs, node, c.bi);
    } else {
      out.list_c[atomicAdd(&out.stats->n_wide, 1ull)] = node;
    }
  }
}

// Exact search, one warp per listed node, f64 arithmetic on the caller's coordinates.  The searched block starts
// as the node's 2x2x2 window and is widened (1, 3, 7, ... cells per side) until the proof holds (or every kept
// particle has been examined).
template <typename T>
__global__ void __launch_bounds__(256) k_search_exact(const rec_t* __restrict__ part, int rs, const uint32_t* __restrict__ start,
                                                       const T* __restrict__ pos, Grid g, Lattice L, SearchOut out) {
  int32_t* __restrict__ nn = out.nn;
  int32_t* __restrict__ nn_pos = out.nn_pos;
  const uint32_t* __restrict__ list = out.list_c;
  vp_nn_stats_dev* __restrict__ stats = out.stats;
  const unsigned long long nw = stats->n_wide;
  const int lane = threadIdx.x & 31;
  const unsigned long long warp0 = (unsigned long long)(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const unsigned long long nwarps = (unsigned long long)(gridDim.x) * (blockDim.x >> 5);
  for (unsigned long long w = warp0; w < nw; w += nwarps) {
    const int64_t node = list[w];
    int k = int(node % L.nz);
    int64_t t = node / L.nz;
    int j = int(t % L.ny), i = int(t / L.ny);
    const double qx = L.qx[i], qy = L.qy[j], qz = L.qz[k];
    const int wx = L.wx[i], wy = L.wy[j], wz = L.wz[k];
    Best b;
    int bpos = -1;
    bool done = false;
    // block = the 2-cell window widened by r cells on every side, r = 0, 1, 3, 7, ...
    for (int r = 0; !done; r = 2 * r + 1) {
      const int x0 = max(wx - r, 0), x1 = min(wx + 1 + r, g.gx - 1);
      const int y0 = max(wy - r, 0), y1 = min(wy + 1 + r, g.gy - 1);
      const int z0 = max(wz - r, 0), z1 = min(wz + 1 + r, g.gz - 1);
      b.d2 = INFINITY;
      b.idx = 0x7fffffff;
      bpos = -1;
      const int nyb = y1 - y0 + 1;
      const int nrows = (x1 - x0 + 1) * nyb;
      if (nrows <= 4) {
        // the 2x2x2 window (almost every listed node ends here): its ~8 particles ONE PER LANE, so that the dependent
        // chain record -> index -> caller's coordinates is walked once instead of once per particle of a row
        const int k4 = lane & 3;
        uint32_t s = 0, e = 0;
        if (k4 < nrows) {
          const size_t row = (size_t(x0 + k4 / nyb) * g.gy + (y0 + k4 % nyb)) * g.gz;
          s = __ldg(start + row + z0);
          e = __ldg(start + row + z1 + 1);
        }
        const uint32_t s0 = __shfl_sync(0xffffffffu, s, 0), s1 = __shfl_sync(0xffffffffu, s, 1);
        const uint32_t s2 = __shfl_sync(0xffffffffu, s, 2), s3 = __shfl_sync(0xffffffffu, s, 3);
        const uint32_t c0 = __shfl_sync(0xffffffffu, e - s, 0), c1 = c0 + __shfl_sync(0xffffffffu, e - s, 1);
        const uint32_t c2 = c1 + __shfl_sync(0xffffffffu, e - s, 2), c3 = c2 + __shfl_sync(0xffffffffu, e - s, 3);
        for (uint32_t t = lane; t < c3; t += 32) {
          const uint32_t p = t < c0 ? s0 + t : (t < c1 ? s1 + (t - c0) : (t < c2 ? s2 + (t - c1) : s3 + (t - c2)));
          const int id = int(__float_as_uint(__ldg(&part[size_t(p) * rs].w)) & ~kFarBit);
          const int before = b.idx;
          const size_t pb = size_t(g.ps) * size_t(id);
          consider(b, qx, qy, qz, double(pos[pb]), double(pos[pb + 1]), double(pos[pb + 2]), id);
          if (b.idx != before) bpos = int(p);
        }
      } else {
        for (int rr = lane; rr < nrows; rr += 32) {
          int X = x0 + rr / nyb, Y = y0 + rr % nyb;
          size_t row = (size_t(X) * g.gy + Y) * g.gz;
          uint32_t s = __ldg(start + row + z0), e = __ldg(start + row + z1 + 1);
          for (uint32_t p = s; p < e; ++p) {
            const int id = int(__float_as_uint(__ldg(&part[size_t(p) * rs].w)) & ~kFarBit);
            const int before = b.idx;
            const size_t pb = size_t(g.ps) * size_t(id);
            consider(b, qx, qy, qz, double(pos[pb]), double(pos[pb + 1]), double(pos[pb + 2]), id);
            if (b.idx != before) bpos = int(p);
          }
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        double od = __shfl_xor_sync(0xffffffffu, b.d2, o);
        int oi = __shfl_xor_sync(0xffffffffu, b.idx, o);
        int op = __shfl_xor_sync(0xffffffffu, bpos, o);
        if (od < b.d2 || (od == b.d2 && oi < b.idx)) { b.d2 = od; b.idx = oi; bpos = op; }
      }
      double m = fmin(axis_margin(qx, g.ox, g.hx, x0, x1, g.gx, g.closed_xlo, g.closed_xhi),
                      fmin(axis_margin(qy, g.oy, g.hy, y0, y1, g.gy, false, false),
                           axis_margin(qz, g.oz, g.hz, z0, z1, g.gz, false, false)));
      if (proven(b, m)) {
        done = true;
      } else if (x0 == 0 && x1 == g.gx - 1 && y0 == 0 && y1 == g.gy - 1 && z0 == 0 && z1 == g.gz - 1) {
        // every kept particle was examined and a closed x face is still nearer than the best one
        if (lane == 0) atomicAdd(&stats->n_unresolved, 1ull);
        done = true;
      }
    }
    if (lane == 0) {
      if (nn) nn[node] = (b.idx == 0x7fffffff) ? -1 : b.idx;
      if (nn_pos) nn_pos[node] = bpos;
      if (out.f.on && bpos >= 0) write_fields(out.f, part, rs, size_t(node), bpos);
    }
  }
}

// fields from the SORTED records: node -> sorted position of its nearest particle -> (v', m) (second half of the record)
// (stride, offset in float4 units: 2, 1 for the sorted 32-byte records; 1, 0 for a plain [np] float4 payload array)
__global__ void __launch_bounds__(256) k_fields_sorted(const int32_t* __restrict__ nn_pos, int64_t n, const float4* __restrict__ srec,
                                                        int stride, int offset, float* vx, float* vy, float* vz, float* px, float* py,
                                                        float* pz, float* e, float* mo) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float4 w = __ldg(srec + size_t(stride) * size_t(nn_pos[t]) + offset);
  if (vx) vx[t] = w.x;
  if (vy) vy[t] = w.y;
  if (vz) vz[t] = w.z;
  if (px) px[t] = w.x * w.w;
  if (py) py[t] = w.y * w.w;
  if (pz) pz[t] = w.z * w.w;
  if (e) e[t] = w.w * (w.x * w.x + w.y * w.y + w.z * w.z);   // interp.py:546 (no 1/2)
  if (mo) mo[t] = w.w;
}

__global__ void __launch_bounds__(256) k_gather_words(const int32_t* __restrict__ idx, int64_t n, const uint32_t* __restrict__ src,
                                                       int rw, uint32_t* __restrict__ dst) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * rw) return;
  int64_t i = t / rw;
  int w = int(t - i * rw);
  dst[t] = src[size_t(idx[i]) * rw + w];
}

template <typename T>
__global__ void __launch_bounds__(256) k_build_fields(const int32_t* __restrict__ nn, int64_t n, const T* __restrict__ vel,
                                                       const T* __restrict__ rho, T lcell3, float* vx, float* vy, float* vz,
                                                       float* px, float* py, float* pz, float* e, float* mo) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n) return;
  size_t i = size_t(nn[t]);
  T a = vel[3 * i], b = vel[3 * i + 1], c = vel[3 * i + 2];
  T m = lcell3;
  if (rho) {
    // reference: payload [rho*v, rho] is gathered, then v = (rho*v)/rho, m = rho*Lcell^3 (interp.py:199-213,272-273)
    T r = rho[i];
    a = (a * r) / r;
    b = (b * r) / r;
    c = (c * r) / r;
    m = r * lcell3;
  }
  if (vx) vx[t] = float(a);
  if (vy) vy[t] = float(b);
  if (vz) vz[t] = float(c);
  if (px) px[t] = float(a * m);
  if (py) py[t] = float(b * m);
  if (pz) pz[t] = float(c * m);
  if (e) e[t] = float(m * (a * a + b * b + c * c));   // interp.py:546 (no 1/2)
  if (mo) mo[t] = float(m);
}

// ------------------------------------------------------------------------------------------ host side
struct AxisPlan {
  double o, h;
  int g;
};

// cell lattice for one axis: `g` cells of size h covering the node range widened by half a node
// spacing on each side (so that for g == n uniform nodes every node sits at a cell centre)
// `corner`: g must be n+1; the cells have the node spacing and every node sits on a cell corner
AxisPlan plan_axis(const double* q, int n, int g, double lo_ext, double hi_ext, bool use_ext, bool corner = false) {
  double qmin = q[0], qmax = q[0];
  for (int i = 1; i < n; ++i) { qmin = fmin(qmin, q[i]); qmax = fmax(qmax, q[i]); }
  double sp = n > 1 ? (qmax - qmin) / (n - 1) : 1.0;
  if (!(sp > 0)) sp = 1.0;
  double lo = qmin - 0.5 * sp, hi = qmax + 0.5 * sp;
  if (corner) { lo = qmin - sp; hi = qmax + sp; }
  if (use_ext) { lo = lo_ext; hi = hi_ext; }
  AxisPlan a;
  a.g = g;
  a.o = lo;
  a.h = (hi - lo) / g;
  return a;
}

static_assert(kWcBuckets == 2048, "k_bin_scatter_wc lays its tables out for kMaxBuckets buckets");
constexpr uint32_t kMaxBuckets = 2048;    // 8 KB of shared-memory counters / cursors per CTA of the bucket pass

Grid plan_grid(int64_t np, const double* qx, int nx, const double* qy, int ny, const double* qz, int nz,
               const vp_nn_opts& o) {
  int gx = o.cells_x, gy = o.cells_y, gz = o.cells_z;
  bool corner_x = false, corner_y = false, corner_z = false;
  if (gx <= 0 || gy <= 0 || gz <= 0) {
    // about one particle per cell inside the lattice volume; when the particle count is comparable to the node
    // count the cells take the node spacing and are shifted so that every node sits on a cell CORNER: the 2x2x2
    // block around a node then proves radius h with 8 candidate cells (a centred 3x3x3 block needs 27 for 1.5 h)
    double per_axis = cbrt(double(np > 0 ? np : 1) / (double(nx) * ny * nz));  // cells per node along an axis
    auto pick = [&](int n, bool& corner) {
      double g = n * per_axis;
      corner = (g > 0.7 * n && g < 1.5 * n && n > 1);
      if (corner) return n + 1;
      int gi = int(g + 0.5);
      return gi < 1 ? 1 : gi;
    };
    gx = pick(nx, corner_x); gy = pick(ny, corner_y); gz = pick(nz, corner_z);
    if (o.use_x_keep) {
      // the x extent is the kept range (closed sides) / the lattice extent (open sides); same cell size as along y
      AxisPlan ay = plan_axis(qy, ny, gy, 0, 0, false, corner_y);
      AxisPlan a0 = plan_axis(qx, nx, nx + 1, 0, 0, false, nx > 1);
      double lo = o.x_lo_is_domain_edge ? a0.o : o.x_keep_lo;
      double hi = o.x_hi_is_domain_edge ? a0.o + a0.h * (nx + 1) : o.x_keep_hi;
      int g = int((hi - lo) / ay.h + 0.5);
      gx = g < 1 ? 1 : g;
      corner_x = false;
    }
  }
  // the linear cell index must fit 32 bits: coarsen until it does
  while (double(gx) * gy * gz >= 4294967295.0) {
    gx = (gx + 1) / 2; gy = (gy + 1) / 2; gz = (gz + 1) / 2;
    corner_x = corner_y = corner_z = false;
  }
  AxisPlan ax = plan_axis(qx, nx, gx, 0, 0, false, corner_x);
  if (o.use_x_keep) {
    AxisPlan a0 = plan_axis(qx, nx, nx + 1, 0, 0, false, nx > 1);
    double lo = o.x_lo_is_domain_edge ? a0.o : o.x_keep_lo;
    double hi = o.x_hi_is_domain_edge ? a0.o + a0.h * (nx + 1) : o.x_keep_hi;
    ax = plan_axis(qx, nx, gx, lo, hi, true);
  }
  AxisPlan ay = plan_axis(qy, ny, gy, 0, 0, false, corner_y);
  AxisPlan az = plan_axis(qz, nz, gz, 0, 0, false, corner_z);
  Grid g;
  g.ox = ax.o; g.oy = ay.o; g.oz = az.o;
  g.hx = ax.h; g.hy = ay.h; g.hz = az.h;
  g.ihx = 1.0 / ax.h; g.ihy = 1.0 / ay.h; g.ihz = 1.0 / az.h;
  g.gx = gx; g.gy = gy; g.gz = gz;
  g.use_keep = o.use_x_keep;
  g.keep_lo = o.x_keep_lo; g.keep_hi = o.x_keep_hi;
  g.closed_xlo = o.use_x_keep && !o.x_lo_is_domain_edge;
  g.closed_xhi = o.use_x_keep && !o.x_hi_is_domain_edge;
  g.ps = o.row_stride > 0 ? o.row_stride : 3;
  g.vs = o.row_stride > 0 ? o.row_stride : 3;
  g.rs = o.row_stride > 0 ? o.row_stride : 1;
  // buckets of 2^bshift consecutive cells holding ~2^20 particles on average (a 32 MB window of records: the counting
  // sort inside it still runs out of L2, profiles/r2_ubench_scatter_gather.jsonl), at most kMaxBuckets of them
  const uint64_t ncells = uint64_t(gx) * gy * gz;
  int blog = 20;
  if (const char* ev = getenv("VP_BUCKET_LOG2")) blog = atoi(ev);
  const double cells_per_bucket = double(ncells) * double(uint64_t(1) << blog) / double(np > 0 ? np : 1);
  int bshift = 0;
  while (bshift < 31 && double(uint64_t(1) << bshift) * 1.4142135623730951 < cells_per_bucket) ++bshift;   // nearest power of two
  if (const char* ev = getenv("VP_BUCKET_SHIFT")) bshift = atoi(ev);
  if (bshift < 0) bshift = 0;
  if (bshift > 31) bshift = 31;
  while (bshift < 31 && ((ncells + (uint64_t(1) << bshift) - 1) >> bshift) > kMaxBuckets) ++bshift;
  g.hxf = float(g.hx); g.hyf = float(g.hy); g.hzf = float(g.hz);
  g.bshift = bshift;
  g.nb = uint32_t((ncells + (uint64_t(1) << bshift) - 1) >> bshift);
  if (g.nb < 1) g.nb = 1;
  return g;
}

size_t vp_scan_scratch_bytes_local(int64_t m) { return vp_align256(size_t((m + 4095) / 4096 + 1) * 4); }

struct NNScratch {
  size_t rec1, spos, tab, hist, sums, tail, total;
};
NNScratch nn_scratch(int64_t np, bool pay, uint64_t ncells, uint32_t nb, int64_t nnodes, bool own_srec = true) {
  NNScratch s;
  s.rec1 = vp_align256(size_t(np) * (pay ? sizeof(Rec32) : sizeof(RecA)));
  s.spos = own_srec ? vp_align256(size_t(np) * (pay ? 2 : 1) * sizeof(rec_t) + 256) : 256;   // (with a payload the caller may supply this array)
  s.tab = vp_align256((ncells + 8) * 4);
  s.hist = 2 * vp_align256((size_t(nb) * kSub + 1) * 4);          // (bucket, sub-stream) histogram and cursors
  s.sums = vp_scan_scratch_bytes_local(int64_t(ncells) + 1);
  const size_t b_lists = 2 * vp_align256(size_t(nnodes) * 4) + vp_align256((size_t(nnodes) / 8 + 4096) * 4);   // + the crowded-brick list (>= the number of bricks of any lattice shape)
  s.tail = s.rec1 > b_lists ? s.rec1 : b_lists;   // the bucketed records are dead once placed; the two node lists reuse them
  s.total = s.tail + s.spos + s.tab + s.hist + s.sums + 4096;
  return s;
}

// Optional payload travelling with the particles (whole-path use): sorted 32-byte records (search half + (v', m) half) and
// the sorted position of every node's nearest particle, so that the field kernel reads the payload almost sequentially.
template <typename T>
struct NNPayload {
  const T* vel = nullptr;
  const T* rho = nullptr;
  double lcell3 = 1.0;
  float4* srec_out = nullptr;    // [np] x 2 float4; null = scratch (only with fused field planes)
  int32_t* nn_pos_out = nullptr;  // [nnodes]; may be null with fused field planes
  FieldOut fo = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
};

template <typename T>
int nn_grid_typed(vp_ctx* ctx, const T* pos, int64_t np, const double* qx, int nx, const double* qy, int ny,
                  const double* qz, int nz, int32_t* nn, const NNPayload<T>* pay, const vp_nn_opts* opts, cudaStream_t st,
                  const vp_host_chunks* host_pos = nullptr) {
  const int64_t nnodes = int64_t(nx) * ny * nz;
  VP_REQUIRE(nnodes < (int64_t(1) << 32), "vp_nn_grid: lattice too large for 32-bit node ids");
  VP_REQUIRE(np < (int64_t(1) << 31), "vp_nn_grid: np must be < 2^31 per device");
  const bool has_pay = pay != nullptr;
  if (has_pay) VP_REQUIRE(pay->vel && ((pay->srec_out && pay->nn_pos_out) || pay->fo.on), "vp_nn_grid: incomplete payload description");
  vp_nn_opts o;
  memset(&o, 0, sizeof o);
  if (opts) o = *opts;
  const Grid g = plan_grid(np, qx, nx, qy, ny, qz, nz, o);
  const int gx = g.gx, gy = g.gy, gz = g.gz;
  const uint64_t ncells = uint64_t(gx) * gy * gz;

  // ---- lattice tables (host -> pinned -> device)
  const size_t nt = size_t(nx) + ny + nz;
  const size_t off_w = vp_align256(nt * 8), off_r = off_w + vp_align256(nt * 4), off_m = off_r + vp_align256(nt * 4);
  const size_t off_t = off_m + vp_align256(nt * 4);
  const size_t tab_bytes = off_t + vp_align256(nt * 4);
  if (ctx->pinned_cap < tab_bytes || ctx->small_cap < tab_bytes) {
    VP_CUDA(cudaStreamSynchronize(st));
    if (ctx->pinned_cap < tab_bytes) {
      if (ctx->pinned_h) cudaFreeHost(ctx->pinned_h);
      VP_CUDA(cudaMallocHost(&ctx->pinned_h, tab_bytes));
      ctx->pinned_cap = tab_bytes;
    }
    if (ctx->small_cap < tab_bytes) {
      if (ctx->small_d) cudaFree(ctx->small_d);
      VP_CUDA(cudaMalloc(&ctx->small_d, tab_bytes));
      ctx->small_cap = tab_bytes;
    }
  }
  // the pinned block may still be in flight from the previous call's upload: wait for that copy only
  if (ctx->ev_tables) VP_CUDA(cudaEventSynchronize(ctx->ev_tables));
  else VP_CUDA(cudaEventCreateWithFlags(&ctx->ev_tables, cudaEventDisableTiming));
  char* hb = static_cast<char*>(ctx->pinned_h);
  double* hq = reinterpret_cast<double*>(hb);
  int* hw = reinterpret_cast<int*>(hb + off_w);
  float* hr = reinterpret_cast<float*>(hb + off_r);
  float* hm = reinterpret_cast<float*>(hb + off_m);
  bool consecutive = true;
  auto fill_axis = [&](const double* q, int n, int at, double o_, double h_, double ih, int gg, bool closed_lo, bool closed_hi) {
    for (int i = 0; i < n; ++i) {
      double f = (q[i] - o_) * ih;
      int c = !(f > 0.0) ? 0 : (f >= double(gg) ? gg - 1 : int(f));
      hq[at + i] = q[i];
      // 2-cell window [w, w+1]: the neighbour on the side of the nearer face of cell c
      int w = (f - double(c) < 0.5) ? c - 1 : c;
      if (w > gg - 2) w = gg - 2;
      if (w < 0) w = 0;
      hw[at + i] = w;
      if (i > 0 && w != hw[at + i - 1] + 1) consecutive = false;
      hr[at + i] = float(q[i] - (o_ + double(w) * h_));
      // same expression as the device axis_margin() for the block [c0, c1]
      const int c0 = w, c1 = w + 1 > gg - 1 ? gg - 1 : w + 1;
      double m = INFINITY;
      if (c0 > 0) m = fmin(m, q[i] - (o_ + double(c0) * h_));
      else if (closed_lo) m = fmin(m, q[i] - o_);
      if (c1 < gg - 1) m = fmin(m, (o_ + double(c1 + 1) * h_) - q[i]);
      else if (closed_hi) m = fmin(m, (o_ + double(gg) * h_) - q[i]);
      float mf = INFINITY;
      if (m != INFINITY) {
        double ms = m * (1.0 - 1.0 / 1048576.0);
        mf = ms > 0.0 ? nextafterf(float(ms), -INFINITY) : 0.f;   // rounded down
        if (!(mf > 0.f)) mf = 0.f;
      }
      hm[at + i] = mf;
    }
  };
  fill_axis(qx, nx, 0, g.ox, g.hx, g.ihx, gx, g.closed_xlo != 0, g.closed_xhi != 0);
  fill_axis(qy, ny, nx, g.oy, g.hy, g.ihy, gy, false, false);
  fill_axis(qz, nz, nx + ny, g.oz, g.hz, g.ihz, gz, false, false);
  // brick kernel: consecutive windows on every axis, z windows starting on a 32-cell boundary, node inside its window
  bool brick = consecutive && (hw[nx + ny] % 32 == 0) && gx >= 2 && gy >= 2 && gz >= 2 && !getenv("VP_SEARCH_ROWS");   // (VP_SEARCH_ROWS: the node-centric kernel, for A/B runs)
  for (size_t t = 0; t < nt && brick; ++t) {
    const double hh = t < size_t(nx) ? g.hx : (t < size_t(nx + ny) ? g.hy : g.hz);
    if (!(hr[t] >= -0.5f * float(hh) && hr[t] <= 2.5f * float(hh))) brick = false;   // keeps the f32 error bound of the brick frame
  }
  // brick kernel: quantisation of d^2 and the per-axis proof thresholds on the quantised value (see k_search_brick)
  const float hmax = float(fmax(fmax(g.hx, g.hy), g.hz));
  const float eps = 2e-5f * hmax;
  BrickQ bq;
  memset(&bq, 0, sizeof bq);
  if (brick) {
    int* ht = reinterpret_cast<int*>(hb + off_t);
    float mmax = 0.f;
    for (size_t t = 0; t < nt; ++t)
      if (hm[t] != INFINITY && hm[t] > mmax) mmax = hm[t];
    const double e = double(eps);
    auto tol = [&](double b) { return 8.0 * sqrt(b) * e + 8.0 * e * e + 1e-6 * b; };     // judge()'s bound, in f64
    const double m2max = mmax > 0.f ? double(mmax) * double(mmax) : 27.0 * double(hmax) * double(hmax);
    const double dclamp = 1.05 * m2max + 4.0 * tol(2.0 * m2max);
    bq.S = float(double(kQ) / dclamp);
    bq.invS = float(1.0 / double(bq.S));
    const double S = double(bq.S), q = 1.6 / S;       // three roundings of 0.5 + the f32 roundings of the scaled squares
    bq.q2 = float(2.0 * q * (1.0 + 1e-6));
    bq.nt = uint32_t(ceil((2.0 * tol(dclamp) + 2.0 * q) * S * (1.0 + 1e-5))) + 2u;
    for (size_t t = 0; t < nt; ++t) {
      int thr = -1;
      if (hm[t] == INFINITY) thr = int(kQ) - 1;
      else {
        const double m2 = double(hm[t]) * double(hm[t]);
        if (m2 > 8.0 * e * e) {
          // largest b with b + tol(b) < m2:  (1 + 1e-6) r^2 + 8 e r + 8 e^2 - m2 = 0,  b = r^2
          const double a2 = 1.0 + 1e-6, r = (-8.0 * e + sqrt(64.0 * e * e + 4.0 * a2 * (m2 - 8.0 * e * e))) / (2.0 * a2);
          const double bsafe = r * r * (1.0 - 1e-6);
          const double nn_ = floor(bsafe * S - 1.6) - 1.0;
          thr = nn_ < 0.0 ? -1 : (nn_ > double(kQ) - 1.0 ? int(kQ) - 1 : int(nn_));
          while (thr >= 0 && !((double(thr) + 1.6) / S + tol((double(thr) + 1.6) / S) < m2)) --thr;   // (belt and braces)
        }
      }
      ht[t] = thr;
    }
    const char* dbt = reinterpret_cast<const char*>(ctx->small_d) + off_t;
    bq.tx = reinterpret_cast<const int*>(dbt); bq.ty = bq.tx + nx; bq.tz = bq.ty + ny;
  }
  VP_CUDA(cudaMemcpyAsync(ctx->small_d, ctx->pinned_h, tab_bytes, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaEventRecord(ctx->ev_tables, st));
  const char* db = reinterpret_cast<const char*>(ctx->small_d);
  Lattice L;
  L.qx = reinterpret_cast<const double*>(db); L.qy = L.qx + nx; L.qz = L.qy + ny;
  L.wx = reinterpret_cast<const int*>(db + off_w); L.wy = L.wx + nx; L.wz = L.wy + ny;
  L.rx = reinterpret_cast<const float*>(db + off_r); L.ry = L.rx + nx; L.rz = L.ry + ny;
  L.mx = reinterpret_cast<const float*>(db + off_m); L.my = L.mx + nx; L.mz = L.my + ny;
  L.nx = nx; L.ny = ny; L.nz = nz;

  // ---- scratch
  const NNScratch sc = nn_scratch(np, has_pay, ncells, g.nb, nnodes, !(has_pay && pay->srec_out));
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, sc.total));
  void* rec1 = vp_arena_alloc(ctx, sc.tail);
  rec_t* srec_own = static_cast<rec_t*>(vp_arena_alloc(ctx, sc.spos));
  rec_t* srec = (has_pay && pay->srec_out) ? reinterpret_cast<rec_t*>(pay->srec_out) : srec_own;
  uint32_t* xtab = static_cast<uint32_t*>(vp_arena_alloc(ctx, sc.tab));
  uint32_t* hist = static_cast<uint32_t*>(vp_arena_alloc(ctx, sc.hist));
  uint32_t* sums = static_cast<uint32_t*>(vp_arena_alloc(ctx, sc.sums));
  VP_REQUIRE(rec1 && srec_own && xtab && hist && sums, "vp_nn_grid: arena carve failed");
  uint32_t* cursor = hist + vp_align256((size_t(g.nb) * kSub + 1) * 4) / 4;
  uint32_t* node_list = static_cast<uint32_t*>(rec1);                                           // -> exact kernel
  uint32_t* list_b = node_list + vp_align256(size_t(nnodes) * 4) / 4;                            // -> wider stage
  uint32_t* crowded = list_b + vp_align256(size_t(nnodes) * 4) / 4;                              // -> node-centric stage A of dense bricks
  const int rs = has_pay ? 2 : 1;     // float4 stride of the sorted records

  VP_CUDA(cudaMemsetAsync(ctx->nn_stats_d, 0, sizeof(vp_nn_stats_dev), st));
  VP_CUDA(cudaMemsetAsync(xtab, 0, (ncells + 8) * 4, st));
  VP_CUDA(cudaMemsetAsync(hist, 0, (size_t(g.nb) * kSub + 1) * 4, st));
  // counts / cursors live at xtab + 4 (16-byte aligned for the scan); after the placement cursor c holds the start of cell
  // c + 1, so the start table is the same array read from xtab + 3 (entry 0 = the zero in front of the counters)
  uint32_t* tab = xtab + 4;
  if (np > 0) {
    const double es = sizeof(T);
    const size_t hsmem = size_t(g.nb) * kSub * 4, ssmem = size_t(g.nb) * 4;
    // thread-block clusters claiming one run per bucket together: measured 55 ms against 26 ms without (four cluster-wide
    // barriers per tile with one 1024-thread CTA per SM), kept behind a switch
    const bool use_clusters = getenv("VP_SCATTER_CLUSTERS") != nullptr;
    static const bool vec_on = !(getenv("VP_BIN_VEC") && atoi(getenv("VP_BIN_VEC")) == 0);
    static const bool wc_on = !(getenv("VP_BIN_WC") && atoi(getenv("VP_BIN_WC")) == 0);
    static bool attr_done = false;
    if (!attr_done) {
      VP_CUDA(cudaFuncSetAttribute(k_bin_hist<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxBuckets * kSub * 4)));
      VP_CUDA(cudaFuncSetAttribute(k_bin_hist<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxBuckets * kSub * 4)));
      VP_CUDA(cudaFuncSetAttribute(k_bin_hist<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxBuckets * kSub * 4)));
      VP_CUDA(cudaFuncSetAttribute(k_bin_hist<double, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxBuckets * kSub * 4)));
      VP_CUDA(cudaFuncSetAttribute(k_bin_hist<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxBuckets * kSub * 4)));
      VP_CUDA(cudaFuncSetAttribute(k_bin_hist<double, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxBuckets * kSub * 4)));
      const int wmax = int(2 * kWin * 16 + 2 * kMaxBuckets * 4 + 33 * 4);
      VP_CUDA(cudaFuncSetAttribute(k_bin_scatter_wc<float, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wmax));
      VP_CUDA(cudaFuncSetAttribute(k_bin_scatter_wc<double, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wmax));
      VP_CUDA(cudaFuncSetAttribute(k_bin_scatter_wc<float, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wmax));
      VP_CUDA(cudaFuncSetAttribute(k_bin_scatter_wc<double, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wmax));
      VP_CUDA(cudaFuncSetAttribute(k_bin_scatter_wc<float, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wmax));
      VP_CUDA(cudaFuncSetAttribute(k_bin_scatter_wc<float, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wmax));
      attr_done = true;
    }
    auto launch_hist = [&](const T* p, int64_t n_c, int64_t i0) {
      // persistent CTAs: as many as the shared-memory histograms allow per SM (ncu: the kernel waits on its position loads --
      // 68 % long-scoreboard stalls at 4 CTAs/SM)
      const int per_sm = int(std::max<size_t>(1, std::min<size_t>(8, (size_t(220) << 10) / (hsmem + 1024))));
      const int64_t nt = (n_c + kBinTile - 1) / kBinTile, cap = int64_t(ctx->sm_count) * per_sm;
      const bool vec = vec_on && g.ps == 3 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
      const bool rows = vec_on && sizeof(T) == 4 && g.ps == 8 && (reinterpret_cast<uintptr_t>(p) & 31) == 0;
      if (rows) k_bin_hist<T, 2><<<unsigned(nt < cap ? nt : cap), 256, hsmem, st>>>(p, n_c, g, hist, uint32_t((i0 / kBinTile) % (kSub * kClu)));
      else if (vec) k_bin_hist<T, 1><<<unsigned(nt < cap ? nt : cap), 256, hsmem, st>>>(p, n_c, g, hist, uint32_t((i0 / kBinTile) % (kSub * kClu)));
      else k_bin_hist<T, 0><<<unsigned(nt < cap ? nt : cap), 256, hsmem, st>>>(p, n_c, g, hist, uint32_t((i0 / kBinTile) % (kSub * kClu)));
    };
    auto launch_scatter = [&](const T* p, const T* v, const T* r, int64_t n_c, int64_t i0) {
      PayloadIn<T> pin;
      pin.vel = v;
      pin.rho = r;
      pin.lcell3 = T(has_pay ? pay->lcell3 : 1.0);
      const unsigned nbk = unsigned((n_c + kBinTile - 1) / kBinTile);
      const uint32_t t0 = uint32_t((i0 / kBinTile) % (kSub * kClu));
      if (use_clusters) {
        const unsigned nbc = (nbk + kClu - 1) / kClu * kClu;      // whole clusters; the surplus CTAs only take part in the syncs
        if (has_pay) k_bin_scatter_clu<T, true><<<nbc, 1024, 2 * ssmem, st>>>(p, pin, n_c, i0, g, cursor, t0, rec1);
        else k_bin_scatter_clu<T, false><<<nbc, 1024, 2 * ssmem, st>>>(p, pin, n_c, i0, g, cursor, t0, rec1);
      } else {
        // (512-thread CTAs on 2048-particle tiles, four per SM instead of two: 25.1 vs 24.7 ms at cfg4 -- the phases of a tile
        // are not what limits this kernel)
        const bool vec = vec_on && g.ps == 3 && (reinterpret_cast<uintptr_t>(p) & 15) == 0 &&
                         (!has_pay || (g.vs == 3 && (reinterpret_cast<uintptr_t>(v) & 15) == 0 && (!r || (g.rs == 1 && (reinterpret_cast<uintptr_t>(r) & 15) == 0))));
        // interleaved 32-byte rows (the slab exchange's output): x y z vx | vy vz rho -, read in place by the write-combining kernel
        const bool rows = vec_on && wc_on && sizeof(T) == 4 && g.ps == 8 && (reinterpret_cast<uintptr_t>(p) & 31) == 0 &&
                          (!has_pay || (g.vs == 8 && v == p + 3 && (!r || (g.rs == 8 && r == p + 6))));
        if (vec || rows) {
          // whole tiles with vector loads, the last partial tile element by element
          const unsigned nfull = unsigned(n_c / kBinTile);
          if (nfull && rows) {
            const size_t wsmem = size_t(has_pay ? 2 : 1) * kWin * 16 + 2 * size_t(kWcBuckets) * 4 + 33 * 4;
            if constexpr (sizeof(T) == 4) {
              if (has_pay) k_bin_scatter_wc<T, true, true><<<nfull, 1024, wsmem, st>>>(p, pin, i0, g, cursor, t0, rec1);
              else k_bin_scatter_wc<T, false, true><<<nfull, 1024, wsmem, st>>>(p, pin, i0, g, cursor, t0, rec1);
            }
          } else if (nfull && wc_on) {
            const size_t wsmem = size_t(has_pay ? 2 : 1) * kWin * 16 + 2 * size_t(kWcBuckets) * 4 + 33 * 4;
            if (has_pay) k_bin_scatter_wc<T, true, false><<<nfull, 1024, wsmem, st>>>(p, pin, i0, g, cursor, t0, rec1);
            else k_bin_scatter_wc<T, false, false><<<nfull, 1024, wsmem, st>>>(p, pin, i0, g, cursor, t0, rec1);
          } else if (nfull) {
            if (has_pay) k_bin_scatter<T, true, 1024, true><<<nfull, 1024, ssmem, st>>>(p, pin, n_c, i0, g, cursor, t0, rec1);
            else k_bin_scatter<T, false, 1024, true><<<nfull, 1024, ssmem, st>>>(p, pin, n_c, i0, g, cursor, t0, rec1);
          }
          if (nfull < nbk) {
            if (nfull) ctx->n_launch += 1;      // (the stage counts one launch; the ragged last tile is a second one)
            const int64_t done = int64_t(nfull) * kBinTile;
            PayloadIn<T> pt = pin;
            if (pt.vel) pt.vel += size_t(g.vs) * done;
            if (pt.rho) pt.rho += size_t(g.rs) * done;
            const uint32_t t1 = uint32_t(((i0 + done) / kBinTile) % (kSub * kClu));
            if (has_pay) k_bin_scatter<T, true, 1024, false><<<1, 1024, ssmem, st>>>(p + size_t(g.ps) * done, pt, n_c - done, i0 + done, g, cursor, t1, rec1);
            else k_bin_scatter<T, false, 1024, false><<<1, 1024, ssmem, st>>>(p + size_t(g.ps) * done, pt, n_c - done, i0 + done, g, cursor, t1, rec1);
          }
        } else {
          if (has_pay) k_bin_scatter<T, true, 1024, false><<<nbk, 1024, ssmem, st>>>(p, pin, n_c, i0, g, cursor, t0, rec1);
          else k_bin_scatter<T, false, 1024, false><<<nbk, 1024, ssmem, st>>>(p, pin, n_c, i0, g, cursor, t0, rec1);
        }
      }
    };
    if (host_pos) {
      // positions arrive from the host in chunks (copy stream); the histogram of chunk c is taken while chunk c+1 moves
      VP_REQUIRE(!has_pay && o.row_stride == 0 && !o.use_x_keep, "vp_nn_grid: host position streaming is the plain compact form");
      VP_REQUIRE(host_pos->chunk % (kBinTile * kClu) == 0 || host_pos->chunk >= np, "vp_nn_grid: host chunks must hold whole cluster tiles");
      VP_TRY(vp_host_streams(ctx));
      const int64_t chunk = host_pos->chunk;
      T* posd = const_cast<T*>(pos);
      VP_CUDA(cudaEventRecord(ctx->ev_used[0], st));   // order the copy stream after everything already queued on st
      VP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_used[0], 0));
      vp_stage stage(ctx, "k1a_bin_hist", st, int((np + chunk - 1) / chunk), double(np) * 3 * es);
      int c = 0;
      for (int64_t i0 = 0; i0 < np; i0 += chunk, ++c) {
        const int64_t n_c = np - i0 < chunk ? np - i0 : chunk;
        VP_CUDA(cudaMemcpyAsync(posd + 3 * i0, static_cast<const T*>(host_pos->pos_h) + 3 * i0, size_t(n_c) * 3 * sizeof(T), cudaMemcpyHostToDevice, ctx->copy_stream));
        VP_CUDA(cudaEventRecord(ctx->ev_h2d[c & 1], ctx->copy_stream));
        VP_CUDA(cudaStreamWaitEvent(st, ctx->ev_h2d[c & 1], 0));
        launch_hist(posd + 3 * i0, n_c, i0);
      }
    } else {
      vp_stage stage(ctx, "k1a_bin_hist", st, 1, double(np) * 3 * es);
      launch_hist(pos, np, 0);
    }
    k_bin_offsets<<<1, 1024, 0, st>>>(hist, cursor, g.nb * kSub, ctx->nn_stats_d);
    ctx->n_launch += 1;
    {
      // read pos (+vel, rho), write one record
      vp_stage stage(ctx, "k1b_bin_scatter", st, 1, double(np) * (has_pay ? (3 + 3 + (pay->rho ? 1 : 0)) * es + 32.0 : 3 * es + 16.0));
      launch_scatter(pos, has_pay ? pay->vel : nullptr, has_pay ? pay->rho : nullptr, np, 0);
    }
    const unsigned nbk = unsigned((np + 256 * kIlp - 1) / (256 * kIlp));
    {
      // the cell index of every record read (one sector), one counter bumped
      vp_stage stage(ctx, "k1c_cell_count", st, 1, double(np) * (has_pay ? 32.0 : 16.0) + double(ncells) * 4.0);
      if (has_pay) k_cell_count<true><<<nbk, 256, 0, st>>>(rec1, ctx->nn_stats_d, tab);
      else k_cell_count<false><<<nbk, 256, 0, st>>>(rec1, ctx->nn_stats_d, tab);
    }
    VP_TRY(vp_scan_exclusive_u32(ctx, tab, int64_t(ncells) + 1, sums, st, "k1d_cell_scan"));
    {
      // record read, cursor bumped, record written in cell order
      vp_stage stage(ctx, "k1e_cell_place", st, 1, double(np) * (has_pay ? 64.0 : 32.0) + double(ncells) * 4.0);
      PlaceGeom pg;
      pg.sx = float(g.hx / 2097152.0); pg.sy = float(g.hy / 2097152.0); pg.sz = float(g.hz / 2097152.0);
      pg.hz = float(g.hz);
      pg.gz = uint32_t(gz);
      if (has_pay) k_cell_place<true><<<nbk, 256, 0, st>>>(rec1, ctx->nn_stats_d, tab, pg, srec);
      else k_cell_place<false><<<nbk, 256, 0, st>>>(rec1, ctx->nn_stats_d, tab, pg, srec);
    }
    VP_CHECK_LAUNCH();
  }
  const uint32_t* start = xtab + 3;
  int32_t* nn_pos = has_pay ? pay->nn_pos_out : nullptr;
  SearchOut so;
  so.nn = nn; so.nn_pos = nn_pos; so.list_b = list_b; so.list_c = node_list; so.crowded = crowded; so.stats = ctx->nn_stats_d;
  so.f = has_pay ? pay->fo : FieldOut{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  {
    // sorted records read once + cell starts read once + one index written per node
    vp_stage stage(ctx, brick ? "k1f_search_brick" : "k1f_search_rows", st, 1,
                   double(np) * (has_pay ? 32.0 : 16.0) + double(ncells) * 4.0 + double(nnodes) * 4.0);
    if (brick) {
      dim3 grid((nz + kBZ - 1) / kBZ, (ny + kBY - 1) / kBY, (nx + kBX - 1) / kBX);
      VP_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "vp_nn_grid: lattice too large for the search launch");
      static bool brick_attr = false;
      if (!brick_attr) {
        VP_CUDA(cudaFuncSetAttribute(k_search_brick, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kBrickSmem)));
        brick_attr = true;
      }
      k_search_brick<<<grid, kBThreads, kBrickSmem, st>>>(srec, rs, start, g, L, eps, bq, so);
      k_search_crowded<<<ctx->sm_count * 6, 256, 0, st>>>(srec, rs, start, g, L, eps, so, grid.x, grid.y);
      ctx->n_launch += 1;
    } else {
      // block = (z nodes, y rows), one x plane per blockIdx.z: no integer division in the kernel
      // A warp = 32 consecutive z nodes; the 8 warps of a CTA take 8 consecutive y rows, so that the cell rows two
      // neighbouring node rows share are fetched into the SM's L1 once (9 cell rows instead of 16 per CTA).
      int bx = 32, by = 8;
      if (const char* ev = getenv("VP_SEARCH_BLOCK")) { bx = atoi(ev); if (bx < 32 || bx > 256 || (bx & (bx - 1))) bx = 32; by = 256 / bx; }
      if (nz > 32 && ny < by) { bx = 256; by = 1; }
      dim3 block(bx, by, 1), grid((nz + bx - 1) / bx, (ny + by - 1) / by, nx);
      VP_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "vp_nn_grid: lattice too large for the search launch");
      // 40 registers without spills (6 CTAs per SM) or 32 with 28 bytes of spills (8 CTAs per SM)
      static const bool occ8 = getenv("VP_SEARCH_OCC6") == nullptr;   // measured: 39.6 ms (8 CTAs/SM) vs 44.0 ms (6)
      if (rs == 2) {
        if (occ8) k_search_rows<2, 8><<<grid, block, 0, st>>>(srec, start, g, L, eps, so);
        else k_search_rows<2, 6><<<grid, block, 0, st>>>(srec, start, g, L, eps, so);
      } else {
        if (occ8) k_search_rows<1, 8><<<grid, block, 0, st>>>(srec, start, g, L, eps, so);
        else k_search_rows<1, 6><<<grid, block, 0, st>>>(srec, start, g, L, eps, so);
      }
    }
  }
  {
    vp_stage stage(ctx, "k1g_search_block4", st, 1);
    // latency bound (ncu: 92 % long-scoreboard stalls), yet more resident warps lose: 10.4 / 11.0 / 12.7 ms at 4 / 5 / 6 CTAs per
    // SM (60 / 48 / 40 registers) -- the spills and the extra DRAM pressure cost more than the occupancy gives
    static const int b4occ = getenv("VP_B4_OCC") ? atoi(getenv("VP_B4_OCC")) : 4;
    if (b4occ <= 4) k_search_block4<4><<<ctx->sm_count * 8, 256, 0, st>>>(srec, rs, start, g, L, eps, so);
    else if (b4occ == 5) k_search_block4<5><<<ctx->sm_count * 8, 256, 0, st>>>(srec, rs, start, g, L, eps, so);
    else k_search_block4<6><<<ctx->sm_count * 8, 256, 0, st>>>(srec, rs, start, g, L, eps, so);
  }
  {
    vp_stage stage(ctx, "k1h_search_exact", st, 1);
    k_search_exact<T><<<ctx->sm_count * 8, 256, 0, st>>>(srec, rs, start, pos, g, L, so);
  }
  VP_CHECK_LAUNCH();
  return VP_OK;
}

template <typename T>
int nn_payload_typed(vp_ctx* ctx, const void* pos, const void* vel, const void* rho, int64_t np, const double* qx, int nx,
                     const double* qy, int ny, const double* qz, int nz, double lcell3, int32_t* nn_idx, int32_t* nn_pos,
                     float* spay, const vp_nn_opts* opts, cudaStream_t st) {
  NNPayload<T> pay;
  pay.vel = static_cast<const T*>(vel);
  pay.rho = static_cast<const T*>(rho);
  pay.lcell3 = lcell3;
  pay.srec_out = reinterpret_cast<float4*>(spay);
  pay.nn_pos_out = nn_pos;
  return nn_grid_typed<T>(ctx, static_cast<const T*>(pos), np, qx, nx, qy, ny, qz, nz, nn_idx, &pay, opts, st);
}

// ------------------------------------------------------------------------------------------ slab bucketing
// Multi-GPU, sharded input: every particle goes to the rank(s) whose kept x range [lo_d, hi_d] contains it.
struct SlabRanges {
  double lo[16], hi[16];
  int n;
};

template <typename T>
__global__ void __launch_bounds__(256) k_bucket_count(const T* __restrict__ pos, int64_t np, SlabRanges R, unsigned long long* __restrict__ counts) {
  __shared__ unsigned sc[16];
  if (threadIdx.x < 16) sc[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool ok = i < np;
  const double x = ok ? double(pos[3 * i]) : 0.0;
  for (int d = 0; d < R.n; ++d) {
    const unsigned m = __ballot_sync(0xffffffffu, ok && x >= R.lo[d] && x <= R.hi[d]);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&sc[d], __popc(m));
  }
  __syncthreads();
  if (threadIdx.x < R.n && sc[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)sc[threadIdx.x]);
}

// rows = [x y z vx vy vz rho 0]: kSlabW = 8 elements (32 bytes in f32, 64 in f64; rho = 0 when absent) -- a padded row is one
// or two whole sectors and every run of rows is 16-byte aligned, which is what the bulk store below needs.
// cursors[d] starts at the first row this rank may write in destination d's buffer (exclusive prefix of counts for the
// single local buffer; with peer stores, the rows of lower ranks for d).
struct SlabDest {
  void* base[16];   // all equal for the local form; peer-mapped receive buffers for the fused exchange
};
constexpr int kSlabW = 8;
// Tiles of 512 x ITEMS particles per CTA (2048 for f32).  Pass 1: which destinations take each particle (one ballot per
// destination and item; only the per-warp counts are kept).  Then ONE contiguous row range is claimed per destination
// (one global atomic each) and the same ballots are formed again to place every row in shared memory, grouped by
// destination (vector stores).  Each group then leaves as ONE bulk copy shared -> global issued by a single thread
// (cp.async.bulk, the TMA engine; SASS UBLKCP): whole lines whether the destination is local HBM or a peer's buffer over
// NVLink, and no thread spends issue slots on the copy.  Rows that do not fit the staging area (very wide halos) are
// stored directly.
template <typename T> struct SlabTile { static constexpr int items = sizeof(T) == 4 ? 4 : 2; };
__device__ __forceinline__ void slab_store_row(float* o, const float (&r)[kSlabW]) {
  reinterpret_cast<float4*>(o)[0] = make_float4(r[0], r[1], r[2], r[3]);
  reinterpret_cast<float4*>(o)[1] = make_float4(r[4], r[5], r[6], r[7]);
}
__device__ __forceinline__ void slab_store_row(double* o, const double (&r)[kSlabW]) {
#pragma unroll
  for (int c = 0; c < kSlabW; c += 2) reinterpret_cast<double2*>(o)[c / 2] = make_double2(r[c], r[c + 1]);
}
template <typename T>
__global__ void __launch_bounds__(512) k_bucket_scatter(const T* __restrict__ pos, const T* __restrict__ vel, const T* __restrict__ rho,
                                                         int64_t np, SlabRanges R, unsigned long long* __restrict__ cursors,
                                                         SlabDest D, int cap_rows) {
  constexpr int ITEMS = SlabTile<T>::items;
  extern __shared__ __align__(128) unsigned char slab_smem[];
  T* stage = reinterpret_cast<T*>(slab_smem);            // [cap_rows][kSlabW]
  __shared__ unsigned wcnt[ITEMS][16][16];                 // [item][warp][destination]: count, then first staged row
  __shared__ unsigned pre[17];                             // exclusive prefix of the rows per destination
  __shared__ unsigned long long sbase[16];
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int64_t i0 = int64_t(blockIdx.x) * (512 * ITEMS);
  double x[ITEMS];
  bool okv[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int64_t i = i0 + r * 512 + tid;
    okv[r] = i < np;
    x[r] = okv[r] ? double(pos[3 * i]) : 0.0;
    for (int d = 0; d < R.n; ++d) {
      const unsigned m = __ballot_sync(0xffffffffu, okv[r] && x[r] >= R.lo[d] && x[r] <= R.hi[d]);
      if (lane == 0) wcnt[r][wp][d] = __popc(m);
    }
  }
  __syncthreads();
  if (tid < R.n) {
    unsigned run = 0;
    for (int r = 0; r < ITEMS; ++r)
      for (int q = 0; q < 16; ++q) {
        const unsigned c = wcnt[r][q][tid];
        wcnt[r][q][tid] = run;
        run += c;
      }
    pre[tid + 1] = run;                                  // (count; turned into the prefix below)
    sbase[tid] = run ? atomicAdd(cursors + tid, (unsigned long long)run) : 0ull;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned a = 0;
    pre[0] = 0;
    for (int d = 0; d < R.n; ++d) { a += pre[d + 1]; pre[d + 1] = a; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int64_t i = i0 + r * 512 + tid;
    T row[kSlabW];
    if (okv[r]) {
      row[0] = pos[3 * i]; row[1] = pos[3 * i + 1]; row[2] = pos[3 * i + 2];
      row[3] = vel[3 * i]; row[4] = vel[3 * i + 1]; row[5] = vel[3 * i + 2];
      row[6] = rho ? rho[i] : T(0);
      row[7] = T(0);
    }
    for (int d = 0; d < R.n; ++d) {
      const bool in = okv[r] && x[r] >= R.lo[d] && x[r] <= R.hi[d];
      const unsigned m = __ballot_sync(0xffffffffu, in);
      if (in) {
        const unsigned rank = wcnt[r][wp][d] + __popc(m & ((1u << lane) - 1u));
        const unsigned p = pre[d] + rank;
        if (p < unsigned(cap_rows)) slab_store_row(stage + size_t(p) * kSlabW, row);
        else slab_store_row(static_cast<T*>(D.base[d]) + (sbase[d] + rank) * size_t(kSlabW), row);   // staging area full
      }
    }
  }
  // the staged rows were written through the generic proxy; the bulk copy reads them through the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid < R.n) {
    const int d = tid;
    const unsigned r0 = min(pre[d], unsigned(cap_rows)), r1 = min(pre[d + 1], unsigned(cap_rows));
    if (r1 > r0) {
      const uint32_t src = uint32_t(__cvta_generic_to_shared(stage + size_t(r0) * kSlabW));
      T* dst = static_cast<T*>(D.base[d]) + (sbase[d] + (r0 - pre[d])) * size_t(kSlabW);
      const uint32_t bytes = (r1 - r0) * uint32_t(kSlabW * sizeof(T));
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // complete (not only read) before the CTA retires
    }
  }
}

__global__ void k_bucket_offsets(const unsigned long long* counts, unsigned long long* cursors, int n) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long a = 0;
    for (int d = 0; d < n; ++d) { cursors[d] = a; a += counts[d]; }
  }
}

template <typename T>
int slab_bucket_typed(vp_ctx* ctx, const T* pos, const T* vel, const T* rho, int64_t np, const SlabRanges& R, T* rows, int64_t cap,
                      int64_t* counts_h, cudaStream_t st) {
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, 1024));
  unsigned long long* cnt = static_cast<unsigned long long*>(vp_arena_alloc(ctx, 512));
  VP_REQUIRE(cnt, "vp_slab_bucket: arena carve failed");
  unsigned long long* cur = cnt + 16;
  VP_CUDA(cudaMemsetAsync(cnt, 0, 512, st));
  const unsigned nb = unsigned((np + 255) / 256);
  const int w = kSlabW;
  if (np > 0) {
    vp_stage stage(ctx, "k0_slab_bucket_count", st, 1, double(np) * 3.0 * sizeof(T));
    k_bucket_count<T><<<nb, 256, 0, st>>>(pos, np, R, cnt);
  }
  k_bucket_offsets<<<1, 32, 0, st>>>(cnt, cur, R.n);
  unsigned long long h[16];
  VP_CUDA(cudaMemcpyAsync(h, cnt, sizeof(unsigned long long) * 16, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaStreamSynchronize(st));   // the split sizes are needed on the host for the all-to-all
  int64_t total = 0;
  for (int d = 0; d < R.n; ++d) { counts_h[d] = int64_t(h[d]); total += counts_h[d]; }
  VP_REQUIRE(total <= cap, "vp_slab_bucket: %lld rows needed, capacity %lld", (long long)total, (long long)cap);
  if (np > 0) {
    vp_stage stage(ctx, "k0_slab_bucket_scatter", st, 1, double(np) * w * sizeof(T) + double(total) * w * sizeof(T));
    SlabDest D;
    for (int d = 0; d < 16; ++d) D.base[d] = rows;
    constexpr int kTile = 512 * SlabTile<T>::items;
    const int cap_rows = kTile + kTile / 4;                 // a particle inside a halo goes to two ranks
    const size_t smem = size_t(cap_rows) * kSlabW * sizeof(T);
    static bool attr = false;
    if (!attr) {
      VP_CUDA(cudaFuncSetAttribute(k_bucket_scatter<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      attr = true;
    }
    k_bucket_scatter<T><<<unsigned((np + kTile - 1) / kTile), 512, smem, st>>>(pos, vel, rho, np, R, cur, D, cap_rows);
  }
  VP_CHECK_LAUNCH();
  VP_CUDA(cudaStreamSynchronize(st));   // cnt/cur live in the scope released on return
  return VP_OK;
}

template <typename T>
int slab_count_typed(vp_ctx* ctx, const T* pos, int64_t np, const SlabRanges& R, int64_t* counts_h, cudaStream_t st) {
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, 1024));
  unsigned long long* cnt = static_cast<unsigned long long*>(vp_arena_alloc(ctx, 256));
  VP_REQUIRE(cnt, "vp_slab_count: arena carve failed");
  VP_CUDA(cudaMemsetAsync(cnt, 0, 256, st));
  if (np > 0) {
    vp_stage stage(ctx, "k0_slab_bucket_count", st, 1, double(np) * 3.0 * sizeof(T));
    k_bucket_count<T><<<unsigned((np + 255) / 256), 256, 0, st>>>(pos, np, R, cnt);
  }
  unsigned long long h[16];
  VP_CUDA(cudaMemcpyAsync(h, cnt, sizeof h, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaStreamSynchronize(st));
  for (int d = 0; d < R.n; ++d) counts_h[d] = int64_t(h[d]);
  return VP_OK;
}

template <typename T>
int slab_scatter_p2p_typed(vp_ctx* ctx, const T* pos, const T* vel, const T* rho, int64_t np, const SlabRanges& R,
                           const int64_t* first_row_h, cudaStream_t st) {
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, 1024));
  unsigned long long* cur = static_cast<unsigned long long*>(vp_arena_alloc(ctx, 256));
  VP_REQUIRE(cur, "vp_slab_scatter_p2p: arena carve failed");
  unsigned long long h[16];
  for (int d = 0; d < 16; ++d) h[d] = d < R.n ? (unsigned long long)first_row_h[d] : 0ull;
  VP_CUDA(cudaMemcpyAsync(cur, h, sizeof h, cudaMemcpyHostToDevice, st));
  const int w = kSlabW;
  if (np > 0) {
    vp_stage stage(ctx, "k0_slab_bucket_scatter", st, 1, double(np) * 2.0 * w * sizeof(T));
    SlabDest D;
    for (int d = 0; d < 16; ++d) D.base[d] = d < R.n ? ctx->slab_peer[d] : nullptr;
    constexpr int kTile = 512 * SlabTile<T>::items;
    const int cap_rows = kTile + kTile / 4;                 // a particle inside a halo goes to two ranks
    const size_t smem = size_t(cap_rows) * kSlabW * sizeof(T);
    static bool attr = false;
    if (!attr) {
      VP_CUDA(cudaFuncSetAttribute(k_bucket_scatter<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      attr = true;
    }
    k_bucket_scatter<T><<<unsigned((np + kTile - 1) / kTile), 512, smem, st>>>(pos, vel, rho, np, R, cur, D, cap_rows);
    VP_CHECK_LAUNCH();
  }
  VP_CUDA(cudaStreamSynchronize(st));   // cur lives in the scope released on return
  return VP_OK;
}

}  // namespace

// ---- sharded particle exchange fused into the bucketing kernel (peer stores over NVLink)
// Teardown order of an IPC-shared buffer (CUDA: freeing an exported allocation while another process still has it mapped is
// undefined):  every rank vp_slab_p2p_close()  ->  barrier across the ranks (caller)  ->  vp_slab_p2p_alloc() may free.
extern "C" int vp_slab_p2p_close(vp_ctx* ctx) {
  VP_REQUIRE(ctx, "vp_slab_p2p_close: null ctx");
  VP_CUDA(cudaSetDevice(ctx->device));
  VP_CUDA(cudaDeviceSynchronize());
  if (ctx->slab_open) {
    for (int d = 0; d < ctx->slab_nranks; ++d)
      if (d != ctx->slab_rank && ctx->slab_peer[d]) { cudaIpcCloseMemHandle(ctx->slab_peer[d]); ctx->slab_peer[d] = nullptr; }
    ctx->slab_open = false;
  }
  return VP_OK;
}

extern "C" int vp_slab_p2p_alloc(vp_ctx* ctx, size_t bytes, unsigned char* handle_out) {
  VP_REQUIRE(ctx && handle_out && bytes > 0, "vp_slab_p2p_alloc: bad argument");
  VP_REQUIRE(!ctx->slab_open, "vp_slab_p2p_alloc: peers are still mapped -- vp_slab_p2p_close() on every rank, then a barrier, first");
  VP_CUDA(cudaSetDevice(ctx->device));
  VP_CUDA(cudaDeviceSynchronize());
  if (ctx->slab_recv) { VP_CUDA(cudaFree(ctx->slab_recv)); ctx->slab_recv = nullptr; }
  VP_CUDA(cudaMalloc(&ctx->slab_recv, bytes));
  ctx->slab_recv_bytes = bytes;
  cudaIpcMemHandle_t h;
  VP_CUDA(cudaIpcGetMemHandle(&h, ctx->slab_recv));
  memcpy(handle_out, &h, 64);
  return VP_OK;
}

extern "C" int vp_slab_p2p_open(vp_ctx* ctx, int nranks, int rank, const unsigned char* all_handles) {
  VP_REQUIRE(ctx && all_handles && ctx->slab_recv && nranks >= 1 && nranks <= 16 && rank >= 0 && rank < nranks,
             "vp_slab_p2p_open: bad argument");
  VP_CUDA(cudaSetDevice(ctx->device));
  for (int d = 0; d < nranks; ++d) {
    if (d == rank) { ctx->slab_peer[d] = ctx->slab_recv; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, all_handles + size_t(d) * 64, 64);
    void* p = nullptr;
    VP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->slab_peer[d] = p;
  }
  ctx->slab_nranks = nranks;
  ctx->slab_rank = rank;
  ctx->slab_open = true;
  return VP_OK;
}

extern "C" int vp_slab_p2p_buffer(vp_ctx* ctx, void** ptr_out, size_t* bytes_out) {
  VP_REQUIRE(ctx && ptr_out && bytes_out, "vp_slab_p2p_buffer: bad argument");
  *ptr_out = ctx->slab_recv;
  *bytes_out = ctx->slab_recv_bytes;
  return VP_OK;
}

static SlabRanges make_ranges(const double* lo_h, const double* hi_h, int nranks) {
  SlabRanges R;
  R.n = nranks;
  for (int d = 0; d < nranks; ++d) { R.lo[d] = lo_h[d]; R.hi[d] = hi_h[d]; }
  return R;
}

extern "C" int vp_slab_count(vp_ctx* ctx, const void* pos_d, int dtype, int64_t np, const double* lo_h, const double* hi_h,
                             int nranks, int64_t* counts_h, void* stream) {
  VP_REQUIRE(ctx && pos_d && lo_h && hi_h && counts_h && nranks >= 1 && nranks <= 16 && np >= 0, "vp_slab_count: bad argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_CUDA(cudaSetDevice(ctx->device));
  const SlabRanges R = make_ranges(lo_h, hi_h, nranks);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == VP_F32) return slab_count_typed<float>(ctx, static_cast<const float*>(pos_d), np, R, counts_h, st);
  if (dtype == VP_F64) return slab_count_typed<double>(ctx, static_cast<const double*>(pos_d), np, R, counts_h, st);
  vp_set_error("vp_slab_count: unknown dtype %d", dtype);
  return VP_ERR_ARG;
}

extern "C" int vp_slab_scatter_p2p(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype, int64_t np,
                                   const double* lo_h, const double* hi_h, int nranks, const int64_t* first_row_h, void* stream) {
  VP_REQUIRE(ctx && pos_d && vel_d && lo_h && hi_h && first_row_h, "vp_slab_scatter_p2p: null argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_REQUIRE(ctx->slab_open && ctx->slab_nranks == nranks, "vp_slab_scatter_p2p: peer buffers not opened for %d ranks", nranks);
  VP_CUDA(cudaSetDevice(ctx->device));
  const SlabRanges R = make_ranges(lo_h, hi_h, nranks);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == VP_F32)
    return slab_scatter_p2p_typed<float>(ctx, static_cast<const float*>(pos_d), static_cast<const float*>(vel_d),
                                         static_cast<const float*>(rho_d), np, R, first_row_h, st);
  if (dtype == VP_F64)
    return slab_scatter_p2p_typed<double>(ctx, static_cast<const double*>(pos_d), static_cast<const double*>(vel_d),
                                          static_cast<const double*>(rho_d), np, R, first_row_h, st);
  vp_set_error("vp_slab_scatter_p2p: unknown dtype %d", dtype);
  return VP_ERR_ARG;
}

extern "C" int vp_slab_bucket(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype, int64_t np,
                              const double* lo_h, const double* hi_h, int nranks, void* rows_d, int64_t cap_rows, int64_t* counts_h,
                              void* stream) {
  VP_REQUIRE(ctx && pos_d && vel_d && lo_h && hi_h && rows_d && counts_h, "vp_slab_bucket: null argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_REQUIRE(nranks >= 1 && nranks <= 16 && np >= 0, "vp_slab_bucket: 1..16 ranks supported");
  VP_CUDA(cudaSetDevice(ctx->device));
  SlabRanges R;
  R.n = nranks;
  for (int d = 0; d < nranks; ++d) { R.lo[d] = lo_h[d]; R.hi[d] = hi_h[d]; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == VP_F32)
    return slab_bucket_typed<float>(ctx, static_cast<const float*>(pos_d), static_cast<const float*>(vel_d), static_cast<const float*>(rho_d),
                                    np, R, static_cast<float*>(rows_d), cap_rows, counts_h, st);
  if (dtype == VP_F64)
    return slab_bucket_typed<double>(ctx, static_cast<const double*>(pos_d), static_cast<const double*>(vel_d),
                                     static_cast<const double*>(rho_d), np, R, static_cast<double*>(rows_d), cap_rows, counts_h, st);
  vp_set_error("vp_slab_bucket: unknown dtype %d", dtype);
  return VP_ERR_ARG;
}

size_t vp_host_chunk_staging_bytes(int64_t chunk, int dtype, bool has_rho) {
  const size_t es = dtype == VP_F64 ? 8 : 4;
  return 2 * (vp_align256(size_t(chunk) * 3 * es) + (has_rho ? vp_align256(size_t(chunk) * es) : vp_align256(size_t(chunk) * es)));
}

int vp_host_streams(vp_ctx* ctx) {
  if (!ctx->copy_stream) {
    VP_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    // highest priority: the small pack kernels must slip in between the blocks of the gridding kernels, or the two staging
    // buffers (and with them the upload) stall behind whatever large grid is resident
    int prio_lo = 0, prio_hi = 0;
    VP_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    VP_CUDA(cudaStreamCreateWithPriority(&ctx->pack_stream, cudaStreamNonBlocking, prio_hi));
    for (int e = 0; e < 2; ++e) {
      VP_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[e], cudaEventDisableTiming));
      VP_CUDA(cudaEventCreateWithFlags(&ctx->ev_used[e], cudaEventDisableTiming));
    }
    VP_CUDA(cudaEventCreateWithFlags(&ctx->ev_pack, cudaEventDisableTiming));
  }
  return VP_OK;
}

int vp_host_fork(vp_ctx* ctx, cudaStream_t st) {
  VP_TRY(vp_host_streams(ctx));
  VP_CUDA(cudaEventRecord(ctx->ev_pack, st));
  VP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_pack, 0));
  VP_CUDA(cudaStreamWaitEvent(ctx->pack_stream, ctx->ev_pack, 0));
  return VP_OK;
}

int vp_nn_grid_host_pos(vp_ctx* ctx, const vp_host_chunks* hc, void* pos_d, int dtype, int64_t np, const double* qx, int nx,
                        const double* qy, int ny, const double* qz, int nz, int32_t* nn_idx_d, cudaStream_t st) {
  VP_REQUIRE(ctx && hc && hc->pos_h && hc->chunk > 0 && pos_d && nn_idx_d, "vp_nn_grid_host_pos: bad argument");
  if (dtype == VP_F32)
    return nn_grid_typed<float>(ctx, static_cast<const float*>(pos_d), np, qx, nx, qy, ny, qz, nz, nn_idx_d, nullptr, nullptr, st, hc);
  return nn_grid_typed<double>(ctx, static_cast<const double*>(pos_d), np, qx, nx, qy, ny, qz, nz, nn_idx_d, nullptr, nullptr, st, hc);
}

// (v', m) of one chunk, the arithmetic of k_keygen_pack (input dtype, then rounded to f32)
template <typename T>
__global__ void __launch_bounds__(256) k_pack_payload(const T* __restrict__ vel, const T* __restrict__ rho, int64_t n, T lcell3,
                                                       float4* __restrict__ pay) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
  T m = lcell3;
  if (rho) {
    const T rr = rho[i];
    vx = (vx * rr) / rr;
    vy = (vy * rr) / rr;
    vz = (vz * rr) / rr;
    m = rr * lcell3;
  }
  pay[i] = make_float4(float(vx), float(vy), float(vz), float(m));
}

template <typename T>
static int pack_payload_host_typed(vp_ctx* ctx, const vp_host_chunks* hc, int64_t np, double lcell3, char* stage_d, float4* pay,
                                   cudaStream_t st) {
  VP_TRY(vp_host_streams(ctx));
  const int64_t chunk = hc->chunk;
  const size_t vb = vp_align256(size_t(chunk) * 3 * sizeof(T)), rb = vp_align256(size_t(chunk) * sizeof(T));
  // (the side streams were ordered behind the caller's earlier work by vp_host_fork; they must NOT wait for the gridding
  // that has just been queued on st)
  int c = 0;
  for (int64_t i0 = 0; i0 < np; i0 += chunk, ++c) {
    const int64_t n_c = np - i0 < chunk ? np - i0 : chunk;
    const int b = c & 1;
    T* vbuf = reinterpret_cast<T*>(stage_d + size_t(b) * (vb + rb));
    T* rbuf = hc->rho_h ? reinterpret_cast<T*>(stage_d + size_t(b) * (vb + rb) + vb) : nullptr;
    if (c >= 2) VP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_used[b], 0));   // staging buffer free again
    VP_CUDA(cudaMemcpyAsync(vbuf, static_cast<const T*>(hc->vel_h) + 3 * i0, size_t(n_c) * 3 * sizeof(T), cudaMemcpyHostToDevice, ctx->copy_stream));
    if (rbuf) VP_CUDA(cudaMemcpyAsync(rbuf, static_cast<const T*>(hc->rho_h) + i0, size_t(n_c) * sizeof(T), cudaMemcpyHostToDevice, ctx->copy_stream));
    VP_CUDA(cudaEventRecord(ctx->ev_h2d[b], ctx->copy_stream));
    VP_CUDA(cudaStreamWaitEvent(ctx->pack_stream, ctx->ev_h2d[b], 0));
    k_pack_payload<T><<<unsigned((n_c + 255) / 256), 256, 0, ctx->pack_stream>>>(vbuf, rbuf, n_c, T(lcell3), pay + i0);
    VP_CUDA(cudaEventRecord(ctx->ev_used[b], ctx->pack_stream));
  }
  VP_CHECK_LAUNCH();
  ctx->n_launch += c;
  VP_CUDA(cudaEventRecord(ctx->ev_pack, ctx->pack_stream));
  VP_CUDA(cudaStreamWaitEvent(st, ctx->ev_pack, 0));
  return VP_OK;
}

int vp_pack_payload_host(vp_ctx* ctx, const vp_host_chunks* hc, int dtype, int64_t np, double lcell3, void* staging_d, float* pay_d,
                         cudaStream_t st) {
  VP_REQUIRE(ctx && hc && hc->vel_h && hc->chunk > 0 && staging_d && pay_d, "vp_pack_payload_host: bad argument");
  if (dtype == VP_F32)
    return pack_payload_host_typed<float>(ctx, hc, np, lcell3, static_cast<char*>(staging_d), reinterpret_cast<float4*>(pay_d), st);
  return pack_payload_host_typed<double>(ctx, hc, np, lcell3, static_cast<char*>(staging_d), reinterpret_cast<float4*>(pay_d), st);
}

size_t vp_nn_grid_scratch_bytes_tables(int64_t np, int pos_dtype, const double* qx, int nx, const double* qy, int ny,
                                       const double* qz, int nz, const vp_nn_opts* opts) {
  vp_nn_opts o;
  memset(&o, 0, sizeof o);
  if (opts) o = *opts;
  Grid g = plan_grid(np, qx, nx, qy, ny, qz, nz, o);
  (void)pos_dtype;
  return nn_scratch(np, true, uint64_t(g.gx) * g.gy * g.gz, g.nb, int64_t(nx) * ny * nz).total + 4096;
}

extern "C" int vp_nn_grid_payload(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype, int64_t np,
                                  const double* qx_h, int nx, const double* qy_h, int ny, const double* qz_h, int nz,
                                  double lcell3, int32_t* nn_idx_d, int32_t* nn_pos_d, float* spay_d, const vp_nn_opts* opts,
                                  void* stream) {
  VP_REQUIRE(ctx && pos_d && vel_d && qx_h && qy_h && qz_h && nn_pos_d && spay_d, "vp_nn_grid_payload: null argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_REQUIRE(np >= 0 && nx > 0 && ny > 0 && nz > 0, "vp_nn_grid_payload: bad sizes");
  VP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == VP_F32)
    return nn_payload_typed<float>(ctx, pos_d, vel_d, rho_d, np, qx_h, nx, qy_h, ny, qz_h, nz, lcell3, nn_idx_d, nn_pos_d, spay_d, opts, st);
  if (dtype == VP_F64)
    return nn_payload_typed<double>(ctx, pos_d, vel_d, rho_d, np, qx_h, nx, qy_h, ny, qz_h, nz, lcell3, nn_idx_d, nn_pos_d, spay_d, opts, st);
  vp_set_error("vp_nn_grid_payload: unknown dtype %d", dtype);
  return VP_ERR_ARG;
}

// K1 + K3 in one call: the search stages write the requested field planes themselves (no nn_pos / record round trip).
extern "C" int vp_nn_grid_fields(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype, int64_t np,
                                 const double* qx_h, int nx, const double* qy_h, int ny, const double* qz_h, int nz, double lcell3,
                                 float* const v_d[3], float* const p_d[3], float* e_d, float* m_d, int32_t* nn_idx_d,
                                 const vp_nn_opts* opts, void* stream) {
  VP_REQUIRE(ctx && pos_d && vel_d && qx_h && qy_h && qz_h, "vp_nn_grid_fields: null argument");
  VP_REQUIRE(np >= 0 && nx > 0 && ny > 0 && nz > 0, "vp_nn_grid_fields: bad sizes");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FieldOut fo;
  fo.vx = v_d ? v_d[0] : nullptr; fo.vy = v_d ? v_d[1] : nullptr; fo.vz = v_d ? v_d[2] : nullptr;
  fo.px = p_d ? p_d[0] : nullptr; fo.py = p_d ? p_d[1] : nullptr; fo.pz = p_d ? p_d[2] : nullptr;
  fo.e = e_d; fo.m = m_d;
  fo.on = (fo.vx || fo.vy || fo.vz || fo.px || fo.py || fo.pz || fo.e || fo.m) ? 1 : 0;
  VP_REQUIRE(fo.on, "vp_nn_grid_fields: no plane requested");
  if (dtype == VP_F32) {
    NNPayload<float> pay;
    pay.vel = static_cast<const float*>(vel_d); pay.rho = static_cast<const float*>(rho_d); pay.lcell3 = lcell3; pay.fo = fo;
    return nn_grid_typed<float>(ctx, static_cast<const float*>(pos_d), np, qx_h, nx, qy_h, ny, qz_h, nz, nn_idx_d, &pay, opts, st);
  }
  if (dtype == VP_F64) {
    NNPayload<double> pay;
    pay.vel = static_cast<const double*>(vel_d); pay.rho = static_cast<const double*>(rho_d); pay.lcell3 = lcell3; pay.fo = fo;
    return nn_grid_typed<double>(ctx, static_cast<const double*>(pos_d), np, qx_h, nx, qy_h, ny, qz_h, nz, nn_idx_d, &pay, opts, st);
  }
  vp_set_error("vp_nn_grid_fields: unknown dtype %d", dtype);
  return VP_ERR_ARG;
}

extern "C" int vp_fields_sorted(vp_ctx* ctx, const int32_t* nn_pos_d, int64_t n_nodes, const float* spay_d, float* const v_d[3],
                                float* const p_d[3], float* e_d, float* m_d, void* stream) {
  return vp_fields_from_records(ctx, nn_pos_d, n_nodes, spay_d, 2, 1, v_d, p_d, e_d, m_d, static_cast<cudaStream_t>(stream));
}

int vp_fields_from_records(vp_ctx* ctx, const int32_t* nn_pos_d, int64_t n_nodes, const float* spay_d, int stride, int offset,
                           float* const v_d[3], float* const p_d[3], float* e_d, float* m_d, cudaStream_t st) {
  VP_REQUIRE(ctx && nn_pos_d && spay_d, "vp_fields_sorted: null argument");
  vp_call_guard guard(ctx, st);
  if (n_nodes == 0) return VP_OK;
  float* v[3] = {v_d ? v_d[0] : nullptr, v_d ? v_d[1] : nullptr, v_d ? v_d[2] : nullptr};
  float* p[3] = {p_d ? p_d[0] : nullptr, p_d ? p_d[1] : nullptr, p_d ? p_d[2] : nullptr};
  int nplanes = (e_d != nullptr) + (m_d != nullptr);
  for (int c = 0; c < 3; ++c) nplanes += (v[c] != nullptr) + (p[c] != nullptr);
  // per node: sorted position read, 16-byte payload record read, 4 B written per plane
  vp_stage stage(ctx, "k3_fields_sorted", st, 1, double(n_nodes) * (4.0 + 16.0 + 4.0 * nplanes));
  k_fields_sorted<<<unsigned((n_nodes + 255) / 256), 256, 0, st>>>(nn_pos_d, n_nodes, reinterpret_cast<const float4*>(spay_d), stride,
                                                                   offset, v[0], v[1], v[2], p[0], p[1], p[2], e_d, m_d);
  VP_CHECK_LAUNCH();
  return VP_OK;
}

extern "C" int vp_nn_grid(vp_ctx* ctx, const void* pos_d, int pos_dtype, int64_t np, const double* qx_h, int nx,
                          const double* qy_h, int ny, const double* qz_h, int nz, int32_t* nn_idx_d,
                          const vp_nn_opts* opts, void* stream) {
  VP_REQUIRE(ctx && pos_d && qx_h && qy_h && qz_h && nn_idx_d, "vp_nn_grid: null argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_REQUIRE(np >= 0 && nx > 0 && ny > 0 && nz > 0, "vp_nn_grid: bad sizes");
  VP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pos_dtype == VP_F32)
    return nn_grid_typed<float>(ctx, static_cast<const float*>(pos_d), np, qx_h, nx, qy_h, ny, qz_h, nz, nn_idx_d, nullptr, opts, st);
  if (pos_dtype == VP_F64)
    return nn_grid_typed<double>(ctx, static_cast<const double*>(pos_d), np, qx_h, nx, qy_h, ny, qz_h, nz, nn_idx_d, nullptr, opts, st);
  vp_set_error("vp_nn_grid: unknown dtype %d", pos_dtype);
  return VP_ERR_ARG;
}

extern "C" int vp_nn_grid_plan(int64_t np, const double* qx_h, int nx, const double* qy_h, int ny, const double* qz_h, int nz,
                               const vp_nn_opts* opts, int64_t* info_out) {
  VP_REQUIRE(qx_h && qy_h && qz_h && info_out, "vp_nn_grid_plan: null argument");
  VP_REQUIRE(np >= 0 && nx > 0 && ny > 0 && nz > 0, "vp_nn_grid_plan: bad sizes");
  vp_nn_opts o;
  memset(&o, 0, sizeof o);
  if (opts) o = *opts;
  const Grid g = plan_grid(np, qx_h, nx, qy_h, ny, qz_h, nz, o);
  const uint64_t ncells = uint64_t(g.gx) * g.gy * g.gz;
  info_out[0] = g.gx; info_out[1] = g.gy; info_out[2] = g.gz;
  info_out[3] = g.bshift; info_out[4] = g.nb; info_out[5] = 0; info_out[6] = 0;
  info_out[7] = 0;
  info_out[8] = int64_t((nn_scratch(np, true, ncells, g.nb, int64_t(nx) * ny * nz).total + (size_t(1) << 20) - 1) >> 20);
  // corner aligned: cells of the node spacing with every node on a cell corner (the 2x2x2 block then proves radius h)
  const double sp = nx > 1 ? (qx_h[nx - 1] - qx_h[0]) / (nx - 1) : 0.0;
  info_out[9] = (g.gx == nx + 1 && nx > 1 && fabs(g.hx - sp) <= 1e-9 * fabs(sp)) ? 1 : 0;
  return VP_OK;
}

extern "C" int vp_nn_grid_stats(vp_ctx* ctx, int64_t* n_wide, int64_t* n_unresolved, int64_t* n_kept, void* stream) {
  VP_REQUIRE(ctx, "vp_nn_grid_stats: null ctx");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  vp_nn_stats_dev h;
  VP_CUDA(cudaMemcpyAsync(&h, ctx->nn_stats_d, sizeof h, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  VP_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  if (n_wide) *n_wide = int64_t(h.n_wide);
  if (n_unresolved) *n_unresolved = int64_t(h.n_unresolved);
  if (n_kept) *n_kept = int64_t(h.n_kept);
  return VP_OK;
}

extern "C" int vp_nn_grid_stats_ex(vp_ctx* ctx, int64_t* out5, void* stream) {
  VP_REQUIRE(ctx && out5, "vp_nn_grid_stats_ex: null argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  vp_nn_stats_dev h;
  VP_CUDA(cudaMemcpyAsync(&h, ctx->nn_stats_d, sizeof h, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  VP_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  out5[0] = int64_t(h.n_wide); out5[1] = int64_t(h.n_unresolved); out5[2] = int64_t(h.n_kept); out5[3] = int64_t(h.n_b);
  out5[4] = int64_t(h.n_far);
  return VP_OK;
}

extern "C" int vp_gather_rows(vp_ctx* ctx, const int32_t* idx_d, int64_t n, const void* src_d, int row_bytes, void* dst_d,
                              void* stream) {
  VP_REQUIRE(ctx && idx_d && src_d && dst_d, "vp_gather_rows: null argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_REQUIRE(row_bytes > 0 && row_bytes % 4 == 0, "vp_gather_rows: row_bytes must be a multiple of 4");
  if (n == 0) return VP_OK;
  int rw = row_bytes / 4;
  int64_t tot = n * rw;
  vp_stage stage(ctx, "gather_rows", static_cast<cudaStream_t>(stream), 1, double(n) * (4.0 + 2.0 * row_bytes));
  k_gather_words<<<unsigned((tot + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      idx_d, n, static_cast<const uint32_t*>(src_d), rw, static_cast<uint32_t*>(dst_d));
  VP_CHECK_LAUNCH();
  return VP_OK;
}

extern "C" int vp_build_fields(vp_ctx* ctx, const int32_t* nn_idx_d, int64_t n_nodes, const void* vel_d, const void* rho_d,
                               int dtype, double lcell3, float* const v_d[3], float* const p_d[3], float* e_d, float* m_d,
                               void* stream) {
  VP_REQUIRE(ctx && nn_idx_d && vel_d, "vp_build_fields: null argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  if (n_nodes == 0) return VP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* v[3] = {v_d ? v_d[0] : nullptr, v_d ? v_d[1] : nullptr, v_d ? v_d[2] : nullptr};
  float* p[3] = {p_d ? p_d[0] : nullptr, p_d ? p_d[1] : nullptr, p_d ? p_d[2] : nullptr};
  unsigned nb = unsigned((n_nodes + 255) / 256);
  int nplanes = 0;
  for (int c = 0; c < 3; ++c) nplanes += (v[c] != nullptr) + (p[c] != nullptr);
  nplanes += (e_d != nullptr) + (m_d != nullptr);
  const double es = dtype == VP_F64 ? 8.0 : 4.0;
  // per node: index read, (v, rho) gathered, 4 B written per plane
  vp_stage stage(ctx, "k3_build_fields", st, 1, double(n_nodes) * (4.0 + (rho_d ? 4.0 : 3.0) * es + 4.0 * nplanes));
  if (dtype == VP_F32)
    k_build_fields<float><<<nb, 256, 0, st>>>(nn_idx_d, n_nodes, static_cast<const float*>(vel_d),
                                              static_cast<const float*>(rho_d), float(lcell3), v[0], v[1], v[2], p[0], p[1],
                                              p[2], e_d, m_d);
  else if (dtype == VP_F64)
    k_build_fields<double><<<nb, 256, 0, st>>>(nn_idx_d, n_nodes, static_cast<const double*>(vel_d),
                                               static_cast<const double*>(rho_d), lcell3, v[0], v[1], v[2], p[0], p[1], p[2],
                                               e_d, m_d);
  else {
    vp_set_error("vp_build_fields: unknown dtype %d", dtype);
    return VP_ERR_ARG;
  }
  VP_CHECK_LAUNCH();
  return VP_OK;
}
