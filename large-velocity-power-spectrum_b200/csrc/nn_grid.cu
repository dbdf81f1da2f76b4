// K1: exact nearest-particle gridding on a sorted cell list.  See include/vpower_b200.h (vp_nn_grid).
//
// Pipeline:  keygen + pack (cell key per particle, optional x filter, one packed record per particle)
//            ->  radix sort of (key, slot) on the ROW bits of the key only (a row = one x plane, a chunk of y, all z)
//            ->  row starts  ->  one CTA per row: counting sort of the row by cell in shared memory, fused with the
//                permutation of the packed records into cell order and with the cell-start table
//            ->  ring-1 search per lattice node (f32 prefilter)  ->  exact f64 search (warp per node, growing ring)
//                for every node the prefilter could not settle.
//
// Exactness: a candidate is accepted only if its distance is strictly below the distance from the
// node to every face of the searched cell block behind which unexamined particles can exist.  All
// distance arithmetic is f64 with the reference's association ((dx*dx+dy*dy)+dz*dz) and no FMA
// contraction (explicit __dmul_rn/__dadd_rn); ties go to the lowest particle index.
#include <math.h>

#include "common.cuh"

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
  // sort key = (row << lb) | local,  row = cx * nyc + (cy >> yb),  local = (cy & (2^yb - 1)) * gz + cz  (< bins <= 2^lb)
  int yb, lb, nyc, bins;
};

// Sorted particle record: position relative to the grid origin rounded to f32 (used only by the f32
// prefilter; every close call is re-decided in f64 from the caller's array) and the particle index.
typedef float4 rec_t;

__device__ __forceinline__ int cell_of(double x, double o, double ih, int g) {
  double f = (x - o) * ih;
  if (!(f > 0.0)) return 0;
  if (f >= double(g)) return g - 1;
  return int(f);
}

// Packed particle record written once, in input order, so that the permutation after the sort is ONE random
// 64-byte access per particle (separate pos/vel/rho arrays would cost three).
//   a = (x-ox, y-oy, z-oz as f32, particle index)        b = (vx', vy', vz', m)  [only with a payload]
// with v' = (rho*v)/rho and m = rho*Lcell^3 evaluated in the input dtype (interp.py:199-213,272-273).
struct __align__(32) rec32_t { float4 a, b; };

template <typename T>
struct PayloadIn {
  const T* vel;   // [np,3] or null
  const T* rho;   // [np] or null (rho = 1)
  T lcell3;
};

// ITEMS particles per thread: 1 without the slab filter (pure streaming), 8 with it so that the compaction needs
// only one global atomic per 2048-particle block
template <typename T, bool PAY, int kKeygenItems>
__global__ void __launch_bounds__(256) k_keygen_pack(const T* __restrict__ pos, PayloadIn<T> pin, int64_t np, int64_t i0, Grid g,
                                                      uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                      void* __restrict__ packed, unsigned long long* __restrict__ kept) {
  // pos / pin.vel / pin.rho point at particle i0 (a chunk); indices and unfiltered slots are global (i0 + local)
  __shared__ unsigned warp_cnt[8];
  __shared__ unsigned long long block_base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t base_i = int64_t(blockIdx.x) * (256 * kKeygenItems);
  bool ok[kKeygenItems];
  uint32_t key[kKeygenItems];
  float4 a[kKeygenItems];
  unsigned mine = 0;
#pragma unroll
  for (int r = 0; r < kKeygenItems; ++r) {
    const int64_t i = base_i + r * 256 + threadIdx.x;
    ok[r] = i < np;
    double x = 0, y = 0, z = 0;
    if (ok[r]) {
      x = pos[size_t(g.ps) * i];
      y = pos[size_t(g.ps) * i + 1];
      z = pos[size_t(g.ps) * i + 2];
      // slab filter: a particle beyond a CLOSED face is dropped; beyond an open (domain-edge) face it is kept and clamped
      if (g.use_keep && ((g.closed_xlo && !(x >= g.keep_lo)) || (g.closed_xhi && !(x <= g.keep_hi)))) ok[r] = false;
    }
    key[r] = 0;
    if (ok[r]) {
      int cx = cell_of(x, g.ox, g.ihx, g.gx), cy = cell_of(y, g.oy, g.ihy, g.gy), cz = cell_of(z, g.oz, g.ihz, g.gz);
      const uint32_t row = uint32_t(cx) * uint32_t(g.nyc) + (uint32_t(cy) >> g.yb);
      const uint32_t loc = (uint32_t(cy) & ((1u << g.yb) - 1u)) * uint32_t(g.gz) + uint32_t(cz);
      key[r] = (row << g.lb) | loc;
      a[r] = make_float4(float(x - g.ox), float(y - g.oy), float(z - g.oz), __int_as_float(int(i0 + i)));
      ++mine;
    }
  }
  // output slot: identity without the filter; with it, a block-wide exclusive scan + one atomic per block
  // (order is irrelevant: the sort follows and ties are decided by particle index)
  unsigned long long slot0 = 0;
  if (g.use_keep) {
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_cnt[w] = incl;
    __syncthreads();
    unsigned woff = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < w) woff += warp_cnt[q];
      tot += warp_cnt[q];
    }
    if (threadIdx.x == 0) block_base = tot ? atomicAdd(kept, (unsigned long long)tot) : 0ull;
    __syncthreads();
    slot0 = block_base + woff + incl - mine;
  }
  unsigned nth = 0;
#pragma unroll
  for (int r = 0; r < kKeygenItems; ++r) {
    if (!ok[r]) continue;
    const int64_t i = base_i + r * 256 + threadIdx.x;
    const int64_t o = g.use_keep ? int64_t(slot0 + nth) : i0 + i;
    ++nth;
    keys[o] = key[r];
    vals[o] = uint32_t(o);
    if (PAY) {
      T vx = pin.vel[size_t(g.vs) * i], vy = pin.vel[size_t(g.vs) * i + 1], vz = pin.vel[size_t(g.vs) * i + 2];
      T m = pin.lcell3;
      if (pin.rho) {
        T rr = pin.rho[size_t(g.rs) * i];
        vx = (vx * rr) / rr;
        vy = (vy * rr) / rr;
        vz = (vz * rr) / rr;
        m = rr * pin.lcell3;
      }
      rec32_t* out = static_cast<rec32_t*>(packed) + o;
      out->a = a[r];
      out->b = make_float4(float(vx), float(vy), float(vz), float(m));
    } else {
      static_cast<float4*>(packed)[o] = a[r];
    }
  }
}

template <bool PAY>
__global__ void __launch_bounds__(256) k_permute(const void* __restrict__ packed, const uint32_t* __restrict__ vals, int64_t n,
                                                  rec_t* __restrict__ spos, float4* __restrict__ spay) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t j = vals[i];
  if (PAY) {
    const float4* src = reinterpret_cast<const float4*>(static_cast<const rec32_t*>(packed) + j);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    spos[i] = a;
    spay[i] = b;
  } else {
    spos[i] = __ldg(static_cast<const float4*>(packed) + j);
  }
}

// row_start[r] = first sorted position whose row (key >> shift) is >= r, for r in [0, nrows]; row_start[nrows] = n.
// Four consecutive keys per thread (one 16-byte load).
__global__ void __launch_bounds__(256) k_row_starts(const uint32_t* __restrict__ keys, int64_t n, int shift, uint32_t nrows,
                                                     uint32_t* __restrict__ start) {
  const int64_t i0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const int lane = threadIdx.x & 31;
  uint32_t k4[4] = {0u, 0u, 0u, 0u};
  if (i0 + 3 < n) {
    const uint4 v = *reinterpret_cast<const uint4*>(keys + i0);
    k4[0] = v.x; k4[1] = v.y; k4[2] = v.z; k4[3] = v.w;
  } else {
    for (int u = 0; u < 4; ++u)
      if (i0 + u < n) k4[u] = keys[i0 + u];
  }
  int64_t kp = (i0 == 0 || i0 >= n) ? -1 : int64_t(keys[i0 - 1] >> shift);
  // almost every warp sees no row boundary at all (2^30 keys, 3e4 rows): leave before the per-key logic
  bool change = false;
  {
    int64_t prev = kp;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u < n) {
        const int64_t k = int64_t(k4[u] >> shift);
        change |= (k != prev) || (i0 + u == n - 1);
        prev = k;
      }
  }
  if (!__any_sync(0xffffffffu, change)) return;
#pragma unroll 1
  for (int u = 0; u < 4; ++u) {
    const int64_t i = i0 + u;
    int64_t lo = 1, hi = 0;  // empty range
    uint32_t v = 0;
    if (i < n) {
      const uint32_t k = k4[u] >> shift;
      if (int64_t(k) != kp) { lo = kp + 1; hi = k; v = uint32_t(i); }
      kp = k;
      if (i == n - 1) {
        // tail: rows after the last key (done by this thread after its own range)
        for (int64_t c = lo; c <= hi; ++c) start[c] = v;
        lo = int64_t(k) + 1; hi = nrows; v = uint32_t(n);
      }
    }
    // long gaps are filled by the whole warp
    unsigned big = __ballot_sync(0xffffffffu, hi - lo >= 32);
    while (big) {
      int src = __ffs(big) - 1;
      big &= big - 1;
      int64_t l = __shfl_sync(0xffffffffu, lo, src), h = __shfl_sync(0xffffffffu, hi, src);
      uint32_t vv = __shfl_sync(0xffffffffu, v, src);
      for (int64_t c = l + lane; c <= h; c += 32) start[c] = vv;
      if (lane == src) { lo = 1; hi = 0; }
    }
    for (int64_t c = lo; c <= hi; ++c) start[c] = v;
  }
}

// One CTA per SM, rows of cells handed out by a global cursor (a row = one x plane, 2^yb consecutive y, all z: `bins`
// cells, contiguous in the linear cell order).  The radix sort has brought the row's (key, slot) pairs together; here
// they are counted per cell in shared memory, the exclusive prefix gives the row's part of the cell-start table, and
// every slot is written to its place inside the row (taken from the per-cell cursors) in a shared-memory order table
// that is then copied out coalesced: this finishes the sort without the two radix passes over the low key bits.
// (Scattering the slots straight to global memory is transaction bound, ~45 G scattered stores/s.  Gathering the packed
// records in the same kernel was tried and lost: 51 ms against 29 ms for the plain k_permute that follows.)
//   short rows (len <= ordcap < 65536): 16-bit counters, two per word, + u32 order table   -- the common case
//   long rows (dense clusters):         32-bit counters, slots scattered to the global array
// The order of the particles INSIDE a cell is whatever the cursors hand out; no result depends on it (ties are decided
// by particle index).
constexpr int kGroupThreads = 1024;
constexpr int kGroupUnroll = 8;
constexpr int kMaxBins = 33 * 1024;         // 132 KB of u32 counters: a 32 x 1025 (y, z) tile of the 1024^3 cell grid fits
constexpr int kGroupSmemMax = 227 * 1024 - 256;   // opt-in dynamic shared memory, minus the static part

template <bool SHORT>
__device__ __forceinline__ uint32_t group_bump(uint32_t* cnt, uint32_t l) {   // previous value of counter l, then +1
  if (SHORT) {
    const uint32_t old = atomicAdd(&cnt[l >> 1], (l & 1u) ? 0x10000u : 1u);
    return (l & 1u) ? (old >> 16) : (old & 0xffffu);
  }
  return atomicAdd(&cnt[l], 1u);
}

template <bool SHORT>
__device__ __forceinline__ void group_row(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                          uint32_t* __restrict__ start, uint32_t* __restrict__ ordg, uint32_t* cnt, uint32_t* ord,
                                          uint32_t* wsum, uint32_t lmask, uint32_t s, uint32_t e, uint32_t used, size_t cell0,
                                          bool last, uint32_t n) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t len = e - s;
  const uint32_t words = SHORT ? (used + 1) / 2 : used;
  for (uint32_t l = tid; l < words; l += kGroupThreads) cnt[l] = 0;
  __syncthreads();
  for (uint32_t p0 = s; p0 < e; p0 += kGroupThreads * kGroupUnroll) {
    uint32_t k[kGroupUnroll];
#pragma unroll
    for (int u = 0; u < kGroupUnroll; ++u) {
      const uint32_t p = p0 + u * kGroupThreads + tid;
      k[u] = p < e ? keys[p] : 0xffffffffu;
    }
#pragma unroll
    for (int u = 0; u < kGroupUnroll; ++u)
      if (p0 + u * kGroupThreads + tid < e) group_bump<SHORT>(cnt, k[u] & lmask);
  }
  __syncthreads();
  // exclusive scan of the counters: thread t owns words [t*ch, (t+1)*ch)
  const uint32_t ch = ((words + kGroupThreads - 1) / kGroupThreads) | 1u;   // odd: conflict-free strided scan
  const uint32_t b0 = min(uint32_t(tid) * ch, words), b1 = min(b0 + ch, words);
  uint32_t sum = 0;
  for (uint32_t l = b0; l < b1; ++l) {
    const uint32_t c = cnt[l];
    sum += SHORT ? (c & 0xffffu) + (c >> 16) : c;
  }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t v = wsum[lane], iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += t;
    }
    wsum[lane] = iv - v;
  }
  __syncthreads();
  uint32_t run = wsum[w] + incl - sum;   // row-relative
  for (uint32_t l = b0; l < b1; ++l) {
    const uint32_t c = cnt[l];
    if (SHORT) {
      const uint32_t c0 = c & 0xffffu, c1 = c >> 16;
      cnt[l] = run | ((run + c0) << 16);
      run += c0 + c1;
    } else {
      cnt[l] = run;
      run += c;
    }
  }
  __syncthreads();
  if (SHORT) {
    const uint16_t* c16 = reinterpret_cast<const uint16_t*>(cnt);
    for (uint32_t l = tid; l < used; l += kGroupThreads) start[cell0 + l] = s + c16[l];
  } else {
    for (uint32_t l = tid; l < used; l += kGroupThreads) start[cell0 + l] = s + cnt[l];
  }
  if (last && tid == 0) start[cell0 + used] = n;
  __syncthreads();   // the cursors move only after the table has been copied out
  for (uint32_t p0 = s; p0 < e; p0 += kGroupThreads * kGroupUnroll) {
    uint32_t k[kGroupUnroll], v[kGroupUnroll];
#pragma unroll
    for (int u = 0; u < kGroupUnroll; ++u) {
      const uint32_t p = p0 + u * kGroupThreads + tid;
      k[u] = p < e ? keys[p] : 0xffffffffu;
      v[u] = p < e ? vals[p] : 0u;
    }
#pragma unroll
    for (int u = 0; u < kGroupUnroll; ++u) {
      if (p0 + u * kGroupThreads + tid >= e) continue;
      const uint32_t d = group_bump<SHORT>(cnt, k[u] & lmask);
      if (SHORT) ord[d] = v[u];
      else ordg[s + d] = v[u];
    }
  }
  __syncthreads();
  if (SHORT) {
    for (uint32_t t = tid; t < len; t += kGroupThreads) ordg[s + t] = ord[t];
  }
  __syncthreads();   // before the next row clears the counters
}

__global__ void __launch_bounds__(kGroupThreads, 1) k_group_rows(const uint32_t* __restrict__ keys,
                                                                  const uint32_t* __restrict__ vals,
                                                                  const uint32_t* __restrict__ row_start,
                                                                  uint32_t* __restrict__ start, uint32_t* __restrict__ ordg, Grid g,
                                                                  uint32_t nrows, uint32_t n, uint32_t ordcap,
                                                                  unsigned long long* __restrict__ cursor) {
  extern __shared__ __align__(16) uint32_t cnt[];   // counters / cursors (u16 pairs or u32), then the order table of short rows
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t next_row;
  uint32_t* ord = cnt + ((uint32_t(g.bins) + 1) / 2 + 3 & ~3u);
  const int tid = threadIdx.x;
  const uint32_t lmask = (1u << g.lb) - 1u;
  if (tid == 0) next_row = uint32_t(min(atomicAdd(cursor, 1ull), (unsigned long long)nrows));
  __syncthreads();
  uint32_t row = next_row;
  while (row < nrows) {
    __syncthreads();   // everybody has read next_row
    if (tid == 0) next_row = uint32_t(min(atomicAdd(cursor, 1ull), (unsigned long long)nrows));
    const uint32_t s = row_start[row], e = row_start[row + 1];
    const uint32_t cx = row / uint32_t(g.nyc), ycb = row - cx * uint32_t(g.nyc);
    const int y0 = int(ycb << g.yb);
    const int ny_here = min(1 << g.yb, g.gy - y0);
    const uint32_t used = uint32_t(ny_here) * uint32_t(g.gz);
    const size_t cell0 = (size_t(cx) * g.gy + size_t(y0)) * g.gz;
    const bool last = row == nrows - 1;
    if (e - s <= ordcap)
      group_row<true>(keys, vals, start, ordg, cnt, ord, wsum, lmask, s, e, used, cell0, last, n);
    else
      group_row<false>(keys, vals, start, ordg, cnt, ord, wsum, lmask, s, e, used, cell0, last, n);
    row = next_row;   // written before the barriers inside group_row
  }
}

__global__ void k_fill_u32(uint32_t* a, int64_t n, uint32_t v) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
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
  const int *cx, *cy, *cz;     // cell of each node coordinate (clamped)
  // ring-1 tables, precomputed on the host: origin-relative f32 coordinates and, per axis, the distance from the
  // node to the nearest face of its ring-1 cell block behind which unexamined particles may exist (rounded down,
  // with the cell-assignment slack already taken off; +inf when the block reaches an open end of the grid)
  const float *fx, *fy, *fz;
  const float *mx, *my, *mz;
  // first cell of the node's 2-cell window per axis: [w, w+1] is the pair of cells whose common face is nearest
  // to the node (with the default grid the nodes sit exactly on cell corners, so the 2x2x2 block proves radius h)
  const int *wx, *wy, *wz;
  int nx, ny, nz;
};

// Stage A of the search, f32 prefilter, one thread per lattice node (consecutive lanes = consecutive z nodes).
// Candidates: the 2x2x2 cell block around the node (4 cell rows x 2 contiguous cells).  Distances are evaluated in f32
// on origin-relative coordinates and the two smallest are tracked.  A node is settled here only if (a) the runner-up
// is farther than the winner by more than a rigorous bound on the f32 error of both and (b) the winner (plus that
// bound) is strictly inside the proof margin of the block.  Unproven nodes go to stage B (4x4x4 block); proven but
// ambiguous ones (near ties, exact ties) go straight to the exact f64 kernel.
//
// f32 error bound: stored coordinate c~ = fl(p-o), query q~ = fl(q-o), d~x = fl(q~-c~):
//   |d~x - dx| <= 2^-24 (|q-o| + |p-o| + |dx|) <= 2^-23 (E + |dx|) =: eps      (E = grid extent)
//   |d~^2 - d^2| <= 2 sqrt(3) d eps + 3 eps^2 + 2^-22 d^2
struct Cand {
  float b1, b2;
  int bi;
};
__device__ __forceinline__ void scan_row(const rec_t* __restrict__ part, uint32_t s, uint32_t e, float qx, float qy, float qz,
                                         Cand& c) {
#pragma unroll 1
  for (uint32_t p = s; p < e; ++p) {
    const float4 q = __ldg(part + p);
    const float dx = qx - q.x, dy = qy - q.y, dz = qz - q.z;
    const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const bool lt = d < c.b1;
    c.b2 = fminf(c.b2, lt ? c.b1 : d);
    c.bi = lt ? int(p) : c.bi;
    c.b1 = lt ? d : c.b1;
  }
}
// 0: settled, 1: proven-or-not but ambiguous / empty -> exact, 2: unambiguous but unproven -> wider block
__device__ __forceinline__ int judge(const Cand& c, float extent, float margin, float* tol_out) {
  if (c.bi < 0) return 2;
  const float rb = sqrtf(c.b2 < INFINITY ? c.b2 : c.b1);
  const float eps = 1.5e-7f * (extent + rb);
  const float tol = 8.f * rb * eps + 8.f * eps * eps + 1e-6f * rb * rb;   // >= err(b1) + err(b2)
  *tol_out = tol;
  const bool proven = (margin == INFINITY) || (margin > 0.f && c.b1 + tol < margin * margin);
  if (!proven) return 2;
  return (c.b2 - c.b1 > tol) ? 0 : 1;
}

__global__ void __launch_bounds__(256) k_search_block2(const rec_t* __restrict__ part, const uint32_t* __restrict__ start, Grid g,
                                                        Lattice L, float extent, int32_t* __restrict__ nn,
                                                        int32_t* __restrict__ nn_pos, uint32_t* __restrict__ list_b,
                                                        uint32_t* __restrict__ list_c, vp_nn_stats_dev* __restrict__ stats) {
  // block = (z nodes, y rows); grid = (z chunks, y chunks, x)
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int i = blockIdx.z;
  if (k >= L.nz || j >= L.ny) return;
  const size_t node = (size_t(i) * L.ny + j) * L.nz + k;
  const float qx = __ldg(L.fx + i), qy = __ldg(L.fy + j), qz = __ldg(L.fz + k);
  const int wx = __ldg(L.wx + i), wy = __ldg(L.wy + j), wz = __ldg(L.wz + k);
  const int z1 = min(wz + 1, g.gz - 1);
  Cand c;
  c.b1 = INFINITY; c.b2 = INFINITY; c.bi = -1;
  uint32_t rs[4], re[4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int X = wx + a, Y = wy + b;
      const bool in = X < g.gx && Y < g.gy;
      const size_t row = in ? (size_t(X) * g.gy + Y) * g.gz : 0;
      rs[a * 2 + b] = in ? __ldg(start + row + wz) : 0u;
      re[a * 2 + b] = in ? __ldg(start + row + z1 + 1) : 0u;
    }
#pragma unroll
  for (int r = 0; r < 4; ++r) scan_row(part, rs[r], re[r], qx, qy, qz, c);
  const float m = fminf(__ldg(L.mx + i), fminf(__ldg(L.my + j), __ldg(L.mz + k)));
  float tol;
  const int verdict = judge(c, extent, m, &tol);
  if (verdict == 0) {
    if (nn) nn[node] = __float_as_int(__ldg(&part[c.bi].w));
    if (nn_pos) nn_pos[node] = c.bi;
  } else if (verdict == 1) {
    list_c[atomicAdd(&stats->n_wide, 1ull)] = uint32_t(node);
  } else {
    list_b[atomicAdd(&stats->pad, 1ull)] = uint32_t(node);
  }
}

// Stage B: the nodes stage A could not prove, one thread per listed node, the 32-cell union of the three 4x2x2 bars
// through the window (proves sqrt(2) h with corner-aligned nodes), same f32 prefilter and verdict.  What is still
// unproven or ambiguous goes to the exact kernel.
__global__ void __launch_bounds__(256) k_search_block4(const rec_t* __restrict__ part, const uint32_t* __restrict__ start, Grid g,
                                                        Lattice L, float extent, int32_t* __restrict__ nn,
                                                        int32_t* __restrict__ nn_pos, const uint32_t* __restrict__ list_b,
                                                        uint32_t* __restrict__ list_c, vp_nn_stats_dev* __restrict__ stats) {
  const unsigned long long nb = stats->pad;
  for (unsigned long long t = (unsigned long long)(blockIdx.x) * blockDim.x + threadIdx.x; t < nb;
       t += (unsigned long long)(gridDim.x) * blockDim.x) {
    const uint32_t node = list_b[t];
    const int k = int(node % uint32_t(L.nz));
    const uint32_t u = node / uint32_t(L.nz);
    const int j = int(u % uint32_t(L.ny)), i = int(u / uint32_t(L.ny));
    const float qx = L.fx[i], qy = L.fy[j], qz = L.fz[k];
    // examined region = union of the three 4x2x2 bars through the 2x2x2 window (32 cells instead of 64):
    //   central rows (x, y both inside the window): 4 cells along z;  rows one step outside in x OR y: 2 cells
    const int wx = L.wx[i], wy = L.wy[j], wz = L.wz[k];
    const int wx1 = min(wx + 1, g.gx - 1), wy1 = min(wy + 1, g.gy - 1), wz1 = min(wz + 1, g.gz - 1);
    const int x0 = max(wx - 1, 0), x1 = min(wx + 2, g.gx - 1);
    const int y0 = max(wy - 1, 0), y1 = min(wy + 2, g.gy - 1);
    const int z0 = max(wz - 1, 0), z1 = min(wz + 2, g.gz - 1);
    Cand c;
    c.b1 = INFINITY; c.b2 = INFINITY; c.bi = -1;
    for (int X = x0; X <= x1; ++X)
      for (int Y = y0; Y <= y1; ++Y) {
        const bool xin = X >= wx && X <= wx1, yin = Y >= wy && Y <= wy1;
        if (!xin && !yin) continue;                                  // corner rows are not part of the union
        const size_t row = (size_t(X) * g.gy + Y) * g.gz;
        const int a = (xin && yin) ? z0 : wz, b = (xin && yin) ? z1 : wz1;
        scan_row(part, __ldg(start + row + a), __ldg(start + row + b + 1), qx, qy, qz, c);
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
    float tol;
    if (judge(c, extent, m, &tol) == 0) {
      if (nn) nn[node] = __float_as_int(__ldg(&part[c.bi].w));
      if (nn_pos) nn_pos[node] = c.bi;
    } else {
      list_c[atomicAdd(&stats->n_wide, 1ull)] = node;
    }
  }
}

// Exact search, one warp per listed node, f64 arithmetic on the caller's coordinates.  The searched block starts
// as the node's 2x2x2 window and is widened (1, 3, 7, ... cells per side) until the proof holds (or every kept
// particle has been examined).
template <typename T>
__global__ void __launch_bounds__(256) k_search_exact(const rec_t* __restrict__ part, const uint32_t* __restrict__ start,
                                                       const T* __restrict__ pos, Grid g, Lattice L, int32_t* __restrict__ nn,
                                                       int32_t* __restrict__ nn_pos, const uint32_t* __restrict__ list,
                                                       vp_nn_stats_dev* __restrict__ stats) {
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
          const int id = __float_as_int(__ldg(&part[p].w));
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
            const int id = __float_as_int(__ldg(&part[p].w));
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
    }
  }
}

// fields from the SORTED payload: node -> sorted position of its nearest particle -> (v', m)
__global__ void __launch_bounds__(256) k_fields_sorted(const int32_t* __restrict__ nn_pos, int64_t n, const float4* __restrict__ spay,
                                                        float* vx, float* vy, float* vz, float* px, float* py, float* pz, float* e,
                                                        float* mo) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float4 w = __ldg(spay + nn_pos[t]);
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
  // key layout: the widest y chunk whose (y, z) tile of cells fits the shared-memory counters of k_group_permute;
  // the grid is coarsened until a z row fits the counters and (row << lb | local) fits 32 bits
  int yb = 0, lb = 0, nyc = 1;
  for (;;) {
    bool fits = gz <= kMaxBins;
    if (fits) {
      yb = 0;
      while ((int64_t(2) << yb) * gz <= kMaxBins && (1 << yb) < gy) ++yb;
      lb = vp_ceil_log2(uint64_t(gz) << yb);
      nyc = (gy + (1 << yb) - 1) >> yb;
      fits = (uint64_t(gx) * uint64_t(nyc)) <= (uint64_t(1) << (32 - lb)) && double(gx) * gy * gz < 4294967295.0;
    }
    if (fits) break;
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
  g.yb = yb; g.lb = lb; g.nyc = nyc; g.bins = gz << yb;
  return g;
}

struct NNScratch {
  size_t keys, packed, spos, start, rows, tail, total;
};
NNScratch nn_scratch(int64_t np, bool pay, uint64_t ncells, uint64_t nrows, int64_t nnodes) {
  NNScratch s;
  s.keys = vp_align256(size_t(np) * 4);
  s.packed = vp_align256(size_t(np) * (pay ? sizeof(rec32_t) : sizeof(float4)));
  s.spos = vp_align256(size_t(np) * sizeof(rec_t));
  s.start = vp_align256((ncells + 1) * 4);
  s.rows = vp_align256((nrows + 1) * 4);
  size_t b_wide = 2 * vp_align256(size_t(nnodes) * 4), b_sort = vp_sort_scratch_bytes(np);
  s.tail = b_sort > b_wide ? b_sort : b_wide;  // the sort scratch is dead once the records are permuted; the two node lists reuse it
  s.total = 2 * s.keys + s.packed + s.spos + s.start + s.rows + s.tail + 2048;
  return s;
}

// Optional payload travelling with the particles (whole-path use): sorted (v', m) records and the sorted position
// of every node's nearest particle, so that the field kernel reads the payload almost sequentially.
template <typename T>
struct NNPayload {
  const T* vel = nullptr;
  const T* rho = nullptr;
  double lcell3 = 1.0;
  float4* spay_out = nullptr;    // [np]
  int32_t* nn_pos_out = nullptr;  // [nnodes]
};

template <typename T>
int nn_grid_typed(vp_ctx* ctx, const T* pos, int64_t np, const double* qx, int nx, const double* qy, int ny,
                  const double* qz, int nz, int32_t* nn, const NNPayload<T>* pay, const vp_nn_opts* opts, cudaStream_t st,
                  const vp_host_chunks* host_pos = nullptr) {
  const int64_t nnodes = int64_t(nx) * ny * nz;
  VP_REQUIRE(nnodes < (int64_t(1) << 32), "vp_nn_grid: lattice too large for 32-bit node ids");
  VP_REQUIRE(np < (int64_t(1) << 31), "vp_nn_grid: np must be < 2^31 per device");
  const bool has_pay = pay != nullptr;
  if (has_pay) VP_REQUIRE(pay->vel && pay->spay_out && pay->nn_pos_out, "vp_nn_grid: incomplete payload description");
  vp_nn_opts o;
  memset(&o, 0, sizeof o);
  if (opts) o = *opts;
  const Grid g = plan_grid(np, qx, nx, qy, ny, qz, nz, o);
  const int gx = g.gx, gy = g.gy, gz = g.gz;
  const uint64_t ncells = uint64_t(gx) * gy * gz;
  const uint64_t nrows = uint64_t(gx) * g.nyc;
  const int row_bits = vp_ceil_log2(nrows);

  // ---- lattice tables (host -> pinned -> device)
  const size_t nt = size_t(nx) + ny + nz;
  const size_t off_c = vp_align256(nt * 8), off_f = off_c + vp_align256(nt * 4), off_m = off_f + vp_align256(nt * 4);
  const size_t off_w = off_m + vp_align256(nt * 4);
  const size_t tab_bytes = off_w + vp_align256(nt * 4);
  if (ctx->pinned_cap < tab_bytes) {
    if (ctx->pinned_h) cudaFreeHost(ctx->pinned_h);
    VP_CUDA(cudaMallocHost(&ctx->pinned_h, tab_bytes));
    ctx->pinned_cap = tab_bytes;
  }
  if (ctx->small_cap < tab_bytes) {
    VP_CUDA(cudaStreamSynchronize(st));
    if (ctx->small_d) cudaFree(ctx->small_d);
    VP_CUDA(cudaMalloc(&ctx->small_d, tab_bytes));
    ctx->small_cap = tab_bytes;
  }
  VP_CUDA(cudaStreamSynchronize(st));  // the pinned block may still be in flight from a previous call
  char* hb = static_cast<char*>(ctx->pinned_h);
  double* hq = reinterpret_cast<double*>(hb);
  int* hc = reinterpret_cast<int*>(hb + off_c);
  float* hf = reinterpret_cast<float*>(hb + off_f);
  float* hm = reinterpret_cast<float*>(hb + off_m);
  int* hw = reinterpret_cast<int*>(hb + off_w);
  auto fill_axis = [&](const double* q, int n, int at, double o_, double h_, double ih, int gg, bool closed_lo, bool closed_hi) {
    for (int i = 0; i < n; ++i) {
      double f = (q[i] - o_) * ih;
      int c = !(f > 0.0) ? 0 : (f >= double(gg) ? gg - 1 : int(f));
      hq[at + i] = q[i];
      hc[at + i] = c;
      hf[at + i] = float(q[i] - o_);
      // 2-cell window [w, w+1]: the neighbour on the side of the nearer face of cell c
      int w = (f - double(c) < 0.5) ? c - 1 : c;
      if (w > gg - 2) w = gg - 2;
      if (w < 0) w = 0;
      hw[at + i] = w;
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
  VP_CUDA(cudaMemcpyAsync(ctx->small_d, ctx->pinned_h, tab_bytes, cudaMemcpyHostToDevice, st));
  const char* db = reinterpret_cast<const char*>(ctx->small_d);
  Lattice L;
  L.qx = reinterpret_cast<const double*>(db); L.qy = L.qx + nx; L.qz = L.qy + ny;
  L.cx = reinterpret_cast<const int*>(db + off_c); L.cy = L.cx + nx; L.cz = L.cy + ny;
  L.fx = reinterpret_cast<const float*>(db + off_f); L.fy = L.fx + nx; L.fz = L.fy + ny;
  L.mx = reinterpret_cast<const float*>(db + off_m); L.my = L.mx + nx; L.mz = L.my + ny;
  L.wx = reinterpret_cast<const int*>(db + off_w); L.wy = L.wx + nx; L.wz = L.wy + ny;
  L.nx = nx; L.ny = ny; L.nz = nz;

  // ---- scratch
  const NNScratch sc = nn_scratch(np, has_pay, ncells, nrows, nnodes);
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, sc.total));
  uint32_t* keys = static_cast<uint32_t*>(vp_arena_alloc(ctx, sc.keys));
  uint32_t* vals = static_cast<uint32_t*>(vp_arena_alloc(ctx, sc.keys));
  void* packed = vp_arena_alloc(ctx, sc.packed);
  rec_t* spos = static_cast<rec_t*>(vp_arena_alloc(ctx, sc.spos));
  uint32_t* start = static_cast<uint32_t*>(vp_arena_alloc(ctx, sc.start));
  uint32_t* row_start = static_cast<uint32_t*>(vp_arena_alloc(ctx, sc.rows));
  void* scratch = vp_arena_alloc(ctx, sc.tail);
  VP_REQUIRE(keys && vals && packed && spos && start && row_start && scratch, "vp_nn_grid: arena carve failed");
  uint32_t* node_list = static_cast<uint32_t*>(scratch);                                      // -> exact kernel
  uint32_t* list_b = node_list + vp_align256(size_t(nnodes) * 4) / 4;                            // -> 4x4x4 stage

  VP_CUDA(cudaMemsetAsync(ctx->nn_stats_d, 0, sizeof(vp_nn_stats_dev), st));
  int64_t n = np;
  if (np > 0) {
    // read pos (+vel, rho), write key, slot and the packed record
    const double es = sizeof(T);
    vp_stage stage(ctx, "k1a_keygen_pack", st, 1, double(np) * (has_pay ? (3 + 3 + (pay->rho ? 1 : 0)) * es + 8.0 + 32.0 : 3 * es + 8.0 + 16.0));
    unsigned long long* kept_d = &ctx->nn_stats_d->n_kept;
    auto launch = [&](const T* p, const T* v, const T* r, int64_t n_c, int64_t i0) {
      PayloadIn<T> pin;
      pin.vel = v;
      pin.rho = r;
      pin.lcell3 = T(has_pay ? pay->lcell3 : 1.0);
      if (o.use_x_keep) {
        const unsigned nb = unsigned((n_c + 256 * 8 - 1) / (256 * 8));
        if (has_pay) k_keygen_pack<T, true, 8><<<nb, 256, 0, st>>>(p, pin, n_c, i0, g, keys, vals, packed, kept_d);
        else k_keygen_pack<T, false, 8><<<nb, 256, 0, st>>>(p, pin, n_c, i0, g, keys, vals, packed, kept_d);
      } else {
        const unsigned nb = unsigned((n_c + 255) / 256);
        if (has_pay) k_keygen_pack<T, true, 1><<<nb, 256, 0, st>>>(p, pin, n_c, i0, g, keys, vals, packed, kept_d);
        else k_keygen_pack<T, false, 1><<<nb, 256, 0, st>>>(p, pin, n_c, i0, g, keys, vals, packed, kept_d);
      }
    };
    if (host_pos) {
      // positions arrive from the host in chunks (copy stream); keys + records of chunk c are made while chunk c+1 moves
      VP_REQUIRE(!has_pay && o.row_stride == 0 && !o.use_x_keep, "vp_nn_grid: host position streaming is the plain compact form");
      VP_TRY(vp_host_streams(ctx));
      const int64_t chunk = host_pos->chunk;
      T* posd = const_cast<T*>(pos);
      VP_CUDA(cudaEventRecord(ctx->ev_used[0], st));   // order the copy stream after everything already queued on st
      VP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_used[0], 0));
      int c = 0;
      for (int64_t i0 = 0; i0 < np; i0 += chunk, ++c) {
        const int64_t n_c = np - i0 < chunk ? np - i0 : chunk;
        VP_CUDA(cudaMemcpyAsync(posd + 3 * i0, static_cast<const T*>(host_pos->pos_h) + 3 * i0, size_t(n_c) * 3 * sizeof(T), cudaMemcpyHostToDevice, ctx->copy_stream));
        VP_CUDA(cudaEventRecord(ctx->ev_h2d[c & 1], ctx->copy_stream));
        VP_CUDA(cudaStreamWaitEvent(st, ctx->ev_h2d[c & 1], 0));
        launch(posd + 3 * i0, nullptr, nullptr, n_c, i0);
      }
    } else {
      launch(pos, has_pay ? pay->vel : nullptr, has_pay ? pay->rho : nullptr, np, 0);
    }
    VP_CHECK_LAUNCH();
  }
  if (o.use_x_keep) {
    unsigned long long kept = 0;
    VP_CUDA(cudaMemcpyAsync(&kept, &ctx->nn_stats_d->n_kept, 8, cudaMemcpyDeviceToHost, st));
    VP_CUDA(cudaStreamSynchronize(st));  // documented: the filtered form syncs once
    n = int64_t(kept);
  } else {
    unsigned long long kept = (unsigned long long)np;
    VP_CUDA(cudaMemcpyAsync(&ctx->nn_stats_d->n_kept, &kept, 8, cudaMemcpyHostToDevice, st));
  }
  // rows brought together by the sort (stable LSD passes over the row bits only), cells inside a row by the group kernel
  VP_TRY(vp_sort_pairs_range(ctx, keys, vals, n, g.lb, row_bits, scratch, st));
  if (n > 0) {
    {
      vp_stage stage(ctx, "k1d_row_starts", st, 1, double(n) * 4.0 + double(nrows) * 4.0);
      k_row_starts<<<unsigned((n + 1023) / 1024), 256, 0, st>>>(keys, n, g.lb, uint32_t(nrows), row_start);
    }
    uint32_t* svals = static_cast<uint32_t*>(scratch);          // first alternate buffer of the finished sort
    {
      // key read twice (count, place), slot read and written, cell table written
      vp_stage stage(ctx, "k1c_group_rows", st, 1, double(n) * 16.0 + double(ncells) * 4.0);
      const size_t ord_off = (((size_t(g.bins) + 1) / 2 + 3) & ~size_t(3)) * 4;      // bytes of the packed 16-bit counters
      const uint32_t room = uint32_t((kGroupSmemMax - ord_off) / 4);
      const uint32_t ordcap = room < 65535u ? room : 65535u;    // a short row's cursors fit 16 bits
      size_t smem = ord_off + size_t(ordcap) * 4;
      if (smem < size_t(g.bins) * 4) smem = size_t(g.bins) * 4;   // long rows: 32-bit counters
      const unsigned nb = unsigned(nrows < uint64_t(ctx->sm_count) ? nrows : uint64_t(ctx->sm_count));
      unsigned long long* cursor = &ctx->nn_stats_d->row_cursor;   // zeroed with the stats block above
      VP_CUDA(cudaFuncSetAttribute(k_group_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, kGroupSmemMax));
      k_group_rows<<<nb, kGroupThreads, smem, st>>>(keys, vals, row_start, start, svals, g, uint32_t(nrows), uint32_t(n), ordcap,
                                                    cursor);
    }
    {
      // slot read, one packed record gathered, sorted records written
      vp_stage stage(ctx, "k1c_permute", st, 1, double(n) * (4.0 + (has_pay ? 64.0 : 32.0)));
      const unsigned nb = unsigned((n + 255) / 256);
      if (has_pay) k_permute<true><<<nb, 256, 0, st>>>(packed, svals, n, spos, pay->spay_out);
      else k_permute<false><<<nb, 256, 0, st>>>(packed, svals, n, spos, nullptr);
    }
  } else {
    k_fill_u32<<<unsigned((ncells + 1 + 255) / 256), 256, 0, st>>>(start, int64_t(ncells + 1), 0u);
  }
  VP_CHECK_LAUNCH();
  int32_t* nn_pos = has_pay ? pay->nn_pos_out : nullptr;
  const float extent = float(fmax(fmax(g.gx * g.hx, g.gy * g.hy), g.gz * g.hz));
  {
    // sorted records read once + cell starts read once + one index written per node
    vp_stage stage(ctx, "k1e_search_block2", st, 1, double(n) * sizeof(rec_t) + double(ncells) * 4.0 + double(nnodes) * 4.0);
    // block = (z nodes, y rows), one x plane per blockIdx.z: no integer division in the kernel
    int bx = nz >= 256 ? 256 : ((nz + 31) / 32) * 32;
    int by = 256 / bx;
    dim3 block(bx, by, 1), grid((nz + bx - 1) / bx, (ny + by - 1) / by, nx);
    VP_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "vp_nn_grid: lattice too large for the search launch");
    k_search_block2<<<grid, block, 0, st>>>(spos, start, g, L, extent, nn, nn_pos, list_b, node_list, ctx->nn_stats_d);
  }
  {
    vp_stage stage(ctx, "k1e_search_block4", st, 1);
    k_search_block4<<<ctx->sm_count * 8, 256, 0, st>>>(spos, start, g, L, extent, nn, nn_pos, list_b, node_list, ctx->nn_stats_d);
  }
  {
    vp_stage stage(ctx, "k1f_search_exact", st, 1);
    k_search_exact<T><<<ctx->sm_count * 8, 256, 0, st>>>(spos, start, pos, g, L, nn, nn_pos, node_list, ctx->nn_stats_d);
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
  pay.spay_out = reinterpret_cast<float4*>(spay);
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

// rows = [x y z vx vy vz (rho)] ; cursors[d] starts at the first row this rank may write in destination d's buffer
// (exclusive prefix of counts for the single local buffer; with peer stores, the rows of lower ranks for d).
struct SlabDest {
  void* base[16];   // all equal for the local form; peer-mapped receive buffers for the fused exchange
};
// 512 particles per block.  Every block claims one contiguous row range per destination (one global atomic each),
// stages its rows in shared memory grouped by destination and writes each group out as one coalesced run -- 128-byte
// store instructions whether the destination is local HBM or a peer's buffer over NVLink.
constexpr int kBktItems = 2;
// staged rows per block (a particle inside a halo goes to two ranks); overflow -> direct stores
template <typename T> struct BktCap { static constexpr int v = sizeof(T) == 4 ? 1024 : 768; };
template <typename T>
__global__ void __launch_bounds__(256) k_bucket_scatter(const T* __restrict__ pos, const T* __restrict__ vel, const T* __restrict__ rho,
                                                         int64_t np, SlabRanges R, unsigned long long* __restrict__ cursors,
                                                         SlabDest D, int w) {
  __shared__ unsigned sc[16];             // rows of this block per destination
  __shared__ unsigned pre[17];            // exclusive prefix of sc
  __shared__ unsigned long long sbase[16];
  constexpr int kBktCap = BktCap<T>::v;
  __shared__ T stage[kBktCap * 7];
  if (threadIdx.x < 16) sc[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t i0 = int64_t(blockIdx.x) * (256 * kBktItems);
  unsigned myoff[kBktItems][16];
  bool okv[kBktItems];
#pragma unroll
  for (int r = 0; r < kBktItems; ++r) {
    const int64_t i = i0 + r * 256 + threadIdx.x;
    okv[r] = i < np;
    const double x = okv[r] ? double(pos[3 * i]) : 0.0;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      myoff[r][d] = 0xffffffffu;
      if (d < R.n) {
        const bool in = okv[r] && x >= R.lo[d] && x <= R.hi[d];
        const unsigned m = __ballot_sync(0xffffffffu, in);
        unsigned wb = 0;
        if (lane == 0 && m) wb = atomicAdd(&sc[d], __popc(m));
        wb = __shfl_sync(0xffffffffu, wb, 0);
        if (in) myoff[r][d] = wb + __popc(m & ((1u << lane) - 1u));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < R.n) sbase[threadIdx.x] = sc[threadIdx.x] ? atomicAdd(cursors + threadIdx.x, (unsigned long long)sc[threadIdx.x]) : 0ull;
  if (threadIdx.x == 0) {
    unsigned a = 0;
    for (int d = 0; d < R.n; ++d) { pre[d] = a; a += sc[d]; }
    for (int d = R.n; d <= 16; ++d) pre[d] = a;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kBktItems; ++r) {
    if (!okv[r]) continue;
    const int64_t i = i0 + r * 256 + threadIdx.x;
    T row[7];
    row[0] = pos[3 * i]; row[1] = pos[3 * i + 1]; row[2] = pos[3 * i + 2];
    row[3] = vel[3 * i]; row[4] = vel[3 * i + 1]; row[5] = vel[3 * i + 2];
    row[6] = rho ? rho[i] : T(0);
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      if (d < R.n && myoff[r][d] != 0xffffffffu) {
        const unsigned p = pre[d] + myoff[r][d];
        if (p < unsigned(kBktCap)) {
          for (int c = 0; c < w; ++c) stage[p * w + c] = row[c];
        } else {   // staging area full (very wide halos): store directly
          T* o = static_cast<T*>(D.base[d]) + (sbase[d] + myoff[r][d]) * size_t(w);
          for (int c = 0; c < w; ++c) o[c] = row[c];
        }
      }
    }
  }
  __syncthreads();
  const unsigned tot = min(pre[16], unsigned(kBktCap));
  int d = 0;
  for (unsigned f = threadIdx.x; f < tot * unsigned(w); f += 256) {
    const unsigned rr = f / unsigned(w), c = f - rr * unsigned(w);
    while (rr >= pre[d + 1]) ++d;                  // f only grows: the destination index is monotone per thread
    static_cast<T*>(D.base[d])[(sbase[d] + (rr - pre[d])) * size_t(w) + c] = stage[f];
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
  const int w = rho ? 7 : 6;
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
    k_bucket_scatter<T><<<unsigned((np + 256 * kBktItems - 1) / (256 * kBktItems)), 256, 0, st>>>(pos, vel, rho, np, R, cur, D, w);
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
  const int w = rho ? 7 : 6;
  if (np > 0) {
    vp_stage stage(ctx, "k0_slab_bucket_scatter", st, 1, double(np) * 2.0 * w * sizeof(T));
    SlabDest D;
    for (int d = 0; d < 16; ++d) D.base[d] = d < R.n ? ctx->slab_peer[d] : nullptr;
    k_bucket_scatter<T><<<unsigned((np + 256 * kBktItems - 1) / (256 * kBktItems)), 256, 0, st>>>(pos, vel, rho, np, R, cur, D, w);
    VP_CHECK_LAUNCH();
  }
  VP_CUDA(cudaStreamSynchronize(st));   // cur lives in the scope released on return
  return VP_OK;
}

}  // namespace

// ---- sharded particle exchange fused into the bucketing kernel (peer stores over NVLink)
extern "C" int vp_slab_p2p_alloc(vp_ctx* ctx, size_t bytes, unsigned char* handle_out) {
  VP_REQUIRE(ctx && handle_out && bytes > 0, "vp_slab_p2p_alloc: bad argument");
  VP_CUDA(cudaSetDevice(ctx->device));
  VP_CUDA(cudaDeviceSynchronize());
  if (ctx->slab_open) {
    for (int d = 0; d < ctx->slab_nranks; ++d)
      if (d != ctx->slab_rank && ctx->slab_peer[d]) { cudaIpcCloseMemHandle(ctx->slab_peer[d]); ctx->slab_peer[d] = nullptr; }
    ctx->slab_open = false;
  }
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
  return nn_scratch(np, true, uint64_t(g.gx) * g.gy * g.gz, uint64_t(g.gx) * g.nyc, int64_t(nx) * ny * nz).total + 4096;
}

extern "C" int vp_nn_grid_payload(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype, int64_t np,
                                  const double* qx_h, int nx, const double* qy_h, int ny, const double* qz_h, int nz,
                                  double lcell3, int32_t* nn_idx_d, int32_t* nn_pos_d, float* spay_d, const vp_nn_opts* opts,
                                  void* stream) {
  VP_REQUIRE(ctx && pos_d && vel_d && qx_h && qy_h && qz_h && nn_pos_d && spay_d, "vp_nn_grid_payload: null argument");
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

extern "C" int vp_fields_sorted(vp_ctx* ctx, const int32_t* nn_pos_d, int64_t n_nodes, const float* spay_d, float* const v_d[3],
                                float* const p_d[3], float* e_d, float* m_d, void* stream) {
  VP_REQUIRE(ctx && nn_pos_d && spay_d, "vp_fields_sorted: null argument");
  if (n_nodes == 0) return VP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* v[3] = {v_d ? v_d[0] : nullptr, v_d ? v_d[1] : nullptr, v_d ? v_d[2] : nullptr};
  float* p[3] = {p_d ? p_d[0] : nullptr, p_d ? p_d[1] : nullptr, p_d ? p_d[2] : nullptr};
  int nplanes = (e_d != nullptr) + (m_d != nullptr);
  for (int c = 0; c < 3; ++c) nplanes += (v[c] != nullptr) + (p[c] != nullptr);
  // per node: sorted position read, 16-byte payload record read, 4 B written per plane
  vp_stage stage(ctx, "k3_fields_sorted", st, 1, double(n_nodes) * (4.0 + 16.0 + 4.0 * nplanes));
  k_fields_sorted<<<unsigned((n_nodes + 255) / 256), 256, 0, st>>>(nn_pos_d, n_nodes, reinterpret_cast<const float4*>(spay_d), v[0],
                                                                   v[1], v[2], p[0], p[1], p[2], e_d, m_d);
  VP_CHECK_LAUNCH();
  return VP_OK;
}

extern "C" int vp_nn_grid(vp_ctx* ctx, const void* pos_d, int pos_dtype, int64_t np, const double* qx_h, int nx,
                          const double* qy_h, int ny, const double* qz_h, int nz, int32_t* nn_idx_d,
                          const vp_nn_opts* opts, void* stream) {
  VP_REQUIRE(ctx && pos_d && qx_h && qy_h && qz_h && nn_idx_d, "vp_nn_grid: null argument");
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
  const uint64_t ncells = uint64_t(g.gx) * g.gy * g.gz, nrows = uint64_t(g.gx) * g.nyc;
  info_out[0] = g.gx; info_out[1] = g.gy; info_out[2] = g.gz;
  info_out[3] = g.yb; info_out[4] = g.lb; info_out[5] = g.nyc; info_out[6] = g.bins;
  info_out[7] = vp_ceil_log2(nrows);
  info_out[8] = int64_t((nn_scratch(np, true, ncells, nrows, int64_t(nx) * ny * nz).total + (size_t(1) << 20) - 1) >> 20);
  // corner aligned: cells of the node spacing with every node on a cell corner (the 2x2x2 block then proves radius h)
  const double sp = nx > 1 ? (qx_h[nx - 1] - qx_h[0]) / (nx - 1) : 0.0;
  info_out[9] = (g.gx == nx + 1 && nx > 1 && fabs(g.hx - sp) <= 1e-9 * fabs(sp)) ? 1 : 0;
  return VP_OK;
}

extern "C" int vp_nn_grid_stats(vp_ctx* ctx, int64_t* n_wide, int64_t* n_unresolved, int64_t* n_kept, void* stream) {
  VP_REQUIRE(ctx, "vp_nn_grid_stats: null ctx");
  vp_nn_stats_dev h;
  VP_CUDA(cudaMemcpyAsync(&h, ctx->nn_stats_d, sizeof h, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  VP_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  if (n_wide) *n_wide = int64_t(h.n_wide);
  if (n_unresolved) *n_unresolved = int64_t(h.n_unresolved);
  if (n_kept) *n_kept = int64_t(h.n_kept);
  return VP_OK;
}

extern "C" int vp_gather_rows(vp_ctx* ctx, const int32_t* idx_d, int64_t n, const void* src_d, int row_bytes, void* dst_d,
                              void* stream) {
  VP_REQUIRE(ctx && idx_d && src_d && dst_d, "vp_gather_rows: null argument");
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
