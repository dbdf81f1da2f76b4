// LSD radix sort of (key,value) u32 pairs on a bit range, 7 or 8 bits per pass -- builds the cell list of K1.
//
// Per pass:  (1) tile digit histograms  (2) exclusive scan, digit-major  (3) stable scatter with
// warp-level match ranking and a shared-memory staged, digit-run coalesced write.
// Algorithmic traffic per pass: 4 B (histogram read) + 16 B (pair read + write) per element.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;  // 4096 keys per CTA
constexpr int kRadix = 256;   // widest digit; the histogram scratch is sized for it
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan across a 256-thread block; returns the exclusive prefix, *total gets the block sum
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* warp_sums /*[8]*/, uint32_t* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < kWarps; ++i) {
    uint32_t s = warp_sums[i];
    if (i < w) woff += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return woff + incl - v;
}

// (1) per-tile digit histogram, written digit-major: hist[d * nblocks + b]
template <int BITS>
__global__ void __launch_bounds__(kThreads) k_tile_hist(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                         uint32_t* __restrict__ hist, int nblocks) {
  constexpr int R = 1 << BITS;
  __shared__ uint32_t sh[R];
  if (threadIdx.x < R) sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = int64_t(blockIdx.x) * kTile;
  if (base + kTile <= n && (reinterpret_cast<uintptr_t>(keys) & 15) == 0) {
    // full tile of a 16-byte aligned array: four 16-byte loads per thread
    const uint4* k4 = reinterpret_cast<const uint4*>(keys + base);
    uint4 v[kItems / 4];
#pragma unroll
    for (int r = 0; r < kItems / 4; ++r) v[r] = k4[r * kThreads + threadIdx.x];
#pragma unroll
    for (int r = 0; r < kItems / 4; ++r) {
      atomicAdd(&sh[(v[r].x >> shift) & (R - 1)], 1u);
      atomicAdd(&sh[(v[r].y >> shift) & (R - 1)], 1u);
      atomicAdd(&sh[(v[r].z >> shift) & (R - 1)], 1u);
      atomicAdd(&sh[(v[r].w >> shift) & (R - 1)], 1u);
    }
  } else {
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
      int64_t i = base + r * kThreads + threadIdx.x;
      if (i < n) atomicAdd(&sh[(keys[i] >> shift) & (R - 1)], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < R) hist[size_t(threadIdx.x) * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// (2) exclusive scan of m u32 values in three kernels (chunk sums, scan of sums, apply)
constexpr int kScanChunk = kThreads * 16;

__global__ void __launch_bounds__(kThreads) k_scan_sums(const uint32_t* __restrict__ a, int64_t m,
                                                         uint32_t* __restrict__ sums) {
  __shared__ uint32_t ws[kWarps];
  const int64_t base = int64_t(blockIdx.x) * kScanChunk;
  uint32_t s = 0;
  if (base + kScanChunk <= m && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
    const uint4* a4 = reinterpret_cast<const uint4*>(a + base);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint4 v = a4[r * kThreads + threadIdx.x];
      s += (v.x + v.y) + (v.z + v.w);
    }
  } else {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      int64_t i = base + r * kThreads + threadIdx.x;
      if (i < m) s += a[i];
    }
  }
  uint32_t tot;
  block_excl_scan_256(s, ws, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kThreads) k_scan_top(uint32_t* __restrict__ sums, int nchunks) {
  __shared__ uint32_t ws[kWarps];
  uint32_t carry = 0;
  int base = 0;
  // 16 consecutive sums per thread and round (four 16-byte accesses): 2^18 chunk sums of a 2^30-cell table take 64 rounds
  // instead of 1024 (0.6 ms -> 0.05 ms; one CTA, every round is two barriers)
  if ((reinterpret_cast<uintptr_t>(sums) & 15) == 0) {
    for (; base + kThreads * 16 <= nchunks; base += kThreads * 16) {
      uint4* p = reinterpret_cast<uint4*>(sums + base + threadIdx.x * 16);
      uint4 q[4];
      uint32_t s = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) { q[r] = p[r]; s += (q[r].x + q[r].y) + (q[r].z + q[r].w); }
      uint32_t tot;
      uint32_t ex = carry + block_excl_scan_256(s, ws, &tot);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint4 v = q[r];
        q[r] = make_uint4(ex, ex + v.x, ex + v.x + v.y, ex + v.x + v.y + v.z);
        ex += (v.x + v.y) + (v.z + v.w);
        p[r] = q[r];
      }
      carry += tot;
    }
  }
  for (; base < nchunks; base += kThreads) {
    int i = base + threadIdx.x;
    uint32_t v = i < nchunks ? sums[i] : 0u, tot;
    uint32_t ex = block_excl_scan_256(v, ws, &tot);
    if (i < nchunks) sums[i] = carry + ex;
    carry += tot;
  }
}

// chunk = 4 sub-blocks of 1024 elements; inside a sub-block thread t owns elements 4t .. 4t+3 (one 16-byte access)
__global__ void __launch_bounds__(kThreads) k_scan_apply(uint32_t* __restrict__ a, int64_t m,
                                                          const uint32_t* __restrict__ sums) {
  __shared__ uint32_t ws[kWarps];
  const int64_t cbase = int64_t(blockIdx.x) * kScanChunk;
  uint32_t carry = sums[blockIdx.x];
  const bool fast = cbase + kScanChunk <= m && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
#pragma unroll 1
  for (int sb = 0; sb < 4; ++sb) {
    const int64_t base = cbase + sb * (kThreads * 4) + int64_t(threadIdx.x) * 4;
    uint32_t v[4];
    if (fast) {
      const uint4 q = *reinterpret_cast<const uint4*>(a + base);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) v[r] = (base + r < m) ? a[base + r] : 0u;
    }
    const uint32_t s = (v[0] + v[1]) + (v[2] + v[3]);
    uint32_t tot;
    uint32_t ex = block_excl_scan_256(s, ws, &tot) + carry;
    carry += tot;
    uint32_t o[4];
    o[0] = ex; o[1] = ex + v[0]; o[2] = o[1] + v[1]; o[3] = o[2] + v[2];
    if (fast) {
      *reinterpret_cast<uint4*>(a + base) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (base + r < m) a[base + r] = o[r];
    }
  }
}

// (3) stable scatter
template <int BITS>
__global__ void __launch_bounds__(kThreads, 4) k_scatter(const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                       uint32_t* __restrict__ kout, uint32_t* __restrict__ vout,
                                                       int64_t n, int shift, const uint32_t* __restrict__ hist,
                                                       int nblocks) {
  constexpr int R = 1 << BITS;
  __shared__ uint32_t wcnt[kWarps][R];
  __shared__ uint32_t dbase[R];   // exclusive prefix of digit totals inside this tile
  __shared__ uint32_t gbase[R];   // global output offset of the digit run minus dbase
  __shared__ uint32_t skey[kTile];
  __shared__ uint32_t sval[kTile];
  __shared__ uint32_t ws[kWarps];

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t tile0 = int64_t(blockIdx.x) * kTile;
  const int64_t rem = n - tile0;
  const int count = rem < kTile ? int(rem) : kTile;

  if (tid < R) {
#pragma unroll
    for (int i = 0; i < kWarps; ++i) wcnt[i][tid] = 0;
  }
  __syncthreads();

  uint32_t key[kItems], val[kItems], rank[kItems];
  const int wbase = w * (kItems * 32);
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    int li = wbase + r * 32 + lane;
    bool ok = li < count;
    key[r] = ok ? kin[tile0 + li] : 0u;
    val[r] = ok ? vin[tile0 + li] : 0u;
  }
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    int li = wbase + r * 32 + lane;
    bool ok = li < count;
    uint32_t d = ok ? ((key[r] >> shift) & (R - 1)) : uint32_t(R);
    // lanes holding the same digit: one ballot per digit bit (MATCH.ANY serialises over distinct values)
    uint32_t peers = __ballot_sync(0xffffffffu, ok);
    if (!ok) peers = ~peers;
#pragma unroll
    for (int bit = 0; bit < BITS; ++bit) {
      const uint32_t m = __ballot_sync(0xffffffffu, (d >> bit) & 1u);
      peers &= ((d >> bit) & 1u) ? m : ~m;
    }
    int leader = __ffs(peers) - 1;
    uint32_t before = __popc(peers & ((1u << lane) - 1u));
    uint32_t old = 0;
    if (ok && lane == leader) {
      old = wcnt[w][d];
      wcnt[w][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + before;
    __syncwarp();
  }
  __syncthreads();

  // digit `tid`: exclusive offsets across warps, then across digits
  uint32_t tot = 0;
  if (tid < R) {
#pragma unroll
    for (int i = 0; i < kWarps; ++i) {
      uint32_t c = wcnt[i][tid];
      wcnt[i][tid] = tot;
      tot += c;
    }
  }
  uint32_t blocktot;
  uint32_t ex = block_excl_scan_256(tot, ws, &blocktot);
  if (tid < R) {
    dbase[tid] = ex;
    gbase[tid] = hist[size_t(tid) * nblocks + blockIdx.x] - ex;
  }
  __syncthreads();

#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    int li = wbase + r * 32 + lane;
    if (li < count) {
      uint32_t d = (key[r] >> shift) & (R - 1);
      uint32_t p = dbase[d] + wcnt[w][d] + rank[r];
      skey[p] = key[r];
      sval[p] = val[r];
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    int i = r * kThreads + tid;
    if (i < count) {
      uint32_t k = skey[i];
      uint32_t d = (k >> shift) & (R - 1);
      size_t o = size_t(gbase[d]) + i;
      kout[o] = k;
      vout[o] = sval[i];
    }
  }
}

}  // namespace

int vp_scan_exclusive_u32(vp_ctx* ctx, uint32_t* a, int64_t m, uint32_t* sums, cudaStream_t st, const char* stage_name) {
  if (m <= 0) return VP_OK;
  const int64_t nchunks64 = (m + kScanChunk - 1) / kScanChunk;
  VP_REQUIRE(nchunks64 < (int64_t(1) << 31), "vp_scan: too many elements");
  const int nchunks = int(nchunks64);
  vp_stage stage(ctx, stage_name, st, 3, double(m) * 12.0);   // read twice, written once
  k_scan_sums<<<nchunks, kThreads, 0, st>>>(a, m, sums);
  k_scan_top<<<1, kThreads, 0, st>>>(sums, nchunks);
  k_scan_apply<<<nchunks, kThreads, 0, st>>>(a, m, sums);
  VP_CHECK_LAUNCH();
  return VP_OK;
}

static inline int64_t sort_nblocks(int64_t n) { return (n + kTile - 1) / kTile; }

size_t vp_sort_scratch_bytes(int64_t n) {
  int64_t nb = sort_nblocks(n < 1 ? 1 : n);
  int64_t m = nb * kRadix;
  int64_t nchunks = (m + kScanChunk - 1) / kScanChunk;
  return vp_align256(size_t(n) * 4) * 2      // alternate key / value buffers
         + vp_align256(size_t(m) * 4)         // tile histograms
         + vp_align256(size_t(nchunks) * 4);  // scan partials
}

template <int BITS>
static void sort_pass(const uint32_t* ka, const uint32_t* va, uint32_t* kb, uint32_t* vb, int64_t n, int shift, uint32_t* hist,
                      uint32_t* sums, int nb, cudaStream_t st) {
  const int64_t m = int64_t(nb) << BITS;
  const int nchunks = int((m + kScanChunk - 1) / kScanChunk);
  k_tile_hist<BITS><<<nb, kThreads, 0, st>>>(ka, n, shift, hist, nb);
  k_scan_sums<<<nchunks, kThreads, 0, st>>>(hist, m, sums);
  k_scan_top<<<1, kThreads, 0, st>>>(sums, nchunks);
  k_scan_apply<<<nchunks, kThreads, 0, st>>>(hist, m, sums);
  k_scatter<BITS><<<nb, kThreads, 0, st>>>(ka, va, kb, vb, n, shift, hist, nb);
}

// Stable sort on key bits [lo, lo + nbits); the other bits travel with the key untouched.
int vp_sort_pairs_range(vp_ctx* ctx, uint32_t* keys, uint32_t* vals, int64_t n, int lo, int nbits, void* scratch,
                        cudaStream_t st) {
  if (n <= 1 || nbits <= 0) return VP_OK;
  VP_REQUIRE(n < (int64_t(1) << 32), "vp_sort_pairs: n=%lld exceeds 2^32", (long long)n);
  VP_REQUIRE(lo >= 0 && lo + nbits <= 32, "vp_sort_pairs: bit range [%d,%d) outside the key", lo, lo + nbits);
  const int nb = int(sort_nblocks(n));
  const int64_t m = int64_t(nb) * kRadix;
  char* p = static_cast<char*>(scratch);
  uint32_t* k2 = reinterpret_cast<uint32_t*>(p);
  p += vp_align256(size_t(n) * 4);
  uint32_t* v2 = reinterpret_cast<uint32_t*>(p);
  p += vp_align256(size_t(n) * 4);
  uint32_t* hist = reinterpret_cast<uint32_t*>(p);
  p += vp_align256(size_t(m) * 4);
  uint32_t* sums = reinterpret_cast<uint32_t*>(p);

  uint32_t *ka = keys, *va = vals, *kb = k2, *vb = v2;
  const int passes = (nbits + 7) / 8;
  const int digit = passes * 7 >= nbits ? 7 : 8;   // 7-bit digits when they cover the range in as many passes (one ballot less per key)
  // algorithmic bytes: 4 (histogram read) + 16 (pair read + write) per element per pass
  vp_stage stage(ctx, "k1b_radix_sort", st, 5 * passes, double(n) * 20.0 * passes);
  for (int pass = 0; pass < passes; ++pass) {
    // a digit reaching past the range (or past bit 31) only re-sorts bits a later pass / nothing overrides: harmless for LSD
    int shift = lo + pass * digit;
    if (shift + digit > 32) shift = 32 - digit;
    if (digit == 7) sort_pass<7>(ka, va, kb, vb, n, shift, hist, sums, nb, st);
    else sort_pass<8>(ka, va, kb, vb, n, shift, hist, sums, nb, st);
    VP_CHECK_LAUNCH();
    uint32_t* t;
    t = ka; ka = kb; kb = t;
    t = va; va = vb; vb = t;
  }
  if (ka != keys) {
    VP_CUDA(cudaMemcpyAsync(keys, ka, size_t(n) * 4, cudaMemcpyDeviceToDevice, st));
    VP_CUDA(cudaMemcpyAsync(vals, va, size_t(n) * 4, cudaMemcpyDeviceToDevice, st));
  }
  return VP_OK;
}

int vp_sort_pairs_impl(vp_ctx* ctx, uint32_t* keys, uint32_t* vals, int64_t n, int bits, void* scratch,
                       cudaStream_t st) {
  return vp_sort_pairs_range(ctx, keys, vals, n, 0, bits, scratch, st);
}
