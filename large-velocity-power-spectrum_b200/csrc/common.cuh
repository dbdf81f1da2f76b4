// Internal helpers shared by the kernels of libvpower_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/vpower_b200.h"

void vp_set_error(const char* fmt, ...);

#define VP_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      vp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return VP_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define VP_CHECK_LAUNCH() VP_CUDA(cudaGetLastError())

#define VP_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      vp_set_error(__VA_ARGS__);   \
      return VP_ERR_ARG;           \
    }                              \
  } while (0)

#define VP_TRY(expr)           \
  do {                         \
    int r__ = (expr);          \
    if (r__ != VP_OK) return r__; \
  } while (0)

// Grow-only bump arena.  alloc() never frees; reset() rewinds.  If a request does not fit the arena
// is re-allocated (after a device sync) -- callers take all their pointers after the last grow by
// following the pattern  reserve(total) ; alloc() ; alloc() ...
struct vp_arena {
  char* base = nullptr;
  size_t cap = 0;
  size_t off = 0;
};

struct vp_nn_stats_dev {
  unsigned long long n_wide;        // nodes sent to the exact stage
  unsigned long long n_unresolved;  // nodes still unproven afterwards
  unsigned long long n_kept;        // particles that survived the x filter
  unsigned long long n_b;           // nodes sent to the wider prefilter stage
  unsigned long long n_far;         // particles outside the cell grid (clamped into end cells)
  unsigned long long n_crowded;     // bricks the particle-centric search handed to the node-centric kernel
  unsigned long long crowded_cursor;  // work cursor of that kernel
};

// optional per-stage timing with CUDA events on the launching stream (vp_profile_enable / vp_profile_report)
struct vp_prof_rec {
  const char* name;
  cudaEvent_t a, b;
  int launches;
  double bytes;  // algorithmic bytes of the stage (0 if not stated)
};

struct vp_pk_plan;
struct vp_ctx {
  // calls on one ctx are serialised by `mu` (threads), and a call on a different stream than the previous one first
  // waits on the device for `ev_last`, the end of the previous call's work: the arena, the small tables and the stats
  // block are shared by all calls of the ctx
  std::recursive_mutex mu;
  int call_depth = 0;
  cudaEvent_t ev_last = nullptr;
  cudaStream_t last_stream = nullptr;
  bool last_valid = false;
  bool prof_on = false;
  std::vector<vp_prof_rec> prof;
  unsigned long long n_launch = 0;        // kernels launched by this ctx since creation
  int device = 0;
  vp_pk_plan* cached_plan = nullptr;      // last plan built by the host-buffer entry point
  double* cached_plan_key = nullptr;      // host copy of (N, nbins, k table, edges) it was built for
  size_t cached_plan_key_len = 0;
  int sm_count = 148;
  vp_arena arena;
  vp_nn_stats_dev* nn_stats_d = nullptr;  // persistent small device block
  double* small_d = nullptr;              // persistent device block for lattice tables etc. (grown on demand)
  size_t small_cap = 0;
  void* pinned_h = nullptr;               // pinned staging for small host->device tables
  size_t pinned_cap = 0;
  cudaEvent_t ev_tables = nullptr;        // upload of the pinned tables (the next call waits for it before rewriting them)
  // sharded particle exchange over peer memory (vp_slab_p2p_*): this rank's receive buffer and every peer's, mapped
  void* slab_recv = nullptr;
  size_t slab_recv_bytes = 0;
  void* slab_peer[16] = {nullptr};
  int slab_nranks = 0, slab_rank = 0;
  bool slab_open = false;
  cudaStream_t copy_stream = nullptr;     // host-buffer entry point: H2D chunks overlap the keygen/pack kernel
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_used[2] = {nullptr, nullptr};
  cudaStream_t pack_stream = nullptr;     // payload chunks are packed here while the gridding runs on the caller's stream
  cudaEvent_t ev_pack = nullptr;
};

// Every compute entry point opens one of these first (see vp_ctx::mu).
struct vp_call_guard {
  vp_ctx* c;
  cudaStream_t st;
  vp_call_guard(vp_ctx* ctx, cudaStream_t s) : c(ctx), st(s) {
    c->mu.lock();
    if (c->call_depth++ == 0 && c->last_valid && c->last_stream != st) cudaStreamWaitEvent(st, c->ev_last, 0);
  }
  ~vp_call_guard() {
    if (--c->call_depth == 0) {
      if (!c->ev_last) cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming);
      if (c->ev_last && cudaEventRecord(c->ev_last, st) == cudaSuccess) { c->last_stream = st; c->last_valid = true; }
    }
    c->mu.unlock();
  }
  vp_call_guard(const vp_call_guard&) = delete;
  vp_call_guard& operator=(const vp_call_guard&) = delete;
};

// Stack discipline: every entry point opens a vp_arena_scope (restores the offset on exit), calls
// vp_arena_reserve(extra) once for everything it will carve, then vp_arena_alloc().  The block can only
// be re-allocated while nothing is carved from it (offset 0); an outer caller that holds buffers across
// inner calls therefore reserves the inner calls' needs up front (see pipeline.cu).
int vp_arena_reserve(vp_ctx* ctx, size_t extra_bytes);
void* vp_arena_alloc(vp_ctx* ctx, size_t bytes);  // 256-byte aligned; nullptr if it does not fit
static inline size_t vp_align256(size_t b) { return (b + 255) & ~size_t(255); }
struct vp_arena_scope {
  vp_ctx* c;
  size_t mark;
  explicit vp_arena_scope(vp_ctx* ctx) : c(ctx), mark(ctx->arena.off) {}
  ~vp_arena_scope() { c->arena.off = mark; }
};

// radix sort (radix_sort.cu).  Sorts in place (result ends in keys/vals); tmp must hold the scratch
// reported by vp_sort_scratch_bytes(n).
size_t vp_sort_scratch_bytes(int64_t n);
int vp_sort_pairs_impl(vp_ctx* ctx, uint32_t* keys, uint32_t* vals, int64_t n, int bits, void* scratch,
                       cudaStream_t st);
// stable sort on key bits [lo, lo + nbits) only
int vp_sort_pairs_range(vp_ctx* ctx, uint32_t* keys, uint32_t* vals, int64_t n, int lo, int nbits, void* scratch,
                        cudaStream_t st);

// exclusive prefix sum of m u32 values in place (three kernels); `sums` holds ceil(m/4096)+1 words of scratch
int vp_scan_exclusive_u32(vp_ctx* ctx, uint32_t* a, int64_t m, uint32_t* sums, cudaStream_t st, const char* stage_name);
// planes from (v', m) records addressed through nn_pos: record i at srec + (stride*i + offset) float4
int vp_fields_from_records(vp_ctx* ctx, const int32_t* nn_pos_d, int64_t n_nodes, const float* srec_d, int stride, int offset,
                           float* const v_d[3], float* const p_d[3], float* e_d, float* m_d, cudaStream_t st);

// Host-resident particle arrays streamed to the device in chunks (vp_host_particles_to_pk): positions land in a
// resident device array (the exact search needs them), velocity/density chunks only pass through two staging buffers.
struct vp_host_chunks {
  const void* pos_h = nullptr;
  const void* vel_h = nullptr;
  const void* rho_h = nullptr;
  int64_t chunk = 0;     // particles per chunk
};
size_t vp_host_chunk_staging_bytes(int64_t chunk, int dtype, bool has_rho);
// Positions only: pos_h chunks -> resident pos_d on the copy stream, keys/records made chunk by chunk behind them, then the
// whole gridding on `st`; nn_idx_d [nx,ny,nz] = ORIGINAL particle index.  Everything is enqueued, nothing is waited for.
int vp_nn_grid_host_pos(vp_ctx* ctx, const vp_host_chunks* hc, void* pos_d, int dtype, int64_t np, const double* qx, int nx,
                        const double* qy, int ny, const double* qz, int nz, int32_t* nn_idx_d, cudaStream_t st);
// vel_h / rho_h chunks -> pay_d [np] float4 (v', m) in INPUT order: copies on the copy stream (behind whatever is queued
// there), packing on a side stream; `st` is made to wait for the last chunk.
int vp_pack_payload_host(vp_ctx* ctx, const vp_host_chunks* hc, int dtype, int64_t np, double lcell3, void* staging_d, float* pay_d,
                         cudaStream_t st);
int vp_host_streams(vp_ctx* ctx);
// copy and pack streams wait for everything queued on st so far (scratch of earlier calls is then free to reuse)
int vp_host_fork(vp_ctx* ctx, cudaStream_t st);

// internal forms used by pipeline.cu (typed device pointers, arena already reserved by the caller)
size_t vp_nn_grid_scratch_bytes_tables(int64_t np, int pos_dtype, const double* qx, int nx, const double* qy, int ny,
                                       const double* qz, int nz, const vp_nn_opts* opts);
size_t vp_pk_fields_scratch_bytes(const vp_pk_plan* plan);

// stage timer: records an event pair around the launches between construction and destruction
struct vp_stage {
  vp_ctx* c;
  cudaStream_t st;
  int idx = -1;
  vp_stage(vp_ctx* ctx, const char* name, cudaStream_t s, int launches, double bytes = 0.0) : c(ctx), st(s) {
    c->n_launch += launches;
    if (!c->prof_on) return;
    vp_prof_rec r;
    r.name = name;
    r.launches = launches;
    r.bytes = bytes;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    c->prof.push_back(r);
    idx = int(c->prof.size()) - 1;
  }
  ~vp_stage() {
    if (idx >= 0) cudaEventRecord(c->prof[idx].b, st);
  }
};

static inline int vp_ceil_log2(uint64_t v) {
  int b = 0;
  while ((uint64_t(1) << b) < v) ++b;
  return b;
}
