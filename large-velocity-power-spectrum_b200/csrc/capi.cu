// Context, scratch arena, error text and the small C-ABI entry points that are not tied to a kernel file.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void vp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

extern "C" const char* vp_last_error(void) { return g_err; }
extern "C" int vp_version(void) { return 100; }

extern "C" int vp_ctx_create(int device, vp_ctx** out) {
  VP_REQUIRE(out, "vp_ctx_create: null out");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    vp_set_error("vp_ctx_create: no CUDA device (%s) -- libvpower_b200 has no CPU fallback",
                 e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return VP_ERR_CUDA;
  }
  VP_REQUIRE(device >= 0 && device < ndev, "vp_ctx_create: device %d out of range (%d devices)", device, ndev);
  VP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  VP_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    vp_set_error("vp_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                 prop.minor);
    return VP_ERR_UNSUPPORTED;
  }
  vp_ctx* c = new vp_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (cudaMalloc(&c->nn_stats_d, sizeof(vp_nn_stats_dev)) != cudaSuccess) {
    delete c;
    vp_set_error("vp_ctx_create: cudaMalloc failed");
    return VP_ERR_NOMEM;
  }
  cudaMemset(c->nn_stats_d, 0, sizeof(vp_nn_stats_dev));
  *out = c;
  return VP_OK;
}

extern "C" int vp_ctx_destroy(vp_ctx* ctx) {
  if (!ctx) return VP_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->arena.base) cudaFree(ctx->arena.base);
  if (ctx->nn_stats_d) cudaFree(ctx->nn_stats_d);
  if (ctx->small_d) cudaFree(ctx->small_d);
  if (ctx->pinned_h) cudaFreeHost(ctx->pinned_h);
  if (ctx->slab_open)
    for (int d = 0; d < ctx->slab_nranks; ++d)
      if (d != ctx->slab_rank && ctx->slab_peer[d]) cudaIpcCloseMemHandle(ctx->slab_peer[d]);
  if (ctx->slab_recv) cudaFree(ctx->slab_recv);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->pack_stream) cudaStreamDestroy(ctx->pack_stream);
  if (ctx->ev_pack) cudaEventDestroy(ctx->ev_pack);
  if (ctx->ev_last) cudaEventDestroy(ctx->ev_last);
  if (ctx->ev_tables) cudaEventDestroy(ctx->ev_tables);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
    if (ctx->ev_used[i]) cudaEventDestroy(ctx->ev_used[i]);
  }
  if (ctx->cached_plan) vp_pk_plan_destroy(ctx->cached_plan);
  if (ctx->cached_plan_key) free(ctx->cached_plan_key);
  delete ctx;
  return VP_OK;
}

extern "C" size_t vp_ctx_arena_bytes(vp_ctx* ctx) { return ctx ? ctx->arena.cap : 0; }

extern "C" int vp_ctx_trim(vp_ctx* ctx) {
  VP_REQUIRE(ctx, "vp_ctx_trim: null ctx");
  VP_CUDA(cudaSetDevice(ctx->device));
  VP_CUDA(cudaDeviceSynchronize());
  if (ctx->arena.base) VP_CUDA(cudaFree(ctx->arena.base));
  ctx->arena = vp_arena();
  return VP_OK;
}

int vp_arena_reserve(vp_ctx* ctx, size_t extra) {
  size_t bytes = ctx->arena.off + vp_align256(extra) + 4096;
  if (ctx->arena.cap >= bytes) return VP_OK;
  if (ctx->arena.off != 0) {
    vp_set_error("vp arena: nested call needs %zu more bytes than the outer call reserved", bytes - ctx->arena.cap);
    return VP_ERR_NOMEM;
  }
  VP_CUDA(cudaDeviceSynchronize());  // nothing may still be using the old block
  if (ctx->arena.base) VP_CUDA(cudaFree(ctx->arena.base));
  ctx->arena = vp_arena();
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ctx->arena.base), bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    vp_set_error("vp arena: cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return VP_ERR_NOMEM;
  }
  ctx->arena.cap = bytes;
  ctx->arena.off = 0;
  return VP_OK;
}

void* vp_arena_alloc(vp_ctx* ctx, size_t bytes) {
  bytes = vp_align256(bytes);
  if (ctx->arena.off + bytes > ctx->arena.cap) return nullptr;
  void* p = ctx->arena.base + ctx->arena.off;
  ctx->arena.off += bytes;
  return p;
}

extern "C" int vp_sort_pairs(vp_ctx* ctx, uint32_t* keys_d, uint32_t* vals_d, int64_t n, int bits, void* stream) {
  VP_REQUIRE(ctx && keys_d && vals_d && n >= 0 && bits >= 0 && bits <= 32, "vp_sort_pairs: bad argument");
  vp_call_guard guard(ctx, static_cast<cudaStream_t>(stream));
  VP_CUDA(cudaSetDevice(ctx->device));
  size_t sb = vp_sort_scratch_bytes(n);
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, sb));
  void* scratch = vp_arena_alloc(ctx, sb);
  return vp_sort_pairs_impl(ctx, keys_d, vals_d, n, bits, scratch, static_cast<cudaStream_t>(stream));
}

extern "C" int vp_profile_enable(vp_ctx* ctx, int on) {
  VP_REQUIRE(ctx, "vp_profile_enable: null ctx");
  ctx->prof_on = on != 0;
  return VP_OK;
}

extern "C" unsigned long long vp_launch_count(vp_ctx* ctx) { return ctx ? ctx->n_launch : 0; }

// JSON: {"stage": {"ms": total, "calls": n, "launches": n, "bytes": total}, ...}; clears the records.  Syncs.
extern "C" int vp_profile_report(vp_ctx* ctx, char* buf, size_t buflen) {
  VP_REQUIRE(ctx && buf && buflen > 2, "vp_profile_report: bad argument");
  VP_CUDA(cudaSetDevice(ctx->device));
  VP_CUDA(cudaDeviceSynchronize());
  struct Agg { std::string name; double ms = 0, bytes = 0; int calls = 0, launches = 0; };
  std::vector<Agg> agg;
  for (auto& r : ctx->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    Agg* g = nullptr;
    for (auto& x : agg) if (x.name == r.name) g = &x;
    if (!g) { agg.push_back(Agg()); g = &agg.back(); g->name = r.name; }
    g->ms += ms; g->bytes += r.bytes; g->calls += 1; g->launches += r.launches;
  }
  ctx->prof.clear();
  std::string out = "{";
  for (size_t i = 0; i < agg.size(); ++i) {
    char tmp[256];
    snprintf(tmp, sizeof tmp, "%s\"%s\": {\"ms\": %.6f, \"calls\": %d, \"launches\": %d, \"bytes\": %.0f}", i ? ", " : "",
             agg[i].name.c_str(), agg[i].ms, agg[i].calls, agg[i].launches, agg[i].bytes);
    out += tmp;
  }
  out += "}";
  VP_REQUIRE(out.size() + 1 <= buflen, "vp_profile_report: buffer too small (%zu needed)", out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return VP_OK;
}
