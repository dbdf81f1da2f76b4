// Whole-path entry points: particles -> (Psum, Nsample) per quantity in one call.
// Host-buffer form = what scripts/parallel_optimized.py main() :272-495 does between loading the snapshot
// and np.savetxt, and GasParticles.ann_interp_to_field(N).spctrm(q) (interp.py:246-277,560-595).
#include <stdlib.h>

#include <vector>

#include "common.cuh"

static int get_plan(vp_ctx* ctx, int N, const double* k_h, const double* edges_h, int nbins, vp_pk_plan** out) {
  size_t len = 2 + size_t(N) + nbins + 1;
  std::vector<double> key(len);
  key[0] = N;
  key[1] = nbins;
  memcpy(&key[2], k_h, sizeof(double) * N);
  memcpy(&key[2 + N], edges_h, sizeof(double) * (nbins + 1));
  if (ctx->cached_plan && ctx->cached_plan_key_len == len && memcmp(ctx->cached_plan_key, key.data(), len * 8) == 0) {
    *out = ctx->cached_plan;
    return VP_OK;
  }
  if (ctx->cached_plan) { vp_pk_plan_destroy(ctx->cached_plan); ctx->cached_plan = nullptr; }
  free(ctx->cached_plan_key);
  ctx->cached_plan_key = nullptr;
  vp_pk_plan* p = nullptr;
  VP_TRY(vp_pk_plan_create(ctx, N, k_h, edges_h, nbins, &p));
  ctx->cached_plan = p;
  ctx->cached_plan_key = static_cast<double*>(malloc(len * 8));
  memcpy(ctx->cached_plan_key, key.data(), len * 8);
  ctx->cached_plan_key_len = len;
  *out = p;
  return VP_OK;
}

static int run_path(vp_ctx* ctx, const void* pos, const void* vel, const void* rho, bool on_host, int dtype, int64_t np,
                    const double* qx, const double* qy, const double* qz, int N, double lcell3, double norm, const double* k_h,
                    const double* edges_h, int nbins, int qmask, int strict, double* psum_h, uint64_t* nsample_h,
                    cudaStream_t st) {
  VP_REQUIRE(ctx && pos && vel && qx && qy && qz && k_h && edges_h && psum_h && nsample_h, "particles_to_pk: null argument");
  vp_call_guard guard(ctx, st);
  // an early error return must not leave H2D / pack work queued on the side streams into arena memory that the scope
  // below has already released
  struct SideDrain {
    vp_ctx* c;
    bool armed = false, ok = false;
    ~SideDrain() {
      if (armed && !ok) {
        if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
        if (c->pack_stream) cudaStreamSynchronize(c->pack_stream);
      }
    }
  } drain{ctx};
  VP_REQUIRE(dtype == VP_F32 || dtype == VP_F64, "particles_to_pk: bad dtype");
  VP_REQUIRE((qmask & 7) != 0 && np > 0 && N > 0, "particles_to_pk: nothing to do");
  VP_CUDA(cudaSetDevice(ctx->device));
  vp_pk_plan* plan = nullptr;
  VP_TRY(get_plan(ctx, N, k_h, edges_h, nbins, &plan));
  const size_t es = dtype == VP_F64 ? 8 : 4;
  const size_t n3 = size_t(N) * N * N;
  const bool want_v = qmask & 1, want_p = qmask & 2, want_e = qmask & 4;
  const int n_pplanes = want_p ? (strict ? 1 : 3) : 0;
  const int nplanes = (want_v ? 3 : 0) + n_pplanes + (want_e ? 1 : 0);

  // host arrays are streamed in chunks: only the positions stay resident (the exact search reads them); velocity and
  // density chunks pass through two staging buffers into the (v', m) records
  const int64_t chunk = np < (int64_t(1) << 24) ? np : (int64_t(1) << 24);
  size_t own = 0;
  if (on_host) own += vp_align256(size_t(np) * 3 * es) + vp_host_chunk_staging_bytes(chunk, dtype, rho != nullptr) + 1024;
  own += vp_align256(n3 * 4) * ((on_host ? 1 : 0) + nplanes) + (on_host ? vp_align256(size_t(np) * 16) : 0) + vp_align256(size_t(nbins) * 16) + 8192;
  size_t inner = vp_nn_grid_scratch_bytes_tables(np, dtype, qx, N, qy, N, qz, N, nullptr);
  size_t inner2 = vp_pk_fields_scratch_bytes(plan);
  vp_arena_scope scope(ctx);
  VP_TRY(vp_arena_reserve(ctx, own + (inner > inner2 ? inner : inner2)));

  const void *pos_d = pos, *vel_d = vel, *rho_d = rho;
  void* pos_res = nullptr;
  if (on_host) {
    pos_res = vp_arena_alloc(ctx, size_t(np) * 3 * es);
    VP_REQUIRE(pos_res, "particles_to_pk: arena carve failed");
  }
  int32_t* nn_pos = on_host ? static_cast<int32_t*>(vp_arena_alloc(ctx, n3 * 4)) : nullptr;
  float* spay = on_host ? static_cast<float*>(vp_arena_alloc(ctx, size_t(np) * 16)) : nullptr;   // host path: (v', m) records in input order
  float* planes[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int i = 0; i < nplanes; ++i) {
    planes[i] = static_cast<float*>(vp_arena_alloc(ctx, n3 * 4));
    VP_REQUIRE(planes[i], "particles_to_pk: arena carve failed");
  }
  double* psum_d = static_cast<double*>(vp_arena_alloc(ctx, size_t(nbins) * 8));
  uint64_t* ns_d = static_cast<uint64_t*>(vp_arena_alloc(ctx, size_t(nbins) * 8));
  VP_REQUIRE((!on_host || (nn_pos && spay)) && psum_d && ns_d, "particles_to_pk: arena carve failed");

  int at = 0;
  float *v3[3] = {nullptr, nullptr, nullptr}, *p3[3] = {nullptr, nullptr, nullptr}, *e1 = nullptr;
  if (want_v) { v3[0] = planes[at++]; v3[1] = planes[at++]; v3[2] = planes[at++]; }
  if (want_p) { p3[0] = planes[at++]; if (!strict) { p3[1] = planes[at++]; p3[2] = planes[at++]; } }
  if (want_e) e1 = planes[at++];
  if (on_host) {
    // Host arrays: the positions cross PCIe first and the whole gridding (bucket pass as the chunks land, counting sort,
    // search) runs while velocity and density follow; those are packed chunk by chunk into (v', m) records in INPUT order
    // on a side stream, and the planes are gathered through the ORIGINAL particle index (nn_pos / spay then hold that index
    // and those records).  Only the plane gather and the transforms remain after the last byte has arrived.
    vp_host_chunks hc;
    hc.pos_h = pos; hc.vel_h = vel; hc.rho_h = rho; hc.chunk = chunk;
    void* staging = vp_arena_alloc(ctx, vp_host_chunk_staging_bytes(chunk, dtype, rho != nullptr) + 512);
    VP_REQUIRE(staging, "particles_to_pk: arena carve failed (staging)");
    drain.armed = true;
    VP_TRY(vp_host_fork(ctx, st));
    VP_TRY(vp_nn_grid_host_pos(ctx, &hc, pos_res, dtype, np, qx, N, qy, N, qz, N, nn_pos, st));
    VP_TRY(vp_pack_payload_host(ctx, &hc, dtype, np, lcell3, staging, spay, st));
    VP_TRY(vp_fields_from_records(ctx, nn_pos, int64_t(n3), spay, 1, 0, v3, p3, e1, nullptr, st));
  } else {
    // device arrays: the search stages write the planes themselves (K3 fused into K1)
    VP_TRY(vp_nn_grid_fields(ctx, pos_d, vel_d, rho_d, dtype, np, qx, N, qy, N, qz, N, lcell3, v3, p3, e1, nullptr, nullptr, nullptr, st));
  }

  std::vector<double> hp(nbins);
  auto one = [&](float** f, int nc, double scale, int row) -> int {
    VP_TRY(vp_pk_fields(plan, f, nc, psum_d, ns_d, st));
    VP_CUDA(cudaMemcpyAsync(hp.data(), psum_d, size_t(nbins) * 8, cudaMemcpyDeviceToHost, st));
    VP_CUDA(cudaMemcpyAsync(nsample_h, ns_d, size_t(nbins) * 8, cudaMemcpyDeviceToHost, st));
    VP_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < nbins; ++j) psum_h[size_t(row) * nbins + j] = hp[j] * scale;
    return VP_OK;
  };
  if (want_v) VP_TRY(one(v3, 3, norm, 0));
  if (want_p) VP_TRY(one(p3, strict ? 1 : 3, strict ? 3.0 * norm : norm, 1));  // strict: three identical components
  if (want_e) VP_TRY(one(&e1, 1, norm, 2));
  drain.ok = true;
  return VP_OK;
}

extern "C" int vp_host_particles_to_pk(vp_ctx* ctx, const void* pos_h, const void* vel_h, const void* rho_h, int dtype,
                                       int64_t np, const double* qx_h, const double* qy_h, const double* qz_h, int N,
                                       double lcell3, double norm, const double* k_h, const double* edges_h, int nbins,
                                       int quantity_mask, int momentum_strict, double* psum_h, uint64_t* nsample_h,
                                       void* stream) {
  return run_path(ctx, pos_h, vel_h, rho_h, true, dtype, np, qx_h, qy_h, qz_h, N, lcell3, norm, k_h, edges_h, nbins,
                  quantity_mask, momentum_strict, psum_h, nsample_h, static_cast<cudaStream_t>(stream));
}

extern "C" int vp_dev_particles_to_pk(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype,
                                      int64_t np, const double* qx_h, const double* qy_h, const double* qz_h, int N,
                                      double lcell3, double norm, const double* k_h, const double* edges_h, int nbins,
                                      int quantity_mask, int momentum_strict, double* psum_h, uint64_t* nsample_h,
                                      void* stream) {
  return run_path(ctx, pos_d, vel_d, rho_d, false, dtype, np, qx_h, qy_h, qz_h, N, lcell3, norm, k_h, edges_h, nbins,
                  quantity_mask, momentum_strict, psum_h, nsample_h, static_cast<cudaStream_t>(stream));
}
