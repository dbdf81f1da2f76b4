/* vpower_b200.h -- C ABI of libvpower_b200.so
 *
 * B200 (sm_100a) implementation of the particles -> P(k) hot path of `vpower`
 * (YujieH3/large-velocity-power-spectrum).  The reference has no FFI of its own:
 * the path sits behind plain Python functions on numpy arrays.
 * Each entry point below names the reference interface it replaces; the Python
 * mirror in large-velocity-power-spectrum_b200/vpower/ binds them with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C, no torch / C++ types.  `stream` is a cudaStream_t passed as void*.
 *   - pointers named *_d are DEVICE pointers owned by the caller; pointers named
 *     *_h are HOST pointers.  No ownership is transferred.
 *   - every call returns 0 on success, <0 on error; vp_last_error() gives text
 *     (thread local).  There is NO CPU fallback: without a CUDA device every
 *     compute call fails with VP_ERR_CUDA.
 *   - calls are asynchronous on `stream` unless the comment says "syncs".
 *   - a vp_ctx owns a grow-only device scratch arena (cudaMalloc) and a few small tables
 *     that the calls below carve their temporaries from.  Calls on one ctx are serialised
 *     by an internal lock, and a call issued on a different stream than the previous call
 *     of the ctx first waits (on the device) for that call's work -- the scratch is shared.
 *     For concurrency between streams or threads use one ctx per stream.
 */
#ifndef VPOWER_B200_H
#define VPOWER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP_OK 0
#define VP_ERR_ARG (-1)
#define VP_ERR_CUDA (-2)
#define VP_ERR_NOMEM (-3)
#define VP_ERR_UNSUPPORTED (-4)
#define VP_ERR_UNRESOLVED (-5) /* NN search could not prove exactness inside the kept particle range */

#define VP_F32 0
#define VP_F64 1

typedef struct vp_ctx vp_ctx;
typedef struct vp_pk_plan vp_pk_plan;

int vp_version(void);
const char* vp_last_error(void);

/* device context + scratch arena */
int vp_ctx_create(int device, vp_ctx** out);
int vp_ctx_destroy(vp_ctx* ctx);
/* Bytes currently held by the ctx arena (grow-only high-water mark). */
size_t vp_ctx_arena_bytes(vp_ctx* ctx);
/* Release the arena (it is re-grown on demand). Syncs the device. */
int vp_ctx_trim(vp_ctx* ctx);

/* Measurement hooks (bench.py): per-stage device time from CUDA events recorded on the launching stream.
 * vp_profile_report writes JSON {"stage": {"ms","calls","launches","bytes"}} into buf, clears the records, syncs.
 * vp_launch_count: kernels launched through this ctx since it was created. */
int vp_profile_enable(vp_ctx* ctx, int on);
int vp_profile_report(vp_ctx* ctx, char* buf, size_t buflen);
unsigned long long vp_launch_count(vp_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * K1  nearest-particle gridding.
 * Replaces: pyann.nn2(data, query, k=1, eps=0, 'kd', 'standard') + `index -= 1`
 *           (vpower/interp.py:1027-1037), the ann/ann_sample CLI (ann/ann_sample.cpp:104-118)
 *           and AnnoyIndex.get_nns_by_vector(q, n=1) (scripts/parallel_optimized.py:302-313,348).
 * Semantics: for every node of the separable lattice (qx[i], qy[j], qz[k]) the 0-based index of
 *   the particle minimising the f64 squared distance ((dx*dx+dy*dy)+dz*dz), NON-periodic, ties
 *   -> lowest particle index.  Result is exact (every answer is proven against the distance to
 *   the unsearched region); it does not depend on the cell-list parameters.
 * pos_d      [np,3] row-major, dtype VP_F32 or VP_F64
 * qx_h/qy_h/qz_h  host f64 tables of node coordinates (the caller evaluates the reference's own
 *   lattice expression: interp.py:1063 linspace, or parallel_optimized.py:343-346 i*LCELL as f32)
 * nn_idx_d   [nx,ny,nz] int32, C order
 * opts       may be NULL (defaults)
 * Multi-GPU: a rank passes its own slice of qx and keeps only particles with
 *   x in [x_keep_lo, x_keep_hi]; answers stay exact or the call reports VP_ERR_UNRESOLVED
 *   through vp_nn_grid_unresolved() so the caller can widen the kept range.
 */
typedef struct vp_nn_opts {
  int cells_x, cells_y, cells_z; /* cell-list resolution per axis; 0 = auto (about 1 particle/cell) */
  int use_x_keep;                /* 0: keep all particles (single GPU) */
  double x_keep_lo, x_keep_hi;   /* particles outside are dropped when use_x_keep */
  int x_lo_is_domain_edge;       /* 1: nothing exists below x_keep_lo (no constraint on that side) */
  int x_hi_is_domain_edge;
  int row_stride;                /* 0: compact arrays ([np,3], [np,3], [np]); >0: pos/vel/rho are columns of one interleaved
                                    row array with this many elements per particle (the layout vp_slab_bucket produces) */
} vp_nn_opts;

int vp_nn_grid(vp_ctx* ctx, const void* pos_d, int pos_dtype, int64_t np, const double* qx_h, int nx,
               const double* qy_h, int ny, const double* qz_h, int nz, int32_t* nn_idx_d,
               const vp_nn_opts* opts, void* stream);
/* Number of lattice nodes the last vp_nn_grid on this ctx needed the wide (ring >= 2) search for,
 * and the number it could not prove at all (only possible with use_x_keep).  Syncs `stream`. */
int vp_nn_grid_stats(vp_ctx* ctx, int64_t* n_wide, int64_t* n_unresolved, int64_t* n_kept, void* stream);
/* The same plus two diagnostics: out5 = { n_wide, n_unresolved, n_kept, nodes the wider prefilter stage took, particles
 * outside the cell grid (clamped into end cells; nodes next to end cells then go to the exact stage) }.  Syncs. */
int vp_nn_grid_stats_ex(vp_ctx* ctx, int64_t* out5, void* stream);
/* The cell list vp_nn_grid would build for these arguments -- host arithmetic only, no device, no ctx (useful for sizing
 * and for checking the key layout at sizes that do not fit a test machine).  info_out[10] = { cells_x, cells_y, cells_z,
 * yb, lb, nyc, bins, row_bits, scratch_MiB, corner_aligned }: the sort key is (row << lb) | local with
 * row = cx*nyc + (cy >> yb) < 2^row_bits and local = (cy mod 2^yb)*cells_z + cz < bins <= 2^lb, row_bits + lb <= 32;
 * scratch_MiB = arena bytes vp_nn_grid_payload needs (MiB, rounded up). */
int vp_nn_grid_plan(int64_t np, const double* qx_h, int nx, const double* qy_h, int ny, const double* qz_h, int nz,
                    const vp_nn_opts* opts, int64_t* info_out);

/* K1 with a payload.  Besides (optionally) nn_idx_d it returns
 *   spay_d   [np] x 8 float: the CELL-SORTED 32-byte records -- floats 0..3 the search half (cell-relative offsets, particle
 *            index), floats 4..7 = (vx', vy', vz', m), v' = (rho*v)/rho, m = rho*lcell3 evaluated in the input dtype
 *            (interp.py:199-213,272-273), rho_d == NULL -> rho = 1;
 *   nn_pos_d [nx,ny,nz] int32 = position in that sorted order of every node's nearest particle,
 * so that vp_fields_sorted reads the payload almost sequentially instead of gathering by particle index. */
int vp_nn_grid_payload(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype, int64_t np,
                       const double* qx_h, int nx, const double* qy_h, int ny, const double* qz_h, int nz, double lcell3,
                       int32_t* nn_idx_d /* may be NULL */, int32_t* nn_pos_d, float* spay_d, const vp_nn_opts* opts,
                       void* stream);
/* K1 + K3 in one call (the whole-path form): the search stages write the requested field planes themselves -- the stage
 * that settles a node reads the payload half of the winner's record (the 32-byte sector the search has just read) and
 * stores vx, vy, vz / px, py, pz / e / m at that node (same planes and arithmetic as vp_build_fields, interp.py:272-273,
 * 501-557).  v_d / p_d may be NULL or hold NULL entries, e_d / m_d may be NULL (at least one plane must be requested);
 * nn_idx_d may be NULL.  The sorted records live in the ctx arena for the duration of the call. */
int vp_nn_grid_fields(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype, int64_t np,
                      const double* qx_h, int nx, const double* qy_h, int ny, const double* qz_h, int nz, double lcell3,
                      float* const v_d[3], float* const p_d[3], float* e_d, float* m_d, int32_t* nn_idx_d,
                      const vp_nn_opts* opts, void* stream);
/* K3 on the sorted payload: same planes as vp_build_fields (interp.py:501-557). */
int vp_fields_sorted(vp_ctx* ctx, const int32_t* nn_pos_d, int64_t n_nodes, const float* spay_d, float* const v_d[3],
                     float* const p_d[3], float* e_d, float* m_d, void* stream);

/* Row gather: dst[i,:] = src[idx[i],:]  (interp.py:1040-1045 `f[index]`), row_bytes in {4,8,12,16,24,32}. */
int vp_gather_rows(vp_ctx* ctx, const int32_t* idx_d, int64_t n, const void* src_d, int row_bytes,
                   void* dst_d, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3  payload + field algebra, fused with the gather.
 * Replaces: GasParticles.density_velocity_vector (interp.py:199-213), the division / mass lines
 *   interp.py:272-273 and the field algebra of BoxField.{velocity,momentum,kinetic_energy}_power
 *   (interp.py:501-557).
 * Per node n with particle i = nn_idx[n]:  w = rho_i*v_i (input dtype), v = w/rho_i, m = rho_i*lcell3.
 * Output planes are float32 [n_nodes] each (FFT input layout, one contiguous cube per component);
 * a NULL pointer skips that plane.
 *   v_d[3]  : vx,vy,vz           p_d[3]: v_c*m  (reference quirk interp.py:523-525 is applied by the
 *   caller: it passes the vx*m plane three times)   e_d : m*(vx^2+vy^2+vz^2)   m_d : m
 * vel_d [np,3], rho_d [np] of `dtype`; rho_d == NULL means rho = 1 (script path: plain velocity).
 */
int vp_build_fields(vp_ctx* ctx, const int32_t* nn_idx_d, int64_t n_nodes, const void* vel_d,
                    const void* rho_d, int dtype, double lcell3, float* const v_d[3], float* const p_d[3],
                    float* e_d, float* m_d, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  NGP deposit.  Replaces deposit_to_grid (interp.py:996-1015):
 *   cell = int((pos // Lcell) % N) per axis with numpy floor-division semantics (periodic wrap),
 *   grid[cell] += w   accumulated in f64.  grid_d is [N,N,N,C] f64 and is ZEROED by the call.
 */
int vp_deposit_ngp(vp_ctx* ctx, const void* pos_d, int pos_dtype, int64_t np, const double* w_d, int ncomp,
                   int N, double Lbox, double* grid_d, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4 + K5  3-D r2c FFT, |F|^2, spherical k-shell binning.
 * Replaces: _vector_power/_scalar_power (interp.py:1372-1421), FFTW_power
 *   (parallel_optimized.py:124-141), _pair_power/_hist_sample (interp.py:1440-1482) and
 *   pair_power/hist_sample (parallel_optimized.py:145-190).
 * The plan holds the geometry: N (power of two, 32..2048), the per-axis angular wavenumber table
 *   k_h[N] = 2 pi fftfreq(N, Lcell) evaluated by the caller with numpy, and the bin edges
 *   edges_h[nbins+1] evaluated by the caller with the reference expression (arange or linspace).
 * Binning is numpy.histogram's: bin j holds edges[j] <= |k| < edges[j+1], last bin closed, where
 *   |k| = sqrt((kx*kx + ky*ky) + kz*kz) in f64 -- reproduced bit-exactly (thresholds on the squared
 *   magnitude are derived on the host so that no rounding differs).
 */
int vp_pk_plan_create(vp_ctx* ctx, int N, const double* k_h, const double* edges_h, int nbins,
                      vp_pk_plan** out);
int vp_pk_plan_destroy(vp_pk_plan* plan);

/* psum_d[j] = sum over ALL N^3 modes in shell j of sum_c |FFT(field_c)|^2 (unnormalised forward DFT,
 * e^{-ikx}); nsample_d[j] = number of modes.  The caller applies 1/2 a^2.  The ncomp real cubes
 * field_d[c] ([N,N,N] f32, C order) are OVERWRITTEN (the transform is in place).
 * psum_d: f64[nbins], nsample_d: u64[nbins]; both are overwritten. */
int vp_pk_fields(vp_pk_plan* plan, float* const* field_d, int ncomp, double* psum_d, uint64_t* nsample_d,
                 void* stream);

/* Slab decomposition over `nranks` processes (one per GPU).  Before the exchange rank r owns the x planes
 * [r*N/nranks, (r+1)*N/nranks) of every real field; after it, the half-spectrum columns kz in
 * [r*kzc, (r+1)*kzc), kzc = N/2/nranks, for ALL x and ky.  The exchange itself (an all-to-all with equal blocks)
 * is the caller's: torch.distributed / NCCL in the Python mirror (vpower/dist.py).
 *   vp_pk_dist_local : z pass in place on field_d[c] ([N/nranks][N][N] f32), then the y pass, whose store is the
 *                      transpose packing: send_d[c] is [nranks][N/nranks][N][kzc] complex64, block d goes to rank d.
 *   vp_pk_dist_final : recv_d[c] is [N][N][kzc] complex64 (= the received blocks in rank order); x pass fused with
 *                      |F|^2 and shell binning -> this rank's PARTIAL psum_d/nsample_d, to be summed over ranks.
 * Replaces the reference's folded DFT across MPI ranks (scripts/parallel_optimized.py:362-389, 455-456). */
int vp_pk_plan_create_dist(vp_ctx* ctx, int N, int nranks, int rank, const double* k_h, const double* edges_h, int nbins,
                           vp_pk_plan** out);
int vp_pk_dist_local(vp_pk_plan* plan, float* const* field_d, int ncomp, float* const* send_d, void* stream);
int vp_pk_dist_final(vp_pk_plan* plan, float* const* recv_d, int ncomp, double* psum_d, uint64_t* nsample_d, void* stream);

/* The same exchange FUSED into the y pass (no NCCL call on the data path): every rank allocates its receive buffers
 * inside the plan and exports them through CUDA IPC; after the handles of all ranks have been gathered (any transport:
 * the Python mirror uses torch.distributed.all_gather) and opened, the y-pass kernel stores every tile straight into
 * the receive buffer of the rank that owns its kz columns -- local HBM for its own block, NVLink peer stores for the
 * others.  The caller separates the passes with a stream-ordered barrier across ranks (a tiny NCCL all-reduce):
 *   local_p2p (all ranks)  ->  barrier  ->  final_p2p  ->  all-reduce of the shells (doubles as the barrier that
 *   protects the receive buffers from the next quantity's stores).
 *   handles_out : ncomp_max * 64 bytes;   all_handles : [nranks][ncomp_max][64] bytes in rank order. */
int vp_pk_dist_p2p_alloc(vp_pk_plan* plan, int ncomp_max, unsigned char* handles_out);
int vp_pk_dist_p2p_open(vp_pk_plan* plan, const unsigned char* all_handles);
int vp_pk_dist_local_p2p(vp_pk_plan* plan, float* const* field_d, int ncomp, void* stream);
int vp_pk_dist_final_p2p(vp_pk_plan* plan, int ncomp, double* psum_d, uint64_t* nsample_d, void* stream);
/* Teardown: every rank vp_pk_dist_p2p_close (unmaps the peers' receive buffers; syncs), then a barrier across the ranks,
 * then vp_pk_plan_destroy (frees this rank's exported buffers). */
int vp_pk_dist_p2p_close(vp_pk_plan* plan);

/* Sharded particle input for the slab decomposition: every particle of this rank's subset is copied into the block of
 * each destination rank d whose kept range lo_h[d] <= x <= hi_h[d] contains it (use -/+ infinity for open ends).
 *   rows_d   [cap_rows, 8] of dtype: x y z vx vy vz rho 0 (rho = 0 when rho_d is NULL; the pad makes a row one or two whole
 *            32-byte sectors and every run 16-byte aligned), blocks in rank order;  counts_h[d] = rows for rank d.
 * One counting pass, one scatter pass (rows grouped in shared memory, one global atomic and ONE bulk copy shared -> global
 * -- cp.async.bulk, the TMA engine -- per tile and destination).  Syncs (the split sizes of the
 * all-to-all are needed on the host).  nranks <= 16. */
int vp_slab_bucket(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype, int64_t np,
                   const double* lo_h, const double* hi_h, int nranks, void* rows_d, int64_t cap_rows, int64_t* counts_h,
                   void* stream);

/* The same exchange FUSED into the bucketing kernel: each rank owns a receive buffer (allocated here, exported with CUDA
 * IPC, mapped by every peer); after vp_slab_count and an exchange of the counts (any transport) every rank knows the
 * first row it may write in each destination's buffer, and vp_slab_scatter_p2p stores the rows (8 elements each, as above)
 * there directly: bulk copies from shared memory into own HBM or, over NVLink, into the peers' buffers.  The caller separates "all ranks have stored" from "this rank reads" with a stream-ordered
 * barrier across ranks.  Rows of one destination arrive grouped by source rank, in rank order. */
/* Lifetime of the shared buffers (CUDA leaves freeing an exported allocation that a peer still maps undefined): to grow or
 * drop them, EVERY rank calls vp_slab_p2p_close (unmaps the peers; syncs the device), the caller runs a barrier across the
 * ranks, and only then vp_slab_p2p_alloc frees and re-allocates (it refuses while peers are mapped). */
int vp_slab_p2p_close(vp_ctx* ctx);
int vp_slab_p2p_alloc(vp_ctx* ctx, size_t bytes, unsigned char* handle_out /* 64 bytes */);
int vp_slab_p2p_open(vp_ctx* ctx, int nranks, int rank, const unsigned char* all_handles /* [nranks][64] */);
int vp_slab_p2p_buffer(vp_ctx* ctx, void** ptr_out, size_t* bytes_out);
int vp_slab_count(vp_ctx* ctx, const void* pos_d, int dtype, int64_t np, const double* lo_h, const double* hi_h, int nranks,
                  int64_t* counts_h, void* stream);                                            /* syncs */
int vp_slab_scatter_p2p(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype, int64_t np,
                        const double* lo_h, const double* hi_h, int nranks, const int64_t* first_row_h, void* stream);

/* Snapshot preamble on the device, in place (the two O(Np) host passes of the reference between reading the snapshot and
 * gridding it):
 *   do_shift: pos[:, c] -= min(pos[:, c])                GasParticles.shift_to_origin (vpower/interp.py:169-175),
 *                                                         scripts/parallel_optimized.py:278-282
 *   do_bulk : v[:, c] -= sum(m * v[:, c]) / sum(m)       GasParticles.remove_bulk_velocity (vpower/interp.py:178-182),
 *                                                         scripts/parallel_optimized.py:284-288
 * pos_d, vel_d [np,3], mass_d [np] of `dtype`.  The minimum is exact; the mass-weighted sums are accumulated in f64 (numpy
 * accumulates pairwise in the array dtype).  min_h[3] / bulk_h[3] (may be NULL) receive what was subtracted.  Syncs. */
int vp_snapshot_preamble(vp_ctx* ctx, void* pos_d, void* vel_d, const void* mass_d, int dtype, int64_t np, int do_shift,
                         int do_bulk, double* min_h, double* bulk_h, void* stream);

/* Test/diagnostic entry points */
/* The blocked half-spectrum layout the y pass writes and the x pass reads, and the tensor map / shared-memory ring of the
 * TMA-fed x pass for a lattice of N with kz_columns half-spectrum columns on this rank (N/2 on one GPU, N/2/nranks in a slab
 * decomposition) -- host arithmetic only, no device, no ctx.  info_out[24] = { C (columns per tile), ky per block, kz tiles,
 * x planes per box, boxes per item, 1 if the TMA kernel serves this N (else per-thread loads), tensor dims[5] (f32 elements,
 * innermost first), strides[4] (bytes, dims 1..4), box[5], bytes of a box slot, items in the ring, dynamic shared memory of
 * the kernel, threads per CTA }.  Element (x, ky, kz = zt*C + c) of the blocked cube is complex number
 * ((((ky / kyb) * N + x) * tiles + zt) * kyb + ky % kyb) * C + c.  (No counterpart in the reference: pyFFTW owns its layout,
 * scripts/parallel_optimized.py:124-141.) */
int vp_fft_x_layout(int N, int kz_columns, int64_t* info_out);
/* In-place 3-D r2c transform only.  Output layout: [N][N][N/2] complex64 where entry (x,y,0) packs
 * (Re F(x,y,0), Re F_zNyquist-line ...) -- see DESIGN.md "half-spectrum layout"; use
 * vp_fft_unpack_half to obtain the conventional [N][N][N/2+1] array. */
int vp_fft_r2c_inplace(vp_pk_plan* plan, float* field_d, void* stream);
int vp_fft_unpack_half(vp_pk_plan* plan, const float* packed_d, float* half_d /* [N][N][N/2+1][2] */,
                       void* stream);
/* Bin a given full power cube P[N,N,N] (f64) -- the standalone form of _pair_power+_hist_sample. */
int vp_power_bin_full(vp_pk_plan* plan, const double* P_d, double* psum_d, uint64_t* nsample_d, void* stream);

/* Full power cube: P_d[N,N,N] (f64) = sum_c |FFT(field_c)|^2 over the whole c2c spectrum -- the array
 * _vector_power/_scalar_power return (interp.py:1372-1421) before the 1/2 a^2 factor.  Fields are overwritten. */
int vp_power_cube(vp_pk_plan* plan, float* const* field_d, int ncomp, double* P_d, void* stream);
/* Folding as an outer stage (the reference's way past the memory of one transform, and its sub-spectrum interchange format).
 *   vp_fold_field : BoxField.fold (interp.py:598-609) = _apply_phase/_get_phase (:1195-1225) + fold_field (:1228-1252):
 *                   folded_d[n,n,n,ncomp] complex128 (n = N/m) = sum over the m^3 sub-blocks of field_c * exp(-i 2 pi beta.x / N),
 *                   divided by m^1.5.  field_d[c]: [N,N,N] f32, untouched.
 *   vp_fold_power : the transform inside FoldedBox.fold_spctrm (interp.py:755-791, _FFTW_vector_power/_FFTW_scalar_power):
 *                   P_d[n,n,n] (f64) = sum_c |FFT_n(folded_c)|^2 of the complex field, before the 1/2 a^2 factor.  The |k| pairing
 *                   with the beta shift and the histogram are vp_k_magnitude + vp_hist_weighted. */
int vp_fold_field(vp_ctx* ctx, const float* const* field_d, int ncomp, int N, int m, const int* beta /* [3] */, double* folded_d,
                  void* stream);
int vp_fold_power(vp_ctx* ctx, const double* folded_d, int ncomp, int n, double* P_d, void* stream);
/* |k| of _pair_power (interp.py:1448-1460): out_d[(i*n+j)*n+l] = sqrt((kx[i]*kx[i] + ky[j]*ky[j]) + kz[l]*kz[l]), f64. */
int vp_k_magnitude(vp_ctx* ctx, const double* kx_h, const double* ky_h, const double* kz_h, int n, double* out_d,
                   void* stream);
/* _hist_sample's two np.histogram calls (interp.py:1474-1477) for arbitrary (k, weight) pairs on the device. */
int vp_hist_weighted(vp_ctx* ctx, const double* k_d, const double* w_d, int64_t n, const double* edges_h, int nbins,
                     double* psum_d, uint64_t* nsample_d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Radix sort (exposed for tests): sorts (key,val) pairs by key, `bits` low bits significant.
 */
int vp_sort_pairs(vp_ctx* ctx, uint32_t* keys_d, uint32_t* vals_d, int64_t n, int bits, void* stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer convenience: the whole path for one particle set in one call, host in / host out.
 * This is what scripts/parallel_optimized.py main() (:272-495) does between "load snapshot" and
 * "savetxt", and what GasParticles.ann_interp_to_field(N).spctrm(q) does in the library.
 *   quantity_mask: bit0 velocity, bit1 momentum, bit2 energy
 *   psum_h [3][nbins] f64 = norm * sum_c |FFT(field_c)|^2 per shell, norm = 1/2 a^2 with
 *          a = (Lbox/2pi)^1.5 / N^3 (interp.py:1380,1386); rows of unrequested quantities untouched
 *   nsample_h [nbins] u64
 *   rho_h may be NULL (rho = 1: the script path, plain velocity); lcell3 = Lcell^3 (interp.py:273)
 *   momentum_strict: 1 = reference behaviour (vx*m used for all three components, interp.py:523-525)
 * The host arrays (pinned for full PCIe speed, pageable works) are uploaded in 2^24-particle chunks on an internal
 * copy stream: positions first -- the cell list and the search run while velocity and density follow -- then velocity
 * and density, packed chunk by chunk into (v', m) records on a high-priority side stream; only the plane gather and
 * the transforms remain after the last byte.  Syncs.
 */
int vp_host_particles_to_pk(vp_ctx* ctx, const void* pos_h, const void* vel_h, const void* rho_h, int dtype,
                            int64_t np, const double* qx_h, const double* qy_h, const double* qz_h, int N,
                            double lcell3, double norm, const double* k_h, const double* edges_h, int nbins,
                            int quantity_mask, int momentum_strict, double* psum_h, uint64_t* nsample_h,
                            void* stream);
/* Same computation with the particle arrays already resident on the device (pos_d/vel_d/rho_d). */
int vp_dev_particles_to_pk(vp_ctx* ctx, const void* pos_d, const void* vel_d, const void* rho_d, int dtype,
                           int64_t np, const double* qx_h, const double* qy_h, const double* qz_h, int N,
                           double lcell3, double norm, const double* k_h, const double* edges_h, int nbins,
                           int quantity_mask, int momentum_strict, double* psum_h, uint64_t* nsample_h,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VPOWER_B200_H */
