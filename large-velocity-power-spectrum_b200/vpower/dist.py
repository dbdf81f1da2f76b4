"""Multi-GPU particles -> P(k): x-slab decomposition, one process per GPU (torch.distributed, NCCL).

What is distributed, and how:
  * the lattice is cut into x slabs; rank r grids the nodes of planes [r*N/P, (r+1)*N/P) from the particles inside
    that slab plus a halo (exact: nodes the halo cannot prove are reported and the halo is widened);
  * every real field is transformed along z and y locally; the store of the y pass is the transpose packing; ONE
    all-to-all per field component moves 8*N^2*(N/2)/P^2 bytes between each pair of ranks; the x pass, |F|^2 and the
    shell binning then run on kz slabs;
  * the per-rank shell sums / counts are summed with one all-reduce.
The reference instead replicates the particles and the search index on every MPI rank and builds each rank's
residue class of k-space by an O(N^3 m^3) phase sum (scripts/parallel_optimized.py:228-236, 362-389, 455-456).

The compute steps go through a small backend object so that the wiring (slab bounds, exchange layout, reductions)
can be exercised on CPU with the gloo backend (tests/test_dist_cpu.py); `CudaBackend` is the product.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def slab_bounds(N, nranks, rank):
    """x planes [x0, x1) and half-spectrum columns [z0, z1) owned by `rank`."""
    if N % nranks or (N // 2) % nranks:
        raise ValueError(f"N={N} is not divisible into {nranks} slabs")
    nx, kzc = N // nranks, N // 2 // nranks
    return rank * nx, (rank + 1) * nx, rank * kzc, (rank + 1) * kzc


def keep_range(ax, x0, x1, nranks, rank, halo_cells):
    """Particle x range a rank keeps: its slab widened by `halo_cells` node spacings (a whole number keeps the
    slab's cell list corner-aligned with the lattice nodes); open at the domain ends."""
    h = (ax[-1] - ax[0]) / (len(ax) - 1)
    lo = ax[x0] - halo_cells * h
    hi = ax[x1 - 1] + halo_cells * h
    return lo, hi, rank == 0, rank == nranks - 1


class CudaBackend:
    """The product path: hand-written kernels behind the C ABI."""

    def __init__(self, N, k_axis, edges, nranks, rank, p2p=False, group=None):
        self.plan = _lib.PkPlan(N, k_axis, edges, nranks=nranks, rank=rank)
        self.p2p = bool(p2p) and nranks > 1
        self.slab_x = None
        if self.p2p:       # both exchanges fused into kernels: peer stores over NVLink instead of NCCL all-to-all
            self.plan.p2p_setup(group)
            self.slab_x = _lib.SlabExchangeP2P(nranks, rank, group)

    def close(self, group=None):
        """Collective: release the peer-mapped buffers in the order CUDA IPC requires (unmap everywhere, barrier, free)."""
        self.plan.close(group)

    def grid_slab(self, pos, vel, rho, ax_loc, ax, lcell3, keep):
        lo, hi, open_lo, open_hi = keep
        o = _lib.NNOpts()
        o.use_x_keep = 1
        o.x_keep_lo, o.x_keep_hi = float(lo), float(hi)
        o.x_lo_is_domain_edge, o.x_hi_is_domain_edge = int(open_lo), int(open_hi)
        if pos.stride(0) != 3:          # column views of the interleaved rows that came out of the particle exchange
            o.row_stride = int(pos.stride(0))
        _, nn_pos, spay = _lib.nn_grid_payload(pos, vel, rho, ax_loc, ax, ax, lcell3, want_idx=False, opts=o)
        st = _lib.nn_grid_stats()
        return (nn_pos, spay), st["n_unresolved"]

    def grid_slab_fields(self, pos, vel, rho, ax_loc, ax, lcell3, keep, quantities, strict):
        """Gridding of the slab with the field planes written by the search itself (vp_nn_grid_fields).
        -> ({quantity: (planes, multiplicity)}, number of unproven nodes)."""
        lo, hi, open_lo, open_hi = keep
        o = _lib.NNOpts()
        o.use_x_keep = 1
        o.x_keep_lo, o.x_keep_hi = float(lo), float(hi)
        o.x_lo_is_domain_edge, o.x_hi_is_domain_edge = int(open_lo), int(open_hi)
        if pos.stride(0) != 3:          # column views of the interleaved rows that came out of the particle exchange
            o.row_stride = int(pos.stride(0))
        want_p = (True, not strict, not strict) if "momentum" in quantities else (False, False, False)
        f, _ = _lib.nn_grid_fields(pos, vel, rho, ax_loc, ax, ax, lcell3, want_v="velocity" in quantities, want_p=want_p,
                                   want_e="energy" in quantities, opts=o)
        out = {}
        if "velocity" in quantities:
            out["velocity"] = ([f["vx"], f["vy"], f["vz"]], 1.0)
        if "momentum" in quantities:
            out["momentum"] = ([f["px"]], 3.0) if strict else ([f["px"], f["py"], f["pz"]], 1.0)
        if "energy" in quantities:
            out["energy"] = ([f["e"]], 1.0)
        return out, _lib.nn_grid_stats()["n_unresolved"]

    def bucket(self, pos, vel, rho, lo, hi):
        return _lib.slab_bucket(pos, vel, rho, lo, hi)

    def fields(self, gridded, quantity, strict):
        nn_pos, spay = gridded
        if quantity == "velocity":
            f = _lib.fields_sorted(nn_pos, spay, want_v=True)
            return [f["vx"], f["vy"], f["vz"]], 1.0
        if quantity == "momentum":
            if strict:
                f = _lib.fields_sorted(nn_pos, spay, want_v=False, want_p=(True, False, False))
                return [f["px"]], 3.0
            f = _lib.fields_sorted(nn_pos, spay, want_v=False, want_p=(True, True, True))
            return [f["px"], f["py"], f["pz"]], 1.0
        f = _lib.fields_sorted(nn_pos, spay, want_v=False, want_e=True)
        return [f["e"]], 1.0

    def fields_all(self, gridded, quantities, strict):
        """Every plane the quantities need from ONE pass over (nn_pos, sorted records) -> {quantity: (planes, multiplicity)}."""
        nn_pos, spay = gridded
        want_p = (True, not strict, not strict) if "momentum" in quantities else (False, False, False)
        f = _lib.fields_sorted(nn_pos, spay, want_v="velocity" in quantities, want_p=want_p, want_e="energy" in quantities)
        out = {}
        if "velocity" in quantities:
            out["velocity"] = ([f["vx"], f["vy"], f["vz"]], 1.0)
        if "momentum" in quantities:
            out["momentum"] = ([f["px"]], 3.0) if strict else ([f["px"], f["py"], f["pz"]], 1.0)
        if "energy" in quantities:
            out["energy"] = ([f["e"]], 1.0)
        return out

    def fft_local(self, slabs):
        return self.plan.dist_local(slabs)

    def fft_final(self, recv):
        return self.plan.dist_final(recv)


def exchange_particles(pos, vel, rho, ax, nranks, rank, halo_cells, group=None, backend=None):
    """Sharded input: every rank holds an arbitrary subset of the particles.  Each particle is sent to the rank(s)
    whose slab + halo contains its x (one all-to-all-v of [pos|vel|rho] rows; a particle inside a halo goes to two
    ranks).  -> (pos, vel, rho) of everything this rank needs.  torch ops only (plumbing around NCCL)."""
    import torch
    import torch.distributed as dist
    N = len(ax)
    los, his = [], []
    for d in range(nranks):
        x0, x1, _, _ = slab_bounds(N, nranks, d)
        lo, hi, open_lo, open_hi = keep_range(ax, x0, x1, nranks, d, halo_cells)
        los.append(-np.inf if open_lo else lo)
        his.append(np.inf if open_hi else hi)
    if backend is not None and getattr(backend, "slab_x", None) is not None:
        recv = backend.slab_x.exchange(pos, vel, rho, los, his)         # rows stored straight into the peers' buffers
        return recv[:, 0:3], recv[:, 3:6], (recv[:, 6] if rho is not None else None)
    if backend is not None and hasattr(backend, "bucket"):
        send, counts = backend.bucket(pos, vel, rho, los, his)          # one counting + one scatter kernel
    else:                                                               # generic torch form (CPU tests)
        rows = torch.cat([pos, vel] + ([rho[:, None]] if rho is not None else []), dim=1)
        x = pos[:, 0]
        parts, counts = [], []
        for d in range(nranks):
            sel = rows[(x >= los[d]) & (x <= his[d])]
            parts.append(sel)
            counts.append(sel.shape[0])
        send = torch.cat(parts, dim=0).contiguous()
    w = send.shape[1]
    send_counts = torch.tensor(counts, dtype=torch.int64, device=pos.device)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    rc = [int(v) for v in recv_counts.tolist()]
    recv = torch.empty((sum(rc), w), dtype=send.dtype, device=pos.device)
    dist.all_to_all_single(recv.reshape(-1), send.reshape(-1), output_split_sizes=[c * w for c in rc],
                           input_split_sizes=[c * w for c in counts], group=group)
    if pos.is_cuda:     # column views: the gridding kernels read the interleaved rows in place (vp_nn_opts.row_stride)
        return recv[:, 0:3], recv[:, 3:6], (recv[:, 6] if rho is not None else None)
    p, v = recv[:, 0:3].contiguous(), recv[:, 3:6].contiguous()
    r = recv[:, 6].contiguous() if rho is not None else None
    return p, v, r


def particles_to_pk_dist(pos, vel, rho, ax, lcell3, norm, k_axis, edges, quantities=("velocity",), momentum_strict=True,
                         group=None, backend=None, halo_cells=4, max_halo_cells=None, timings=None, sharded=False):
    """Whole path on `world_size` ranks.  `sharded=False`: every rank passes the SAME particle arrays (replicated, as in
    the reference MPI script).  `sharded=True`: every rank passes its own subset; the particles are first exchanged so that
    each rank holds its slab + halo.  Tensors live on the rank's device.
    -> (dict quantity -> Psum[nbins] numpy, Nsample[nbins] numpy), identical on every rank."""
    import torch
    import torch.distributed as dist

    nranks = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N = len(ax)
    x0, x1, _, _ = slab_bounds(N, nranks, rank)
    if backend is None:
        backend = CudaBackend(N, k_axis, edges, nranks, rank)
    if max_halo_cells is None:
        max_halo_cells = N
    shard = (pos, vel, rho)
    marks = []

    def mark(name):
        if timings is not None and isinstance(pos, torch.Tensor) and pos.is_cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    # ---- K1 on the slab; widen the halo until every node is proven (rarely more than once)
    halo = halo_cells
    while True:
        mark("start")
        if sharded and nranks > 1:
            pos, vel, rho = exchange_particles(*shard, ax, nranks, rank, halo, group, backend)
        mark("exchange_particles")
        if hasattr(backend, "grid_slab_fields"):       # the search writes the planes itself
            gridded = None
            planes, unresolved = backend.grid_slab_fields(pos, vel, rho, ax[x0:x1], ax, lcell3,
                                                          keep_range(ax, x0, x1, nranks, rank, halo), quantities, momentum_strict)
        else:
            planes = None
            gridded, unresolved = backend.grid_slab(pos, vel, rho, ax[x0:x1], ax, lcell3,
                                                    keep_range(ax, x0, x1, nranks, rank, halo))
        mark("grid_slab")
        flag = torch.tensor([int(unresolved > 0)], dtype=torch.int64, device=_device_of(pos))
        if nranks > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()) == 0:
            break
        if halo >= max_halo_cells:
            raise _lib.VPowerError("nearest-particle search could not be proven inside the widest halo")
        halo = min(2 * halo, max_halo_cells)
    if timings is not None:
        timings["halo_cells"] = halo
        # bytes this rank stores into the other ranks' buffers (exact, from the exchange tables): the particle rows, and per
        # transformed component the kz columns of its x slab that other ranks own
        sx = getattr(backend, "slab_x", None)
        ncomp_total = sum(3 if (q == "velocity" or (q == "momentum" and not momentum_strict)) else 1 for q in quantities)
        timings["nvlink_bytes"] = {
            "particle_exchange": int(getattr(sx, "last_peer_bytes", 0)) if (sharded and sx is not None) else 0,
            "transpose": int(ncomp_total * (x1 - x0) * N * (N // 2) * 8 * (nranks - 1) // nranks) if getattr(backend, "p2p", False) else 0}

    out = {}
    ns_total = None
    if planes is None and hasattr(backend, "fields_all"):
        planes = backend.fields_all(gridded, quantities, momentum_strict)
    for q in quantities:
        slabs, mult = planes.pop(q) if planes is not None else backend.fields(gridded, q, momentum_strict)
        if getattr(backend, "p2p", False):
            # fused exchange: the y pass stores into every rank's receive buffer; a stream-ordered tiny all-reduce is
            # the barrier between "all ranks have written" and "this rank reads"
            backend.plan.dist_local_p2p(slabs)
            mark("fields+fft_local")
            tok = torch.zeros(1, device=slabs[0].device)
            dist.all_reduce(tok, group=group)
            mark("all_to_all")
            psum, ns = backend.plan.dist_final_p2p(len(slabs), slabs[0].device)
            mark("fft_final")
            dist.all_reduce(psum, op=dist.ReduceOp.SUM, group=group)      # also protects the buffers from the next stores
            dist.all_reduce(ns, op=dist.ReduceOp.SUM, group=group)
            out[q] = psum.cpu().numpy() * (mult * norm)
            ns_total = ns.cpu().numpy().astype(np.int64)
            mark("all_reduce+d2h")
            continue
        send = backend.fft_local(slabs)                              # [P, nx, N, kzc] complex64 per component
        mark("fields+fft_local")
        recv = []
        for s in send:
            r = torch.empty_like(s)
            if nranks > 1:
                dist.all_to_all_single(torch.view_as_real(r).reshape(-1), torch.view_as_real(s).reshape(-1), group=group)
            else:
                r.copy_(s)
            recv.append(r.reshape(N, N, -1))                         # [x][ky][kz_local]
        mark("all_to_all")
        psum, ns = backend.fft_final(recv)
        mark("fft_final")
        if nranks > 1:
            dist.all_reduce(psum, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(ns, op=dist.ReduceOp.SUM, group=group)
        out[q] = psum.cpu().numpy() * (mult * norm)
        ns_total = ns.cpu().numpy().astype(np.int64)
        mark("all_reduce+d2h")
    if marks:
        torch.cuda.synchronize()
        acc = {}
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            acc[name] = acc.get(name, 0.0) + a.elapsed_time(b)
        timings["phases_ms"] = acc
    return out, ns_total


def _device_of(t):
    import torch
    return t.device if isinstance(t, torch.Tensor) else torch.device("cpu")
