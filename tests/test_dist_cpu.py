"""world_size-2 (and 4) gloo runs of the multi-GPU wiring (vpower/dist.py) on CPU.

The compute backend is replaced by a numpy restatement of the kernels' CONTRACT (slab gridding with a kept range,
packed half-spectrum layout, exchange layout, partial shell sums), so that what is tested is the decomposition:
slab bounds, halo widening, all-to-all block layout and the final reductions must reproduce the single-process oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """numpy model of CudaBackend (same method contract, same buffer layouts)."""

    def __init__(self, orc, N, L, k_axis, edges, centres, nranks, rank):
        self.orc, self.N, self.L, self.k, self.edges, self.centres = orc, N, L, k_axis, edges, centres
        self.P, self.rank = nranks, rank

    def grid_slab(self, pos, vel, rho, ax_loc, ax, lcell3, keep):
        lo, hi, open_lo, open_hi = keep
        p = pos.numpy().astype(np.float64)
        sel = np.ones(len(p), bool)
        if not open_lo:
            sel &= p[:, 0] >= lo
        if not open_hi:
            sel &= p[:, 0] <= hi
        ids = np.nonzero(sel)[0]
        idx = ids[self.orc.nn_exact_lattice(p[ids], ax_loc, ax, ax)]
        # a node is proven only if its nearest kept particle is nearer than every closed face of the kept range
        g = np.meshgrid(ax_loc, ax, ax, indexing="ij")
        d = np.sqrt(((np.stack(g, -1) - p[idx]) ** 2).sum(-1))
        margin = np.full(d.shape, np.inf)
        if not open_lo:
            margin = np.minimum(margin, g[0] - lo)
        if not open_hi:
            margin = np.minimum(margin, hi - g[0])
        unresolved = int((d >= margin).sum())
        v = vel.numpy().astype(np.float64)
        r = rho.numpy().astype(np.float64)
        vv = (v * r[:, None]) / r[:, None]
        return (idx, vv, r * lcell3), unresolved

    def fields(self, gridded, quantity, strict):
        idx, v, m = gridded
        vg, mg = v[idx], m[idx]
        comps = self.orc.field_components(vg, mg, quantity, strict_reference=False)
        if quantity == "momentum" and strict:
            return [comps[0]], 3.0
        return comps, 1.0

    def fft_local(self, slabs):
        import torch
        N, P = self.N, self.P
        kzc = N // 2 // P
        out = []
        for s in slabs:
            R = np.fft.rfft(np.asarray(s, dtype=np.float64), axis=2)
            packed = R[..., : N // 2].copy()
            packed[..., 0] = R[..., 0].real + 1j * R[..., N // 2].real          # DESIGN.md "half-spectrum layout"
            Y = np.fft.fft(packed, axis=1)
            send = np.stack([Y[:, :, d * kzc:(d + 1) * kzc] for d in range(P)], axis=0)
            out.append(torch.from_numpy(np.ascontiguousarray(send).astype(np.complex64)))
        return out

    def fft_final(self, recv):
        import torch
        N, P, rank = self.N, self.P, self.rank
        kzc = N // 2 // P
        k2 = self.k ** 2
        psum = np.zeros(len(self.edges) - 1)
        ns = np.zeros(len(self.edges) - 1, dtype=np.int64)
        kz = np.arange(rank * kzc, (rank + 1) * kzc)
        kmag = np.sqrt((k2[:, None, None] + k2[None, :, None]) + k2[kz][None, None, :])
        Pw = np.zeros((N, N, kzc))
        planes = []
        for r in recv:
            X = np.fft.fft(r.numpy().astype(np.complex128), axis=0)
            Pw += np.abs(X) ** 2
            planes.append(X[:, :, 0])
        on = np.ones(kzc, bool)
        if rank == 0:
            on[0] = False
        h, _ = np.histogram(kmag[:, :, on].ravel(), bins=self.edges, weights=2 * Pw[:, :, on].ravel())
        c, _ = np.histogram(kmag[:, :, on].ravel(), bins=self.edges)
        psum += h
        ns += 2 * c
        if rank == 0:
            idx = (-np.arange(N)) % N
            pa, pb = np.zeros((N, N)), np.zeros((N, N))
            for Z in planes:
                Zm = np.conj(Z[np.ix_(idx, idx)])
                pa += np.abs(0.5 * (Z + Zm)) ** 2
                pb += np.abs((Z - Zm) / 2j) ** 2
            w = k2[:, None] + k2[None, :]
            for pw, kzv in ((pa, k2[0]), (pb, k2[N // 2])):
                km = np.sqrt(w + kzv).ravel()
                psum += np.histogram(km, bins=self.edges, weights=pw.ravel())[0]
                ns += np.histogram(km, bins=self.edges)[0]
        return torch.from_numpy(psum), torch.from_numpy(ns)


def _worker(rank, world, port, N, Np, halo, q, ret, sharded=False):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "large-velocity-power-spectrum_b200"))
    import vpower_oracle as orc
    from vpower import dist as vd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = 1.0
    pos, vel, dens, _ = orc.synth_particles(4, Np, L)
    ax, k = orc.lattice_axis_lib(L, N), orc.k_axis(L, N)
    centres, edges = orc.edges_lib(2 * np.pi / L, np.pi * N / L, 2 * np.pi / L)
    a = (L / (2 * np.pi)) ** 1.5 / N ** 3
    be = OracleBackend(orc, N, L, k, edges, centres, world, rank)
    tm = {}
    sl = slice(rank * Np // world, (rank + 1) * Np // world) if sharded else slice(None)
    out, ns = vd.particles_to_pk_dist(torch.from_numpy(pos[sl]), torch.from_numpy(vel[sl]), torch.from_numpy(dens[sl]), ax,
                                      (L / N) ** 3, 0.5 * a * a, k, edges, quantities=q, backend=be, halo_cells=halo, timings=tm,
                                      sharded=sharded)
    if rank == 0:
        ret["out"], ret["ns"], ret["halo"] = {k_: v.tolist() for k_, v in out.items()}, ns.tolist(), tm["halo_cells"]
    dist.destroy_process_group()


@pytest.mark.parametrize("world,halo,sharded", [(2, 4, False), (4, 1, False), (2, 0.25, False), (2, 2, True), (4, 0.25, True)])
def test_slab_pipeline_matches_single_process_oracle(orc, world, halo, sharded):
    import torch.multiprocessing as mp
    N, Np = 16, 3000
    q = ("velocity", "momentum", "energy")
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + world + (10 if sharded else 0)
    mp.spawn(_worker, args=(world, port, N, Np, halo, q, ret, sharded), nprocs=world, join=True)
    pos, vel, dens, _ = orc.synth_particles(4, Np, 1.0)
    v, m, Lcell = orc.ann_interp_to_field(pos.astype(np.float64), dens.astype(np.float64), vel.astype(np.float64), 1.0, N)
    for name in q:
        ref = orc.spctrm(v, m, Lcell, name)
        assert ret["ns"] == ref[:, 3].astype(np.int64).tolist()
        assert np.allclose(ret["out"][name], ref[:, 2], rtol=1e-5)
    if halo < 1:
        assert ret["halo"] > halo          # the widening loop had to run


def test_slab_bounds():
    sys.path.insert(0, os.path.join(ROOT, "large-velocity-power-spectrum_b200"))
    from vpower import dist as vd
    assert vd.slab_bounds(1024, 8, 3) == (384, 512, 192, 256)
    assert vd.slab_bounds(64, 1, 0) == (0, 64, 0, 32)
    with pytest.raises(ValueError):
        vd.slab_bounds(24, 5, 0)
    ax = np.linspace(0.5, 15.5, 16)
    lo, hi, ol, oh = vd.keep_range(ax, 4, 8, 4, 1, 2)
    assert np.isclose(lo, 4.5 - 2.0) and np.isclose(hi, 7.5 + 2.0) and not ol and not oh
    assert vd.keep_range(ax, 0, 4, 4, 0, 2)[2] and vd.keep_range(ax, 12, 16, 4, 3, 2)[3]
