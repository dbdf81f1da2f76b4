"""Multi-GPU check, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py

N-rank result (NCCL all-to-all + all-reduce) vs the single-GPU path on rank 0 and vs the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "large-velocity-power-spectrum_b200"))
import vpower_oracle as orc  # noqa: E402
from vpower import _lib, dist as vd  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for N, Np in ((256, 1 << 22), (512, 1 << 24)):
        L = 1.0
        pos, vel, dens, _ = orc.synth_particles(8, Np, L)
        ax, k = orc.lattice_axis_lib(L, N), orc.k_axis(L, N)
        centres, edges = orc.edges_lib(2 * np.pi / L, np.pi * N / L, 2 * np.pi / L)
        a = (L / (2 * np.pi)) ** 1.5 / N ** 3
        d = [_lib.to_device(x) for x in (pos, vel, dens)]
        qs = ("velocity", "momentum", "energy")
        out, ns = vd.particles_to_pk_dist(*d, ax, (L / N) ** 3, 0.5 * a * a, k, edges, quantities=qs)
        # sharded input: every rank passes only its own part of the particle list
        sl = slice(rank * Np // world, (rank + 1) * Np // world)
        out_s, ns_s = vd.particles_to_pk_dist(*[t[sl].contiguous() for t in d], ax, (L / N) ** 3, 0.5 * a * a, k, edges,
                                              quantities=qs, sharded=True)
        # exchange fused into the y pass (CUDA IPC peer stores) instead of the NCCL all-to-all
        be = vd.CudaBackend(N, k, edges, world, rank, p2p=True)
        out_p, ns_p = vd.particles_to_pk_dist(*d, ax, (L / N) ** 3, 0.5 * a * a, k, edges, quantities=qs, backend=be)
        sl = slice(rank * Np // world, (rank + 1) * Np // world)      # sharded input + peer-store particle exchange, twice
        shard = [t[sl].contiguous() for t in d]
        vd.particles_to_pk_dist(*shard, ax, (L / N) ** 3, 0.5 * a * a, k, edges, quantities=qs, backend=be, sharded=True)
        out_p2, ns_p2 = vd.particles_to_pk_dist(*shard, ax, (L / N) ** 3, 0.5 * a * a, k, edges, quantities=qs, backend=be,
                                                sharded=True)
        if rank == 0:
            same_p = all(np.array_equal(x, ns) for x in (ns_p, ns_p2)) and all(
                np.allclose(o[q], out[q], rtol=1e-12) for o in (out_p, out_p2) for q in qs)
            print(f"N={N} world={world}: fused peer-store transpose and particle exchange == NCCL paths: {same_p}", flush=True)
            ok &= same_p
        be.close()                                                    # collective: unmap peers, barrier, free
        del be
        if rank == 0:
            same_s = np.array_equal(ns_s, ns) and all(np.allclose(out_s[q], out[q], rtol=1e-12) for q in qs)
            print(f"N={N} world={world}: sharded input == replicated input: {same_s}", flush=True)
            ok &= same_s
            ref, ref_ns = _lib.particles_to_pk(*d, ax, ax, ax, N, (L / N) ** 3, 0.5 * a * a, k, edges, quantities=qs)
            same = np.array_equal(ns, ref_ns)
            err = max(np.max(np.abs(out[q] / ref[q] - 1)) for q in qs)
            print(f"N={N} world={world}: Nsample identical={same}, max rel dPsum vs 1 GPU = {err:.2e}", flush=True)
            ok &= same and err < 1e-6
            if N == 256:
                v, m, Lcell = orc.ann_interp_to_field(pos.astype(np.float64), dens.astype(np.float64), vel.astype(np.float64), L, N)
                for q in qs:
                    r = orc.spctrm(v, m, Lcell, q)
                    e = np.max(np.abs(out[q] / r[:, 2] - 1))
                    print(f"   {q}: vs CPU oracle Nsample exact={np.array_equal(ns, r[:, 3].astype(np.int64))} max rel dPsum={e:.2e}", flush=True)
                    ok &= np.array_equal(ns, r[:, 3].astype(np.int64)) and e < 1e-4
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST CHECK", "PASS" if int(flag.item()) else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
