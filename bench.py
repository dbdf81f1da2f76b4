#!/usr/bin/env python
"""bench.py -- particles -> P(k) hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg3|cfg2|cfg1] [--impl reference]

One "step" = one pass of the whole hot path (nearest-particle gridding, field algebra, 3-D FFTs, |F|^2,
k-shell binning) over one synthetic particle set, for velocity + momentum + kinetic-energy spectra.
Prints ONE JSON line (rank 0).  `value` = particles per second through the whole path with the particle
arrays resident in HBM; `e2e` = the same through the host-buffer C-ABI call (pinned host arrays, H2D and
D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "large-velocity-power-spectrum_b200"))

WORKLOADS = {  # BASELINE.json configs
    "cfg1": dict(N=64, Np=1 << 18, quantities=("velocity",)),
    "cfg2": dict(N=256, Np=1 << 24, quantities=("velocity", "momentum")),
    "cfg3": dict(N=512, Np=1 << 27, quantities=("energy",)),
    "cfg4": dict(N=1024, Np=1 << 30, quantities=("velocity", "momentum", "energy")),
}
METRIC = "P(k) end-to-end throughput, particles -> binned velocity+momentum+KE spectra (Gpart/s = Np / end-to-end s)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def synth_on_device(torch, Np, seed, L=1.0):
    """Synthetic particle set generated on the device (uniform positions, smooth flow + noise, ~uniform density)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(1234 + seed)
    pos = torch.rand((Np, 3), generator=g, device="cuda", dtype=torch.float32) * L
    vel = torch.empty((Np, 3), device="cuda", dtype=torch.float32)
    rs = np.random.default_rng(seed)
    chunk = 1 << 26
    for s in range(0, Np, chunk):
        x = pos[s:s + chunk]
        v = 0.1 * 3.4641 * (torch.rand(x.shape, generator=g, device="cuda", dtype=torch.float32) - 0.5)
        for _ in range(4):
            kv = torch.tensor(rs.integers(1, 7, size=3), device="cuda", dtype=torch.float32)
            a = torch.tensor(rs.normal(size=3) / 4.0, device="cuda", dtype=torch.float32)
            ph = float(rs.uniform(0, 2 * np.pi))
            v += a[None, :] * torch.sin(2 * np.pi * (x @ kv) / L + ph)[:, None]
        vel[s:s + chunk] = v
    rho = (Np / L ** 3) * (1.0 + 0.1 * torch.rand(Np, generator=g, device="cuda", dtype=torch.float32))
    return pos, vel, rho


def synth_clustered_on_device(torch, n_lat, seed, L=1.0, rms_cells=2.0):
    """cfg3: n_lat^3 lattice particles displaced by a sum of long-wave sinusoids (Zel'dovich-like), rms ~2 cells:
    voids and dense sheets -> exercises the wide stages of the nearest-particle search."""
    rs = np.random.default_rng(seed)
    g1 = (torch.arange(n_lat, device="cuda", dtype=torch.float32) + 0.5) / n_lat
    q = torch.stack(torch.meshgrid(g1, g1, g1, indexing="ij"), dim=-1).reshape(-1, 3)
    disp = torch.zeros_like(q)
    for _ in range(64):     # wavenumbers up to 48: displacement gradients of order one -> shell crossing, voids, sheets
        kv = torch.tensor(rs.integers(1, 49, size=3) * rs.choice([-1, 1], size=3), device="cuda", dtype=torch.float32)
        kk = float(torch.linalg.norm(kv))
        ph = float(rs.uniform(0, 2 * np.pi))
        disp += (1.0 / kk) * (kv / kk)[None, :] * torch.sin(2 * np.pi * (q @ kv) + ph)[:, None]
    disp *= (rms_cells / n_lat) / float(torch.sqrt((disp ** 2).sum(1).mean()))
    pos = torch.remainder(q + disp, 1.0) * L
    del q, disp
    g = torch.Generator(device="cuda")
    g.manual_seed(99 + seed)
    Np = pos.shape[0]
    vel = torch.randn((Np, 3), generator=g, device="cuda", dtype=torch.float32)
    rho = (Np / L ** 3) * (1.0 + 0.1 * torch.rand(Np, generator=g, device="cuda", dtype=torch.float32))
    return pos.contiguous(), vel, rho


def geometry(orc_like, N, L):
    Lcell = L / N
    ax = np.linspace(Lcell / 2, L + Lcell / 2, N)                         # library lattice, interp.py:1063
    k = 2 * np.pi * np.fft.fftfreq(N, Lcell)
    kmin, kmax = 2 * np.pi / L, np.pi / Lcell
    edges = np.arange(kmin - kmin / 2, kmax + 3 * kmin / 2, kmin)         # interp.py:1473
    a = (L / (2 * np.pi)) ** 1.5 / N ** 3
    return ax, k, edges, Lcell ** 3, 0.5 * a * a


def cpu_reference_run(N, Np, quantities, seed=0):
    """The path on host cores through the oracle port (kd-tree search, pocketfft, numpy.histogram), all threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vpower_oracle as orc
    pos, vel, dens, _ = orc.synth_particles(seed, Np, 1.0)
    p64, v64, d64 = pos.astype(np.float64), vel.astype(np.float64), dens.astype(np.float64)
    t0 = time.perf_counter()
    v, m, Lcell = orc.ann_interp_to_field(p64, d64, v64, 1.0, N)
    t1 = time.perf_counter()
    out = [orc.spctrm(v, m, Lcell, q) for q in quantities]
    t2 = time.perf_counter()
    return {"s_total": t2 - t0, "s_nn": t1 - t0, "s_spectra": t2 - t1, "nbins": len(out[0])}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload or "cfg4"]
    sN, sNp = 128, 1 << 21        # bounded sample: same density (1 particle per node) and the same quantities
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_run(sN, sNp, wl["quantities"], seed=i)
        if i >= args.warmup:
            times.append(r["s_total"])
    t = float(np.mean(times))
    val = sNp / t / 1e9
    cores = os.cpu_count()
    sample = f"{sN}^3 lattice / 2^21 particles per step, quantities {'+'.join(wl['quantities'])}, oracle port (cKDTree workers=-1, scipy.fft workers=-1)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gpart/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload or 'cfg4'} ({wl['N']}^3 lattice, {wl['Np']} particles) -- timed on a bounded sample",
                       "sample": sample},
            "cpu_baseline": {"value": val, "unit": "Gpart/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Gpart/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL all-to-all instead of the fused peer-store transpose")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import __graft_entry__ as ge
    ge.build()
    from vpower import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL prints its version banner to stdout when the first communicator
        # comes up, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.all_reduce(torch.zeros(1, device="cuda"))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    free_b, total_b = torch.cuda.mem_get_info()
    wname = args.workload or ("cfg4" if free_b > 140e9 else "cfg2")
    wl = WORKLOADS[wname]
    N, Np, quantities = wl["N"], wl["Np"], ("velocity", "momentum", "energy")
    L = 1.0
    ax, k, edges, lc3, norm = geometry(None, N, L)
    hbm_peak, peak_src = peaks()

    if wname == "cfg3":
        pos, vel, rho = synth_clustered_on_device(torch, round(Np ** (1 / 3)), seed=2)
    else:
        pos, vel, rho = synth_on_device(torch, Np, seed=3)
    torch.cuda.synchronize()

    if world > 1:
        from vpower import dist as vd
        backend = vd.CudaBackend(N, k, edges, world, rank, p2p=not args.no_p2p)
        # sharded input: rank r owns particles [r*Np/P, (r+1)*Np/P) of the synthetic set; the slab exchange is timed
        lo_i, hi_i = rank * (Np // world), (rank + 1) * (Np // world)
        pos, vel, rho = pos[lo_i:hi_i].clone(), vel[lo_i:hi_i].clone(), rho[lo_i:hi_i].clone()
        torch.cuda.empty_cache()

    phase_t = {}

    def step_dev():
        if world > 1:      # x-slab gridding, slab FFT with one all-to-all per field, all-reduce of the shells
            return vd.particles_to_pk_dist(pos, vel, rho, ax, lc3, norm, k, edges, quantities=quantities, backend=backend,
                                           sharded=True, timings=phase_t)
        return _lib.particles_to_pk(pos, vel, rho, ax, ax, ax, N, lc3, norm, k, edges, quantities=quantities)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        out, ns = step_dev()
    barrier()
    _lib.profile_enable(True)
    _lib.profile_report()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            out, ns = step_dev()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (_lib.launch_count() - l0) // args.steps
    stages = _lib.profile_report()
    _lib.profile_enable(False)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = Np / (ms * 1e-3) / 1e9

    # per-stage roofline (algorithmic bytes stated by the library per launch / CUDA-event time of the stage)
    table = {}
    for name, s in stages.items():
        per_ms = s["ms"] / args.steps
        gbs = (s["bytes"] / args.steps) / (per_ms * 1e-3) / 1e9 if per_ms > 0 and s["bytes"] > 0 else None
        table[name] = {"ms_per_step": round(per_ms, 4), "launches_per_step": s["launches"] // args.steps,
                       "alg_GB_per_step": round(s["bytes"] / args.steps / 1e9, 4),
                       "GBps": None if gbs is None else round(gbs, 1), "frac": None if gbs is None else round(gbs / hbm_peak, 4)}
    dom = max(table, key=lambda n: table[n]["ms_per_step"])
    kernel_ms = sum(v["ms_per_step"] for v in table.values())
    # DRAM bytes of that kernel's launch from a committed `ncu --set full` capture of the same workload (single GPU only)
    traffic, traffic_src = None, None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(wname, {}) if world == 1 else {}
        if dom in cap:
            traffic, traffic_src = float(cap[dom]), cap.get("source")
    except (OSError, ValueError):
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": table[dom]["GBps"], "peak": hbm_peak, "unit": "GB/s",
            "frac": table[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "share_of_step": round(table[dom]["ms_per_step"] / kernel_ms, 4) if kernel_ms > 0 else None,
            "alg_bytes_per_launch": table[dom]["alg_GB_per_step"] * 1e9 / max(1, table[dom]["launches_per_step"])}

    # end to end through the host-buffer C ABI: pinned host arrays in, spectra out
    e2e = None
    if not args.no_e2e and rank == 0 and world == 1:
        try:
            hp = torch.empty(pos.shape, dtype=pos.dtype, pin_memory=True).copy_(pos)
            hv = torch.empty(vel.shape, dtype=vel.dtype, pin_memory=True).copy_(vel)
            hr = torch.empty(rho.shape, dtype=rho.dtype, pin_memory=True).copy_(rho)
            torch.cuda.synchronize()
            a_pos, a_vel, a_rho = hp.numpy(), hv.numpy(), hr.numpy()
            del pos, vel, rho
            torch.cuda.empty_cache()
            n_e2e = max(1, min(args.steps, 3))
            _lib.particles_to_pk(a_pos, a_vel, a_rho, ax, ax, ax, N, lc3, norm, k, edges, quantities=quantities)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                out2, ns2 = _lib.particles_to_pk(a_pos, a_vel, a_rho, ax, ax, ax, N, lc3, norm, k, edges, quantities=quantities)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n_e2e
            e2e = {"value": Np / dt / 1e9, "unit": "Gpart/s", "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": int(a_pos.nbytes + a_vel.nbytes + a_rho.nbytes),
                   "d2h_bytes_per_step": int(len(quantities) * (len(edges) - 1) * 16), "steps": n_e2e,
                   "api": "vp_host_particles_to_pk (pinned host arrays)"}
            assert np.array_equal(ns2, ns)
        except Exception as ex:  # noqa: BLE001
            e2e = {"value": None, "unit": "Gpart/s", "error": str(ex)[:200]}

    if not args.no_e2e and world > 1:
        # multi-GPU end to end: every rank's shard starts in pinned host memory; H2D + exchange + path inside the timed region
        try:
            hp = torch.empty(pos.shape, dtype=pos.dtype, pin_memory=True).copy_(pos)
            hv = torch.empty(vel.shape, dtype=vel.dtype, pin_memory=True).copy_(vel)
            hr = torch.empty(rho.shape, dtype=rho.dtype, pin_memory=True).copy_(rho)
            del pos, vel, rho
            torch.cuda.empty_cache()

            def step_e2e():
                dp, dv, dr = hp.cuda(non_blocking=True), hv.cuda(non_blocking=True), hr.cuda(non_blocking=True)
                return vd.particles_to_pk_dist(dp, dv, dr, ax, lc3, norm, k, edges, quantities=quantities, backend=backend, sharded=True)

            step_e2e()
            n_e2e = max(1, min(args.steps, 3))
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                out2, ns2 = step_e2e()
            barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], device="cuda", dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dt = float(dt.item())
            e2e = {"value": Np / dt / 1e9, "unit": "Gpart/s", "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": int(world * (hp.numel() + hv.numel() + hr.numel()) * 4),
                   "d2h_bytes_per_step": int(len(quantities) * (len(edges) - 1) * 16), "steps": n_e2e,
                   "api": "vpower.dist.particles_to_pk_dist (sharded, pinned host shards per rank)"}
        except Exception as ex:  # noqa: BLE001
            e2e = {"value": None, "unit": "Gpart/s", "error": str(ex)[:200]}

    cpu = None
    if not args.no_cpu and rank == 0:
        sN, sNp = 128, 1 << 21
        r = cpu_reference_run(sN, sNp, quantities)
        cpu = {"value": sNp / r["s_total"] / 1e9, "unit": "Gpart/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{sN}^3 lattice / 2^21 particles, V+M+KE, oracle port (cKDTree workers=-1, scipy.fft workers=-1); "
                         f"nn {r['s_nn']:.2f}s spectra {r['s_spectra']:.2f}s"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Gpart/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{wname}: {N}^3 lattice, {Np} particles (2^{int(np.log2(Np))}), velocity+momentum+energy P(k), "
                                       f"library lattice and edges", "l2": "inputs larger than L2" if Np * 28 > 2e8 else "inputs fit L2",
                           "momentum": "reference-strict (vx*m x3)"},
                "nn_stats": _lib.nn_grid_stats() if world == 1 else None, "seconds_per_pk": ms * 1e-3, "nn_gridding_gpart_s": None, "gpu_launches": int(launches),
                "roofline": roof, "stages": table, "dist_phases_ms_last_step": phase_t.get("phases_ms"), "cpu_baseline": cpu, "e2e": e2e, "clocks": clk.summary()}
        nn_ms = sum(table[n]["ms_per_step"] for n in table if n.startswith("k1"))
        if nn_ms > 0:
            line["nn_gridding_gpart_s"] = Np / (nn_ms * 1e-3) / 1e9
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
