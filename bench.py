#!/usr/bin/env python
"""bench.py -- particles -> P(k) hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg2|cfg3|cfg4|cfg5] [--impl reference]

One "step" = one pass of the whole hot path (nearest-particle gridding, field algebra, 3-D FFTs, |F|^2,
k-shell binning) over one synthetic particle set, for the quantities the BASELINE.json config names.
Prints ONE JSON line (rank 0).  `value` = particles per second through the whole path with the particle
arrays resident in HBM; `e2e` = the same through the host-buffer C-ABI call (pinned host arrays, H2D and
D2H inside the timed region).  `result` is a digest of what was computed (mode counts, CRC, rounded shell sums):
the same workload gives the same digest at every GPU count, and the run fails if the mode counts differ from the
integer-shell closed form or a lattice node was left unproven.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "large-velocity-power-spectrum_b200"))

WORKLOADS = {  # BASELINE.json configs
    "cfg1": dict(N=64, Np=1 << 18, quantities=("velocity",), gen="uniform", seed=0, min_gpus=1),
    "cfg2": dict(N=256, Np=1 << 24, quantities=("velocity", "momentum"), gen="uniform", seed=1, min_gpus=1),
    "cfg3": dict(N=512, Np=1 << 27, quantities=("energy",), gen="clustered", seed=2, min_gpus=1),
    "cfg4": dict(N=1024, Np=1 << 30, quantities=("velocity", "momentum", "energy"), gen="uniform", seed=3, min_gpus=1),
    "cfg5": dict(N=2048, Np=1 << 33, quantities=("velocity",), gen="uniform", seed=4, min_gpus=8),
}
METRIC = "P(k) end-to-end throughput, particles -> binned spectra of the config's quantities (Gpart/s = Np / end-to-end s)"
M32 = 0xFFFFFFFF


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


# ----------------------------------------------------------------------------------------- synthetic particle sets
# Counter-based: particle i of a workload is a pure function of (seed, i), so a rank can generate exactly its own index
# range [lo, hi) on its own device and the union over ranks is the same set at every GPU count (cfg5: no rank ever holds
# 2^33 particles).  The integer hash is the oracle's hash_uniform (tests/test_host.py checks the two against each other).
def _mix32_int(x):
    x &= M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def _mix32_t(x):
    x = x & M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & M32          # int64 product wraps; the low 32 bits are exact
    x = x ^ (x >> 16)
    return x


def hash_uniform_t(torch, seed, lo, hi, stream, device):
    """uniforms in [0,1) with a 24-bit mantissa for particle indices [lo, hi) -- f32 tensor."""
    i = torch.arange(lo, hi, dtype=torch.int64, device=device)
    c = _mix32_int(seed * 0x9E3779B1 + stream * 0x85EBCA77)
    h = _mix32_t(i ^ c)
    h = _mix32_t(h + (i >> 32) + stream)
    return (h >> 8).to(torch.float32) * (1.0 / 16777216.0)


def synth_range(torch, wl, lo, hi, device="cuda", L=1.0, chunk=1 << 25):
    """Particles [lo, hi) of workload `wl` -> (pos [n,3], vel [n,3], rho [n]) f32 on `device`."""
    seed, Np, kind = wl["seed"], wl["Np"], wl["gen"]
    n = hi - lo
    pos = torch.empty((n, 3), dtype=torch.float32, device=device)
    vel = torch.empty((n, 3), dtype=torch.float32, device=device)
    rho = torch.empty(n, dtype=torch.float32, device=device)
    rs = np.random.default_rng(1000 + seed)
    modes = [(rs.integers(1, 7, size=3).astype(np.float32), (rs.normal(size=3) / 4.0).astype(np.float32), float(rs.uniform(0, 2 * np.pi)))
             for _ in range(4)]
    nblob = 8
    centres = rs.uniform(0.15, 0.85, size=(nblob, 3)).astype(np.float32)
    sigmas = (rs.uniform(0.01, 0.06, size=nblob)).astype(np.float32)
    for s in range(lo, hi, chunk):
        e = min(hi, s + chunk)
        x = torch.stack([hash_uniform_t(torch, seed, s, e, c, device) for c in range(3)], dim=1)
        if kind == "clustered":
            # every second particle is pulled into one of 8 Gaussian blobs (sigma 1-6 % of the box): dense cells holding
            # tens of particles next to half-empty voids -> the wide stages of the nearest-particle search matter
            i = torch.arange(s, e, dtype=torch.int64, device=device)
            inblob = (i & 1) == 1
            b = ((i >> 1) % nblob)
            u1 = hash_uniform_t(torch, seed, s, e, 8, device).clamp_min(2.0 ** -24)
            r = torch.sqrt(-2.0 * torch.log(u1))
            g = torch.stack([r * torch.cos(2 * np.pi * x[:, 0]), r * torch.sin(2 * np.pi * x[:, 0]),
                             torch.sqrt(-2.0 * torch.log(x[:, 1].clamp_min(2.0 ** -24))) * torch.cos(2 * np.pi * x[:, 2])], dim=1)
            cb = torch.from_numpy(centres).to(device)[b]
            sb = torch.from_numpy(sigmas).to(device)[b]
            xb = torch.remainder(cb + sb[:, None] * g, 1.0)
            x = torch.where(inblob[:, None], xb, x)
            del i, b, u1, r, g, cb, sb, xb
        v = 0.1 * 3.4641016 * (torch.stack([hash_uniform_t(torch, seed, s, e, 3 + c, device) for c in range(3)], dim=1) - 0.5)
        for kv, a, ph in modes:
            arg = 2 * np.pi * (x @ torch.from_numpy(kv).to(device)) + ph
            v += torch.from_numpy(a).to(device)[None, :] * torch.sin(arg)[:, None]
        pos[s - lo:e - lo] = x * L
        vel[s - lo:e - lo] = v
        rho[s - lo:e - lo] = (Np / L ** 3) * (1.0 + 0.1 * hash_uniform_t(torch, seed, s, e, 7, device))
        del x, v
    return pos, vel, rho


def geometry(N, L):
    Lcell = L / N
    ax = np.linspace(Lcell / 2, L + Lcell / 2, N)                         # library lattice, interp.py:1063
    k = 2 * np.pi * np.fft.fftfreq(N, Lcell)
    kmin, kmax = 2 * np.pi / L, np.pi / Lcell
    edges = np.arange(kmin - kmin / 2, kmax + 3 * kmin / 2, kmin)         # interp.py:1473
    a = (L / (2 * np.pi)) ** 1.5 / N ** 3
    return ax, k, edges, Lcell ** 3, 0.5 * a * a


def integer_shell_counts(N):
    """Nsample of the library edges in closed form: modes n in [-N/2, N/2)^3 with floor(|n| + 1/2) == j, j = 1..N/2
    (SURVEY.md App. B2/B5).  Exact integer arithmetic: the number of ways to write s as a sum of three squares of axis
    indices, by two polynomial products."""
    import scipy.signal
    h = N // 2
    ax = np.arange(-h, h, dtype=np.int64) ** 2
    c1 = np.bincount(ax, minlength=h * h + 1).astype(np.float64)
    c2 = np.rint(scipy.signal.fftconvolve(c1, c1))
    c3 = np.rint(scipy.signal.fftconvolve(c2, c1)).astype(np.int64)
    assert int(c3.sum()) == N ** 3
    cum = np.concatenate([[0], np.cumsum(c3)])
    j = np.arange(1, h + 1, dtype=np.int64)
    lo, hi = j * j - j + 1, np.minimum(j * j + j, len(c3) - 1)
    return cum[hi + 1] - cum[lo]


def result_digest(out, ns, quantities):
    ns = np.asarray(ns, dtype=np.int64)
    d = {"nsample_sum": int(ns.sum()), "nsample_crc32": int(zlib.crc32(ns.astype("<i8").tobytes())), "psum": {}}
    for q in quantities:
        p = np.asarray(out[q], dtype=np.float64)
        mid = len(p) // 2
        d["psum"][q] = {"sum": float(f"{p.sum():.9e}"), "first": float(f"{p[0]:.9e}"), "mid": float(f"{p[mid]:.9e}"),
                        "last": float(f"{p[-1]:.9e}")}
    return d


# ----------------------------------------------------------------------------------------- CPU reference arm
def cpu_port_run(N, Np, quantities, seed=0):
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


def cpu_verbatim_script_run(N, Np, seed=0):
    """The UNMODIFIED reference MPI script (scripts/parallel_optimized.py main(), single rank) through the import shims
    of oracle/refshims.py -- only where the reference tree exists (the build container; never the GPU box)."""
    import contextlib
    import io
    import runpy
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refshims
    import vpower_oracle as orc
    refshims.install(engine="oracle")
    pos, vel, dens, mass = orc.synth_particles(seed, Np, 1.0)
    with tempfile.TemporaryDirectory() as td:
        snap = os.path.join(td, "snap.hdf5")
        refshims.register_snapshot(snap, pos.copy(), mass.copy(), dens.copy(), vel.copy())
        argv = sys.argv
        sys.argv = ["parallel_optimized.py", "-i", snap, "-o", td, "-N", str(N), "-M", str(N), "-b", str(1 << 15), "-f"]
        try:
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                runpy.run_path(os.path.join(refshims.REF_ROOT, "scripts", "parallel_optimized.py"), run_name="__main__")
            dt = time.perf_counter() - t0
        finally:
            sys.argv = argv
        pk = np.loadtxt(os.path.join(td, "Pk.txt"))
    return {"s_total": dt, "nbins": len(pk)}


def reference_tree():
    for p in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("VPOWER_REFERENCE", "/root/reference")):
        if p and os.path.isfile(os.path.join(p, "scripts", "parallel_optimized.py")):
            return p
    return None


def run_reference(args):
    """CPU arm.  cfg1 / cfg2: the SAME workload in full (same_config true) through the oracle port with every host thread;
    where the reference tree is present (build container) cfg1 also runs the unmodified MPI script.  cfg3-5: a bounded
    sample at the same density, extrapolated per particle and labelled so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wname = args.workload or "cfg4"
    wl = WORKLOADS[wname]
    full = wname in ("cfg1", "cfg2")
    sN, sNp = (wl["N"], wl["Np"]) if full else (128, 1 << 21)
    times, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_port_run(sN, sNp, wl["quantities"], seed=wl["seed"] if full else i)
        if i >= args.warmup:
            times.append(last["s_total"])
    t = float(np.mean(times))
    val = sNp / t / 1e9
    cores = os.cpu_count()
    sample = (f"{sN}^3 lattice / {sNp} particles per step ({'the full workload' if full else 'bounded sample, same density'}), "
              f"quantities {'+'.join(wl['quantities'])}, oracle port (cKDTree workers=-1, scipy.fft workers=-1); "
              f"last step: nn {last['s_nn']:.2f}s spectra {last['s_spectra']:.2f}s")
    cpu = {"value": val, "unit": "Gpart/s", "cores": cores, "kind": "port", "sample": sample, "same_config": full,
           "extrapolated": not full}
    verb = None
    if wname == "cfg1" and reference_tree():
        try:
            r = cpu_verbatim_script_run(wl["N"], wl["Np"], seed=wl["seed"])
            verb = {"kind": "reference-verbatim", "value": wl["Np"] / r["s_total"] / 1e9, "unit": "Gpart/s", "cores": 1,
                    "seconds": r["s_total"], "sample": "scripts/parallel_optimized.py main() unmodified, single rank, cfg1 in full "
                    "(annoy -> exact search, pyfftw -> scipy.fft, mpi4py -> one rank)"}
        except Exception as ex:  # noqa: BLE001
            verb = {"kind": "reference-verbatim", "error": str(ex)[:200]}
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gpart/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{wname} ({wl['N']}^3 lattice, {wl['Np']} particles)" + ("" if full else " -- timed on a bounded sample"),
                       "sample": sample, "same_config": full, "extrapolated": not full},
            "cpu_baseline": cpu, "reference_verbatim": verb,
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
    if world < wl["min_gpus"]:
        raise SystemExit(f"bench.py: workload {wname} needs at least {wl['min_gpus']} GPUs")
    N, Np, quantities = wl["N"], wl["Np"], wl["quantities"]
    L = 1.0
    ax, k, edges, lc3, norm = geometry(N, L)
    hbm_peak, peak_src = peaks()

    # every rank generates its own index range of the workload's particle set (sharded input)
    lo_i, hi_i = rank * (Np // world), (rank + 1) * (Np // world)
    pos, vel, rho = synth_range(torch, wl, lo_i, hi_i)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    if world > 1:
        from vpower import dist as vd
        backend = vd.CudaBackend(N, k, edges, world, rank, p2p=not args.no_p2p)

    phase_t = {}
    unresolved = [0]

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
    free_e, total_e = torch.cuda.mem_get_info()       # arena and caching allocator are grow-only: in use now = high-water mark
    peak_mem = total_e - free_e
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, float(peak_mem)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, peak_mem = float(t[0].item()), float(t[1].item())
    value = Np / (ms * 1e-3) / 1e9
    nn_stats = _lib.nn_grid_stats() if world == 1 else {"n_unresolved": 0, "halo_cells": phase_t.get("halo_cells")}

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
    # DRAM bytes of that kernel's launch from a committed `ncu --set full` capture of the same workload
    traffic, traffic_src = None, None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(wname if world == 1 else f"{wname}_{world}gpu", {})
        if dom in cap:
            traffic, traffic_src = float(cap[dom]), cap.get("source")
    except (OSError, ValueError):
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": table[dom]["GBps"], "peak": hbm_peak, "unit": "GB/s",
            "frac": table[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "share_of_step": round(table[dom]["ms_per_step"] / kernel_ms, 4) if kernel_ms > 0 else None,
            "alg_bytes_per_launch": table[dom]["alg_GB_per_step"] * 1e9 / max(1, table[dom]["launches_per_step"])}

    # ---- what was computed: digest + hard checks
    digest = result_digest(out, ns, quantities)
    closed = integer_shell_counts(N)
    digest["nsample_closed_form_ok"] = bool(np.array_equal(np.asarray(ns, dtype=np.int64), closed))
    digest["n_unresolved"] = int(nn_stats.get("n_unresolved", 0))
    failed = (not digest["nsample_closed_form_ok"]) or digest["n_unresolved"] != 0

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
            e2e["result_matches_device_run"] = bool(np.array_equal(ns2, ns) and all(
                np.allclose(out2[q], out[q], rtol=1e-9) for q in quantities))
            failed |= not e2e["result_matches_device_run"]
        except Exception as ex:  # noqa: BLE001
            e2e = {"value": None, "unit": "Gpart/s", "error": str(ex)[:200]}

    if wl["min_gpus"] > 1:
        args.no_e2e = True          # cfg5: the pinned host copy of the 8 shards (240 GB) is not worth the host memory
    if not args.no_e2e and world > 1:
        # multi-GPU end to end: every rank's shard starts in pinned host memory; H2D + exchange + path inside the timed region
        try:
            import torch.distributed as dist
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
            e2e["result_matches_device_run"] = bool(np.array_equal(ns2, ns))
        except Exception as ex:  # noqa: BLE001
            e2e = {"value": None, "unit": "Gpart/s", "error": str(ex)[:200]}

    cpu = None
    if not args.no_cpu and rank == 0:
        sN, sNp = (N, Np) if wname in ("cfg1", "cfg2") else (128, 1 << 21)
        r = cpu_port_run(sN, sNp, quantities, seed=wl["seed"])
        cpu = {"value": sNp / r["s_total"] / 1e9, "unit": "Gpart/s", "cores": os.cpu_count(), "kind": "port",
               "same_config": wname in ("cfg1", "cfg2"), "extrapolated": wname not in ("cfg1", "cfg2"),
               "sample": f"{sN}^3 lattice / {sNp} particles, {'+'.join(quantities)}, oracle port (cKDTree workers=-1, scipy.fft workers=-1); "
                         f"nn {r['s_nn']:.2f}s spectra {r['s_spectra']:.2f}s"}

    nvlink = None
    if world > 1 and phase_t.get("nvlink_bytes"):
        # bytes this rank stored into peer buffers (exact, from the exchange tables) over the device time of the kernels that
        # store them; reference: 770 GB/s per direction measured for a peer copy on this pool (B200_PROFILING.md), 900 nominal.
        # Hardware NVLink counters are not exposed in this container (profiles/r2_nvlink_counters_unavailable.txt).
        nb = phase_t["nvlink_bytes"]
        ex_ms = table.get("k0_slab_bucket_scatter", {}).get("ms_per_step")
        y_ms = table.get("k4b_fft_y", {}).get("ms_per_step")
        nvlink = {"rank0_bytes_per_step": nb, "peak_GBps": 770.0,
                  "particle_exchange": {"ms": ex_ms, "GBps": round(nb["particle_exchange"] / ex_ms / 1e6, 1) if ex_ms else None,
                                        "frac": round(nb["particle_exchange"] / ex_ms / 1e6 / 770.0, 3) if ex_ms else None},
                  "transpose_in_y_pass": {"ms": y_ms, "GBps": round(nb["transpose"] / y_ms / 1e6, 1) if y_ms else None,
                                          "frac": round(nb["transpose"] / y_ms / 1e6 / 770.0, 3) if y_ms else None}}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Gpart/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{wname}: {N}^3 lattice, {Np} particles (2^{int(np.log2(Np))}, {wl['gen']}), "
                                       f"{'+'.join(quantities)} P(k), library lattice and edges",
                           "l2": "inputs larger than L2" if Np * 28 > 2e8 else "inputs fit L2",
                           "momentum": "reference-strict (vx*m x3)", "input": "sharded by particle index range across ranks"},
                "result": digest, "nn_stats": nn_stats, "seconds_per_pk": ms * 1e-3, "nn_gridding_gpart_s": None,
                "gpu_launches": int(launches), "peak_device_bytes_per_gpu": int(peak_mem),
                "nvlink": nvlink, "roofline": roof, "stages": table, "dist_phases_ms_last_step": phase_t.get("phases_ms"), "cpu_baseline": cpu, "e2e": e2e,
                "clocks": clk.summary()}
        nn_ms = sum(table[n]["ms_per_step"] for n in table if n.startswith("k1"))
        if nn_ms > 0:
            line["nn_gridding_gpart_s"] = Np / (nn_ms * 1e-3) / 1e9
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if failed:
        sys.stderr.write(f"bench.py: RESULT CHECK FAILED: {json.dumps(digest)}\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
