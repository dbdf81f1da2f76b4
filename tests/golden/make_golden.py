"""Generate tests/golden/reference_golden.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What runs: /root/reference/vpower/interp.py, spctrm.py and
/root/reference/scripts/parallel_optimized.py, imported / runpy'd verbatim with
the import shims of oracle/refshims.py standing in for the third-party packages
that are not installed (pyfftw->scipy.fft, pyann->the reference's own
ann/ann_sample ELF = ANN 1.1.2, annoy->exact search, mpi4py->single rank,
h5py->dict).  The arrays it returns are the golden vectors the oracle
(oracle/vpower_oracle.py) and the CUDA path are pinned against.
"""
import contextlib
import io
import json
import os
import runpy
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import refshims  # noqa: E402
import vpower_oracle as orc  # noqa: E402

out = {}
manifest = {"reference": "YujieH3/large-velocity-power-spectrum @ /root/reference", "cases": {}}


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ---------------------------------------------------------------- library path, true ANN engine
state = refshims.install(engine="ann_sample")
import interp as ref_interp  # noqa: E402  (the reference module)
import spctrm as ref_spctrm  # noqa: E402


def lib_case(name, seed, Np, N, Lbox, quantise16, clustered=False):
    pos32, vel32, dens32, mass32 = orc.synth_particles(seed, Np, Lbox)
    pos = pos32.astype(np.float64)
    if clustered:   # pull half of the particles into a blob -> voids + dense cells
        rng = np.random.default_rng(seed)
        sel = rng.random(Np) < 0.5
        pos[sel] = 0.3 * Lbox + 0.08 * Lbox * rng.normal(size=(sel.sum(), 3))
        pos = np.abs(pos)
    if quantise16:
        pos = np.floor(pos / Lbox * 65536.0) / 65536.0 * Lbox
    vel, dens, mass = vel32.astype(np.float64), dens32.astype(np.float64), mass32.astype(np.float64)
    gp = ref_interp.GasParticles(pos.copy(), mass.copy(), dens.copy(), vel.copy(), Lbox)
    with quiet():
        bf = gp.ann_interp_to_field(N)
    nn = state["last_nn"]
    idx_ann = nn["idx0"].reshape(N, N, N)
    # what ANN actually saw (text round trip), as separable axis tables
    dpar = nn["data_parsed"]
    qpar = nn["query_parsed"].reshape(N, N, N, 3)
    ax = [qpar[:, 0, 0, 0].copy(), qpar[0, :, 0, 1].copy(), qpar[0, 0, :, 2].copy()]
    assert np.array_equal(qpar[..., 0], np.broadcast_to(ax[0][:, None, None], (N, N, N)))
    assert np.array_equal(qpar[..., 2], np.broadcast_to(ax[2][None, None, :], (N, N, N)))
    idx_orc, ties = orc.nn_exact_lattice(dpar, ax[0], ax[1], ax[2], return_ties=True)
    mism = int((idx_orc != idx_ann).sum())
    mism_nontie = int(((idx_orc != idx_ann) & ~ties).sum())
    print(f"{name}: ANN vs oracle mismatches {mism} (non-tied {mism_nontie}), ties {int(ties.sum())}")
    assert mism_nontie == 0
    out[f"{name}/pos"] = pos
    out[f"{name}/pos_parsed_differs"] = np.array(not np.array_equal(dpar, pos))
    if not np.array_equal(dpar, pos):
        out[f"{name}/pos_parsed"] = dpar
    out[f"{name}/vel"], out[f"{name}/dens"], out[f"{name}/mass"] = vel32, dens32, mass32
    for c in range(3):
        out[f"{name}/axis_parsed{c}"] = ax[c]
    out[f"{name}/axis_lib"] = orc.lattice_axis_lib(Lbox, N)
    out[f"{name}/nn_ann"] = idx_ann.astype(np.int32)
    out[f"{name}/nn_ties"] = np.packbits(ties.ravel())
    out[f"{name}/v_grid"] = np.stack([bf.vx, bf.vy, bf.vz], axis=-1)
    out[f"{name}/m_grid"] = np.asarray(bf.mass)
    for q in ("velocity", "momentum", "energy"):
        with quiet():
            sp = bf.spctrm(q)
        out[f"{name}/spctrm_{q}"] = sp.data()
    # raw power grids and the k pairing, for the small case only
    if N <= 16:
        with quiet():
            out[f"{name}/Pgrid_velocity"] = bf.velocity_power()
            out[f"{name}/Pgrid_energy"] = bf.kinetic_energy_power()
        out[f"{name}/pairs_k"] = ref_interp._pair_power(out[f"{name}/Pgrid_velocity"], bf.Lbox, bf.Nsize)[:, 0]
    manifest["cases"][name] = {"kind": "lib", "seed": seed, "Np": Np, "N": N, "Lbox": Lbox,
                               "engine": "ann/ann_sample (ANN 1.1.2)", "mismatch_total": mism,
                               "mismatch_nontied": mism_nontie, "ties": int(ties.sum())}


lib_case("lib16", seed=11, Np=4096, N=16, Lbox=1.0, quantise16=True)
lib_case("lib24", seed=12, Np=6000, N=24, Lbox=2.5, quantise16=False)
lib_case("lib32c", seed=13, Np=8000, N=32, Lbox=1.0, quantise16=False, clustered=True)

# ---------------------------------------------------------------- deposit_to_grid
rng = np.random.default_rng(5)
for tag, dt in (("f64", np.float64), ("f32", np.float32)):
    Np, N, L = 5000, 12, 1.5
    p = (rng.random((Np, 3)) * 1.4 * L - 0.2 * L).astype(dt)       # some outside the box, some negative
    p[:7] = np.array([[0, 0, 0], [L, L, L], [L / N, 2 * L / N, 3 * L / N], [-L / N, 0, L],
                      [0.3, 0.7, 0.1], [1.5, 1.5, 1.5], [-1e-9, 1e-9, L - 1e-9]], dtype=dt)
    w1 = rng.integers(1, 9, size=Np).astype(np.float64)
    w4 = rng.normal(size=(Np, 4))
    out[f"deposit_{tag}/pos"], out[f"deposit_{tag}/w1"], out[f"deposit_{tag}/w4"] = p, w1, w4
    out[f"deposit_{tag}/grid1"] = ref_interp.deposit_to_grid(w1, p, N, L)
    out[f"deposit_{tag}/grid4"] = ref_interp.deposit_to_grid(w4, p, N, L)
    manifest["cases"][f"deposit_{tag}"] = {"kind": "deposit", "Np": Np, "N": N, "Lbox": L}

# ---------------------------------------------------------------- shell geometry (edges + counts)
for N, L in ((16, 1.0), (32, 2.5), (64, 1.0), (48, 0.7)):
    P = np.ones((N, N, N))
    pairs = ref_interp._pair_power(P, L, N)
    kmin, kmax = 2 * np.pi / L, np.pi / (L / N)
    h = ref_interp._hist_sample(pairs, kmin, kmax, kmin)
    out[f"shells_lib_{N}_{L}/centres"] = h[:, 0]
    out[f"shells_lib_{N}_{L}/Nsample"] = h[:, 3]
    manifest["cases"][f"shells_lib_{N}_{L}"] = {"kind": "shells", "N": N, "Lbox": L, "sum": float(h[:, 3].sum())}

# ---------------------------------------------------------------- MPI script, verbatim, single rank
state = refshims.install(engine="oracle")


def script_case(name, seed, Np, NTOT, MAXNBOX, nbuffer):
    pos32, vel32, dens32, mass32 = orc.synth_particles(seed, Np, 1.0)
    pos32 = pos32 + np.float32(0.013)                                # exercised by the min-corner shift
    mass = (1.0 + orc.hash_uniform(seed, Np, 9)).astype(np.float32)
    with tempfile.TemporaryDirectory() as td:
        snap = os.path.join(td, "snap.hdf5")
        refshims.register_snapshot(snap, pos32.copy(), mass.copy(), dens32.copy(), vel32.copy())
        argv = sys.argv
        sys.argv = ["parallel_optimized.py", "-i", snap, "-o", td, "-N", str(NTOT), "-M", str(MAXNBOX),
                    "-b", str(nbuffer), "-f"]
        try:
            with quiet(), contextlib.redirect_stderr(io.StringIO()):
                runpy.run_path(os.path.join(refshims.REF_ROOT, "scripts", "parallel_optimized.py"),
                               run_name="__main__")
        finally:
            sys.argv = argv
        pk = np.loadtxt(os.path.join(td, "Pk.txt"))
    out[f"{name}/pos"], out[f"{name}/vel"], out[f"{name}/mass"] = pos32, vel32, mass
    out[f"{name}/Pk"] = pk
    manifest["cases"][name] = {"kind": "script", "seed": seed, "Np": Np, "NTOT": NTOT, "MAXNBOX": MAXNBOX,
                               "NBUFFER": nbuffer, "ranks": 1}
    print(name, "Pk.txt", pk.shape, "Nsample", pk[:, 3].astype(int).tolist())


script_case("script16", seed=21, Np=3000, NTOT=16, MAXNBOX=16, nbuffer=512)
script_case("script16_fold2", seed=21, Np=3000, NTOT=16, MAXNBOX=8, nbuffer=512)   # n_loops=8, m=2 folding

np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
    json.dump(manifest, f, indent=1)
print("wrote", os.path.join(HERE, "reference_golden.npz"),
      os.path.getsize(os.path.join(HERE, "reference_golden.npz")) // 1024, "KiB")
