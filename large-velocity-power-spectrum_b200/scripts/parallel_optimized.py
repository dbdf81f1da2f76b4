#!/usr/bin/env python
"""Drop-in for the reference driver `scripts/parallel_optimized.py` on B200.

Same command line (-i -o -N -M -l -b -f), same output file `<output>/Pk.txt` = np.savetxt of [nbins,4]
(k, P, Psum, Nsample), same helper names (`planner`, `FFTW_vector_power`, `FFTW_power`, `pair_power`, `hist_sample`,
`main`).

    python parallel_optimized.py -i snapshot.hdf5 -o out/ -N 1024 -f                      # one GPU
    python -m torch.distributed.run --nproc-per-node 8 parallel_optimized.py -i ... -f     # one process per GPU

What changed underneath (reference lines: scripts/parallel_optimized.py):
  * the Annoy index and the pure-Python loop over lattice nodes (:300-358) -> exact nearest-particle gridding on the GPU
    (cell list + proven search); positions and node coordinates enter as float32 values exactly as there (:346, add_item);
  * the folded DFT spread over MPI ranks (:362-389: allgather + phase sum, one residue class of k-space per rank and
    loop) -> ONE full N^3 transform (they are identical); with several GPUs the lattice is cut into
    x slabs and the transform does one all-to-all (vpower/dist.py).  -M/--maxnbox and -b/--nbuffer are therefore accepted
    and only reported: no box has to be folded to fit memory and nothing is queued;
  * three pyFFTW complex64 transforms + |.|^2 (:124-141, 409-411), k pairing (:145-172) and the two np.histogram calls
    (:176-190) -> r2c FFT with |F|^2 and the shell binning fused into its last pass; bin edges are the script's own
    `linspace` expression (including its 511-bin quirk at N=1024);
  * MPI.Reduce of float32 shell sums (:455-456) -> f64 / u64 all-reduce; the float32 cast the script applies before
    writing (:436-438) is kept so that Pk.txt is interchangeable.
"""
import argparse
import datetime
import os
import sys
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from vpower import _lib  # noqa: E402

# ------------------------------------CONFIG------------------------------------ (defaults of the reference :26-35)
SNAPSHOT = "snapshot_550.hdf5"
SAVEDIR = "../output/"
NBUFFER = 5000
NTOT = 1000
MAXNBOX = 500
LTOT = 1
remove_bulk_velocity = True


def _parse(argv=None):
    p = argparse.ArgumentParser(description="Compute the velocity power spectrum on B200 (drop-in for the MPI driver).",
                                usage="python %(prog)s [options]   |   torchrun --nproc-per-node <gpus> %(prog)s [options]")
    p.add_argument("-i", "--input", nargs="?", type=str, default=SNAPSHOT, help="Path to the snapshot file (.hdf5 or .npz).")
    p.add_argument("-o", "--output", nargs="?", type=str, default=SAVEDIR, help="Directory to save the power spectrum.")
    p.add_argument("-N", "--ntot", nargs="?", type=int, default=NTOT, help="Total resolution.")
    p.add_argument("-M", "--maxnbox", nargs="?", type=int, default=MAXNBOX, help="Accepted for compatibility (planner report only).")
    p.add_argument("-l", "--ltot", nargs="?", type=int, default=LTOT, help="Total length of the box.")
    p.add_argument("-b", "--nbuffer", nargs="?", type=int, default=NBUFFER, help="Accepted for compatibility (unused).")
    p.add_argument("-f", action="store_true", help="Skip confirmation and start the computation.")
    return p.parse_args(argv)


# -----------------------------------FUNCTIONS----------------------------------
def planner(n_total_res, l_total_length, n_box_affordable, n_total_threads):
    """The reference's work plan (:70-88): loops, threads per axis, box size, box length.  Kept for the report."""
    tpa = round(n_total_threads ** (1 / 3))
    assert tpa ** 3 == n_total_threads, "Number of threads must be a cube of an integer. Support for any number is not yet implemented."
    tpa = int(tpa)
    n_loops_per_axis = 1
    n_full_box = n_total_res / tpa
    assert n_full_box.is_integer(), "Divided Nbox must be an integer."
    n_box = n_full_box
    while n_box > n_box_affordable or not float(n_box).is_integer():
        n_loops_per_axis += 1
        n_box = n_full_box / n_loops_per_axis
    return n_loops_per_axis ** 3, tpa, int(n_box), int(n_box) / n_total_res * l_total_length


def _real_cube(f, torch):
    f = np.asarray(f)
    if np.iscomplexobj(f):
        if np.abs(f.imag).max() > 0:
            raise Exception("vpower_b200: the power helpers take the real (unfolded) field")
        f = f.real
    return _lib.to_device(np.ascontiguousarray(f), dtype=torch.float32)


def FFTW_vector_power(fx, fy, fz, Lbox, Nsize):
    """sum over the three components of 1/2 |const * FFT(f_c)|^2, const = (Lbox/2pi)^1.5 / Nsize^3 (:92-121): one
    in-place r2c transform per component, accumulated into one power cube on the device."""
    import torch
    const = (Lbox / (2 * np.pi)) ** 1.5 / Nsize ** 3
    plan = _lib.PkPlan(Nsize, 2 * np.pi * np.fft.fftfreq(Nsize, Lbox / float(Nsize)), np.array([0.0, 1.0]))
    P = plan.power_cube([_real_cube(f, torch) for f in (fx, fy, fz)])
    return (P * (0.5 * const * const)).cpu().numpy().astype(np.float32)


def FFTW_power(f, Lbox, Nsize):
    """1/2 |const * FFT(f)|^2 with const = (Lbox/2pi)^1.5 / Nsize^3 (:124-141).  f: real [N,N,N] cube
    (the unfolded path never forms the complex folded field)."""
    import torch
    const = (Lbox / (2 * np.pi)) ** 1.5 / Nsize ** 3
    plan = _lib.PkPlan(Nsize, 2 * np.pi * np.fft.fftfreq(Nsize, Lbox / float(Nsize)), np.array([0.0, 1.0]))
    P = plan.power_cube([_real_cube(f, torch)])
    return (P * (0.5 * const * const)).cpu().numpy().astype(np.float32)


def pair_power(Pk, Lbox, Nbox, shift=np.array([0, 0, 0])):
    """(|k|, P) pairs, shape (n,2); the script subtracts a non-zero shift from each axis (:145-172)."""
    Lcell = Lbox / float(Nbox)
    ks = 2 * np.pi * np.fft.fftfreq(Nbox, Lcell)
    axes = [ks - shift[c] if shift[c] != 0 else ks for c in range(3)]
    k = _lib.k_magnitude(*axes).cpu().numpy()
    return np.column_stack((k, np.ravel(Pk)))


def _script_edges(kmin, kmax, spacing):
    n_bins = int((kmax - kmin) / spacing) + 1                                        # :178
    return np.linspace(kmin, kmax, n_bins), np.linspace(kmin - spacing / 2, kmax + spacing / 2, n_bins + 1)   # :179-180


def hist_sample(Pk_pair, kmin, kmax, spacing):
    """Mean power per shell with the script's linspace edges; empty shells give NaN as there (:176-190)."""
    import torch
    centres, edges = _script_edges(kmin, kmax, spacing)
    pairs = _lib.to_device(np.ascontiguousarray(Pk_pair), dtype=torch.float64)
    Psum, ns = _lib.hist_weighted(pairs[:, 0].contiguous(), pairs[:, 1].contiguous(), edges)
    with warnings.catch_warnings(), np.errstate(invalid="ignore", divide="ignore"):
        warnings.simplefilter("ignore")
        P = Psum / ns
    return np.column_stack((centres, P, Psum, ns.astype(np.float64)))


def _load(path):
    if path.endswith(".npz"):
        z = np.load(path)
        get = lambda k: z[k] if k in z.files else z["PartType0/" + k]  # noqa: E731
        return get("Coordinates"), get("Masses"), get("Velocities")
    import h5py
    with h5py.File(path, "r") as f:
        return f["PartType0/Coordinates"][:], f["PartType0/Masses"][:], f["PartType0/Velocities"][:]


# -----------------------------------MAIN---------------------------------------
def main(argv=None):
    args = _parse(argv)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    NTOT_, LTOT_ = args.ntot, args.ltot
    LCELL = LTOT_ / NTOT_
    outputfile = os.path.join(args.output, "Pk.txt")
    assert os.path.isdir(args.output), "Output directory does not exist."
    assert os.path.isfile(args.input), "Snapshot file does not exist."

    if rank == 0:
        try:
            n_loops, tpa, Nbox, _ = planner(NTOT_, LTOT_, args.maxnbox, 1)
            print(f"[{datetime.datetime.now()}] Reference planner would use {n_loops} loops of {Nbox}^3 boxes per thread; "
                  f"this run: one {NTOT_}^3 transform on {world} GPU(s).", flush=True)
        except AssertionError as e:
            print(f"[{datetime.datetime.now()}] (reference planner: {e})", flush=True)
        if not args.f:
            print("Accept plan? (y/n)", flush=True)
            if input() != "y":
                print("Plan rejected.", flush=True)
                sys.exit(0)
        print(f"Snapshot: {args.input}\nOutput file: {outputfile}\nNTOT: {NTOT_}\nLTOT: {LTOT_}", flush=True)

    # --------------------------------LOAD DATA--------------------------------- (:272-291)
    coords, mass, velocity = _load(args.input)
    coords = np.array(coords)
    velocity = np.array(velocity)
    for c in range(3):
        coords[:, c] -= np.min(coords[:, c])
    if remove_bulk_velocity:
        M = np.sum(mass)
        for c in range(3):
            velocity[:, c] -= np.sum(mass * velocity[:, c]) / M

    # --------------------------------QUERY + FFT + SAMPLE----------------------
    ax = (np.arange(NTOT_) * LCELL).astype(np.float32).astype(np.float64)            # :343-346 nodes as float32
    pos32 = np.ascontiguousarray(coords, dtype=np.float32)                          # Annoy stores float32 items (:306)
    vel32 = np.ascontiguousarray(velocity, dtype=np.float32)
    kmin, kmax, spacing = 2 * np.pi / LTOT_, np.pi / LCELL, 2 * np.pi / LTOT_          # :430
    centres, edges = _script_edges(kmin, kmax, spacing)
    const = (LTOT_ / (2 * np.pi)) ** 1.5 / NTOT_ ** 3                                 # :131
    kax = 2 * np.pi * np.fft.fftfreq(NTOT_, LCELL)
    if world > 1:
        from vpower import dist as vd
        out, ns = vd.particles_to_pk_dist(_lib.to_device(pos32), _lib.to_device(vel32), None, ax, LCELL ** 3,
                                          0.5 * const * const, kax, edges, quantities=("velocity",))
    else:
        out, ns = _lib.particles_to_pk(pos32, vel32, None, ax, ax, ax, NTOT_, LCELL ** 3, 0.5 * const * const, kax, edges,
                                       quantities=("velocity",))
    if rank == 0:
        # :436-461 -- the reference holds the table as float32 and forms P in float32 from the float32-rounded k
        Pkk = np.empty((len(centres), 4), dtype=np.float32)
        Pkk[:, 0] = centres
        Pkk[:, 2] = out["velocity"]
        Pkk[:, 3] = ns
        with warnings.catch_warnings(), np.errstate(invalid="ignore", divide="ignore"):
            warnings.simplefilter("ignore")
            Pkk[:, 1] = Pkk[:, 2] / Pkk[:, 3] * (4 * np.pi * Pkk[:, 0] ** 2)
        np.savetxt(outputfile, Pkk)                                                  # :473
        print(f"[{datetime.datetime.now()}] Saved: {outputfile}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    assert main() == 0, "Program stopped before completion."
