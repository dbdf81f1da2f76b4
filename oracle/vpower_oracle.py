"""CPU oracle for the particles -> P(k) hot path of `vpower`.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file.
Allowed importers: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs -- and there only as the checker / the CPU arm.

What this is: a plain numpy/scipy restatement of the reference algorithm,
function by function, each citing the reference file:line it follows
(paths relative to the upstream repo root).  It is written for clarity and for
sizes up to ~256^3, not for speed.

Parity pin status (see tests/test_oracle_golden.py, tests/golden/make_golden.py):
  * lattice, payload algebra, field algebra, FFT normalisation, |k| pairing,
    shell histogram (both the library `arange` edges and the script `linspace`
    edges): PINNED against the reference's own Python
    (vpower/interp.py, scripts/parallel_optimized.py) executed verbatim in the
    build container through import shims (oracle/refshims); outputs committed
    under tests/golden/.
  * nearest neighbour: the arithmetic lives in ANN 1.1.2 (C++, not vendored
    upstream, reached through `pyann` which is not vendored either).  PINNED
    against the reference's prebuilt `ann/ann_sample` ELF (libANN 1.1.2 linked
    statically) run in the build container; the indices it printed are
    committed under tests/golden/.  Tie order differs by contract: ANN returns
    the first point met in tree order, this oracle (and the CUDA path) return
    the lowest particle index (BASELINE.json north_star); golden cases carry
    the list of exact ties so that they can be excluded.
  * FFT: the reference calls FFTW 3.3.10 through pyFFTW 0.12.0 (not installed
    here); scipy.fft (pocketfft) is the stand-in, identical to rounding.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
from scipy.spatial import cKDTree

# --------------------------------------------------------------------------
# a1  query lattice
# --------------------------------------------------------------------------

def lattice_axis_lib(Lbox, Nsize):
    """Per-axis node coordinates of the library lattice.

    vpower/interp.py:1060-1069 (`make_grid_coords`): `linspace(Lcell/2,
    Lbox + Lcell/2, Nsize)` -- spacing Lbox/(Nsize-1), last node outside the box.
    """
    Lcell = Lbox / Nsize
    return np.linspace(Lcell / 2, Lbox + Lcell / 2, Nsize)


def lattice_axis_script(NTOT, LTOT):
    """Per-axis node coordinates of the MPI script lattice.

    scripts/parallel_optimized.py:206,343-346: x = ((r+n1)*Nbox + i) * LCELL with
    LCELL = LTOT/NTOT, then `np.array([x,y,z], dtype=np.float32)`.  Over all
    ranks/loops the integer factor runs over 0..NTOT-1.  The float32 cast is part
    of the reference; the returned float64 values are float32-representable.
    """
    LCELL = LTOT / NTOT
    x = np.arange(NTOT) * LCELL
    return x.astype(np.float32).astype(np.float64)


def make_grid_coords(Lbox, Nsize):
    """[Nsize^3, 3] f64 query points, C order, 'ij' meshgrid.  interp.py:1060-1069."""
    ax = lattice_axis_lib(Lbox, Nsize)
    g = np.meshgrid(ax, ax, ax, indexing="ij")
    return np.reshape(g, (3, Nsize ** 3)).T


# --------------------------------------------------------------------------
# a2  exact nearest particle
# --------------------------------------------------------------------------

def _d2(q, p):
    """Squared distance with ANN's association: ((dx*dx + dy*dy) + dz*dz), f64.

    ANN 1.1.2 kd_search.cpp (leaf visit): dist = 0; for d: t = q[d]-p[d]; dist += t*t.
    """
    dx = q[..., 0] - p[..., 0]
    dy = q[..., 1] - p[..., 1]
    dz = q[..., 2] - p[..., 2]
    return (dx * dx + dy * dy) + dz * dz


def nn_exact_points(data_pos, query_pos, kcand=4, return_ties=False, tree=None):
    """Index of the nearest data point for every query point.

    Contract (BASELINE.json north_star; reference call interp.py:1027-1037 with
    k=1, eps=0 and the `-1`): argmin over particles of the f64 squared Euclidean
    distance, NON-periodic, ties broken by the LOWEST particle index.

    A kd-tree only proposes candidates; the decision is made here with the
    explicit f64 formula on the `kcand` nearest, and every query whose runner-up
    is within 1e-9 relative of the winner (or whose candidate list is not
    provably complete) is re-decided by brute force over all particles.
    """
    data = np.ascontiguousarray(data_pos, dtype=np.float64)
    qry = np.ascontiguousarray(query_pos, dtype=np.float64)
    n = data.shape[0]
    k = int(min(kcand, n))
    if tree is None:                                           # (callers that ask query by query pass their tree of `data`)
        tree = cKDTree(data)
    _, cand = tree.query(qry, k=k, workers=-1 if len(qry) > 64 else 1)
    if k == 1:
        cand = cand[:, None]
    d2 = _d2(qry[:, None, :], data[cand])                      # [Nq, k]
    order = np.lexsort((cand, d2), axis=1)                      # by d2 then index
    best = np.take_along_axis(cand, order[:, :1], axis=1)[:, 0]
    ties = np.zeros(len(qry), dtype=bool)
    if k > 1:
        d2s = np.take_along_axis(d2, order, axis=1)
        near = d2s[:, 1] <= d2s[:, 0] * (1 + 1e-9) + 1e-300
        # if all k candidates are (nearly) equidistant the list may be incomplete
        for qi in np.nonzero(near)[0]:
            dd = _d2(qry[qi][None, :], data)
            m = dd.min()
            w = np.nonzero(dd == m)[0]
            best[qi] = w[0]
            ties[qi] = len(w) > 1
    if return_ties:
        return best.astype(np.int64), ties
    return best.astype(np.int64)


def nn_exact_lattice(data_pos, qx, qy, qz, return_ties=False):
    """Nearest particle for the separable lattice (qx[i], qy[j], qz[k]); -> [nx,ny,nz]."""
    g = np.meshgrid(qx, qy, qz, indexing="ij")
    q = np.stack([a.ravel() for a in g], axis=1)
    r = nn_exact_points(data_pos, q, return_ties=return_ties)
    shp = (len(qx), len(qy), len(qz))
    if return_ties:
        return r[0].reshape(shp), r[1].reshape(shp)
    return r.reshape(shp)


def ann_interpolate(data_pos, query_pos, f, Nsize, eps=0.0, treetype="kd", searchtype="standard"):
    """interp.py:1018-1049: gather f at the nearest particle, reshape to the cube."""
    index = nn_exact_points(data_pos, query_pos)
    f = np.asarray(f)
    if f.ndim == 1:
        return np.reshape(f[index], (Nsize, Nsize, Nsize))
    if f.ndim == 2:
        return np.reshape(f[index, :], (Nsize, Nsize, Nsize, f.shape[1]))
    raise Exception("Unsupported data shape.")


# --------------------------------------------------------------------------
# a4  payload algebra / gridding front end
# --------------------------------------------------------------------------

def density_velocity_vector(velocity, density):
    """interp.py:199-213: [rho*vx, rho*vy, rho*vz, rho]."""
    return np.stack((velocity[:, 0] * density, velocity[:, 1] * density,
                     velocity[:, 2] * density, density), axis=1)


def ann_interp_to_field(pos, density, velocity, Lbox, Nsize):
    """interp.py:246-277 -> (v_grid [N,N,N,3], m_grid [N,N,N], Lcell)."""
    Lcell = Lbox / Nsize
    vec = ann_interpolate(pos, make_grid_coords(Lbox, Nsize),
                          density_velocity_vector(velocity, density), Nsize)
    v_grid = vec[..., :3] / vec[..., 3, None]
    m_grid = vec[..., 3] * Lcell ** 3
    return v_grid, m_grid, Lcell


# --------------------------------------------------------------------------
# a5  NGP deposit
# --------------------------------------------------------------------------

def deposit_to_grid(f, pos, Nsize, Lbox):
    """interp.py:996-1015: periodic NGP scatter-add into an f64 grid."""
    f = np.asarray(f)
    if f.ndim == 1:
        grid = np.zeros((Nsize, Nsize, Nsize))
    else:
        grid = np.zeros((Nsize, Nsize, Nsize, f.shape[1]))
    Lcell = Lbox / float(Nsize)
    index = np.array((pos // Lcell) % Nsize, dtype=int)
    np.add.at(grid, tuple(index.T), f)
    return grid


def deposit_cell_index(pos, Nsize, Lbox):
    """The integer cell triple used by `deposit_to_grid` (interp.py:1010-1011)."""
    Lcell = Lbox / float(Nsize)
    return np.array((pos // Lcell) % Nsize, dtype=int)


# --------------------------------------------------------------------------
# a6 / a7  field algebra, FFT, power
# --------------------------------------------------------------------------

def _fftn(a):
    # pyfftw.interfaces.numpy_fft.fftn is dtype preserving (f32->c64, f64->c128),
    # and so is scipy.fft.fftn.
    return scipy.fft.fftn(a, workers=-1)


def vector_power(fx, fy, fz, Lbox, Nsize):
    """interp.py:1372-1387: P = 1/2 sum_c |a FFT(f_c)|^2, a = (Lbox/2pi)^1.5 / N^3."""
    a = (Lbox / (2 * np.pi)) ** 1.5 / Nsize ** 3
    P = 0.5 * np.abs(_fftn(fx) * a) ** 2
    P = P + 0.5 * np.abs(_fftn(fy) * a) ** 2
    P = P + 0.5 * np.abs(_fftn(fz) * a) ** 2
    return P


def scalar_power(f, Lbox, Nsize):
    """interp.py:1408-1421."""
    a = (Lbox / (2 * np.pi)) ** 1.5 / Nsize ** 3
    return 0.5 * np.abs(_fftn(f) * a) ** 2


def field_components(v_grid, m_grid, quantity, strict_reference=True):
    """The real fields that get transformed for each quantity.

    interp.py:501-518 velocity (vx,vy,vz); :521-541 momentum -- the reference
    multiplies `vx` three times (:523-525), reproduced when strict_reference;
    :544-557 energy E = m (vx^2+vy^2+vz^2), no 1/2.
    """
    vx, vy, vz = v_grid[..., 0], v_grid[..., 1], v_grid[..., 2]
    if quantity == "velocity":
        return [vx, vy, vz]
    if quantity == "momentum":
        if strict_reference:
            return [vx * m_grid, vx * m_grid, vx * m_grid]
        return [vx * m_grid, vy * m_grid, vz * m_grid]
    if quantity == "energy":
        return [m_grid * (vx ** 2 + vy ** 2 + vz ** 2)]
    raise Exception("Unrecognized physical quantity name.")


def power_grid(v_grid, m_grid, Lcell, quantity, strict_reference=True):
    N = m_grid.shape[0]
    Lbox = N * Lcell
    comps = field_components(v_grid, m_grid, quantity, strict_reference)
    if len(comps) == 3:
        return vector_power(comps[0], comps[1], comps[2], Lbox, N)
    return scalar_power(comps[0], Lbox, N)


# --------------------------------------------------------------------------
# a8 / a9  |k| pairing and shell histogram
# --------------------------------------------------------------------------

def k_axis(Lbox, Nsize):
    """interp.py:1448-1449: kSpace = 2 pi fftfreq(N, Lcell)."""
    Lcell = Lbox / float(Nsize)
    return 2 * np.pi * np.fft.fftfreq(Nsize, Lcell)


def k_magnitude(Lbox, Nsize):
    """interp.py:1451-1460: k = sqrt(kx*kx + ky*ky + kz*kz), C order, flattened."""
    ks = k_axis(Lbox, Nsize)
    kx, ky, kz = np.meshgrid(ks, ks, ks, indexing="ij")
    return np.ravel(np.sqrt(kx * kx + ky * ky + kz * kz))


def edges_lib(kmin, kmax, spacing):
    """interp.py:1472-1473 -> (centres, edges)."""
    centres = np.arange(kmin, kmax + spacing, spacing)
    edges = np.arange(kmin - spacing / 2, kmax + 3 * spacing / 2, spacing)
    return centres, edges


def edges_script(kmin, kmax, spacing):
    """scripts/parallel_optimized.py:178-180 -> (centres, edges)."""
    n_bins = int((kmax - kmin) / spacing) + 1
    centres = np.linspace(kmin, kmax, n_bins)
    edges = np.linspace(kmin - spacing / 2, kmax + spacing / 2, n_bins + 1)
    return centres, edges


def hist_sample(k, P, centres, edges, empty_to_zero=True):
    """interp.py:1474-1482 / parallel_optimized.py:181-188 -> [nbins,4] (k,P,Psum,Nsample)."""
    Psum, _ = np.histogram(k, bins=edges, weights=np.ravel(P))
    Nsample, _ = np.histogram(k, bins=edges)
    with np.errstate(invalid="ignore", divide="ignore"):
        Pm = Psum / Nsample
    if empty_to_zero:
        Pm[Nsample == 0] = 0
    return np.column_stack((centres, Pm, Psum, Nsample))


def spctrm(v_grid, m_grid, Lcell, quantity="velocity", kmin=None, kmax=None, kres=None,
           strict_reference=True, edges="lib"):
    """interp.py:560-595 -> [nbins,4] (k, P*4 pi k^2, Psum, Nsample)."""
    N = m_grid.shape[0]
    Lbox = N * Lcell
    if kmin is None:
        kmin = 2 * np.pi / Lbox
    if kmax is None:
        kmax = np.pi / Lcell
    if kres is None:
        kres = kmin
    P = power_grid(v_grid, m_grid, Lcell, quantity, strict_reference)
    k = k_magnitude(Lbox, N)
    c, e = (edges_lib if edges == "lib" else edges_script)(kmin, kmax, kres)
    Pkk = hist_sample(k, P, c, e, empty_to_zero=(edges == "lib"))
    Pkk[:, 1] *= 4 * np.pi * Pkk[:, 0] ** 2
    return Pkk


def shell_counts(Lbox, Nsize, edges):
    """Nsample only (geometry): histogram of |k| over all N^3 modes."""
    n, _ = np.histogram(k_magnitude(Lbox, Nsize), bins=edges)
    return n.astype(np.int64)


# --------------------------------------------------------------------------
# f3  folding (outer stage): BoxField.fold -> FoldedBox.fold_spctrm
#     pinned by tests/golden/reference_golden_fold.npz (tests/golden/make_golden_fold.py)
# --------------------------------------------------------------------------

def fold_phase(beta, Nsize):
    """interp.py:1215-1225 (_get_phase with x0=y0=z0=0, Nphase=totalNsize=Nsize):
    exp(-i (2 pi / N) (beta . x)) on the [N,N,N] lattice, complex128."""
    n = np.arange(Nsize)
    xxx, yyy, zzz = np.meshgrid(n, n, n, indexing="ij")
    return np.exp(-1j * (2 * np.pi / Nsize) * (beta[0] * xxx + beta[1] * yyy + beta[2] * zzz))


def fold_field(f, m):
    """interp.py:1228-1252: sum of the m^3 sub-blocks of size N/m, added in (i, j, k) order."""
    if m == 1:
        return f
    n1, n2, n3 = f.shape[0] // m, f.shape[1] // m, f.shape[2] // m
    r = 0.0
    for i in range(m):
        for j in range(m):
            for k in range(m):
                r = r + f[i * n1:(i + 1) * n1, j * n2:(j + 1) * n2, k * n3:(k + 1) * n3]
    return r


def fold_velocity(v_grid, m, beta):
    """BoxField.fold, interp.py:598-609 (+ _apply_phase :1195-1212): complex128 [N/m,N/m,N/m,3] = fold(v * phase) / m^1.5."""
    N = v_grid.shape[0]
    phi = np.array(v_grid, dtype=np.complex128) * fold_phase(beta, N)[..., None]
    return fold_field(phi, m) / m ** 1.5


def fold_spctrm(folded, m, beta, totalLbox, kmin=None, kmax=None, kres=None):
    """FoldedBox.fold_spctrm, interp.py:755-791 -> [nbins,4] (k, P*4 pi k^2, Psum, Nsample).
    folded: [n,n,n,3] (vector) or [n,n,n] (scalar) complex; Lbox = totalLbox/m, Lcell = Lbox/n."""
    n = folded.shape[0]
    Lbox = totalLbox / m
    Lcell = Lbox / n
    if kmin is None:
        kmin = 2 * np.pi / totalLbox
    if kmax is None:
        kmax = np.pi / Lcell
    if kres is None:
        kres = kmin
    a = (Lbox / (2 * np.pi)) ** 1.5 / n ** 3                                   # :1398
    fk = scipy.fft.fftn(folded, axes=(0, 1, 2), workers=-1) * a
    P = 0.5 * np.sum(np.abs(fk) ** 2, axis=3) if folded.ndim == 4 else 0.5 * np.abs(fk) ** 2
    ks = k_axis(Lbox, n)
    shift = 2 * np.pi * np.asarray(beta) / totalLbox                           # :780
    ax = [ks + shift[c] if shift[c] > 0 else ks for c in range(3)]             # :1453-1458
    kx, ky, kz = np.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
    k = np.ravel(np.sqrt(kx * kx + ky * ky + kz * kz))
    c, e = edges_lib(kmin, kmax, kres)
    Pkk = hist_sample(k, P, c, e, empty_to_zero=True)
    Pkk[:, 1] *= 4 * np.pi * Pkk[:, 0] ** 2
    return Pkk


# --------------------------------------------------------------------------
# whole path, library flavour and script flavour
# --------------------------------------------------------------------------

def particles_to_pk_lib(pos, density, velocity, Lbox, Nsize, quantity="velocity",
                        strict_reference=True):
    """GasParticles.ann_interp_to_field(N).spctrm(quantity)  (interp.py:246-277, 560-595)."""
    v, m, Lcell = ann_interp_to_field(pos, density, velocity, Lbox, Nsize)
    return spctrm(v, m, Lcell, quantity, strict_reference=strict_reference)


def particles_to_pk_script(coords, velocity, NTOT, LTOT):
    """What scripts/parallel_optimized.py main() computes, with the folded DFT
    replaced by the full transform it is equal to and the Annoy
    approximate search replaced by the exact one (north_star).

    Lattice :343-346 (f32 nodes i*LCELL), velocity sampled at the nearest
    particle :348-351, three complex64 transforms :409-411 with
    const=(L/2pi)^1.5/N^3 :131, k pairing :145-172, script edges :176-190,
    P*4 pi k^2 :434.  Psum/Nsample are cast to f32 by the script (:436-438);
    this oracle keeps f64 (the cast is applied by the Pk.txt writer).
    """
    ax = lattice_axis_script(NTOT, LTOT)
    # Annoy stores items as float32 (add_item): positions enter as f32 values.
    p32 = np.asarray(coords, dtype=np.float32).astype(np.float64)
    idx = nn_exact_lattice(p32, ax, ax, ax)
    f = np.asarray(velocity, dtype=np.float32)[idx]                  # [N,N,N,3] f32
    const = (LTOT / (2 * np.pi)) ** 1.5 / NTOT ** 3
    P = np.zeros((NTOT,) * 3, dtype=np.float32)
    for c in range(3):
        b = scipy.fft.fftn(f[..., c].astype(np.complex64), workers=-1)
        P = P + (0.5 * np.abs(b * const) ** 2).astype(np.float32)
    k = k_magnitude(LTOT, NTOT)
    LCELL = LTOT / NTOT
    c_, e_ = edges_script(2 * np.pi / LTOT, np.pi / LCELL, 2 * np.pi / LTOT)
    Pkk = hist_sample(k, P, c_, e_, empty_to_zero=False)
    Pkk[:, 1] *= 4 * np.pi * Pkk[:, 0] ** 2
    return Pkk


# --------------------------------------------------------------------------
# synthetic inputs (shared by tests and bench; integer hash -> identical anywhere)
# --------------------------------------------------------------------------

def _mix32(x):
    x = np.asarray(x, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def hash_uniform(seed, n, stream):
    """n uniforms in [0,1) with 24-bit mantissa (exact in f32), counter based."""
    i = np.arange(n, dtype=np.uint64)
    h = _mix32(i ^ _mix32(np.uint64(seed) * np.uint64(0x9E3779B1) + np.uint64(stream) * np.uint64(0x85EBCA77)))
    h = _mix32(h + (i >> np.uint64(32)) + np.uint64(stream))
    return ((h >> np.uint64(8)).astype(np.float64) * (1.0 / 16777216.0)).astype(np.float32)


def synth_particles(seed, Np, Lbox=1.0, clustered=False, lattice_n=None):
    """Synthetic particle set (pos f32 [Np,3], vel f32 [Np,3], density f32 [Np], mass f32 [Np]).

    uniform: pos = U[0,L).  clustered: a lattice_n^3 lattice displaced by a sum of
    long-wave sinusoids with rms ~2 cells (Zel'dovich-like).
    velocity: 8 smooth modes + 0.1 white noise.
    """
    rng = np.random.default_rng(1000 + seed)
    if clustered:
        n = lattice_n
        assert n ** 3 == Np
        g = (np.arange(n, dtype=np.float64) + 0.5) / n
        q = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
        disp = np.zeros_like(q)
        for _ in range(24):
            kv = rng.integers(1, 5, size=3) * rng.choice([-1, 1], size=3)
            kk = np.sqrt((kv ** 2).sum())
            amp = 1.0 / kk ** 2
            ph = rng.uniform(0, 2 * np.pi)
            disp += amp * (kv / kk)[None, :] * np.sin(2 * np.pi * (q @ kv) + ph)[:, None]
        disp *= (2.0 / n) / np.sqrt((disp ** 2).sum(1).mean())
        pos = np.mod(q + disp, 1.0) * Lbox
        pos = pos.astype(np.float32)
    else:
        pos = np.stack([hash_uniform(seed, Np, c) for c in range(3)], axis=1) * np.float32(Lbox)
    x = pos.astype(np.float64) / Lbox
    vel = np.zeros((Np, 3))
    for _ in range(8):
        kv = rng.integers(1, 7, size=3)
        a = rng.normal(size=3) / np.sqrt((kv ** 2).sum())
        ph = rng.uniform(0, 2 * np.pi)
        vel += a[None, :] * np.sin(2 * np.pi * (x @ kv) + ph)[:, None]
    noise = np.stack([hash_uniform(seed, Np, 3 + c) for c in range(3)], axis=1).astype(np.float64) - 0.5
    vel = (vel + 0.1 * noise * np.sqrt(12.0)).astype(np.float32)
    dens = (Np / Lbox ** 3 * (1.0 + 0.1 * hash_uniform(seed, Np, 7).astype(np.float64))).astype(np.float32)
    mass = np.ones(Np, dtype=np.float32)
    return pos, vel, dens, mass
