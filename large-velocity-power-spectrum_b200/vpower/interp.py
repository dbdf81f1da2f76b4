"""Particles -> gridded fields -> P(k): host-side mirror of the reference `vpower/interp.py` hot path.

Every public name keeps the reference's signature, array shapes, dtypes and return conventions
(SURVEY.md 8(b)); the arithmetic runs in hand-written sm_100a kernels behind the C ABI of
libvpower_b200.so (include/vpower_b200.h), reached through `_lib` (ctypes).  There is no CPU
fallback: without the shared library or a CUDA device the compute calls raise `VPowerError`.

Reference line numbers (vpower/interp.py) are cited per function.  Out of scope here: the Voxelize
path, the ANN command-line path, the brick inventory on disk (`BrickInventory`, `interp_to_brick`) and plotting.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .spctrm import PowerSpectrum

__all__ = ["load_snapshot", "GasParticles", "BoxField", "FoldedBox", "ann_interpolate", "make_grid_coords", "deposit_to_grid",
           "check_conservation", "_vector_power", "_scalar_power", "_pair_power", "_hist_sample", "_get_phase", "_apply_phase",
           "fold_field"]


# ------------------------------------------------------------------------------------------ snapshot
def _is_dev(a):
    return type(a).__module__.startswith("torch") and getattr(a, "is_cuda", False)


def load_snapshot(file, Lbox=1.0, remove_bulk_velocity=True, shift_to_origin=True, device=False):
    """PartType0 {Coordinates, Masses, Density, Velocities} -> GasParticles.  interp.py:84-131.

    HDF5 needs h5py (not in this image); a `.npz` with the same four keys
    ('PartType0/Coordinates', ...) or bare ('Coordinates', ...) is accepted as well.
    device=True (not in the reference): the four arrays are uploaded once and stay CUDA tensors; the shift and the
    bulk-velocity removal then run as device reductions (vp_snapshot_preamble) and `ann_interp_to_field` grids them in
    place -- nothing O(Np) happens on the host after the file has been read.
    """
    if str(file).endswith(".npz"):
        z = np.load(file)
        get = lambda k: z[k] if k in z.files else z["PartType0/" + k]  # noqa: E731
        c, m, d, v = get("Coordinates"), get("Masses"), get("Density"), get("Velocities")
    else:
        import h5py
        with h5py.File(file, "r") as f:
            g = f["PartType0"]
            c, m, d, v = g["Coordinates"][:], g["Masses"][:], g["Density"][:], g["Velocities"][:]
    if device:
        dt = np.float64 if np.asarray(c).dtype == np.float64 else np.float32
        c, m, d, v = (_lib.to_device(np.asarray(a, dtype=dt)) for a in (c, m, d, v))
    gp = GasParticles(c, m, d, v, Lbox=Lbox)
    if remove_bulk_velocity is True:
        gp.remove_bulk_velocity()
    if shift_to_origin is True:
        gp.shift_to_origin()
    return gp


# ------------------------------------------------------------------------------------------ particles
class GasParticles:
    """interp.py:137-450 (the part on the hot path)."""

    def __init__(self, pos, mass, density, velocity, Lbox) -> None:
        self.pos = pos
        self.mass = mass
        self.density = density
        self.velocity = velocity
        self.Lbox = Lbox
        self.v = self.velocity

    def __len__(self) -> int:
        return len(self.pos)

    def __getitem__(self, index):
        return GasParticles(self.pos[index], self.mass[index], self.density[index], self.v[index], self.Lbox)

    def shift_to_origin(self) -> None:
        """interp.py:169-175.  CUDA tensors: one device reduction + one update (vp_snapshot_preamble)."""
        if _is_dev(self.pos):
            _lib.snapshot_preamble(self.pos, None, None, shift=True, bulk=False)
            return
        for c in range(3):
            self.pos[:, c] -= np.min(self.pos[:, c])

    def remove_bulk_velocity(self) -> None:
        """Subtract the mass-weighted mean velocity.  interp.py:178-182.  CUDA tensors: on the device."""
        if _is_dev(self.v):
            _lib.snapshot_preamble(None, self.v, self.mass, shift=False, bulk=True)
            return
        M = np.sum(self.mass)
        for c in range(3):
            self.v[:, c] -= np.sum(self.mass * self.v[:, c]) / M

    def rho(self, smoothing_rate=1.0):
        return self.density / smoothing_rate ** 3

    def h(self, smoothing_rate=1.0):
        """Smoothing length from mass and density.  interp.py:190-196."""
        V = self.mass / (self.density / smoothing_rate ** 3)
        return ((3 * V) / (4 * np.pi)) ** (1 / 3)

    @property
    def r(self):
        return self.h()

    def density_velocity_vector(self):
        """[Np,4] payload [rho*vx, rho*vy, rho*vz, rho].  interp.py:199-213."""
        return np.stack((self.v[:, 0] * self.density, self.v[:, 1] * self.density, self.v[:, 2] * self.density,
                         self.density), axis=1)

    def ann_interp_to_field(self, Nsize, eps=0.0, treetype="kd", searchtype="standard"):
        """Nearest-particle sampling of velocity and density on the Nsize^3 lattice -> BoxField.
        interp.py:246-277.  `treetype`/`searchtype` are accepted and ignored, as in the reference (:268-269);
        eps must be 0 (the search is exact).

        The returned BoxField keeps the nearest-particle indices and the particle arrays on the device, so
        that `spctrm()` runs without a host round trip; `.vx/.vy/.vz/.mass` materialise numpy arrays on demand.
        """
        if eps != 0.0:
            raise Exception("vpower_b200: only the exact search (eps=0) is implemented")
        Lcell = self.Lbox / Nsize
        ax = _lattice_axis(self.Lbox, Nsize)
        if _is_dev(self.pos):
            pos_t, vel_t, rho_t = self.pos.contiguous(), self.v.contiguous(), self.density.contiguous()
        else:
            dt = np.float64 if np.asarray(self.pos).dtype == np.float64 else np.float32
            pos_t = _lib.to_device(np.asarray(self.pos, dtype=dt))
            vel_t = _lib.to_device(np.asarray(self.v, dtype=dt))
            rho_t = _lib.to_device(np.asarray(self.density, dtype=dt))
        nn, nn_pos, spay = _lib.nn_grid_payload(pos_t, vel_t, rho_t, ax, ax, ax, Lcell ** 3)
        del pos_t
        return BoxField._from_device(nn, vel_t, rho_t, Lcell, nn_pos, spay)

    # totals used by check_conservation (interp.py:424-450)
    def total_mass(self):
        return np.sum(self.mass)

    def total_momentum(self):
        return np.array([np.sum(self.mass * self.v[:, c]) for c in range(3)])

    def total_kinetic_energy(self):
        return 0.5 * np.sum(self.mass * (self.v[:, 0] ** 2 + self.v[:, 1] ** 2 + self.v[:, 2] ** 2))

    def specific_kinetic_energy(self):
        return self.total_kinetic_energy() / self.total_mass()


# ------------------------------------------------------------------------------------------ gridded field
class BoxField:
    """Velocity + mass on a regular lattice.  interp.py:456-666 (hot-path part)."""

    strict_reference = True   # momentum_power uses vx for all three components, as interp.py:523-525 does

    def __init__(self, v, mass, Lcell) -> None:
        self.Lcell = Lcell
        self._dev = None
        self._host = {"vx": v[..., 0], "vy": v[..., 1], "vz": v[..., 2], "mass": mass}
        self.Nsize = len(mass)
        self.Lbox = self.Nsize * self.Lcell

    @classmethod
    def _from_device(cls, nn_t, vel_t, rho_t, Lcell, nn_pos_t=None, spay_t=None):
        self = cls.__new__(cls)
        self.Lcell = Lcell
        self._dev = {"nn": nn_t, "vel": vel_t, "rho": rho_t, "nn_pos": nn_pos_t, "spay": spay_t}
        self._host = {}
        self.Nsize = int(nn_t.shape[0])
        self.Lbox = self.Nsize * self.Lcell
        return self

    # -- lazily materialised numpy views (reference attributes vx, vy, vz, mass) ------------------
    def _materialise(self):
        d = self._dev
        N = self.Nsize
        idx = d["nn"].reshape(-1)
        rho = d["rho"]
        payload = (d["vel"] * rho[:, None])                      # rho*v in the input dtype, interp.py:203-208
        w = _lib.gather_rows(idx, payload.contiguous())         # f[index], interp.py:1043
        r = _lib.gather_rows(idx, rho)
        v = (w / r[:, None]).reshape(N, N, N, 3).cpu().numpy()  # interp.py:272
        m = (r * self.Lcell ** 3).reshape(N, N, N).cpu().numpy()  # interp.py:273
        self._host = {"vx": v[..., 0], "vy": v[..., 1], "vz": v[..., 2], "mass": m}

    def _get(self, name):
        if name not in self._host:
            self._materialise()
        return self._host[name]

    def _set(self, name, value):
        # assigning a plane (trim, down_sample, user code) detaches the field from its device-resident form
        if self._dev is not None:
            self._materialise()
            self._dev = None
        self._host[name] = value

    vx = property(lambda self: self._get("vx"), lambda self, a: self._set("vx", a))
    vy = property(lambda self: self._get("vy"), lambda self, a: self._set("vy", a))
    vz = property(lambda self: self._get("vz"), lambda self, a: self._set("vz", a))
    mass = property(lambda self: self._get("mass"), lambda self, a: self._set("mass", a))

    def __getitem__(self, index):                                   # interp.py:474
        return BoxField(self.get_v()[index], self.mass[index], self.Lcell)

    def __array__(self, dtype=None, copy=None):                     # interp.py:478
        a = self.get_data()
        return a if dtype is None else a.astype(dtype)

    def __array_wrap__(self, arr, context=None, return_scalar=False):   # interp.py:482 (mass = arr[..., :3], as there)
        return BoxField(arr[..., :3], arr[..., :3], self.Lcell)

    def get_v(self):
        return np.stack((self.vx, self.vy, self.vz), axis=3)

    def get_density(self):
        return self.mass / self.Lcell ** 3

    def get_data(self):
        return np.stack((self.vx, self.vy, self.vz, self.mass), axis=3)

    # -- lattice surgery (interp.py:611-636) -----------------------------------------------------------
    def trim(self, Nmargin, Nbrick) -> None:
        """Keep the central Nbrick^3 block (drop Nmargin nodes on every side); Nsize and Lbox follow as in the reference."""
        n1, n2 = Nmargin, Nmargin + Nbrick
        vx, vy, vz, m = self.vx, self.vy, self.vz, self.mass
        self.vx, self.vy, self.vz = vx[n1:n2, n1:n2, n1:n2], vy[n1:n2, n1:n2, n1:n2], vz[n1:n2, n1:n2, n1:n2]
        self.mass = m[n1:n2, n1:n2, n1:n2]
        self.Nsize = self.Nsize - 2 * Nmargin
        self.Lbox = self.Lbox * Nbrick / (Nbrick + 2 * Nmargin)

    def down_sample(self, n) -> None:
        """Average n^3 blocks of mass and momentum; the velocity becomes the mass-weighted one.  `Nsize /= n` turns Nsize
        into a float exactly as interp.py:635 does."""
        px, py, pz = (down_sample(c * self.mass, n) for c in (self.vx, self.vy, self.vz))
        m = down_sample(self.mass, n)
        m[np.where(m == 0)] = 1e-10                                   # avoid zero mass (:630)
        self.mass = m
        self.vx, self.vy, self.vz = px / m, py / m, pz / m
        self.Nsize /= n
        self.Lcell *= n

    def mean_kinetic_energy(self) -> float:                          # interp.py:639
        return 0.5 * np.mean(self.mass * (self.vx ** 2 + self.vy ** 2 + self.vz ** 2))

    # -- totals (interp.py:639-666) -----------------------------------------------------------------
    def total_mass(self):
        return np.sum(self.mass)

    def total_momentum(self):
        return np.array([np.sum(self.mass * self.vx), np.sum(self.mass * self.vy), np.sum(self.mass * self.vz)])

    def total_kinetic_energy(self):
        return 0.5 * np.sum(self.mass * (self.vx ** 2 + self.vy ** 2 + self.vz ** 2))

    def specific_kinetic_energy(self):
        return self.total_kinetic_energy() / self.total_mass()

    # -- device planes for one quantity --------------------------------------------------------------
    def _planes(self, quantity):
        """-> (list of float32 CUDA cubes to transform, multiplicity).  Field algebra interp.py:501-557."""
        torch = _lib._torch()
        strict = self.strict_reference
        if self._dev is not None and self._dev.get("spay") is not None:
            d = self._dev          # payload already in cell-sorted order: near-sequential reads
            if quantity == "velocity":
                f = _lib.fields_sorted(d["nn_pos"], d["spay"], want_v=True)
                return [f["vx"], f["vy"], f["vz"]], 1.0
            if quantity == "momentum":
                if strict:
                    f = _lib.fields_sorted(d["nn_pos"], d["spay"], want_v=False, want_p=(True, False, False))
                    return [f["px"]], 3.0          # three identical components (interp.py:523-525)
                f = _lib.fields_sorted(d["nn_pos"], d["spay"], want_v=False, want_p=(True, True, True))
                return [f["px"], f["py"], f["pz"]], 1.0
            f = _lib.fields_sorted(d["nn_pos"], d["spay"], want_v=False, want_e=True)
            return [f["e"]], 1.0
        if self._dev is not None:
            d = self._dev
            lc3 = float(self.Lcell) ** 3
            if quantity == "velocity":
                f = _lib.build_fields(d["nn"], d["vel"], d["rho"], lc3, want_v=True)
                return [f["vx"], f["vy"], f["vz"]], 1.0
            if quantity == "momentum":
                if strict:
                    f = _lib.build_fields(d["nn"], d["vel"], d["rho"], lc3, want_v=False, want_p=(True, False, False))
                    return [f["px"]], 3.0          # three identical components (interp.py:523-525)
                f = _lib.build_fields(d["nn"], d["vel"], d["rho"], lc3, want_v=False, want_p=(True, True, True))
                return [f["px"], f["py"], f["pz"]], 1.0
            f = _lib.build_fields(d["nn"], d["vel"], d["rho"], lc3, want_v=False, want_e=True)
            return [f["e"]], 1.0
        up = lambda a: _lib.to_device(np.ascontiguousarray(a), dtype=torch.float32)  # noqa: E731
        if quantity == "velocity":
            return [up(self.vx), up(self.vy), up(self.vz)], 1.0
        if quantity == "momentum":
            if strict:
                return [up(self.vx * self.mass)], 3.0
            return [up(self.vx * self.mass), up(self.vy * self.mass), up(self.vz * self.mass)], 1.0
        return [up(self.mass * (self.vx ** 2 + self.vy ** 2 + self.vz ** 2))], 1.0

    def _norm(self):
        a = (self.Lbox / (2 * np.pi)) ** 1.5 / self.Nsize ** 3      # interp.py:1380
        return 0.5 * a * a

    def _power_grid(self, quantity):
        planes, mult = self._planes(quantity)
        plan = _plan_for(self.Lbox, self.Nsize, None)
        return (plan.power_cube(planes) * (mult * self._norm())).cpu().numpy()

    def velocity_power(self) -> np.ndarray:
        """Full power cube 1/2 sum_c |a FFT(v_c)|^2.  interp.py:501-518."""
        return self._power_grid("velocity")

    def momentum_power(self) -> np.ndarray:
        """interp.py:521-541."""
        return self._power_grid("momentum")

    def kinetic_energy_power(self) -> np.ndarray:
        """interp.py:544-557."""
        return self._power_grid("energy")

    def fold(self, m, beta, quantity="velocity"):
        """Folded velocity field for the residue class `beta` of the folding factor `m` (interp.py:598-609):
        phase multiply, sum of the m^3 sub-blocks, / m^1.5 -- one kernel; the result stays on the device until `.f` is read."""
        if quantity != "velocity":
            raise Exception("""Unsupported physical quantity name.""")
        planes, _ = self._planes("velocity")
        phi = _lib.fold_field(planes, int(m), [int(b) for b in beta])
        return FoldedBox(phi, m, beta, self.Lbox / m, self.Nsize // m)

    def spctrm(self, quantity="velocity", kmin=None, kmax=None, kres=None) -> PowerSpectrum:
        """Shell-averaged spectrum of `quantity` ('velocity' | 'momentum' | 'energy').  interp.py:560-595.
        FFT, |F|^2 and the k-shell histogram run fused on the device (no power cube is formed)."""
        if kmin is None:
            kmin = 2 * np.pi / self.Lbox
        if kmax is None:
            kmax = np.pi / self.Lcell
        if kres is None:
            kres = kmin
        if quantity not in ("velocity", "momentum", "energy"):
            raise Exception("""Unrecognized physical quantity name.
        Supported: 'velocity', 'momentum', 'energy'.""")
        centres, edges = _edges_lib(kmin, kmax, kres)
        planes, mult = self._planes(quantity)
        plan = _plan_for(self.Lbox, self.Nsize, edges)
        raw, ns = plan.fields(planes)
        Psum = raw * (mult * self._norm())
        Nsample = ns.astype(np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            P = Psum / Nsample
        P[Nsample == 0] = 0                                   # interp.py:1479
        Pkk = np.column_stack((centres, P, Psum, Nsample))
        Pkk[:, 1] *= 4 * np.pi * Pkk[:, 0] ** 2               # interp.py:590
        return PowerSpectrum(Pkk)


class FoldedBox:
    """A folded (complex) field and the sub-spectrum it yields.  interp.py:740-800.
    `f` is a numpy complex128 array [n,n,n,3] (vector) or [n,n,n] (scalar), or the CUDA tensor BoxField.fold produced --
    reading the attribute `f` returns numpy either way."""

    def __init__(self, f, m, beta, Lbox, Nsize) -> None:
        self._f = f
        self.Lbox = Lbox
        self.Nsize = Nsize
        self.Lcell = Lbox / Nsize
        self.m = m
        self.beta = beta
        self.totalLbox = Lbox * m

    @property
    def f(self):
        if not isinstance(self._f, np.ndarray):
            self._f = self._f.cpu().numpy()
        return self._f

    @f.setter
    def f(self, value):
        self._f = value

    def fold_spctrm(self, fft_object=None, beta=np.array([0, 0, 0]), kmin=None, kmax=None, kres=None) -> PowerSpectrum:
        """Sub-spectrum of this residue class (interp.py:755-791): c2c transform of the folded field, 1/2 sum |a F|^2, |k|
        pairing shifted by 2 pi beta / totalLbox, shell histogram -- all on the device.  `fft_object` (a pyFFTW plan in the
        reference) is ignored.  Unlike the reference, `f` is not overwritten by the power grid."""
        torch = _lib._torch()
        if kmin is None:
            kmin = 2 * np.pi / self.totalLbox
        if kmax is None:
            kmax = np.pi / self.Lcell
        if kres is None:
            kres = kmin
        f = self._f
        if isinstance(f, np.ndarray):
            if f.ndim not in (3, 4):
                raise Exception("""Unrecognized field shape.
        Supported: (N, N, N, 3), (N, N, N).""")
            f = _lib.to_device(np.ascontiguousarray(f, dtype=np.complex128))
        n = int(self.Nsize)
        a = (self.Lbox / (2 * np.pi)) ** 1.5 / n ** 3                      # interp.py:1398
        P = _lib.fold_power(f) * (0.5 * a * a)
        ks = _k_axis(self.Lbox, n)
        shift = 2 * np.pi * np.asarray(beta) / self.totalLbox              # interp.py:780
        axes = [ks + shift[c] if shift[c] > 0 else ks for c in range(3)]   # interp.py:1453-1458
        k = _lib.k_magnitude(*axes)
        centres, edges = _edges_lib(kmin, kmax, kres)
        Psum, ns = _lib.hist_weighted(k, P.reshape(-1), edges)
        Nsample = ns.astype(np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            Pm = Psum / Nsample
        Pm[Nsample == 0] = 0
        Pkk = np.column_stack((centres, Pm, Psum, Nsample))
        Pkk[:, 1] *= 4 * np.pi * Pkk[:, 0] ** 2
        return PowerSpectrum(Pkk, m=self.m, beta=beta)

    def save(self, run_output_dir) -> None:
        """Pickle under `run_output_dir` as folded_field_b{beta}.pkl (interp.py:793-800)."""
        import os
        import pickle
        _ = self.f                                                          # materialise on the host before pickling
        with open(os.path.join(run_output_dir, "folded_field_b{}{}{}.pkl".format(*self.beta)), "wb") as file:
            pickle.dump(self, file)


# ------------------------------------------------------------------------------------------ functions
def _get_phase(beta, totalNsize, Nphase, x0, y0, z0) -> np.ndarray:
    """exp(-i (2 pi / totalNsize) beta.x) on the brick [x0, x0+Nphase) x ..., complex128.  interp.py:1215-1225 (host helper;
    BoxField.fold evaluates the phase inside its kernel)."""
    x, y, z = (np.arange(o, o + Nphase) for o in (x0, y0, z0))
    xxx, yyy, zzz = np.meshgrid(x, y, z, indexing="ij")
    return np.exp(-1j * (2 * np.pi / totalNsize) * (beta[0] * xxx + beta[1] * yyy + beta[2] * zzz))


def _apply_phase(f, phase) -> np.ndarray:
    """complex128 copy of f times the phase (every vector component).  interp.py:1195-1212."""
    phi = np.array(f, dtype=np.complex128)
    phi *= phase if phi.shape == phase.shape else phase[..., None]
    return phi


def fold_field(f, m):
    """Sum of the m^3 sub-blocks of size N/m, in (i, j, k) order.  interp.py:1228-1252."""
    if m == 1:
        return f
    n1, n2, n3 = f.shape[0] // m, f.shape[1] // m, f.shape[2] // m
    r = 0.0
    for i in range(m):
        for j in range(m):
            for k in range(m):
                r = r + f[i * n1:(i + 1) * n1, j * n2:(j + 1) * n2, k * n3:(k + 1) * n3, ...]
    return r


def _lattice_axis(Lbox, Nsize):
    Lcell = Lbox / Nsize
    return np.linspace(Lcell / 2, Lbox + Lcell / 2, Nsize)       # interp.py:1062-1063


def make_grid_coords(Lbox, Nsize) -> np.ndarray:
    """[Nsize^3, 3] lattice node coordinates, C order.  interp.py:1060-1069."""
    ax = _lattice_axis(Lbox, Nsize)
    grid = np.meshgrid(ax, ax, ax, indexing="ij")
    return np.reshape(grid, (3, Nsize ** 3)).T


def _separable_axes(query_pos, Nsize):
    """Recover the three axis tables from an [N^3,3] C-order lattice; None if it is not separable."""
    q = np.asarray(query_pos, dtype=np.float64)
    if q.shape != (Nsize ** 3, 3):
        return None
    q = q.reshape(Nsize, Nsize, Nsize, 3)
    ax = (q[:, 0, 0, 0], q[0, :, 0, 1], q[0, 0, :, 2])
    ok = (np.array_equal(q[..., 0], np.broadcast_to(ax[0][:, None, None], q.shape[:3])) and
          np.array_equal(q[..., 1], np.broadcast_to(ax[1][None, :, None], q.shape[:3])) and
          np.array_equal(q[..., 2], np.broadcast_to(ax[2][None, None, :], q.shape[:3])))
    return ax if ok else None


def ann_interpolate(data_pos, query_pos, f, Nsize, eps, treetype="kd", searchtype="standard"):
    """f sampled at the nearest particle of every query node, reshaped to the cube.  interp.py:1018-1049.
    The query set must be a separable lattice (what make_grid_coords produces)."""
    if eps != 0.0:
        raise Exception("vpower_b200: only the exact search (eps=0) is implemented")
    ax = _separable_axes(query_pos, Nsize)
    if ax is None:
        raise Exception("vpower_b200: query_pos must be an [Nsize^3,3] separable lattice in C order")
    pos = np.asarray(data_pos)
    dt = np.float64 if pos.dtype == np.float64 else np.float32
    nn = _lib.nn_grid(_lib.to_device(np.asarray(pos, dtype=dt)), *ax)
    f = np.asarray(f)
    if f.ndim not in (1, 2):
        raise Exception("Unsupported data shape.")
    if f.dtype.itemsize * (1 if f.ndim == 1 else f.shape[1]) % 4 != 0:
        raise Exception("Unsupported data shape.")
    g = _lib.gather_rows(nn.reshape(-1), _lib.to_device(f)).cpu().numpy()
    if f.ndim == 1:
        return np.reshape(g, (Nsize, Nsize, Nsize))
    return np.reshape(g, (Nsize, Nsize, Nsize, f.shape[1]))


def deposit_to_grid(f, pos, Nsize, Lbox):
    """Periodic nearest-grid-point deposit, f64 grid.  interp.py:996-1015."""
    f = np.asarray(f)
    pos = np.asarray(pos)
    dt = np.float64 if pos.dtype == np.float64 else np.float32
    g = _lib.deposit_ngp(_lib.to_device(np.asarray(pos, dtype=dt)), _lib.to_device(f), Nsize, Lbox)
    return g.cpu().numpy()


def _k_axis(Lbox, Nsize):
    Lcell = Lbox / float(Nsize)
    return 2 * np.pi * np.fft.fftfreq(Nsize, Lcell)              # interp.py:1448-1449


def _edges_lib(kmin, kmax, spacing):
    centres = np.arange(kmin, kmax + spacing, spacing)           # interp.py:1472
    edges = np.arange(kmin - spacing / 2, kmax + 3 * spacing / 2, spacing)   # interp.py:1473
    return centres, edges


_plans = {}


def _plan_for(Lbox, Nsize, edges):
    if edges is None:   # power-cube use: the edges are irrelevant, any valid pair will do
        edges = np.array([0.0, 1.0])
    key = (float(Lbox), int(Nsize), edges.tobytes())
    if key not in _plans:
        if len(_plans) > 8:
            _plans.clear()
        _plans[key] = _lib.PkPlan(Nsize, _k_axis(Lbox, Nsize), edges)
    return _plans[key]


def _power_cube(fields, Lbox, Nsize, mult=1.0):
    torch = _lib._torch()
    a = (Lbox / (2 * np.pi)) ** 1.5 / Nsize ** 3
    planes = [_lib.to_device(np.ascontiguousarray(f), dtype=torch.float32) for f in fields]
    P = _plan_for(Lbox, Nsize, None).power_cube(planes)
    return (P * (0.5 * a * a * mult)).cpu().numpy()


def _vector_power(fx, fy, fz, Lbox, Nsize):
    """1/2 (|a F_x|^2 + |a F_y|^2 + |a F_z|^2), full [N,N,N] cube.  interp.py:1372-1387."""
    return _power_cube([fx, fy, fz], Lbox, Nsize)


def _scalar_power(f, Lbox, Nsize):
    """1/2 |a F|^2.  interp.py:1408-1421."""
    return _power_cube([f], Lbox, Nsize)


def _FFTW_vector_power(f, Lbox, Nsize, fft_object=None):
    """interp.py:1390-1405 with the transform on the GPU: f is [N,N,N,3]; `fft_object` (a pyFFTW plan there) is ignored."""
    f = np.asarray(f)
    return _vector_power(f[..., 0], f[..., 1], f[..., 2], Lbox, Nsize)


def _FFTW_scalar_power(f, Lbox, Nsize, fft_object=None):
    """interp.py:1424-1437; `fft_object` is ignored."""
    return _scalar_power(f, Lbox, Nsize)


def _pair_power(Pk, Lbox, Nsize, shift=np.array([0, 0, 0])):
    """[N^3,2] (|k|, P) pairs, C order.  interp.py:1440-1467 (shift applied only where > 0, as there)."""
    ks = _k_axis(Lbox, Nsize)
    axes = [ks + shift[c] if shift[c] > 0 else ks for c in range(3)]
    k = _lib.k_magnitude(*axes).cpu().numpy()
    return np.transpose(np.stack((k, np.ravel(Pk))))


def _hist_sample(Pk_pair, kmin, kmax, spacing):
    """Mean power per k shell -> [nbins,4] (centre, P, Psum, Nsample).  interp.py:1470-1482."""
    torch = _lib._torch()
    centres, edges = _edges_lib(kmin, kmax, spacing)
    pairs = _lib.to_device(np.ascontiguousarray(Pk_pair), dtype=torch.float64)
    Psum, ns = _lib.hist_weighted(pairs[:, 0].contiguous(), pairs[:, 1].contiguous(), edges)
    Nsample = ns.astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        P = Psum / Nsample
    P[Nsample == 0] = 0
    return np.column_stack((centres, P, Psum, Nsample))


def down_sample(r, n):
    """Mean over n^3 blocks (interp.py:1255-1267).  The reference indexes a fourth axis unconditionally; here trailing axes
    are optional, so the 3-D planes BoxField.down_sample passes work as well."""
    if n == 1:
        return r
    d = 0.0
    for i in range(n):
        for j in range(n):
            for k in range(n):
                d = d + r[i::n, j::n, k::n, ...]
    d /= n ** 3
    return d


def _vec_to_vm_grid(vec_grid, Lcell):
    """[N,N,N,4] (rho*v, rho) -> (v [N,N,N,3], m [N,N,N]) in place, interp.py:970-992 (deprecated there, kept for callers)."""
    rho_grid = vec_grid[:, :, :, 3]
    m_grid = rho_grid * Lcell ** 3
    vec_grid[:, :, :, 0] /= rho_grid
    vec_grid[:, :, :, 1] /= rho_grid
    vec_grid[:, :, :, 2] /= rho_grid
    return vec_grid[:, :, :, 0:3], m_grid


def check_conservation(gasParticles, boxField) -> tuple:
    """Ratios (field / particles) of total mass, momentum, kinetic energy and specific kinetic energy.
    interp.py:1269-1319 (the reference also prints them)."""
    mass = boxField.total_mass() / gasParticles.total_mass()
    mom = boxField.total_momentum() / gasParticles.total_momentum()
    en = boxField.total_kinetic_energy() / gasParticles.total_kinetic_energy()
    sp = boxField.specific_kinetic_energy() / gasParticles.specific_kinetic_energy()
    print("Total mass restored by {:.3%}".format(mass))
    print("Total momentum restored by ({:.3%}, {:.3%}, {:.3%})".format(*mom))
    print("Total kinetic energy restored by {:.3%}".format(en))
    print("Specific kinetic energy restored by {:.3%}".format(sp))
    return mass, mom, en, sp
