"""CPU-only checks: the C-ABI library loads and exports every symbol of include/vpower_b200.h, the host-side
mirror keeps the reference's conventions, and the product path fails loudly without a GPU."""
import ctypes
import os
import re
import pickle

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from vpower import _lib
    return _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vpower_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    names = _declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(built.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vpower_b200.h but not exported"
    assert sorted(built.SIGNATURES) == names, "ctypes prototypes out of sync with the header"
    assert built.load_library().vp_version() >= 100


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import vpower.interp as vi
    h = ctypes.c_void_p()
    rc = built.load_library().vp_ctx_create(0, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in built.load_library().vp_last_error()
    with pytest.raises(built.VPowerError):
        vi.GasParticles(np.zeros((8, 3)), np.ones(8), np.ones(8), np.zeros((8, 3)), 1.0).ann_interp_to_field(4)
    with pytest.raises(built.VPowerError):
        vi.deposit_to_grid(np.ones(8), np.zeros((8, 3)), 4, 1.0)
    with pytest.raises(built.VPowerError):
        vi.BoxField(np.zeros((4, 4, 4, 3)), np.ones((4, 4, 4)), 0.25).spctrm("velocity")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "large-velocity-power-spectrum_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "vpower_oracle" not in txt and "refshims" not in txt, f


def test_host_expressions_match_oracle(built, orc):
    import vpower.interp as vi
    for N, L in ((16, 1.0), (24, 2.5), (1024, 1.0), (500, 0.7)):
        assert np.array_equal(vi._lattice_axis(L, N), orc.lattice_axis_lib(L, N))
        assert np.array_equal(vi._k_axis(L, N), orc.k_axis(L, N))
        kmin, kmax = 2 * np.pi / L, np.pi / (L / N)
        for a, b in zip(vi._edges_lib(kmin, kmax, kmin), orc.edges_lib(kmin, kmax, kmin)):
            assert np.array_equal(a, b)
    assert np.array_equal(vi.make_grid_coords(2.5, 6), orc.make_grid_coords(2.5, 6))
    # the script's linspace edges give 511 bins at (1024, 1)
    c, e = orc.edges_script(2 * np.pi, np.pi * 1024, 2 * np.pi)
    assert len(c) == 511 and len(e) == 512


def test_separable_axes_detection(built):
    import vpower.interp as vi
    q = vi.make_grid_coords(1.0, 5)
    ax = vi._separable_axes(q, 5)
    assert ax is not None and np.array_equal(ax[0], vi._lattice_axis(1.0, 5))
    q2 = q.copy()
    q2[7, 1] += 1e-3
    assert vi._separable_axes(q2, 5) is None
    assert vi._separable_axes(q[:-1], 5) is None


def test_gasparticles_preprocessing(built):
    """shift_to_origin / remove_bulk_velocity / payload (interp.py:169-213)."""
    import vpower.interp as vi
    rng = np.random.default_rng(0)
    pos, vel = rng.random((100, 3)) + 3.0, rng.normal(size=(100, 3)) + 5.0
    mass, dens = rng.random(100) + 0.5, rng.random(100) + 1.0
    gp = vi.GasParticles(pos.copy(), mass, dens, vel.copy(), 1.0)
    gp.shift_to_origin()
    gp.remove_bulk_velocity()
    assert np.allclose(gp.pos.min(axis=0), 0)
    assert np.allclose((mass[:, None] * gp.v).sum(0), 0, atol=1e-12)
    pay = gp.density_velocity_vector()
    assert pay.shape == (100, 4) and np.array_equal(pay[:, 3], dens) and np.array_equal(pay[:, 0], gp.v[:, 0] * dens)
    assert len(gp) == 100 and len(gp[:10]) == 10
    assert np.isclose(gp.total_kinetic_energy(), 0.5 * np.sum(mass * (gp.v ** 2).sum(1)))


def test_power_spectrum_container(built, tmp_path):
    """vpower.spctrm keeps the reference semantics (spctrm.py:55-380)."""
    import vpower.spctrm as vs
    k = np.arange(1, 6) * 2 * np.pi
    a = vs.PowerSpectrum(np.column_stack((k, np.ones(5), np.full(5, 2.0), np.full(5, 4.0))))
    b = vs.PowerSpectrum(np.column_stack((k, np.ones(5), np.full(5, 6.0), np.full(5, 4.0))))
    assert len(a) == 5 and np.isclose(a.Lbox(), 1.0) and np.isclose(a.kres(), 2 * np.pi)
    c = a.copy()
    c.add(b)
    assert np.allclose(c.Psum, 8) and np.allclose(c.Nsample, 8) and np.allclose(c.P, 4 * np.pi * k ** 2)
    c.remove(b)
    assert np.allclose(c.Psum, 2) and np.allclose(c.Nsample, 4)
    with pytest.raises(ValueError):
        c.remove(b)
    with pytest.raises(Exception):
        a.add(vs.PowerSpectrum(np.ones((3, 4))))
    assert np.isclose(a.energy(), np.sum(a.P[:-1] * np.diff(k)))
    e = vs.empty_spectrum_like(a)
    assert np.all(e.Psum == 0) and e.m == 0
    a.save(str(tmp_path))
    assert os.path.isfile(tmp_path / "full_spctrm.pkl")
    assert np.array_equal(vs.PowerSpectrum.load(str(tmp_path)).Psum, a.Psum)
    s1 = vs.PowerSpectrum(a.data(), m=2, beta=np.array([0, 1, 1]))
    s1.save(str(tmp_path))
    assert os.path.isfile(tmp_path / "sub_spctrm_b011.pkl")
    sl = vs.SpectrumList.load(str(tmp_path))
    assert len(sl) == 1 and sl.m == 2 and np.array_equal(sl[np.array([0, 1, 1])].Psum, a.Psum)
    assert vs.init_beta_space(2).shape == (8, 3)
    assert np.isclose(vs.relative_diff(a.copy(), b.copy()), 0.0)
    pw = vs.PowerSpectrum(np.column_stack((k, k ** -2.0, k, k)))
    assert np.isclose(pw.index(), -2.0)
    pickle.loads(pickle.dumps(a))


def test_fft_index_model():
    """numpy model of the kernel's stage/thread/register index maps (csrc/fft_core.cuh LineFFT) vs np.fft."""
    def dft(v):
        r = len(v)
        kk = np.arange(r)
        return np.exp(-2j * np.pi * np.outer(kk, kk) / r) @ v

    def model(x, R2, R3):
        L = len(x)
        T = L // 16
        W = np.exp(-2j * np.pi * np.arange(L) / L)
        sm = np.zeros(L, complex)
        for t in range(T):
            v = dft(np.array([x[j * T + t] for j in range(16)]))
            for k1 in range(16):
                sm[k1 * T + t] = v[k1] * W[t * k1]
        M2 = 16 // R2
        regs = np.zeros((T, 16), complex)
        sm2 = np.zeros(L, complex)
        for t in range(T):
            for b in range(M2):
                g = t * M2 + b
                k1, d3 = g // R3, g % R3
                u = dft(np.array([sm[k1 * T + a * R3 + d3] for a in range(R2)]))
                for k2 in range(R2):
                    val = u[k2] * (W[16 * d3 * k2] if R3 > 1 else 1)
                    regs[t, k2 + R2 * b] = val
                    sm2[k1 * T + k2 * R3 + d3] = val
        X = np.zeros(L, complex)
        if R3 == 1:
            for t in range(T):
                for j in range(16):
                    X[(t * M2 + j // R2) + 16 * (j % R2)] = regs[t, j]          # kout, two-stage form
            return X
        M3 = 16 // R3
        for t in range(T):
            for b in range(M3):
                g = t + T * b
                u = dft(np.array([sm2[(g % 16) * T + (g // 16) * R3 + a] for a in range(R3)]))
                for a in range(R3):
                    X[t + T * (j := b) + 16 * R2 * a] = u[a]                      # kout, three-stage form
        return X

    rng = np.random.default_rng(0)
    for R2, R3 in ((2, 1), (4, 1), (8, 1), (16, 1), (16, 2), (16, 4), (16, 8)):
        L = 16 * R2 * R3
        x = rng.normal(size=L) + 1j * rng.normal(size=L)
        assert np.abs(model(x, R2, R3) - np.fft.fft(x)).max() < 1e-11 * L



def test_boxfield_host_helpers_match_reference_expressions():
    """trim / down_sample / mean_kinetic_energy / slicing / _vec_to_vm_grid (interp.py:474-482, 611-641, 970-992, 1255-1267):
    pure host code, checked against the reference's expressions written out independently."""
    import vpower.interp as vi
    rng = np.random.default_rng(3)
    N = 12
    v = rng.standard_normal((N, N, N, 3))
    m = 1.0 + rng.random((N, N, N))
    bf = vi.BoxField(v.copy(), m.copy(), 0.25)
    assert bf.Nsize == N and bf.Lbox == N * 0.25
    assert np.isclose(bf.mean_kinetic_energy(), 0.5 * np.mean(m * (v ** 2).sum(-1)))
    sub = bf[2:6]
    assert isinstance(sub, vi.BoxField) and sub.Nsize == 4 and np.array_equal(sub.vy, v[2:6, ..., 1])
    assert np.array_equal(np.asarray(bf), np.concatenate([v, m[..., None]], axis=3))
    # trim
    t = vi.BoxField(v.copy(), m.copy(), 0.25)
    t.trim(2, 8)
    assert t.Nsize == 8 and np.isclose(t.Lbox, N * 0.25 * 8 / 12)
    assert np.array_equal(t.vx, v[2:10, 2:10, 2:10, 0]) and np.array_equal(t.mass, m[2:10, 2:10, 2:10])
    # down_sample: block means of mass and momentum, mass-weighted velocity
    d = vi.BoxField(v.copy(), m.copy(), 0.25)
    d.down_sample(2)
    blk = lambda a: a.reshape(N // 2, 2, N // 2, 2, N // 2, 2).mean(axis=(1, 3, 5))
    assert np.allclose(d.mass, blk(m)) and np.allclose(d.vz, blk(v[..., 2] * m) / blk(m))
    assert d.Nsize == N / 2 and d.Lcell == 0.5
    assert np.allclose(vi.down_sample(np.stack([m, m], axis=3), 3)[..., 1], m.reshape(4, 3, 4, 3, 4, 3).mean(axis=(1, 3, 5)))
    assert vi.down_sample(m, 1) is m
    # _vec_to_vm_grid
    rho = 1.0 + rng.random((N, N, N))
    vec = np.concatenate([v * rho[..., None], rho[..., None]], axis=3)
    vv, mm = vi._vec_to_vm_grid(vec, 0.5)
    assert np.allclose(vv, v) and np.allclose(mm, rho * 0.125)


def test_cell_list_plan_invariants_without_a_device():
    """vp_nn_grid_plan is host arithmetic: the linear cell index must fit 32 bits, the buckets of 2^bucket_shift consecutive
    cells must cover the grid with at most 2048 of them (the shared-memory counters of the bucket pass), and the
    1-particle-per-node lattices must come out corner aligned -- for every BASELINE configuration (cfg5 cannot be run on one
    device) and for awkward shapes."""
    from vpower import _lib

    def check(np_particles, qx, qy, qz, opts=None):
        p = _lib.nn_grid_plan(np_particles, qx, qy, qz, opts)
        gx, gy, gz = p["cells_x"], p["cells_y"], p["cells_z"]
        assert gx >= 1 and gy >= 1 and gz >= 1
        ncells = gx * gy * gz
        assert ncells < 2 ** 32 - 1
        assert 1 <= p["n_buckets"] <= 2048 and 0 <= p["bucket_shift"] <= 31
        assert p["n_buckets"] == -(-ncells // (1 << p["bucket_shift"]))           # the buckets cover every cell
        return p

    for N, Np in ((64, 1 << 18), (256, 1 << 24), (512, 1 << 27), (1024, 1 << 30)):
        ax = np.linspace(0.5 / N, 1.0 + 0.5 / N, N)                            # library lattice
        p = check(Np, ax, ax, ax)
        assert (p["cells_x"], p["cells_y"], p["cells_z"]) == (N + 1,) * 3 and p["corner_aligned"] == 1
        # about 2^20 particles per bucket: the counting sort inside a bucket's window of records runs out of L2
        assert Np <= (1 << 21) or 1 << 19 <= Np / p["n_buckets"] <= 1 << 21
    p = check(1 << 30, *(np.linspace(0.5 / 1024, 1.0 + 0.5 / 1024, 1024),) * 3)
    assert p["bucket_shift"] == 20 and p["n_buckets"] == 1028 and p["scratch_MiB"] < 150 * 1024
    # cfg5 (2048^3): one device cannot hold it; a slab of it (8 ranks) must plan fine, the whole lattice gets coarser cells
    ax5 = np.linspace(0.5 / 2048, 1.0 + 0.5 / 2048, 2048)
    o = _lib.NNOpts()
    o.use_x_keep, o.x_keep_lo, o.x_keep_hi = 1, ax5[256] - 4 / 2048, ax5[511] + 4 / 2048
    o.x_lo_is_domain_edge = o.x_hi_is_domain_edge = 0
    p = check((1 << 30) + (1 << 26), ax5[256:512], ax5, ax5, o)
    assert p["cells_y"] == 2049 and p["cells_z"] == 2049 and 256 <= p["cells_x"] <= 280
    p = check((1 << 31) - 1, ax5, ax5, ax5)
    assert p["cells_x"] * p["cells_y"] * p["cells_z"] < 2 ** 32 - 1
    # awkward shapes: a single long line of nodes, a plane, very few / very many particles per node, explicit cell counts
    line = np.linspace(0.0, 1.0, 100000)
    check(1000, np.array([0.5]), np.array([0.5]), line)
    check(1 << 26, np.array([0.5]), np.array([0.5]), line)                      # many cells along z: coarsened to fit
    check(5, line[:300], line[:200], np.array([0.1, 0.2]))
    check(1 << 31 - 1, np.linspace(0, 1, 16), np.linspace(0, 1, 16), np.linspace(0, 1, 16))
    o2 = _lib.NNOpts()
    o2.cells_x, o2.cells_y, o2.cells_z = 7, 100000, 3
    check(12345, line[:50], line[:60], line[:70], o2)



def test_bench_generator_matches_oracle_hash_and_is_rank_count_independent():
    """bench.py's counter-based particle generator: same integer hash as the oracle's hash_uniform (positions bit-equal to
    orc.synth_particles), and a rank's index range is a pure function of (seed, index) -- the union over ranks is the same
    particle set at every GPU count."""
    import torch
    import bench
    import vpower_oracle as orc
    for seed, stream in ((0, 0), (3, 2), (4, 7)):
        assert np.array_equal(orc.hash_uniform(seed, 50000, stream), bench.hash_uniform_t(torch, seed, 0, 50000, stream, "cpu").numpy())
    wl = bench.WORKLOADS["cfg1"]
    p, v, r = bench.synth_range(torch, wl, 0, 4000, device="cpu")
    assert np.array_equal(p.numpy(), orc.synth_particles(wl["seed"], wl["Np"], 1.0)[0][:4000])
    for lo, hi in ((0, 1000), (1000, 4000)):
        p2, v2, r2 = bench.synth_range(torch, wl, lo, hi, device="cpu", chunk=700)
        assert torch.equal(p[lo:hi], p2) and torch.equal(v[lo:hi], v2) and torch.equal(r[lo:hi], r2)
    wl3 = dict(bench.WORKLOADS["cfg3"], Np=1 << 18)
    pc, vc, rc = bench.synth_range(torch, wl3, 0, 1 << 18, device="cpu")
    assert float(pc.min()) >= 0.0 and float(pc.max()) < 1.0 and bool(torch.isfinite(vc).all())
    h = np.histogramdd(pc.numpy(), bins=16, range=[(0, 1)] * 3)[0]
    assert h.max() > 20 * h.mean()                      # genuinely clustered


def test_bench_integer_shell_closed_form():
    import bench
    for N in (16, 48, 64):
        n1 = np.fft.fftfreq(N, 1.0 / N)
        n2 = (n1[:, None, None] ** 2 + n1[None, :, None] ** 2 + n1[None, None, :] ** 2).ravel()
        ref = np.bincount(np.floor(np.sqrt(n2) + 0.5).astype(np.int64), minlength=N)[1:N // 2 + 1]
        assert np.array_equal(bench.integer_shell_counts(N), ref)
    assert bench.integer_shell_counts(16).tolist()[:8] == [18, 62, 98, 210, 350, 450, 602, 687]
    assert bench.integer_shell_counts(64).sum() == 143457 and bench.integer_shell_counts(256).sum() == 8886577


@pytest.mark.parametrize("N,nranks", [(64, 1), (128, 2), (256, 1), (512, 4), (1024, 1), (1024, 2), (1024, 8), (500, 1), (1000, 1),
                                      (2048, 8), (250, 1)])
def test_x_pass_tensor_map_addresses_the_blocked_layout(built, N, nranks):
    """The tensor map the TMA-fed x pass encodes (vp_fft_x_layout: host arithmetic, the same function the launcher uses) must
    address exactly the blocked half spectrum the y pass stores (csrc/fft3d.cu YDest): for every (x, ky, kz) the box
    coordinates the kernel issues -- (0, ky % kyb, zt, q*box_x, ky / kyb) -- plus the position inside the box reach complex
    number ((((ky/kyb)*N + x)*tiles + zt)*kyb + ky%kyb)*C + c; the map obeys the TMA rules (16-byte strides and box rows,
    box dims <= 256) and the ring fits the 227 KB of a CTA whenever the TMA kernel is selected."""
    from vpower import _lib
    kzc = N // 2 // nranks
    g = _lib.fft_x_layout(N, kzc)
    C, kyb, tiles, bx = g["C"], g["kyb"], g["tiles"], g["box_x"]
    assert tiles * C == kzc and N % kyb == 0 and N % bx == 0 and g["boxes"] == N // bx
    dims, strides, box = g["dims"], [4] + g["strides"], g["box"]
    assert dims == [2 * C, kyb, tiles, N, N // kyb] and box == [2 * C, 1, 1, bx, 1]
    assert np.prod(dims) * 4 == N * N * kzc * 8                      # the map covers the rank's cube exactly
    expect_tma = N not in (250, 2048)
    assert g["tma"] == expect_tma
    if g["tma"]:
        assert all(s % 16 == 0 for s in strides[1:]) and (box[0] * 4) % 16 == 0 and max(box) <= 256
        assert g["smem_bytes"] <= 227 * 1024 and g["box_slot_bytes"] % 128 == 0 and g["box_slot_bytes"] >= bx * C * 8
        assert g["ring_items"] in (1, 2) and g["threads"] <= 1024
    rng = np.random.default_rng(N + nranks)
    n = 4000
    x, ky, zt, c = rng.integers(0, N, n), rng.integers(0, N, n), rng.integers(0, tiles, n), rng.integers(0, C, n)
    blocked = ((((ky // kyb) * N + x) * tiles + zt) * kyb + ky % kyb) * C + c          # complex index, YDest / k_fft_x_pow
    q, xin = x // bx, x % bx                                                           # box of the item, row inside the box
    coord = np.stack([2 * c, ky % kyb, zt, q * bx + xin, ky // kyb], axis=1).astype(np.int64)
    byte = (coord * np.asarray(strides, dtype=np.int64)).sum(axis=1)
    assert np.array_equal(byte, blocked * 8)
    # shared-memory side: box q lands at q*slot, dense [box_x][C] complex -- what the kernel's register loads index
    smem = q * g["box_slot_bytes"] + (xin * C + c) * 8
    assert np.all(smem + 8 <= g["boxes"] * g["box_slot_bytes"])
    # corner cases of the index range
    for (xx, kk, zz, cc) in ((0, 0, 0, 0), (N - 1, N - 1, tiles - 1, C - 1)):
        b = ((((kk // kyb) * N + xx) * tiles + zz) * kyb + kk % kyb) * C + cc
        co = np.array([2 * cc, kk % kyb, zz, xx, kk // kyb], dtype=np.int64)
        assert int((co * np.asarray(strides, dtype=np.int64)).sum()) == b * 8 and b < N * N * kzc
