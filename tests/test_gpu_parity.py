"""Parity of the CUDA path (through the C ABI / the vpower mirror) against the CPU oracle and the golden
vectors of the unmodified reference.  All tests need a B200: `pytest -m gpu`.

Bars (BASELINE.json north_star): nearest-particle indices and per-bin mode counts bit-exact; binned P(k)
within 1e-5 relative where the oracle is f64 (library flavour), 1e-3 where the reference itself is f32
(script flavour).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vp():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    import vpower
    from vpower import _lib
    _lib.load_library()
    return vpower


@pytest.fixture(scope="module")
def lib(vp):
    from vpower import _lib
    return _lib


def _pos_seen_by_ann(g, name):
    return g[f"{name}/pos_parsed"] if bool(g[f"{name}/pos_parsed_differs"]) else g[f"{name}/pos"]


# ------------------------------------------------------------------------------------------ radix sort
@pytest.mark.parametrize("n,bits", [(1, 8), (1000, 8), (4096, 16), (100003, 24), (1 << 20, 30), ((1 << 21) + 17, 32),
                                    (5000, 7), (70001, 14), (300007, 21), (1 << 19, 28)])
def test_sort_pairs(lib, n, bits):
    import torch
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << bits, size=n, dtype=np.uint64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    kt = torch.from_numpy(keys.view(np.int32)).cuda()
    vt = torch.from_numpy(vals.view(np.int32)).cuda()
    lib.sort_pairs(kt, vt, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(kt.cpu().numpy().view(np.uint32), keys[order])
    assert np.array_equal(vt.cpu().numpy().view(np.uint32), vals[order])          # stable


def test_sort_pairs_unaligned_views(lib):
    """Key/value arrays that start 4 bytes into an allocation (tensor views): the 16-byte load path must not be taken."""
    import torch
    n, bits = 50001, 20
    rng = np.random.default_rng(7)
    keys = rng.integers(0, 1 << bits, size=n, dtype=np.uint64).astype(np.uint32)
    kt = torch.zeros(n + 1, dtype=torch.int32, device="cuda")[1:]
    vt = torch.zeros(n + 1, dtype=torch.int32, device="cuda")[1:]
    kt.copy_(torch.from_numpy(keys.view(np.int32)))
    vt.copy_(torch.arange(n, dtype=torch.int32))
    assert kt.data_ptr() % 16 == 4
    lib.sort_pairs(kt, vt, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(kt.cpu().numpy().view(np.uint32), keys[order])
    assert np.array_equal(vt.cpu().numpy().view(np.uint32), order.astype(np.uint32))


# ------------------------------------------------------------------------------------------ K1 nearest particle
@pytest.mark.parametrize("name,N", [("lib16", 16), ("lib24", 24), ("lib32c", 32)])
def test_nn_vs_ann_golden(lib, golden, name, N):
    """Same inputs the reference's ANN 1.1.2 engine saw; bit-exact on every non-tied query."""
    pos = _pos_seen_by_ann(golden, name)
    ax = [golden[f"{name}/axis_parsed{c}"] for c in range(3)]
    nn = lib.nn_grid(lib.to_device(pos), *ax).cpu().numpy()
    ties = np.unpackbits(golden[f"{name}/nn_ties"])[: N ** 3].astype(bool).reshape(N, N, N)
    assert np.array_equal(nn[~ties], golden[f"{name}/nn_ann"][~ties])
    st = lib.nn_grid_stats()
    assert st["n_unresolved"] == 0 and st["n_kept"] == len(pos)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("Np,N,clustered", [(1 << 15, 32, False), (1 << 18, 64, False), (40 ** 3, 40, True), (5000, 48, False),
                                            (300000, 24, False)])
def test_nn_vs_oracle(lib, orc, dtype, Np, N, clustered):
    pos, _, _, _ = orc.synth_particles(3, Np, 1.0, clustered=clustered, lattice_n=40 if clustered else None)
    pos = pos.astype(dtype)
    if dtype == np.float64:
        pos = pos + 1e-9 * np.random.default_rng(0).random(pos.shape)             # genuinely 64-bit coordinates
    ax = orc.lattice_axis_lib(1.0, N)
    ref, ties = orc.nn_exact_lattice(pos.astype(np.float64), ax, ax, ax, return_ties=True)
    nn = lib.nn_grid(lib.to_device(pos), ax, ax, ax).cpu().numpy()
    assert np.array_equal(nn, ref)                       # ties included: both sides pick the lowest index
    st = lib.nn_grid_stats()
    assert st["n_unresolved"] == 0 and st["n_far"] == 0  # every particle lies inside the cell grid of a lattice covering the box


def test_nn_script_lattice_and_ties(lib, orc):
    """f32 node coordinates i*LCELL (parallel_optimized.py:343-346) and constructed exact ties."""
    N = 32
    ax = orc.lattice_axis_script(N, 1)
    rng = np.random.default_rng(5)
    pos = rng.integers(0, 64, size=(4000, 3)).astype(np.float32) / 64.0           # coarse lattice -> many exact ties
    ref, ties = orc.nn_exact_lattice(pos.astype(np.float64), ax, ax, ax, return_ties=True)
    assert ties.sum() > 100
    nn = lib.nn_grid(lib.to_device(pos), ax, ax, ax).cpu().numpy()
    assert np.array_equal(nn, ref)


def test_nn_edge_cases(lib, orc):
    ax = orc.lattice_axis_lib(1.0, 8)
    # one particle, far outside the lattice
    pos = np.array([[5.0, -3.0, 0.5]], dtype=np.float64)
    assert (lib.nn_grid(lib.to_device(pos), ax, ax, ax).cpu().numpy() == 0).all()
    # all particles in one corner: every node needs the wide search
    pos = (np.random.default_rng(1).random((500, 3)) * 0.05).astype(np.float32)
    ref = orc.nn_exact_lattice(pos.astype(np.float64), ax, ax, ax)
    assert np.array_equal(lib.nn_grid(lib.to_device(pos), ax, ax, ax).cpu().numpy(), ref)
    assert lib.nn_grid_stats()["n_wide"] > 0
    # rectangular lattice with different tables per axis
    qx, qy, qz = np.linspace(0.1, 0.9, 5), np.linspace(0.0, 1.0, 17), np.linspace(0.3, 0.4, 9)
    pos = np.random.default_rng(2).random((3000, 3))
    ref = orc.nn_exact_lattice(pos, qx, qy, qz)
    assert np.array_equal(lib.nn_grid(lib.to_device(pos), qx, qy, qz).cpu().numpy(), ref)


def test_nn_dense_cluster_long_rows(lib, orc):
    """A blob holding far more particles than one row of cells can order in shared memory (> 65536 per row):
    the cell-list build takes its global-scratch path; the answers stay exact."""
    rng = np.random.default_rng(21)
    N = 24
    blob = 0.5 + 0.004 * rng.standard_normal((260000, 3))
    pos = np.concatenate([blob, rng.random((N ** 3 - 1000, 3))]).astype(np.float32)
    rng.shuffle(pos)
    ax = orc.lattice_axis_lib(1.0, N)
    ref = orc.nn_exact_lattice(pos.astype(np.float64), ax, ax, ax)
    nn = lib.nn_grid(lib.to_device(pos), ax, ax, ax).cpu().numpy()
    assert np.array_equal(nn, ref)
    assert lib.nn_grid_stats()["n_unresolved"] == 0


def test_nn_slab_matches_full(lib, orc):
    """Multi-GPU building block: a slab of the lattice with only nearby particles kept gives the same answer."""
    N, Np = 64, 1 << 17
    pos, _, _, _ = orc.synth_particles(9, Np, 1.0)
    ax = orc.lattice_axis_lib(1.0, N)
    full = lib.nn_grid(lib.to_device(pos), ax, ax, ax).cpu().numpy()
    h = ax[1] - ax[0]
    for r in range(4):
        sl = slice(r * N // 4, (r + 1) * N // 4)
        o = lib.NNOpts()
        o.use_x_keep = 1
        o.x_keep_lo = ax[sl][0] - 4 * h
        o.x_keep_hi = ax[sl][-1] + 4 * h
        o.x_lo_is_domain_edge = 0
        o.x_hi_is_domain_edge = 0
        part = lib.nn_grid(lib.to_device(pos), ax[sl], ax, ax, o).cpu().numpy()
        st = lib.nn_grid_stats()
        assert st["n_unresolved"] == 0 and st["n_kept"] < Np
        assert np.array_equal(part, full[sl])


def test_ann_interpolate_api(vp, orc, golden):
    """vpower.interp.ann_interpolate keeps the reference's signature and gather semantics (interp.py:1018-1049)."""
    name, N, L = "lib16", 16, 1.0
    pos = golden[f"{name}/pos"]
    f = orc.density_velocity_vector(golden[f"{name}/vel"].astype(np.float64), golden[f"{name}/dens"].astype(np.float64))
    got = vp.interp.ann_interpolate(pos, vp.interp.make_grid_coords(L, N), f, N, 0.0)
    ref = f[golden[f"{name}/nn_ann"].astype(np.int64).ravel()].reshape(N, N, N, 4)
    assert got.dtype == f.dtype and np.array_equal(got, ref)
    got1 = vp.interp.ann_interpolate(pos, vp.interp.make_grid_coords(L, N), f[:, 3].copy(), N, 0.0)
    assert np.array_equal(got1, ref[..., 3])
    assert np.array_equal(vp.interp.make_grid_coords(L, N), orc.make_grid_coords(L, N))


# ------------------------------------------------------------------------------------------ K2 deposit
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_deposit_golden(vp, golden, tag):
    pos, w1, w4 = golden[f"deposit_{tag}/pos"], golden[f"deposit_{tag}/w1"], golden[f"deposit_{tag}/w4"]
    g1 = vp.interp.deposit_to_grid(w1, pos, 12, 1.5)
    assert g1.dtype == np.float64 and np.array_equal(g1, golden[f"deposit_{tag}/grid1"])       # integer weights: exact
    g4 = vp.interp.deposit_to_grid(w4, pos, 12, 1.5)
    assert np.allclose(g4, golden[f"deposit_{tag}/grid4"], rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------------------------------ K4 FFT
@pytest.mark.parametrize("N", [64, 128, 256])
def test_fft_half_spectrum(lib, N):
    import torch
    rng = np.random.default_rng(N)
    f = rng.normal(size=(N, N, N)).astype(np.float32)
    ks = 2 * np.pi * np.fft.fftfreq(N, 1.0 / N)
    plan = lib.PkPlan(N, ks, np.array([0.0, 1.0]))
    got = plan.fft_half(torch.from_numpy(f.copy()).cuda()).cpu().numpy()
    ref = np.fft.rfftn(f.astype(np.float64))
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err


# ------------------------------------------------------------------------------------------ K4+K5 fused
@pytest.mark.parametrize("N,L", [(64, 1.0), (128, 2.5), (256, 1.0), (48, 0.7), (16, 1.0), (250, 0.7), (500, 1.0)])
def test_pk_fields_vs_oracle(lib, orc, N, L):
    import torch
    rng = np.random.default_rng(100 + N)
    f = [rng.normal(size=(N, N, N)).astype(np.float32) for _ in range(3)]
    # give the spectrum some dynamic range
    x = (np.arange(N) + 0.5) / N
    f[0] += 3 * np.sin(2 * np.pi * 2 * x)[:, None, None].astype(np.float32)
    f[1] += 2 * np.cos(2 * np.pi * 3 * x)[None, :, None].astype(np.float32)
    kmin, kmax = 2 * np.pi / L, np.pi / (L / N)
    for maker in (orc.edges_lib, orc.edges_script):
        centres, edges = maker(kmin, kmax, kmin)
        plan = lib.PkPlan(N, orc.k_axis(L, N), edges)
        k = orc.k_magnitude(L, N)
        for comps in ([0, 1, 2], [1]):
            P = sum(np.abs(np.fft.fftn(f[c].astype(np.float64))) ** 2 for c in comps)
            ref = orc.hist_sample(k, P, centres, edges, empty_to_zero=True)
            psum, ns = plan.fields([torch.from_numpy(f[c].copy()).cuda() for c in comps])
            assert np.array_equal(ns, ref[:, 3].astype(np.int64))                 # mode counts bit-exact
            ok = ref[:, 3] > 0
            rel = np.abs(psum[ok] - ref[ok, 2]) / ref[ok, 2]
            # north_star: 1e-3 relative per bin for an f32 transform.  The power-of-two and N = 250 lines stay below 1e-5; at
            # N = 500 the first shell (26 modes) sits next to the 1e8-amplitude test tone and collects its coherent f32
            # round-off through the radix-10/5 butterflies (1.6e-5 observed), every other shell is at 2e-7.
            assert rel.max() < (5e-5 if N == 500 else 1e-5), (N, comps, rel.max())


def test_shell_counts_match_reference(lib, orc, golden):
    import torch
    for N, L in ((16, 1.0), (32, 2.5), (64, 1.0), (48, 0.7)):
        kmin, kmax = 2 * np.pi / L, np.pi / (L / N)
        c, e = orc.edges_lib(kmin, kmax, kmin)
        plan = lib.PkPlan(N, orc.k_axis(L, N), e)
        _, ns = plan.fields([torch.ones((N, N, N), dtype=torch.float32, device="cuda")])
        assert np.array_equal(ns, golden[f"shells_lib_{N}_{L}/Nsample"].astype(np.int64))
    # full-size property: every mode lands in exactly one shell or outside [first,last] edge
    for N in (256, 512):
        kmin, kmax = 2 * np.pi, np.pi * N
        c, e = orc.edges_lib(kmin, kmax, kmin)
        plan = lib.PkPlan(N, orc.k_axis(1.0, N), e)
        _, ns = plan.fields([torch.ones((N, N, N), dtype=torch.float32, device="cuda")])
        if N == 256:
            assert ns.sum() == 8886577
        j = np.arange(N // 2) + 1
        # integer-shell closed form: modes with floor(|n|+1/2) == j
        n1 = np.fft.fftfreq(N, 1.0 / N)
        n2 = (n1[:, None, None] ** 2 + n1[None, :, None] ** 2 + n1[None, None, :] ** 2).ravel()
        shell = np.floor(np.sqrt(n2) + 0.5).astype(np.int64)
        ref = np.bincount(shell, minlength=N)[1:N // 2 + 1]
        assert np.array_equal(ns, ref)


def test_power_cube_and_pairs_api(vp, orc, golden):
    """_vector_power / _scalar_power / _pair_power / _hist_sample keep the reference's return conventions."""
    v, m = golden["lib16/v_grid"], golden["lib16/m_grid"]
    P = vp.interp._vector_power(v[..., 0], v[..., 1], v[..., 2], 1.0, 16)
    assert P.shape == (16, 16, 16)
    assert np.allclose(P, golden["lib16/Pgrid_velocity"], rtol=2e-4, atol=1e-6 * golden["lib16/Pgrid_velocity"].max())
    E = m * (v[..., 0] ** 2 + v[..., 1] ** 2 + v[..., 2] ** 2)
    Pe = vp.interp._scalar_power(E, 1.0, 16)
    assert np.allclose(Pe, golden["lib16/Pgrid_energy"], rtol=2e-4, atol=1e-6 * golden["lib16/Pgrid_energy"].max())
    pairs = vp.interp._pair_power(golden["lib16/Pgrid_velocity"], 1.0, 16)
    assert pairs.shape == (4096, 2) and np.array_equal(pairs[:, 0], golden["lib16/pairs_k"])   # |k| bit-exact
    h = vp.interp._hist_sample(pairs, 2 * np.pi, np.pi * 16, 2 * np.pi)
    ref = golden["lib16/spctrm_velocity"]
    assert np.array_equal(h[:, 0], ref[:, 0]) and np.array_equal(h[:, 3], ref[:, 3])
    assert np.allclose(h[:, 2], ref[:, 2], rtol=1e-12)


# ------------------------------------------------------------------------------------------ whole path, library flavour
@pytest.mark.parametrize("name,N,L", [("lib16", 16, 1.0), ("lib24", 24, 2.5), ("lib32c", 32, 1.0)])
def test_library_path_vs_reference_golden(vp, golden, name, N, L):
    """GasParticles(...).ann_interp_to_field(N).spctrm(q) against the unmodified reference (true ANN engine)."""
    pos = _pos_seen_by_ann(golden, name)
    gp = vp.interp.GasParticles(pos.copy(), golden[f"{name}/mass"].astype(np.float64),
                                golden[f"{name}/dens"].astype(np.float64), golden[f"{name}/vel"].astype(np.float64), L)
    bf = gp.ann_interp_to_field(N)
    if name == "lib16":     # the lattice the reference used is the un-parsed one only when positions round trip
        assert np.array_equal(np.stack([bf.vx, bf.vy, bf.vz], -1), golden[f"{name}/v_grid"])
        assert np.array_equal(bf.mass, golden[f"{name}/m_grid"])
    for q in ("velocity", "momentum", "energy"):
        sp = bf.spctrm(q)
        ref = golden[f"{name}/spctrm_{q}"]
        assert np.array_equal(sp.k, ref[:, 0])
        assert np.array_equal(sp.Nsample, ref[:, 3])
        if name == "lib16":
            assert np.allclose(sp.Psum, ref[:, 2], rtol=1e-5, atol=0)
            assert np.allclose(sp.P, ref[:, 1], rtol=1e-5, atol=0)


@pytest.mark.parametrize("N,Np", [(64, 1 << 18), (128, 1 << 19)])
def test_library_path_vs_oracle(vp, orc, N, Np):
    """cfg1-sized run of the object API against the oracle (64^3 / 2^18 = BASELINE configs[0])."""
    L = 1.0
    pos, vel, dens, mass = orc.synth_particles(0, Np, L)
    p64, v64, d64 = pos.astype(np.float64), vel.astype(np.float64), dens.astype(np.float64)
    gp = vp.interp.GasParticles(p64, mass.astype(np.float64), d64, v64, L)
    bf = gp.ann_interp_to_field(N)
    v_ref, m_ref, Lcell = orc.ann_interp_to_field(p64, d64, v64, L, N)
    assert np.array_equal(bf.get_v(), v_ref) and np.array_equal(bf.mass, m_ref)
    for q in ("velocity", "momentum", "energy"):
        sp = bf.spctrm(q)
        ref = orc.spctrm(v_ref, m_ref, Lcell, q)
        assert np.array_equal(sp.Nsample, ref[:, 3])
        assert np.allclose(sp.P, ref[:, 1], rtol=1e-5, atol=0), np.max(np.abs(sp.P / ref[:, 1] - 1))
    bf.strict_reference = False
    sp = bf.spctrm("momentum")
    ref = orc.spctrm(v_ref, m_ref, Lcell, "momentum", strict_reference=False)
    assert np.allclose(sp.P, ref[:, 1], rtol=1e-5, atol=0)
    # Parseval (interp.py:1377-1378): sum(P) (2pi/L)^3 == 1/2 mean(v^2)
    Pg = bf.velocity_power()
    assert np.isclose(Pg.sum() * (2 * np.pi / L) ** 3, 0.5 * np.mean((v_ref ** 2).sum(-1)), rtol=1e-5)


def test_boxfield_from_arrays(vp, orc, golden):
    """BoxField(v, mass, Lcell) built from host arrays, as a reference user would (interp.py:456-471)."""
    v, m = golden["lib32c/v_grid"], golden["lib32c/m_grid"]
    bf = vp.interp.BoxField(v, m, 1.0 / 32)
    for q in ("velocity", "momentum", "energy"):
        sp = bf.spctrm(q)
        ref = golden[f"lib32c/spctrm_{q}"]
        assert np.array_equal(sp.Nsample, ref[:, 3])
        assert np.allclose(sp.P, ref[:, 1], rtol=2e-5, atol=0)
    with pytest.raises(Exception):
        bf.spctrm("vorticity")
    # custom k range / resolution (interp.py:560-570)
    sp = bf.spctrm("velocity", kmin=4 * np.pi, kmax=40 * np.pi, kres=np.pi)
    ref = orc.spctrm(v, m, 1.0 / 32, "velocity", kmin=4 * np.pi, kmax=40 * np.pi, kres=np.pi)
    assert np.array_equal(sp.Nsample, ref[:, 3]) and np.allclose(sp.Psum, ref[:, 2], rtol=2e-5)


# ------------------------------------------------------------------------------------------ whole path, one C call
@pytest.mark.parametrize("host", [True, False])
def test_one_call_path(lib, orc, host):
    N, Np, L = 64, 1 << 18, 1.0
    pos, vel, dens, _ = orc.synth_particles(0, Np, L)
    ax = orc.lattice_axis_lib(L, N)
    Lcell = L / N
    centres, edges = orc.edges_lib(2 * np.pi / L, np.pi / Lcell, 2 * np.pi / L)
    a = (L / (2 * np.pi)) ** 1.5 / N ** 3
    args = (pos, vel, dens) if host else tuple(lib.to_device(x) for x in (pos, vel, dens))
    out, ns = lib.particles_to_pk(*args, ax, ax, ax, N, Lcell ** 3, 0.5 * a * a, orc.k_axis(L, N), edges,
                                  quantities=("velocity", "momentum", "energy"))
    v_ref, m_ref, _ = orc.ann_interp_to_field(pos.astype(np.float64), dens.astype(np.float64), vel.astype(np.float64), L, N)
    for q in ("velocity", "momentum", "energy"):
        ref = orc.spctrm(v_ref, m_ref, Lcell, q)
        assert np.array_equal(ns, ref[:, 3].astype(np.int64))
        # f32 particle arithmetic (rho*v)/rho differs from the f64 oracle by ulps of f32
        assert np.allclose(out[q], ref[:, 2], rtol=1e-4, atol=0), (q, np.max(np.abs(out[q] / ref[:, 2] - 1)))


@pytest.mark.parametrize("name", ["script16", "script16_fold2"])
def test_script_flavour_vs_reference_golden(lib, orc, golden, name):
    """The unfolded full transform equals the script's folded pipeline (verbatim single-rank run, Pk.txt)."""
    pos, vel, mass = golden[f"{name}/pos"], golden[f"{name}/vel"].copy(), golden[f"{name}/mass"]
    pos = pos - pos.min(axis=0)
    M = np.sum(mass)
    for c in range(3):
        vel[:, c] -= np.sum(mass * vel[:, c]) / M
    NTOT, LTOT = 16, 1
    ax = orc.lattice_axis_script(NTOT, LTOT)
    LCELL = LTOT / NTOT
    centres, edges = orc.edges_script(2 * np.pi / LTOT, np.pi / LCELL, 2 * np.pi / LTOT)
    const = (LTOT / (2 * np.pi)) ** 1.5 / NTOT ** 3
    out, ns = lib.particles_to_pk(pos.astype(np.float32), vel.astype(np.float32), None, ax, ax, ax, NTOT, LCELL ** 3,
                                  0.5 * const * const, orc.k_axis(LTOT, NTOT), edges, quantities=("velocity",))
    ref = golden[f"{name}/Pk"]
    assert np.array_equal(ns, ref[:, 3].astype(np.int64))
    assert np.allclose(out["velocity"], ref[:, 2], rtol=1e-3, atol=0)


def test_errors_are_loud(lib, vp):
    import torch
    with pytest.raises(lib.VPowerError):
        lib.PkPlan(64, np.arange(64.0), np.array([1.0, 0.5]))          # edges must increase
    with pytest.raises(Exception):
        vp.interp.ann_interpolate(np.zeros((4, 3)), np.zeros((7, 3)), np.zeros(4), 2, 0.0)
    with pytest.raises(Exception):
        vp.interp.GasParticles(np.zeros((4, 3)), np.ones(4), np.ones(4), np.zeros((4, 3)), 1.0).ann_interp_to_field(8, eps=0.5)


# ------------------------------------------------------------------------------------------ slab decomposition (multi-GPU path)
@pytest.mark.parametrize("P", [1, 2, 4])
def test_slab_path_emulated_on_one_gpu(lib, orc, P):
    """All kernels of the multi-GPU path with nranks = P, the ranks run one after the other on this GPU and the
    all-to-all done by slicing: must reproduce the single-GPU result (Nsample bit-identical)."""
    import torch
    from vpower import dist as vd
    N, Np, L = 256, 1 << 21, 1.0
    pos, vel, dens, _ = orc.synth_particles(6, Np, L)
    ax, k = orc.lattice_axis_lib(L, N), orc.k_axis(L, N)
    centres, edges = orc.edges_lib(2 * np.pi / L, np.pi * N / L, 2 * np.pi / L)
    a = (L / (2 * np.pi)) ** 1.5 / N ** 3
    dpos, dvel, drho = (lib.to_device(x) for x in (pos, vel, dens))
    qs = ("velocity", "momentum", "energy")
    ref, ref_ns = lib.particles_to_pk(dpos, dvel, drho, ax, ax, ax, N, (L / N) ** 3, 0.5 * a * a, k, edges, quantities=qs)
    backends = [vd.CudaBackend(N, k, edges, P, r) for r in range(P)]
    gridded = []
    for r in range(P):
        x0, x1, _, _ = vd.slab_bounds(N, P, r)
        g, unresolved = backends[r].grid_slab(dpos, dvel, drho, ax[x0:x1], ax, (L / N) ** 3, vd.keep_range(ax, x0, x1, P, r, 4))
        assert unresolved == 0
        gridded.append(g)
    for q in qs:
        sends, mult = [], 1.0
        for r in range(P):
            slabs, mult = backends[r].fields(gridded[r], q, True)
            sends.append(backends[r].fft_local(slabs))                # per comp: [P, nx, N, kzc]
        psum = torch.zeros(len(edges) - 1, dtype=torch.float64, device="cuda")
        ns = torch.zeros(len(edges) - 1, dtype=torch.int64, device="cuda")
        for d in range(P):
            recv = [torch.cat([sends[r][c][d] for r in range(P)], dim=0).contiguous() for c in range(len(sends[0]))]
            ps, n_ = backends[d].fft_final(recv)
            psum += ps
            ns += n_
        assert np.array_equal(ns.cpu().numpy(), ref_ns)
        got = psum.cpu().numpy() * (mult * 0.5 * a * a)
        assert np.allclose(got, ref[q], rtol=1e-6, atol=0), (q, np.max(np.abs(got / ref[q] - 1)))


# ------------------------------------------------------------------------------------------ script drop-in
def test_script_dropin_writes_reference_pk_txt(lib, golden, tmp_path):
    """scripts/parallel_optimized.py drop-in: same flags, same Pk.txt as the verbatim reference run (single rank)."""
    import importlib.util
    import os
    name = "script16"
    snap = tmp_path / "snap.npz"
    np.savez(snap, **{"PartType0/Coordinates": golden[f"{name}/pos"], "PartType0/Masses": golden[f"{name}/mass"],
                      "PartType0/Velocities": golden[f"{name}/vel"]})
    path = os.path.join(os.path.dirname(lib.LIB_PATH), "scripts", "parallel_optimized.py")
    spec = importlib.util.spec_from_file_location("po_b200", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(["-i", str(snap), "-o", str(tmp_path), "-N", "16", "-M", "8", "-b", "512", "-f"]) == 0
    got = np.loadtxt(tmp_path / "Pk.txt")
    ref = golden[f"{name}/Pk"]
    assert got.shape == ref.shape
    assert np.allclose(got[:, 0], ref[:, 0], rtol=1e-6)
    assert np.array_equal(got[:, 3], ref[:, 3])
    assert np.allclose(got[:, 2], ref[:, 2], rtol=1e-3) and np.allclose(got[:, 1], ref[:, 1], rtol=1e-3)
    # helper functions keep the script's conventions
    assert mod.planner(1000, 1, 500, 8) == (1, 2, 500, 0.5)
    P = np.random.default_rng(0).random((16, 16, 16))
    pairs = mod.pair_power(P, 1.0, 16)
    h = mod.hist_sample(pairs, 2 * np.pi, np.pi * 16, 2 * np.pi)
    assert pairs.shape == (4096, 2) and h.shape == (8, 4) and np.array_equal(h[:, 3], ref[:, 3])
    # FFTW_vector_power / FFTW_power (:92-141): 1/2 |const FFT|^2 summed over components, against numpy's c2c transform
    f3 = np.random.default_rng(1).standard_normal((3, 16, 16, 16)).astype(np.float32)
    const = (2.0 / (2 * np.pi)) ** 1.5 / 16 ** 3
    want = sum(0.5 * np.abs(np.fft.fftn(f.astype(np.float64)) * const) ** 2 for f in f3)
    got3 = mod.FFTW_vector_power(f3[0], f3[1], f3[2], 2.0, 16)
    assert got3.dtype == np.float32 and np.allclose(got3, want, rtol=2e-4, atol=1e-6 * want.max())
    want1 = 0.5 * np.abs(np.fft.fftn(f3[0].astype(np.float64)) * const) ** 2
    assert np.allclose(mod.FFTW_power(f3[0].astype(np.complex64), 2.0, 16), want1, rtol=2e-4, atol=1e-6 * want1.max())


def test_host_chunk_streaming_equals_device_path(lib, orc):
    """vp_host_particles_to_pk streams the host arrays in 2^24-particle chunks: positions first (gridding overlaps the
    rest of the upload), then velocity/density packed into input-order records on a side stream, planes gathered through
    the original particle index.  Three chunks here.  Must reproduce the device-resident path (same planes)."""
    import torch
    N, Np, L = 256, (1 << 25) + 777, 1.0
    g = torch.Generator(device="cuda").manual_seed(5)
    pos = torch.rand((Np, 3), generator=g, device="cuda")
    vel = torch.randn((Np, 3), generator=g, device="cuda")
    rho = 1.0 + torch.rand(Np, generator=g, device="cuda")
    ax, k = orc.lattice_axis_lib(L, N), orc.k_axis(L, N)
    centres, edges = orc.edges_lib(2 * np.pi / L, np.pi * N / L, 2 * np.pi / L)
    a = (L / (2 * np.pi)) ** 1.5 / N ** 3
    qs = ("velocity", "energy")
    ref, ref_ns = lib.particles_to_pk(pos, vel, rho, ax, ax, ax, N, (L / N) ** 3, 0.5 * a * a, k, edges, quantities=qs)
    out, ns = lib.particles_to_pk(pos.cpu().numpy(), vel.cpu().numpy(), rho.cpu().numpy(), ax, ax, ax, N, (L / N) ** 3,
                                  0.5 * a * a, k, edges, quantities=qs)
    assert np.array_equal(ns, ref_ns)
    for q in qs:
        # same planes bit for bit; the shell sums may differ in the order of the per-CTA f64 flushes
        assert np.allclose(out[q], ref[q], rtol=1e-13, atol=0)


def test_slab_bucket_kernel(lib):
    """vp_slab_bucket: every particle lands (once) in the block of each rank whose kept range contains it."""
    import torch
    n = 200003
    g = torch.Generator(device="cuda").manual_seed(9)
    pos = torch.rand((n, 3), generator=g, device="cuda")
    vel = torch.randn((n, 3), generator=g, device="cuda")
    rho = torch.rand(n, generator=g, device="cuda")
    lo = [-np.inf, 0.2, 0.45, 0.7]
    hi = [0.3, 0.55, 0.8, np.inf]
    rows, counts = lib.slab_bucket(pos, vel, rho, lo, hi)
    x = pos[:, 0].cpu().numpy()
    ref = torch.cat([pos, vel, rho[:, None]], 1).cpu().numpy()
    got = rows.cpu().numpy()
    at = 0
    for d in range(4):
        m = (x >= lo[d]) & (x <= hi[d])
        assert counts[d] == int(m.sum())
        blk = got[at:at + counts[d]]
        at += counts[d]
        assert blk.shape[1] == 8 and not blk[:, 7].any()                 # padded rows: whole 32-byte sectors
        blk = blk[:, :7]
        a = blk[np.lexsort(blk.T[::-1])]
        b = ref[m][np.lexsort(ref[m].T[::-1])]
        assert np.array_equal(a, b)
    rows6, counts6 = lib.slab_bucket(pos, vel, None, lo, hi)
    assert counts6 == counts and rows6.shape[1] == 8 and not rows6[:sum(counts6), 6:].any().item()


def test_fft_2048_single_mode_and_low_shells(lib, orc):
    """N = 2048 (cfg5's lattice): the 1024-thread y/x kernels and the non-prefetch x pass.  No CPU oracle can transform
    2048^3, so: a single plane wave must put N^6/2 into its shell and (to rounding) nothing elsewhere, and the mode counts
    of the first 40 shells must equal a brute-force count of integer lattice points."""
    import torch
    N, L = 2048, 1.0
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~40 GB of device memory")
    f = torch.empty((N, N, N), dtype=torch.float32, device="cuda")
    ar = torch.arange(N, device="cuda", dtype=torch.float64)
    for x0 in range(0, N, 32):
        ph = (3 * ar[x0:x0 + 32, None, None] + 5 * ar[None, :, None] + 7 * ar[None, None, :]) * (2 * np.pi / N)
        f[x0:x0 + 32] = torch.cos(ph).to(torch.float32)
        del ph
    centres, edges = orc.edges_lib(2 * np.pi / L, np.pi * N / L, 2 * np.pi / L)
    plan = lib.PkPlan(N, orc.k_axis(L, N), edges)
    psum, ns = plan.fields([f])
    del f
    j0 = int(np.floor(np.sqrt(9 + 25 + 49) + 0.5)) - 1            # shell of |k| = sqrt(83) kf
    expect = float(N) ** 6 / 2
    assert abs(psum[j0] / expect - 1) < 1e-5
    assert (np.delete(psum, j0).sum()) < 1e-6 * expect
    m = 41
    n1 = np.arange(-m, m + 1)
    r = np.sqrt(n1[:, None, None] ** 2 + n1[None, :, None] ** 2 + n1[None, None, :] ** 2)
    shell = np.floor(r + 0.5).astype(int)
    ref = np.bincount(shell.ravel(), minlength=m + 1)[1:m]
    assert np.array_equal(ns[:m - 1], ref)
    assert ns.sum() > 0.5 * (4 / 3) * np.pi * (N / 2) ** 3


def test_fft_binning_parseval_and_linearity_full_size(lib, orc):
    """cfg4 lattice (1024^3), where no CPU oracle finishes: size-independent properties of K4+K5.  Parseval for the
    unnormalised DFT (sum over ALL modes of |F|^2 = N^3 sum f^2, every mode counted once), shells partition the modes
    (mode counts of the library edges + the modes outside them = N^3), and exact linearity under scaling by 2."""
    import torch
    N, L = 1024, 1.0
    free, _ = torch.cuda.mem_get_info()
    if free < 30e9:
        pytest.skip("needs ~15 GB of device memory")
    g = torch.Generator(device="cuda").manual_seed(3)
    f = torch.randn((N, N, N), generator=g, device="cuda", dtype=torch.float32)
    f[:, :, ::2] += 0.5                                             # some power at the z Nyquist plane and at DC
    want = float((f.double() ** 2).sum().item()) * N ** 3
    k = orc.k_axis(L, N)
    one = lib.PkPlan(N, k, np.array([0.0, 1e30]))                    # one shell holding every mode, DC included
    psum, ns = one.fields([f.clone()])
    assert int(ns[0]) == N ** 3
    assert abs(psum[0] - want) / want < 1e-5
    centres, edges = orc.edges_lib(2 * np.pi / L, np.pi * N / L, 2 * np.pi / L)
    plan = lib.PkPlan(N, k, edges)
    p1, n1 = plan.fields([f.clone()])
    p2, n2 = plan.fields([2.0 * f])
    # scaling by a power of two is exact in every f32/f64 operation; only the order of the per-CTA f64 flushes differs
    assert np.array_equal(n1, n2) and np.allclose(p2, 4.0 * p1, rtol=1e-13, atol=0)
    inside = plan_inside = int(n1.sum())
    lo, hi = lib.PkPlan(N, k, np.array([0.0, edges[0]])), lib.PkPlan(N, k, np.array([edges[-1], 1e30]))
    below, above = lo.fields([f.clone()])[1], hi.fields([f.clone()])[1]
    # the library's last bin is closed on the right and the next plan's first bin is closed on the left: modes exactly
    # on edges[-1] would be counted twice; there are none for these edges (|k| = edges[-1] needs an irrational ratio)
    assert inside + int(below[0]) + int(above[0]) == N ** 3
    assert abs((p1.sum() + lo.fields([f.clone()])[0][0] + hi.fields([f.clone()])[0][0]) - want) / want < 1e-5


def _spot_check_nn(torch, pos, ax, nn, rs, n_random, n_far, n_sample=1 << 22):
    """Re-decide lattice nodes independently of the library: particles are binned into 64^3 coarse cells with torch.sort;
    for every picked node the particles of all coarse cells touching the cube of half-width R = 1.0001 |node - claimed NN|
    are fetched and the exact f64 argmin (ties -> lowest index) is taken on the host.  Picks: `n_random` uniformly random
    nodes + the `n_far` nodes with the LARGEST claimed distance out of `n_sample` random ones (the nodes the wide search
    stages had to settle) + the eight lattice corners.  -> (number of nodes checked, number of wrong answers, max R / h)."""
    N = len(ax)
    G = 64
    lo, hi = float(pos.min().item()), float(pos.max().item())
    w = (hi - lo) / G * (1 + 1e-6) + 1e-30
    cid = torch.zeros(pos.shape[0], dtype=torch.int32, device=pos.device)
    for c in range(3):
        cid = cid * G + ((pos[:, c] - lo) / w).to(torch.int32).clamp_(0, G - 1)
    sid, order = torch.sort(cid)
    del cid
    starts = torch.searchsorted(sid, torch.arange(G ** 3 + 1, dtype=torch.int32, device=pos.device)).cpu().numpy()
    del sid
    axt = torch.from_numpy(ax).to(pos.device)
    samp = torch.from_numpy(rs.integers(0, N, size=(n_sample, 3))).to(pos.device)
    got = nn[samp[:, 0], samp[:, 1], samp[:, 2]].long()
    d = ((axt[samp[:, 0]] - pos[got, 0].double()) ** 2 + (axt[samp[:, 1]] - pos[got, 1].double()) ** 2
         + (axt[samp[:, 2]] - pos[got, 2].double()) ** 2)
    far = samp[torch.topk(d, n_far).indices].cpu().numpy()
    corners = np.array([[a, b, c] for a in (0, N - 1) for b in (0, N - 1) for c in (0, N - 1)])
    pick = np.concatenate([samp[:n_random].cpu().numpy(), far, corners])
    h = ax[1] - ax[0]
    bad, rmax = 0, 0.0
    for i, j, k in pick:
        node = np.array([ax[i], ax[j], ax[k]])
        g = int(nn[i, j, k])
        pg = pos[g].cpu().numpy().astype(np.float64)
        dg = ((node[0] - pg[0]) ** 2 + (node[1] - pg[1]) ** 2) + (node[2] - pg[2]) ** 2
        R = np.sqrt(dg) * 1.0001 + 1e-7
        rmax = max(rmax, R / h)
        rng_c = [range(max(0, int((node[c] - R - lo) / w)), min(G - 1, int((node[c] + R - lo) / w)) + 1) for c in range(3)]
        idx = []
        for a in rng_c[0]:
            for b in rng_c[1]:
                c0, c1 = (a * G + b) * G + rng_c[2][0], (a * G + b) * G + rng_c[2][-1]
                idx.append(order[starts[c0]:starts[c1 + 1]])          # coarse cells along z are contiguous
        idx = torch.cat(idx)
        c = pos[idx].cpu().numpy().astype(np.float64)
        idx = idx.cpu().numpy()
        d2 = ((node[0] - c[:, 0]) ** 2 + (node[1] - c[:, 1]) ** 2) + (node[2] - c[:, 2]) ** 2
        best = idx[np.lexsort((idx, d2))[0]]
        bad += int(best != g)
    return len(pick), bad, rmax


def test_nn_cfg3_clustered_spot_check(lib, orc):
    """cfg3 (512^3 lattice, 2^27 particles, half of them in eight Gaussian blobs -- bench.py's generator): cells holding
    thousands of particles next to half-empty voids, so the long-row path of the cell-list build and the wide stages of the
    search all carry real load.  No CPU oracle builds a kd-tree of 1.3e8 points in test time: 2600 nodes are re-decided."""
    import torch
    import bench
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~40 GB of device memory")
    wl = bench.WORKLOADS["cfg3"]
    N = wl["N"]
    pos, _, _ = bench.synth_range(torch, wl, 0, wl["Np"])
    ax = orc.lattice_axis_lib(1.0, N)
    nn = lib.nn_grid(pos, ax, ax, ax)
    st = lib.nn_grid_stats()
    assert st["n_unresolved"] == 0 and st["n_kept"] == wl["Np"]
    assert st["n_wide"] > 100000                       # the exact stage is genuinely exercised
    lib.trim()
    n, bad, rmax = _spot_check_nn(torch, pos, ax, nn, np.random.default_rng(3), 1500, 1100)
    assert n >= 2600 and bad == 0
    assert rmax > 1.5                                   # nodes in voids were among the picks


def test_nn_cfg4_full_size_spot_check(lib, orc):
    """cfg4 at its own size: 1024^3 lattice, 2^30 uniform-random particles (bench.py's generator).  Exercises the two-pass
    row sort, the 32-bit slot arithmetic and every search stage at 2^30; 2100 nodes re-decided independently, 1000 of them
    the farthest-from-their-particle nodes of a 4M sample (what stages B and C settled)."""
    import torch
    import bench
    free, _ = torch.cuda.mem_get_info()
    if free < 150e9:
        pytest.skip("needs a whole B200 (~120 GB of device memory)")
    wl = bench.WORKLOADS["cfg4"]
    N = wl["N"]
    pos = torch.empty((wl["Np"], 3), dtype=torch.float32, device="cuda")
    step = 1 << 27
    for s in range(0, wl["Np"], step):                 # positions only: the same particle set bench.py times
        x = torch.stack([bench.hash_uniform_t(torch, wl["seed"], s, s + step, c, "cuda") for c in range(3)], dim=1)
        pos[s:s + step] = x
        del x
    ax = orc.lattice_axis_lib(1.0, N)
    nn = lib.nn_grid(pos, ax, ax, ax)
    st = lib.nn_grid_stats()
    assert st["n_unresolved"] == 0 and st["n_kept"] == wl["Np"]
    lib.trim()
    torch.cuda.empty_cache()
    assert int(nn.min()) >= 0 and int(nn.max()) < wl["Np"]
    n, bad, rmax = _spot_check_nn(torch, pos, ax, nn, np.random.default_rng(4), 1100, 1000)
    assert n >= 2100 and bad == 0
    assert rmax > 1.0                                   # nodes beyond the 2x2x2 proof radius were among the picks


def test_cfg2_full_vs_oracle(vp, lib, orc):
    """BASELINE cfg2 in full against the CPU oracle: 256^3 lattice, 2^24 particles, velocity + momentum spectra through the
    whole-path C-ABI call (device-resident and host-buffer forms).  Nearest-particle indices and mode counts bit-exact,
    binned P(k) within 1e-5 of the f64 oracle."""
    import torch
    import bench
    wl = bench.WORKLOADS["cfg2"]
    N, Np, L = wl["N"], wl["Np"], 1.0
    pos, vel, rho = bench.synth_range(torch, wl, 0, Np)
    p64, v64, d64 = (t.cpu().numpy().astype(np.float64) for t in (pos, vel, rho))
    ax, k, edges, lc3, norm = bench.geometry(N, L)
    ref_idx = orc.nn_exact_lattice(p64, ax, ax, ax)
    nn = lib.nn_grid(pos, ax, ax, ax).cpu().numpy()
    assert np.array_equal(nn, ref_idx)
    assert lib.nn_grid_stats()["n_unresolved"] == 0
    v_ref, m_ref, Lcell = orc.ann_interp_to_field(p64, d64, v64, L, N)
    out, ns = lib.particles_to_pk(pos, vel, rho, ax, ax, ax, N, lc3, norm, k, edges, quantities=wl["quantities"])
    out_h, ns_h = lib.particles_to_pk(p64.astype(np.float32), v64.astype(np.float32), d64.astype(np.float32), ax, ax, ax, N, lc3,
                                      norm, k, edges, quantities=wl["quantities"])
    assert np.array_equal(ns, bench.integer_shell_counts(N))
    for q in wl["quantities"]:
        ref = orc.spctrm(v_ref, m_ref, Lcell, q)
        assert np.array_equal(ns, ref[:, 3].astype(np.int64)) and np.array_equal(ns_h, ns)
        assert np.max(np.abs(out[q] / ref[:, 2] - 1)) < 1e-5
        assert np.max(np.abs(out_h[q] / ref[:, 2] - 1)) < 1e-5


def test_two_rank_dist_check_under_torchrun():
    """N>1 on real GPUs: tests/dist_gpu_check.py under torchrun with 2 ranks (NCCL + peer-store exchanges == replicated ==
    single GPU == CPU oracle).  Needs two visible B200s; the single-GPU test box skips it (the gloo tests in
    tests/test_dist_cpu.py cover the wiring there)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "dist_gpu_check.py")],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "DIST CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_nn_full_size_spot_check(lib, orc):
    """cfg3 scale (512^3 lattice, 2^27 particles, clustered: voids and sheets).  The CPU oracle cannot build a kd-tree of
    1.3e8 points in test time, so 600 random lattice nodes are re-decided independently: all particles inside a generous
    cube around the node (torch mask on the device) -> exact f64 argmin with lowest-index ties on the host."""
    import torch
    N, L = 512, 1.0
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~25 GB of device memory")
    g = torch.Generator(device="cuda").manual_seed(11)
    g1 = (torch.arange(N, device="cuda", dtype=torch.float32) + 0.5) / N
    q = torch.stack(torch.meshgrid(g1, g1, g1, indexing="ij"), dim=-1).reshape(-1, 3)
    rs = np.random.default_rng(11)
    disp = torch.zeros_like(q)
    for _ in range(24):
        kv = torch.tensor(rs.integers(1, 40, size=3) * rs.choice([-1, 1], size=3), device="cuda", dtype=torch.float32)
        kk = float(torch.linalg.norm(kv))
        disp += (1.0 / kk) * (kv / kk)[None, :] * torch.sin(2 * np.pi * (q @ kv) + float(rs.uniform(0, 6.28)))[:, None]
    disp *= (2.0 / N) / float(torch.sqrt((disp ** 2).sum(1).mean()))
    pos = torch.remainder(q + disp, 1.0).contiguous()
    del q, disp
    ax = orc.lattice_axis_lib(L, N)
    nn = lib.nn_grid(pos, ax, ax, ax)
    st = lib.nn_grid_stats()
    assert st["n_unresolved"] == 0 and st["n_kept"] == N ** 3
    h = ax[1] - ax[0]
    pick = rs.integers(0, N, size=(600, 3))
    bad = 0
    for i, j, k in pick:
        node = np.array([ax[i], ax[j], ax[k]])
        got = int(nn[i, j, k])
        pg = pos[got].cpu().numpy().astype(np.float64)
        dg = ((node[0] - pg[0]) ** 2 + (node[1] - pg[1]) ** 2) + (node[2] - pg[2]) ** 2
        R = np.sqrt(dg) * 1.0001 + 1e-6                       # anything nearer than the answer lies inside this cube
        m = ((pos[:, 0] - float(node[0])).abs() <= R) & ((pos[:, 1] - float(node[1])).abs() <= R) & ((pos[:, 2] - float(node[2])).abs() <= R)
        idx = torch.nonzero(m).reshape(-1).cpu().numpy()
        c = pos[m].cpu().numpy().astype(np.float64)
        d2 = ((node[0] - c[:, 0]) ** 2 + (node[1] - c[:, 1]) ** 2) + (node[2] - c[:, 2]) ** 2
        best = idx[np.lexsort((idx, d2))[0]]
        bad += int(best != got)
    assert bad == 0
    assert R < 20 * h


@pytest.mark.parametrize("Np,N,clustered", [(1 << 18, 64, False), (40 ** 3, 40, True), (5000, 48, False)])
def test_fused_gridding_fields_equals_two_step(lib, orc, Np, N, clustered):
    """vp_nn_grid_fields (the search stages write the planes) == vp_nn_grid_payload + vp_fields_sorted == the oracle's
    gather, bit for bit, and the indices it returns are the oracle's."""
    pos, vel, dens, _ = orc.synth_particles(5, Np, 1.0, clustered=clustered, lattice_n=40 if clustered else None)
    ax = orc.lattice_axis_lib(1.0, N)
    lc3 = (1.0 / N) ** 3
    dp, dv, dr = lib.to_device(pos), lib.to_device(vel), lib.to_device(dens)
    f1, nn = lib.nn_grid_fields(dp, dv, dr, ax, ax, ax, lc3, want_v=True, want_p=(True, True, True), want_e=True, want_m=True,
                                want_idx=True)
    _, nn_pos, srec = lib.nn_grid_payload(dp, dv, dr, ax, ax, ax, lc3)
    f2 = lib.fields_sorted(nn_pos, srec, want_v=True, want_p=(True, True, True), want_e=True, want_m=True)
    ref = orc.nn_exact_lattice(pos.astype(np.float64), ax, ax, ax)
    assert np.array_equal(nn.cpu().numpy(), ref)
    for k in f2:
        assert np.array_equal(f1[k].cpu().numpy(), f2[k].cpu().numpy()), k
    v32 = (vel * dens[:, None]) / dens[:, None]
    assert np.array_equal(f1["vx"].cpu().numpy(), v32[ref, 0]) and np.array_equal(f1["m"].cpu().numpy(), (dens * np.float32(lc3))[ref])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_snapshot_preamble_on_device(lib, vp, orc, dtype):
    """shift_to_origin / remove_bulk_velocity as device reductions (vp_snapshot_preamble) vs the reference's numpy
    expressions (interp.py:169-182): the shift is exact, the bulk velocity agrees to the rounding of the summation order."""
    import torch
    pos, vel, dens, _ = orc.synth_particles(9, 100003, 1.0)
    mass = (1.0 + orc.hash_uniform(9, 100003, 11)).astype(dtype)
    pos, vel = (pos + 0.37).astype(dtype), (vel + 0.25).astype(dtype)
    gp_ref = vp.interp.GasParticles(pos.copy(), mass.copy(), dens.astype(dtype), vel.copy(), 1.0)
    gp_ref.remove_bulk_velocity()
    gp_ref.shift_to_origin()
    gp = vp.interp.GasParticles(lib.to_device(pos), lib.to_device(mass), lib.to_device(dens.astype(dtype)), lib.to_device(vel), 1.0)
    gp.remove_bulk_velocity()
    gp.shift_to_origin()
    assert np.array_equal(gp.pos.cpu().numpy(), gp_ref.pos)                       # minimum and subtraction are exact
    tol = 2e-6 if dtype == np.float32 else 1e-13
    assert np.max(np.abs(gp.v.cpu().numpy() - gp_ref.v)) < tol * np.max(np.abs(vel))
    bf = gp.ann_interp_to_field(16)                                               # device-resident particles are gridded in place
    ax = orc.lattice_axis_lib(1.0, 16)
    assert np.array_equal(bf._dev["nn"].cpu().numpy(), orc.nn_exact_lattice(gp_ref.pos.astype(np.float64), ax, ax, ax))


def test_fft_binning_N1000_reference_default(lib, orc):
    """N = 1000, the reference MPI script's default NTOT (scripts/parallel_optimized.py:30), runs the mixed-radix line
    transform (10 x 10 x 10 points; the z pass 10 x 10 x 5 on the half length), not a fallback: Parseval over all modes,
    shells partitioning the N^3 modes (library edges: 500 shells; script edges: 499), exact linearity,
    and a plane wave landing in its shell."""
    import torch
    N, L = 1000, 1.0
    free, _ = torch.cuda.mem_get_info()
    if free < 30e9:
        pytest.skip("needs ~20 GB of device memory")
    g = torch.Generator(device="cuda").manual_seed(5)
    f = torch.randn((N, N, N), generator=g, device="cuda", dtype=torch.float32)
    f[:, :, ::2] += 0.5
    want = float((f.double() ** 2).sum().item()) * N ** 3
    k = orc.k_axis(L, N)
    one = lib.PkPlan(N, k, np.array([0.0, 1e30]))
    psum, ns = one.fields([f.clone()])
    assert int(ns[0]) == N ** 3 and abs(psum[0] - want) / want < 1e-5
    centres, edges = orc.edges_lib(2 * np.pi / L, np.pi * N / L, 2 * np.pi / L)
    assert len(centres) == 500
    plan = lib.PkPlan(N, k, edges)
    p1, n1 = plan.fields([f.clone()])
    p2, n2 = plan.fields([2.0 * f])
    assert np.array_equal(n1, n2) and np.allclose(p2, 4.0 * p1, rtol=1e-13, atol=0)
    import bench
    assert np.array_equal(n1, bench.integer_shell_counts(N))
    cs, es = orc.edges_script(2 * np.pi / L, np.pi * N / L, 2 * np.pi / L)
    assert len(cs) == 499
    x = (torch.arange(N, device="cuda", dtype=torch.float64) / N)
    wave = torch.cos(2 * np.pi * (3 * x[:, None, None] + 5 * x[None, :, None] + 7 * x[None, None, :])).float()
    pw, nw = plan.fields([wave])
    j0 = int(np.floor(np.sqrt(9 + 25 + 49) + 0.5)) - 1
    assert abs(pw[j0] / (float(N) ** 6 / 2) - 1) < 1e-5 and np.delete(pw, j0).sum() < 1e-6 * float(N) ** 6 / 2


# ------------------------------------------------------------------------------------------ folding (outer stage)
@pytest.mark.parametrize("name,seed,N,L,m", [("fold16_m2", 31, 16, 1.0, 2), ("fold24_m3", 32, 24, 2.5, 3), ("fold128_m2", 33, 128, 1.0, 2)])
def test_fold_vs_reference(vp, golden_fold, name, seed, N, L, m):
    """BoxField.fold(m, beta) and FoldedBox.fold_spctrm on the device against the unmodified reference (interp.py:598-609,
    755-791): folded field to 1e-12, mode counts bit-exact, Psum to 1e-9 where the folded transform runs in f64 (n = 8) and
    1e-5 where it runs through the f32 line FFT (n = 64)."""
    import vpower.interp as vi
    from conftest import fold_case_field
    g = golden_fold
    v, mass = fold_case_field(seed, N)
    bf = vi.BoxField(v, mass, L / N)
    for tag in sorted(k.split("_b")[-1] for k in g.files if k.startswith(f"{name}/spctrm_b")):
        beta = np.array([int(t) for t in tag])
        fb = bf.fold(m, beta)
        assert fb.Nsize == N // m and fb.m == m and abs(fb.totalLbox - L) < 1e-15
        sp = fb.fold_spctrm(None, beta=beta)
        f = fb.f
        assert f.dtype == np.complex128 and f.shape == (N // m,) * 3 + (3,)
        if f"{name}/folded_b{tag}" in g.files:
            assert np.abs(f - g[f"{name}/folded_b{tag}"]).max() < 1e-12
        else:
            assert np.abs(f[::8, ::8, ::8, :] - g[f"{name}/folded_b{tag}_sample"]).max() < 1e-12
        ref = g[f"{name}/spctrm_b{tag}"]
        got = sp.data()
        assert sp.m == m and tuple(sp.beta) == tuple(beta)
        assert np.array_equal(got[:, 0], ref[:, 0])
        assert np.array_equal(got[:, 3], ref[:, 3])                                   # mode counts bit-exact
        ok = ref[:, 3] > 0
        rel = np.abs(got[ok, 2] - ref[ok, 2]) / ref[ok, 2]
        assert rel.max() < (1e-9 if N // m < 64 else 1e-5), (name, tag, rel.max())
        assert np.allclose(got[ok, 1], ref[ok, 1], rtol=1e-5 if N // m >= 64 else 1e-9)


def test_fold_from_numpy_folded_box(vp, golden_fold):
    """A FoldedBox built by hand from a numpy array (as reference-produced pickles are) goes through the same device path."""
    import vpower.interp as vi
    g = golden_fold
    f = g["fold16_m2/folded_b100"]
    fb = vi.FoldedBox(f.copy(), 2, np.array([1, 0, 0]), 0.5, 8)
    got = fb.fold_spctrm(None, beta=np.array([1, 0, 0])).data()
    ref = g["fold16_m2/spctrm_b100"]
    assert np.array_equal(got[:, 3], ref[:, 3])
    ok = ref[:, 3] > 0
    assert (np.abs(got[ok, 2] - ref[ok, 2]) / ref[ok, 2]).max() < 1e-9


@pytest.mark.parametrize("Np,N,with_rho", [(1 << 18, 64, True), ((1 << 17) + 4099, 48, True), (3 * 4096, 32, False)])
def test_interleaved_rows_equal_compact(lib, orc, Np, N, with_rho):
    """vp_nn_opts.row_stride = 8: positions, velocities and densities read in place from the 32-byte rows the slab exchange
    delivers (x y z vx | vy vz rho -) -- the row forms of the bucket kernels -- give the planes and indices of the compact
    arrays, bit for bit (whole tiles and a ragged last tile)."""
    import torch
    pos, vel, dens, _ = orc.synth_particles(9, Np, 1.0)
    ax = orc.lattice_axis_lib(1.0, N)
    lc3 = (1.0 / N) ** 3
    dp, dv, dr = lib.to_device(pos), lib.to_device(vel), lib.to_device(dens)
    rows = torch.zeros((Np, 8), dtype=torch.float32, device="cuda")
    rows[:, 0:3], rows[:, 3:6], rows[:, 6] = dp, dv, dr
    o = lib.NNOpts()
    o.row_stride = 8
    kw = dict(want_v=True, want_p=(True, False, False), want_e=True, want_m=True, want_idx=True)
    f1, nn1 = lib.nn_grid_fields(dp, dv, dr if with_rho else None, ax, ax, ax, lc3, **kw)
    f2, nn2 = lib.nn_grid_fields(rows[:, 0:3], rows[:, 3:6], rows[:, 6] if with_rho else None, ax, ax, ax, lc3, opts=o, **kw)
    assert np.array_equal(nn1.cpu().numpy(), nn2.cpu().numpy())
    assert np.array_equal(nn1.cpu().numpy(), orc.nn_exact_lattice(pos.astype(np.float64), ax, ax, ax))
    for k in f1:
        assert np.array_equal(f1[k].cpu().numpy(), f2[k].cpu().numpy()), k
