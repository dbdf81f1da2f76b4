"""Pin the CPU oracle (oracle/vpower_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

LIB_CASES = [("lib16", 16, 1.0), ("lib24", 24, 2.5), ("lib32c", 32, 1.0)]


def _pos_seen_by_ann(g, name):
    return g[f"{name}/pos_parsed"] if bool(g[f"{name}/pos_parsed_differs"]) else g[f"{name}/pos"]


@pytest.mark.parametrize("name,N,L", LIB_CASES)
def test_nn_matches_ann_engine(golden, orc, name, N, L):
    pos = _pos_seen_by_ann(golden, name)
    ax = [golden[f"{name}/axis_parsed{c}"] for c in range(3)]
    idx, ties = orc.nn_exact_lattice(pos, *ax, return_ties=True)
    gties = np.unpackbits(golden[f"{name}/nn_ties"])[: N ** 3].astype(bool).reshape(N, N, N)
    assert np.array_equal(ties, gties)
    ann = golden[f"{name}/nn_ann"]
    assert np.array_equal(idx[~ties], ann[~ties])      # bit-exact on every non-tied query


@pytest.mark.parametrize("name,N,L", LIB_CASES)
def test_lattice_axis(golden, orc, name, N, L):
    assert np.array_equal(orc.lattice_axis_lib(L, N), golden[f"{name}/axis_lib"])
    # text round trip used by the ANN CLI moves nodes by < 1e-16
    assert np.allclose(orc.lattice_axis_lib(L, N), golden[f"{name}/axis_parsed0"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("name,N,L", LIB_CASES)
def test_fields_from_ann_indices(golden, orc, name, N, L):
    """Payload algebra interp.py:199-213,272-273 applied to the golden ANN indices."""
    idx = golden[f"{name}/nn_ann"].astype(np.int64)
    vel, dens = golden[f"{name}/vel"].astype(np.float64), golden[f"{name}/dens"].astype(np.float64)
    vec = orc.density_velocity_vector(vel, dens)[idx.ravel()].reshape(N, N, N, 4)
    v = vec[..., :3] / vec[..., 3, None]
    m = vec[..., 3] * (L / N) ** 3
    assert np.array_equal(v, golden[f"{name}/v_grid"])
    assert np.array_equal(m, golden[f"{name}/m_grid"])


@pytest.mark.parametrize("name,N,L", LIB_CASES)
@pytest.mark.parametrize("q", ["velocity", "momentum", "energy"])
def test_spctrm(golden, orc, name, N, L, q):
    ref = golden[f"{name}/spctrm_{q}"]
    got = orc.spctrm(golden[f"{name}/v_grid"], golden[f"{name}/m_grid"], L / N, q)
    assert got.shape == ref.shape == (N // 2, 4)
    assert np.array_equal(got[:, 0], ref[:, 0])                     # bin centres
    assert np.array_equal(got[:, 3], ref[:, 3])                     # Nsample bit-exact
    assert np.allclose(got[:, 2], ref[:, 2], rtol=1e-12, atol=0)
    assert np.allclose(got[:, 1], ref[:, 1], rtol=1e-12, atol=0)


def test_power_grid_and_pairs(golden, orc):
    v, m = golden["lib16/v_grid"], golden["lib16/m_grid"]
    assert np.allclose(orc.power_grid(v, m, 1.0 / 16, "velocity"), golden["lib16/Pgrid_velocity"], rtol=1e-12)
    assert np.allclose(orc.power_grid(v, m, 1.0 / 16, "energy"), golden["lib16/Pgrid_energy"], rtol=1e-12)
    assert np.array_equal(orc.k_magnitude(1.0, 16), golden["lib16/pairs_k"])


def test_whole_library_path(golden, orc):
    """particles -> P(k) entirely inside the oracle vs the reference run (true ANN engine)."""
    name, N, L = "lib16", 16, 1.0
    assert not bool(golden[f"{name}/pos_parsed_differs"])
    for q in ("velocity", "momentum", "energy"):
        got = orc.particles_to_pk_lib(golden[f"{name}/pos"], golden[f"{name}/dens"].astype(np.float64),
                                      golden[f"{name}/vel"].astype(np.float64), L, N, q)
        ref = golden[f"{name}/spctrm_{q}"]
        assert np.array_equal(got[:, 3], ref[:, 3])
        assert np.allclose(got[:, 1:3], ref[:, 1:3], rtol=1e-10, atol=0)


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_deposit(golden, orc, tag):
    pos, w1, w4 = golden[f"deposit_{tag}/pos"], golden[f"deposit_{tag}/w1"], golden[f"deposit_{tag}/w4"]
    assert np.array_equal(orc.deposit_to_grid(w1, pos, 12, 1.5), golden[f"deposit_{tag}/grid1"])
    assert np.array_equal(orc.deposit_to_grid(w4, pos, 12, 1.5), golden[f"deposit_{tag}/grid4"])


@pytest.mark.parametrize("N,L", [(16, 1.0), (32, 2.5), (64, 1.0), (48, 0.7)])
def test_shell_geometry(golden, orc, N, L):
    kmin, kmax = 2 * np.pi / L, np.pi / (L / N)
    c, e = orc.edges_lib(kmin, kmax, kmin)
    assert np.array_equal(c, golden[f"shells_lib_{N}_{L}/centres"])
    assert np.array_equal(orc.shell_counts(L, N, e), golden[f"shells_lib_{N}_{L}/Nsample"].astype(np.int64))
    if (N, L) == (64, 1.0):
        assert orc.shell_counts(L, N, e).sum() == 143457
    if (N, L) == (16, 1.0):
        assert orc.shell_counts(L, N, e).tolist() == [18, 62, 98, 210, 350, 450, 602, 687]


@pytest.mark.parametrize("name", ["script16", "script16_fold2"])
def test_script_path(golden, orc, name):
    """Oracle's full-transform restatement vs the script's folded, f32 pipeline (verbatim run)."""
    pos, vel, mass = golden[f"{name}/pos"], golden[f"{name}/vel"], golden[f"{name}/mass"]
    # script preamble parallel_optimized.py:280-288: min-corner shift, mass-weighted bulk removal
    pos = pos - pos.min(axis=0)
    M = np.sum(mass)
    vel = vel.copy()
    for c in range(3):
        vel[:, c] -= np.sum(mass * vel[:, c]) / M
    got = orc.particles_to_pk_script(pos, vel, 16, 1)
    ref = golden[f"{name}/Pk"]
    assert got.shape == ref.shape
    assert np.allclose(got[:, 0], ref[:, 0], rtol=1e-6)             # script casts to f32
    assert np.array_equal(got[:, 3], ref[:, 3])
    assert np.allclose(got[:, 2], ref[:, 2], rtol=2e-5)
    assert np.allclose(got[:, 1], ref[:, 1], rtol=2e-5)


# ------------------------------------------------------------------------------------------ folding (outer stage)
FOLD_CASES = [("fold16_m2", 31, 16, 1.0, 2), ("fold24_m3", 32, 24, 2.5, 3), ("fold128_m2", 33, 128, 1.0, 2)]


@pytest.mark.parametrize("name,seed,N,L,m", FOLD_CASES)
def test_fold_oracle_vs_reference(golden_fold, orc, name, seed, N, L, m):
    """oracle fold_velocity / fold_spctrm == BoxField.fold / FoldedBox.fold_spctrm of the unmodified reference."""
    from conftest import fold_case_field
    g = golden_fold
    v, _ = fold_case_field(seed, N)
    if f"{name}/v" in g.files:
        assert np.array_equal(v, g[f"{name}/v"])
    tags = sorted(k.split("_b")[-1] for k in g.files if k.startswith(f"{name}/spctrm_b"))
    assert tags
    for tag in tags:
        beta = [int(t) for t in tag]
        f = orc.fold_velocity(v, m, beta)
        if f"{name}/folded_b{tag}" in g.files:
            assert np.abs(f - g[f"{name}/folded_b{tag}"]).max() < 1e-13
        else:
            assert np.abs(f[::8, ::8, ::8, :] - g[f"{name}/folded_b{tag}_sample"]).max() < 1e-13
            assert abs(f.sum() - g[f"{name}/folded_b{tag}_sum"][0]) < 1e-9 * max(1.0, abs(f.sum()))
        sp = orc.fold_spctrm(f, m, beta, L)
        ref = g[f"{name}/spctrm_b{tag}"]
        assert np.array_equal(sp[:, 0], ref[:, 0]) and np.array_equal(sp[:, 3], ref[:, 3])
        ok = ref[:, 3] > 0
        assert np.abs(sp[ok, 2] - ref[ok, 2]).max() <= 1e-12 * np.abs(ref[ok, 2]).max()


def test_fold_subspectra_partition_the_full_spectrum(golden_fold):
    """The reference's own identity behind folding: over all m^3 residue classes the sub-spectra hold every mode exactly once
    (checked here on the classes the golden file carries: their mode counts never exceed the full spectrum's)."""
    g = golden_fold
    full = g["fold16_m2/spctrm_full"]
    total = sum(g[k][:, 3].sum() for k in g.files if k.startswith("fold16_m2/spctrm_b"))
    assert total <= full[:, 3].sum() + 16 ** 3       # shells of the shifted lattices overlap the unshifted ones only partly
    assert total > 0
