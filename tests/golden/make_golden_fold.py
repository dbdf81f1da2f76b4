"""Generate tests/golden/reference_golden_fold.npz by running the UNMODIFIED reference's folding stage.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_fold.py

What runs: /root/reference/vpower/interp.py -- BoxField.fold (:598-609), _get_phase / _apply_phase / fold_field
(:1195-1252), FoldedBox.fold_spctrm (:755-791) with _FFTW_vector_power, _pair_power(shift) and _hist_sample -- imported
verbatim through the import shims of oracle/refshims.py (no third-party engine is involved in this stage: the
`fft_object` the reference expects from the caller is scipy.fft.fftn over the three lattice axes, which is what a
pyFFTW plan for a [n,n,n,3] array computes).  Also voxelize_interp_to_field-free: nothing here touches the NN engine.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import scipy.fft

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import refshims  # noqa: E402

refshims.install(engine="oracle")
import interp as ref_interp  # noqa: E402  (the reference module)

out = {}
manifest = {"reference": "YujieH3/large-velocity-power-spectrum @ /root/reference", "cases": {}}


def fft_object(f):
    return scipy.fft.fftn(f, axes=(0, 1, 2))


def fold_case(name, seed, N, Lbox, m, betas):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(N, N, N, 3)).astype(np.float32)
    x = (np.arange(N) + 0.5) / N
    v[..., 0] += (2.0 * np.sin(2 * np.pi * 3 * x)[:, None, None]).astype(np.float32)
    v[..., 2] += (1.5 * np.cos(2 * np.pi * 5 * x)[None, :, None]).astype(np.float32)
    mass = (1.0 + rng.random((N, N, N))).astype(np.float32)
    bf = ref_interp.BoxField(v.copy(), mass.copy(), Lbox / N)
    out[f"{name}/v"] = v
    out[f"{name}/mass"] = mass
    with contextlib.redirect_stdout(io.StringIO()):
        full = bf.spctrm("velocity").data()
    out[f"{name}/spctrm_full"] = full
    for beta in betas:
        b = np.array(beta)
        tag = "".join(str(int(t)) for t in beta)
        with contextlib.redirect_stdout(io.StringIO()):
            fb = bf.fold(m, b)
            out[f"{name}/folded_b{tag}"] = np.asarray(fb.f).copy()          # complex128 [n,n,n,3]
            sp = fb.fold_spctrm(fft_object, beta=b)
        out[f"{name}/spctrm_b{tag}"] = sp.data()
        assert sp.m == m and tuple(sp.beta) == tuple(beta)
    manifest["cases"][name] = {"kind": "fold", "seed": seed, "N": N, "Lbox": Lbox, "m": m, "betas": [list(b) for b in betas]}
    print(name, "done:", len(betas), "sub-spectra,", "n =", N // m)


fold_case("fold16_m2", seed=31, N=16, Lbox=1.0, m=2, betas=[(0, 0, 0), (1, 0, 0), (0, 1, 1), (1, 1, 1)])
fold_case("fold24_m3", seed=32, N=24, Lbox=2.5, m=3, betas=[(0, 0, 0), (2, 1, 0), (1, 2, 2)])
fold_case("fold128_m2", seed=33, N=128, Lbox=1.0, m=2, betas=[(0, 0, 0), (1, 0, 1)])

# only the small cases keep their full arrays; the 128^3 case keeps the spectra and a checksum of the folded field
for k in list(out):
    if k.startswith("fold128_m2/folded_"):
        a = out.pop(k)
        out[k + "_sample"] = a[::8, ::8, ::8, :].copy()
        out[k + "_sum"] = np.array([a.sum()])
    if k == "fold128_m2/v" or k == "fold128_m2/mass":
        out.pop(k)       # regenerated from the seed in the test (same recipe as above)

np.savez_compressed(os.path.join(HERE, "reference_golden_fold.npz"), **out)
with open(os.path.join(HERE, "reference_golden_fold.json"), "w") as f:
    json.dump(manifest, f, indent=1)
print("wrote", os.path.join(HERE, "reference_golden_fold.npz"),
      os.path.getsize(os.path.join(HERE, "reference_golden_fold.npz")) // 1024, "KiB")
