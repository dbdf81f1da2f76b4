"""Import shims that let the UNMODIFIED reference Python run in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/vpower_oracle.py header).  Used by
tests/golden/make_golden.py to produce the committed golden vectors, and by
nothing that runs on the GPU box (the reference tree is not there).

The reference imports third-party packages that are not installed here and
cannot be installed (no network): pyfftw, pyann, annoy, mpi4py, h5py,
matplotlib, voxelize, memory_profiler.  Each fake below provides exactly the
attributes the reference touches (SURVEY.md Appendix C lists the call sites).

  pyfftw  -> scipy.fft (dtype preserving like pyFFTW)
  pyann   -> the reference's own prebuilt `ann/ann_sample` ELF (true ANN 1.1.2
             engine) when `engine="ann_sample"`, else the exact oracle search
  annoy   -> exact nearest neighbour (Annoy's random-projection forest is
             approximate and unseeded; the contract is the exact search)
  mpi4py  -> single-rank communicator
  h5py    -> dict-backed file registered with `register_snapshot`
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile
import types

import numpy as np
import scipy.fft

REF_ROOT = os.environ.get("VPOWER_REFERENCE", "/root/reference")
ANN_SAMPLE = os.path.join(REF_ROOT, "ann", "ann_sample")
# The ELF is an unaudited binary from the public reference tree.  It is executed ONLY when the maintainer regenerates
# the golden vectors (tests/golden/make_golden.py) and has opted in with VPOWER_ALLOW_REFERENCE_ELF=1, and only if its
# sha256 is the one that was inspected when the vectors were made.  Nothing on the GPU box, in build() or in pytest runs it.
ANN_SAMPLE_SHA256 = "5128f74259b161021e7e7766b0a911c1cd5ec6b08fef44361cae30847bff009a"

_snapshots = {}
_state = {"engine": "oracle", "last_nn": None}


def register_snapshot(path, coords, masses, density, velocities):
    _snapshots[path] = {
        "PartType0/Coordinates": coords, "PartType0/Masses": masses,
        "PartType0/Density": density, "PartType0/Velocities": velocities,
    }
    if not os.path.isfile(path):
        with open(path, "wb") as f:
            f.write(b"placeholder for the h5py shim\n")


def run_ann_sample(data, query):
    """Run the reference ELF exactly as vpower/interp.py:1052-1134 would drive it:
    text files written with '%.16f' tab separated; returns (idx0, data_parsed, query_parsed)
    where *_parsed are the doubles the binary actually saw (text round trip)."""
    import hashlib
    if os.environ.get("VPOWER_ALLOW_REFERENCE_ELF") != "1":
        raise RuntimeError("refusing to execute the reference's prebuilt ann/ann_sample: set VPOWER_ALLOW_REFERENCE_ELF=1 "
                           "(only needed to regenerate tests/golden/reference_golden.npz)")
    with open(ANN_SAMPLE, "rb") as fh:
        digest = hashlib.sha256(fh.read()).hexdigest()
    if digest != ANN_SAMPLE_SHA256:
        raise RuntimeError(f"{ANN_SAMPLE}: sha256 {digest} is not the pinned {ANN_SAMPLE_SHA256}")
    with tempfile.TemporaryDirectory() as td:
        dp, qp = os.path.join(td, "data.pts"), os.path.join(td, "query.pts")
        np.savetxt(dp, data, delimiter="\t", fmt="%.16f")
        np.savetxt(qp, query, delimiter="\t", fmt="%.16f")
        out = subprocess.run([ANN_SAMPLE, "-d", "3", "-e", "0", "-max", str(len(data)), "-nn", "1",
                              "-df", dp, "-qf", qp], check=True, capture_output=True, text=True).stdout
        data_parsed = np.loadtxt(dp, delimiter="\t", ndmin=2)
        query_parsed = np.loadtxt(qp, delimiter="\t", ndmin=2)
    rows = np.array([ln.split("\t") for ln in out.strip().split("\n")])
    return rows[:, 1].astype(np.int64), data_parsed, query_parsed


def install(engine="oracle"):
    """Install the fake modules.  engine: 'oracle' | 'ann_sample' for pyann.nn2."""
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import vpower_oracle as orc

    _state["engine"] = engine

    # ---- pyfftw -----------------------------------------------------------
    pyfftw = types.ModuleType("pyfftw")
    interfaces = types.ModuleType("pyfftw.interfaces")
    cache = types.ModuleType("pyfftw.interfaces.cache")
    cache.enable = lambda *a, **k: None
    numpy_fft = types.ModuleType("pyfftw.interfaces.numpy_fft")
    numpy_fft.fftn = lambda a, threads=1, overwrite_input=False, **k: scipy.fft.fftn(a)
    interfaces.cache, interfaces.numpy_fft = cache, numpy_fft
    pyfftw.interfaces = interfaces
    pyfftw.empty_aligned = lambda shape, dtype="float64", **k: np.empty(shape, dtype=dtype)

    class FFTW:
        def __init__(self, a, b, axes=(0, 1, 2), **k):
            self.a, self.b, self.axes = a, b, axes

        def __call__(self, x=None):
            if x is not None:
                return scipy.fft.fftn(x, axes=self.axes)
            self.b[...] = scipy.fft.fftn(self.a, axes=self.axes)
            return self.b

    pyfftw.FFTW = FFTW
    builders = types.ModuleType("pyfftw.builders")
    builders.fftn = lambda a, threads=1, **k: (lambda x=None: scipy.fft.fftn(a if x is None else x))
    pyfftw.builders = builders
    sys.modules.update({"pyfftw": pyfftw, "pyfftw.interfaces": interfaces,
                        "pyfftw.interfaces.cache": cache, "pyfftw.interfaces.numpy_fft": numpy_fft,
                        "pyfftw.builders": builders})

    # ---- pyann ------------------------------------------------------------
    pyann = types.ModuleType("pyann")

    def nn2(data, query, k=1, eps=0.0, treetype="kd", searchtype="standard"):
        d, q = np.asarray(data, dtype=np.float64), np.asarray(query, dtype=np.float64)
        if _state["engine"] == "ann_sample":
            idx0, dpar, qpar = run_ann_sample(d, q)
            _state["last_nn"] = {"idx0": idx0, "data_parsed": dpar, "query_parsed": qpar}
        else:
            idx0 = orc.nn_exact_points(d, q)
            _state["last_nn"] = {"idx0": idx0}
        res = types.SimpleNamespace()
        res.nn_idx = np.matrix((idx0 + 1).reshape(-1, 1))          # pyann is 1-based (interp.py:1037)
        return res

    pyann.nn2 = nn2
    sys.modules["pyann"] = pyann

    # ---- annoy ------------------------------------------------------------
    annoy = types.ModuleType("annoy")

    class AnnoyIndex:
        def __init__(self, f, metric):
            self.items, self.tree = {}, None

        def add_item(self, i, vec):
            self.items[i] = np.asarray(vec, dtype=np.float32)        # Annoy stores float32

        def build(self, n_trees, n_jobs=-1):
            n = len(self.items)
            self.data = np.stack([self.items[i] for i in range(n)]).astype(np.float64)
            self.tree = None

        def save(self, path):
            np.save(path + ".npy", self.data)
            open(path, "wb").close()

        def load(self, path):
            self.data = np.load(path + ".npy")
            self.tree = None

        def get_nns_by_vector(self, q, n=1, search_k=-1, include_distances=False):
            qq = np.asarray(q, dtype=np.float32).astype(np.float64)[None, :]
            if self.tree is None:                 # one kd-tree per index, not one per query (the decision stays the oracle's)
                from scipy.spatial import cKDTree
                self.tree = cKDTree(self.data)
            return [int(orc.nn_exact_points(self.data, qq, tree=self.tree)[0])]

    annoy.AnnoyIndex = AnnoyIndex
    sys.modules["annoy"] = annoy

    # ---- mpi4py -----------------------------------------------------------
    mpi4py = types.ModuleType("mpi4py")
    MPI = types.ModuleType("mpi4py.MPI")

    class _Comm:
        def Get_rank(self): return 0
        def Get_size(self): return 1
        def Barrier(self): pass
        def allgather(self, obj): return [obj]
        def Reduce(self, sendbuf, recvbuf, op=None, root=0): recvbuf[...] = sendbuf

    MPI.COMM_WORLD, MPI.SUM = _Comm(), "SUM"
    mpi4py.MPI = MPI
    sys.modules.update({"mpi4py": mpi4py, "mpi4py.MPI": MPI})

    # ---- h5py -------------------------------------------------------------
    h5py = types.ModuleType("h5py")

    class _Group(dict):
        def __getitem__(self, key):
            if dict.__contains__(self, key):
                return dict.__getitem__(self, key)
            sub = {k[len(key) + 1:]: v for k, v in self.items() if k.startswith(key + "/")}
            if not sub:
                raise KeyError(key)
            return _Group(sub)

    class File(_Group):
        def __init__(self, path, mode="r"):
            super().__init__({k: np.array(v, copy=True) for k, v in _snapshots[path].items()})

        def close(self): pass

    h5py.File = File
    sys.modules["h5py"] = h5py

    # ---- stubs ------------------------------------------------------------
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.Axes = object
    colors = types.ModuleType("matplotlib.colors")
    colors.LogNorm = object
    mpl.pyplot, mpl.colors = plt, colors
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.colors": colors})

    vox = types.ModuleType("voxelize")

    class Voxelize:
        def __init__(self=None, *a, **k): pass

    vox.Voxelize = Voxelize
    sys.modules["voxelize"] = vox

    mp = types.ModuleType("memory_profiler")
    mp.profile = lambda f: f
    sys.modules["memory_profiler"] = mp

    # reference does `from spctrm import PowerSpectrum` (interp.py:42)
    vp = os.path.join(REF_ROOT, "vpower")
    if vp not in sys.path:
        sys.path.insert(0, vp)
    return _state
