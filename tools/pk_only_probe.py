"""Probe: stage times and a result digest of vp_pk_fields alone (K4 + K5) on smooth-plus-noise cubes.
usage: pk_only_probe.py N ncomp [reps]   (environment: VP_X_TMA=0 selects the per-thread-load x pass)"""
import json, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "large-velocity-power-spectrum_b200"))
import numpy as np
import torch
from vpower import _lib
N = int(sys.argv[1]); ncomp = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
k = 2 * np.pi * np.fft.fftfreq(N, d=1.0 / N)
edges = np.arange(0.5, N // 2 + 1, 1.0) * 2 * np.pi
plan = _lib.PkPlan(N, k, edges)
g = torch.Generator(device="cuda"); g.manual_seed(5)
src = [torch.randn((N, N, N), dtype=torch.float32, device="cuda", generator=g) for _ in range(ncomp)]
def run():
    cubes = [s.clone() for s in src]
    return plan.fields(cubes)
for _ in range(2):
    ps, ns = run()
torch.cuda.synchronize()
_lib.profile_enable(True); _lib.profile_report()
for _ in range(reps):
    ps, ns = run()
st = _lib.profile_report()
print(json.dumps({"probe": "vp_pk_fields", "N": N, "ncomp": ncomp, "tma": os.environ.get("VP_X_TMA", "1"), "promo": os.environ.get("VP_X_L2PROMO", "1"),
                  "nsample_crc32": zlib.crc32(ns.tobytes()), "psum_sum": float(ps.sum()), "psum_crc32": zlib.crc32(ps.astype(np.float32).tobytes()),
                  "stages_ms": {kk: round(v["ms"] / reps, 4) for kk, v in st.items()}}))
