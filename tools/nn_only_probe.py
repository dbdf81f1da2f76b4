"""Probe: stage times of vp_nn_grid (no payload: 16-byte sorted records) at cfg4 size -- compares the search on 16-byte
records with the search on the 32-byte (search half | payload half) records of the whole path."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "large-velocity-power-spectrum_b200"))
import torch
import bench
from vpower import _lib
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
N, Np = wl["N"], wl["Np"]
pos = torch.empty((Np, 3), dtype=torch.float32, device="cuda")
step = 1 << 26
for s in range(0, Np, step):
    e = min(Np, s + step)
    pos[s:e] = torch.stack([bench.hash_uniform_t(torch, wl["seed"], s, e, c, "cuda") for c in range(3)], dim=1)
ax = bench.geometry(N, 1.0)[0]
for _ in range(2):
    nn = _lib.nn_grid(pos, ax, ax, ax)
torch.cuda.synchronize()
_lib.profile_enable(True); _lib.profile_report()
for _ in range(3):
    nn = _lib.nn_grid(pos, ax, ax, ax)
st = _lib.profile_report()
print(json.dumps({"probe": "nn_grid without payload", "stats": _lib.nn_grid_stats(), "stages_ms": {k: round(v["ms"] / 3, 3) for k, v in st.items()}}))
