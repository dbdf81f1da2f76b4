"""K2 (vp_deposit_ngp, the device form of deposit_to_grid, vpower/interp.py:996-1015) timed on its own: it is not on the
particles -> P(k) path of any BASELINE configuration, so bench.py has no line for it.  One JSON line per case, CUDA-event
timed, inputs resident and far larger than L2; roofline against the measured HBM peak with the kernel's stated algorithmic
bytes (positions read + weights read per particle; the f64 grid cells are read-modify-written by the atomics and count as
traffic, not as algorithmic bytes)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "large-velocity-power-spectrum_b200"))
import torch
import bench
from vpower import _lib

peak = 6539.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
except Exception:
    pass
for N, lg, C in ((512, 27, 1), (512, 27, 4), (1024, 28, 1)):
    Np = 1 << lg
    pos = torch.stack([bench.hash_uniform_t(torch, 5, 0, Np, c, "cuda") for c in range(3)], dim=1).contiguous()
    w = torch.rand((Np, C) if C > 1 else (Np,), dtype=torch.float64, device="cuda")
    for _ in range(2):
        g = _lib.deposit_ngp(pos, w, N, 1.0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    a.record()
    for _ in range(steps):
        g = _lib.deposit_ngp(pos, w, N, 1.0)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    alg = Np * (12.0 + 8.0 * C)
    print(json.dumps({"kernel": "k2_deposit_ngp", "N": N, "Np": Np, "ncomp": C, "ms": round(ms, 3), "Gpart_s": round(Np / ms / 1e6, 2),
                      "roofline": {"bound": "hbm", "achieved": round(alg / ms / 1e6, 1), "peak": peak, "unit": "GB/s",
                                   "frac": round(alg / ms / 1e6 / peak, 3)},
                      "check_sum_of_grid_minus_sum_of_weights": float((g.sum() - w.sum()).abs() / w.sum())}))
    del pos, w, g
