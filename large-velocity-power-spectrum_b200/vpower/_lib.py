"""ctypes binding of libvpower_b200.so (include/vpower_b200.h) + thin torch plumbing.

PyTorch is used only for device memory and streams.  There is no CPU fallback: importing this
module succeeds anywhere (so that the host-side logic can be tested), but every compute call
raises `VPowerError` when the shared library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libvpower_b200.so")

VP_F32, VP_F64 = 0, 1


class VPowerError(RuntimeError):
    pass


class NNOpts(C.Structure):
    _fields_ = [("cells_x", C.c_int), ("cells_y", C.c_int), ("cells_z", C.c_int), ("use_x_keep", C.c_int),
                ("x_keep_lo", C.c_double), ("x_keep_hi", C.c_double), ("x_lo_is_domain_edge", C.c_int),
                ("x_hi_is_domain_edge", C.c_int), ("row_stride", C.c_int)]


_P, _D, _I, _L = C.c_void_p, C.c_double, C.c_int, C.c_int64
_dp = C.POINTER(C.c_double)

# name -> (restype, argtypes): every symbol declared in include/vpower_b200.h
SIGNATURES = {
    "vp_version": (_I, []),
    "vp_last_error": (C.c_char_p, []),
    "vp_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "vp_ctx_destroy": (_I, [_P]),
    "vp_ctx_arena_bytes": (C.c_size_t, [_P]),
    "vp_ctx_trim": (_I, [_P]),
    "vp_profile_enable": (_I, [_P, _I]),
    "vp_profile_report": (_I, [_P, C.c_char_p, C.c_size_t]),
    "vp_launch_count": (C.c_ulonglong, [_P]),
    "vp_nn_grid": (_I, [_P, _P, _I, _L, _dp, _I, _dp, _I, _dp, _I, _P, C.POINTER(NNOpts), _P]),
    "vp_nn_grid_stats": (_I, [_P, C.POINTER(_L), C.POINTER(_L), C.POINTER(_L), _P]),
    "vp_nn_grid_stats_ex": (_I, [_P, C.POINTER(_L), _P]),
    "vp_nn_grid_plan": (_I, [_L, _dp, _I, _dp, _I, _dp, _I, C.POINTER(NNOpts), C.POINTER(_L)]),
    "vp_fft_x_layout": (_I, [_I, _I, C.POINTER(_L)]),
    "vp_nn_grid_payload": (_I, [_P, _P, _P, _P, _I, _L, _dp, _I, _dp, _I, _dp, _I, _D, _P, _P, _P, C.POINTER(NNOpts), _P]),
    "vp_nn_grid_fields": (_I, [_P, _P, _P, _P, _I, _L, _dp, _I, _dp, _I, _dp, _I, _D, C.POINTER(_P), C.POINTER(_P), _P, _P, _P,
                               C.POINTER(NNOpts), _P]),
    "vp_fields_sorted": (_I, [_P, _P, _L, _P, C.POINTER(_P), C.POINTER(_P), _P, _P, _P]),
    "vp_slab_bucket": (_I, [_P, _P, _P, _P, _I, _L, _dp, _dp, _I, _P, _L, C.POINTER(_L), _P]),
    "vp_slab_p2p_close": (_I, [_P]),
    "vp_slab_p2p_alloc": (_I, [_P, C.c_size_t, C.c_char_p]),
    "vp_slab_p2p_open": (_I, [_P, _I, _I, C.c_char_p]),
    "vp_slab_p2p_buffer": (_I, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "vp_slab_count": (_I, [_P, _P, _I, _L, _dp, _dp, _I, C.POINTER(_L), _P]),
    "vp_slab_scatter_p2p": (_I, [_P, _P, _P, _P, _I, _L, _dp, _dp, _I, C.POINTER(_L), _P]),
    "vp_gather_rows": (_I, [_P, _P, _L, _P, _I, _P, _P]),
    "vp_build_fields": (_I, [_P, _P, _L, _P, _P, _I, _D, C.POINTER(_P), C.POINTER(_P), _P, _P, _P]),
    "vp_snapshot_preamble": (_I, [_P, _P, _P, _P, _I, _L, _I, _I, _dp, _dp, _P]),
    "vp_deposit_ngp": (_I, [_P, _P, _I, _L, _P, _I, _I, _D, _P, _P]),
    "vp_pk_plan_create": (_I, [_P, _I, _dp, _dp, _I, C.POINTER(_P)]),
    "vp_pk_plan_destroy": (_I, [_P]),
    "vp_pk_plan_create_dist": (_I, [_P, _I, _I, _I, _dp, _dp, _I, C.POINTER(_P)]),
    "vp_pk_dist_local": (_I, [_P, C.POINTER(_P), _I, C.POINTER(_P), _P]),
    "vp_pk_dist_final": (_I, [_P, C.POINTER(_P), _I, _P, _P, _P]),
    "vp_pk_dist_p2p_alloc": (_I, [_P, _I, C.c_char_p]),
    "vp_pk_dist_p2p_open": (_I, [_P, C.c_char_p]),
    "vp_pk_dist_p2p_close": (_I, [_P]),
    "vp_pk_dist_local_p2p": (_I, [_P, C.POINTER(_P), _I, _P]),
    "vp_pk_dist_final_p2p": (_I, [_P, _I, _P, _P, _P]),
    "vp_pk_fields": (_I, [_P, C.POINTER(_P), _I, _P, _P, _P]),
    "vp_fft_r2c_inplace": (_I, [_P, _P, _P]),
    "vp_fft_unpack_half": (_I, [_P, _P, _P, _P]),
    "vp_power_bin_full": (_I, [_P, _P, _P, _P, _P]),
    "vp_power_cube": (_I, [_P, C.POINTER(_P), _I, _P, _P]),
    "vp_fold_field": (_I, [_P, C.POINTER(_P), _I, _I, _I, C.POINTER(_I), _P, _P]),
    "vp_fold_power": (_I, [_P, _P, _I, _I, _P, _P]),
    "vp_k_magnitude": (_I, [_P, _dp, _dp, _dp, _I, _P, _P]),
    "vp_hist_weighted": (_I, [_P, _P, _P, _L, _dp, _I, _P, _P, _P]),
    "vp_sort_pairs": (_I, [_P, _P, _P, _L, _I, _P]),
    "vp_host_particles_to_pk": (_I, [_P, _P, _P, _P, _I, _L, _dp, _dp, _dp, _I, _D, _D, _dp, _dp, _I, _I, _I, _P, _P, _P]),
    "vp_dev_particles_to_pk": (_I, [_P, _P, _P, _P, _I, _L, _dp, _dp, _dp, _I, _D, _D, _dp, _dp, _I, _I, _I, _P, _P, _P]),
}

_lib = None
_lock = threading.Lock()
_ctx = {}


def load_library():
    """dlopen the shared object and declare every prototype.  No GPU needed for this step."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise VPowerError(f"{LIB_PATH} not found -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                                  "vpower_b200 has no CPU fallback")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        raise VPowerError(f"libvpower_b200 error {rc}: {load_library().vp_last_error().decode()}")


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise VPowerError("no CUDA device: vpower_b200 runs on B200 (sm_100a) only and has no CPU fallback")
    return torch


def ctx(device=None):
    torch = _torch()
    dev = torch.cuda.current_device() if device is None else int(device)
    if dev not in _ctx:
        h = _P()
        _check(load_library().vp_ctx_create(dev, C.byref(h)))
        _ctx[dev] = h
    return _ctx[dev]


def profile_enable(on=True):
    _check(load_library().vp_profile_enable(ctx(), 1 if on else 0))


def profile_report():
    import json
    buf = C.create_string_buffer(1 << 16)
    _check(load_library().vp_profile_report(ctx(), buf, len(buf)))
    return json.loads(buf.value.decode())


def launch_count():
    return int(load_library().vp_launch_count(ctx()))


def arena_bytes():
    """Bytes of device scratch the context holds (grow-only high-water mark of the arena)."""
    return int(load_library().vp_ctx_arena_bytes(ctx()))


def trim():
    """Give the context's scratch arena back to the driver (it is re-grown on demand).  Syncs the device."""
    _check(load_library().vp_ctx_trim(ctx()))


def stream_ptr():
    return _P(_torch().cuda.current_stream().cuda_stream)


def _dtype_code(t):
    torch = _torch()
    if t.dtype == torch.float32:
        return VP_F32
    if t.dtype == torch.float64:
        return VP_F64
    raise VPowerError(f"unsupported dtype {t.dtype}; use float32 or float64")


def _as_dp(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def to_device(a, dtype=None):
    """numpy / torch -> contiguous CUDA tensor (plumbing)."""
    torch = _torch()
    if isinstance(a, torch.Tensor):
        t = a
    else:
        arr = np.ascontiguousarray(a)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("="))
        t = torch.from_numpy(arr)
    if dtype is not None:
        t = t.to(dtype)
    return t.to("cuda", non_blocking=False).contiguous()


# --------------------------------------------------------------------------- K1
def nn_grid(pos_t, qx, qy, qz, opts: NNOpts | None = None):
    """pos_t: CUDA tensor [Np,3] f32/f64 -> CUDA int32 tensor [nx,ny,nz] (0-based nearest particle)."""
    torch = _torch()
    assert pos_t.is_cuda and pos_t.dim() == 2 and pos_t.shape[1] == 3 and pos_t.is_contiguous()
    qx_a, qx_p = _as_dp(qx)
    qy_a, qy_p = _as_dp(qy)
    qz_a, qz_p = _as_dp(qz)
    out = torch.empty((len(qx_a), len(qy_a), len(qz_a)), dtype=torch.int32, device=pos_t.device)
    _check(load_library().vp_nn_grid(ctx(), _P(pos_t.data_ptr()), _dtype_code(pos_t), pos_t.shape[0], qx_p, len(qx_a),
                                     qy_p, len(qy_a), qz_p, len(qz_a), _P(out.data_ptr()),
                                     C.byref(opts) if opts is not None else None, stream_ptr()))
    return out


def slab_bucket(pos_t, vel_t, rho_t, lo, hi):
    """Sharded input -> rows [n,8] = [x y z vx vy vz rho 0] grouped by destination rank + counts per rank (see vp_slab_bucket;
    rho = 0 when absent)."""
    torch = _torch()
    P = len(lo)
    n = pos_t.shape[0]
    w = SLAB_ROW
    lo_a, lop = _as_dp(lo)
    hi_a, hip = _as_dp(hi)
    cap = n + n // 2 + 1024          # halo duplicates; if that is too small the call reports the exact need in `counts`
    for attempt in range(2):
        rows = torch.empty((cap, w), dtype=pos_t.dtype, device=pos_t.device)
        counts = (_L * P)()
        rc = load_library().vp_slab_bucket(ctx(), _P(pos_t.data_ptr()), _P(vel_t.data_ptr()),
                                           _P(rho_t.data_ptr()) if rho_t is not None else None, _dtype_code(pos_t), n, lop, hip, P,
                                           _P(rows.data_ptr()), cap, counts, stream_ptr())
        if rc == 0:
            break
        need = sum(int(c) for c in counts)
        if attempt == 1 or need <= cap:          # a real error (CUDA, arguments), not the capacity
            _check(rc)
        cap = need                               # wide halos can send a particle to every rank: up to P * n rows
    counts = [int(c) for c in counts]
    return rows[:sum(counts)], counts


SLAB_ROW = 8          # elements per exchanged particle row (padded: whole 32-byte sectors, 16-byte aligned runs)


class _RawCuda:
    """Expose a raw device allocation of the library to torch (zero copy) through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class SlabExchangeP2P:
    """Sharded particle exchange with peer stores: receive buffers live in the library, mapped into every rank."""

    def __init__(self, nranks, rank, group=None):
        self.nranks, self.rank, self.group, self.cap_bytes = nranks, rank, group, 0

    def _ensure(self, need_bytes):
        """Collective: (re)allocate all receive buffers when any rank needs more room, then re-map the peers."""
        torch = _torch()
        import torch.distributed as dist
        if need_bytes <= self.cap_bytes:
            return
        cap = int(need_bytes * (1.15 if need_bytes < (8 << 30) else 1.03)) + (1 << 20)
        # an exported buffer may only be freed once no peer maps it any more: unmap everywhere, barrier, then re-allocate
        _check(load_library().vp_slab_p2p_close(ctx()))
        dist.barrier(group=self.group)
        buf = C.create_string_buffer(64)
        _check(load_library().vp_slab_p2p_alloc(ctx(), cap, buf))
        mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(mine) for _ in range(self.nranks)]
        dist.all_gather(allh, mine, group=self.group)
        raw = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        _check(load_library().vp_slab_p2p_open(ctx(), self.nranks, self.rank, raw))
        self.cap_bytes = cap

    def exchange(self, pos_t, vel_t, rho_t, lo, hi):
        """-> torch tensor [rows, 8] ALIASING this rank's receive buffer (all particles of its slab + halo).  It is valid
        until the next exchange() on this object (which overwrites, and may re-allocate, the buffer): clone it to keep it."""
        torch = _torch()
        import torch.distributed as dist
        P = self.nranks
        lo_a, lop = _as_dp(lo)
        hi_a, hip = _as_dp(hi)
        n = pos_t.shape[0]
        w = SLAB_ROW
        es = pos_t.element_size()
        counts = (_L * P)()
        _check(load_library().vp_slab_count(ctx(), _P(pos_t.data_ptr()), _dtype_code(pos_t), n, lop, hip, P, counts, stream_ptr()))
        mine = torch.tensor([int(c) for c in counts], dtype=torch.int64, device=pos_t.device)
        mat = [torch.empty_like(mine) for _ in range(P)]
        dist.all_gather(mat, mine, group=self.group)                  # mat[src][dst]
        M = torch.stack(mat).cpu().numpy()
        need = int(M.sum(axis=0).max()) * w * es                       # the fullest receive buffer, same on every rank
        self._ensure(need)
        first = (_L * P)(*[int(M[:self.rank, d].sum()) for d in range(P)])
        self.last_peer_bytes = int((M[self.rank].sum() - M[self.rank, self.rank]) * w * es)    # rows this rank stores into peers
        # barrier: every rank has finished reading its receive buffer from the previous exchange (stream ordered)
        tok = torch.zeros(1, device=pos_t.device)
        dist.all_reduce(tok, group=self.group)
        _check(load_library().vp_slab_scatter_p2p(ctx(), _P(pos_t.data_ptr()), _P(vel_t.data_ptr()),
                                                  _P(rho_t.data_ptr()) if rho_t is not None else None, _dtype_code(pos_t), n,
                                                  lop, hip, P, first, stream_ptr()))
        dist.all_reduce(tok, group=self.group)                         # all ranks have stored
        rows = int(M[:, self.rank].sum())
        ptr, nbytes = _P(), C.c_size_t()
        _check(load_library().vp_slab_p2p_buffer(ctx(), C.byref(ptr), C.byref(nbytes)))
        typestr = "<f8" if es == 8 else "<f4"
        return torch.as_tensor(_RawCuda(ptr.value, (rows, w), typestr), device=pos_t.device)


def nn_grid_payload(pos_t, vel_t, rho_t, qx, qy, qz, lcell3, want_idx=True, opts: NNOpts | None = None):
    """K1 with the payload sorted alongside: -> (nn_idx or None, nn_pos, srec[np,8] f32 = the cell-sorted 32-byte records:
    floats 0..3 the search half, floats 4..7 the payload (v'x, v'y, v'z, m)).
    With opts.row_stride > 0 the three tensors are column views of one interleaved row tensor."""
    torch = _torch()
    assert pos_t.is_cuda and vel_t.dtype == pos_t.dtype
    if opts is None or opts.row_stride == 0:
        assert pos_t.is_contiguous() and vel_t.is_contiguous()
    qx_a, qx_p = _as_dp(qx)
    qy_a, qy_p = _as_dp(qy)
    qz_a, qz_p = _as_dp(qz)
    shape = (len(qx_a), len(qy_a), len(qz_a))
    nn_idx = torch.empty(shape, dtype=torch.int32, device=pos_t.device) if want_idx else None
    nn_pos = torch.empty(shape, dtype=torch.int32, device=pos_t.device)
    spay = torch.empty((pos_t.shape[0], 8), dtype=torch.float32, device=pos_t.device)
    _check(load_library().vp_nn_grid_payload(
        ctx(), _P(pos_t.data_ptr()), _P(vel_t.data_ptr()), _P(rho_t.data_ptr()) if rho_t is not None else None,
        _dtype_code(pos_t), pos_t.shape[0], qx_p, shape[0], qy_p, shape[1], qz_p, shape[2], float(lcell3),
        _P(nn_idx.data_ptr()) if want_idx else None, _P(nn_pos.data_ptr()), _P(spay.data_ptr()),
        C.byref(opts) if opts is not None else None, stream_ptr()))
    return nn_idx, nn_pos, spay


def nn_grid_fields(pos_t, vel_t, rho_t, qx, qy, qz, lcell3, want_v=True, want_p=(False, False, False), want_e=False,
                   want_m=False, want_idx=False, opts: NNOpts | None = None):
    """K1 + K3 fused: -> (dict of float32 CUDA cubes vx,vy,vz,px,py,pz,e,m as requested, nn_idx or None)."""
    torch = _torch()
    assert pos_t.is_cuda and vel_t.dtype == pos_t.dtype
    if opts is None or opts.row_stride == 0:
        assert pos_t.is_contiguous() and vel_t.is_contiguous()
    qx_a, qx_p = _as_dp(qx)
    qy_a, qy_p = _as_dp(qy)
    qz_a, qz_p = _as_dp(qz)
    shape = (len(qx_a), len(qy_a), len(qz_a))
    out = {}

    def cube(name, on):
        if on:
            out[name] = torch.empty(shape, dtype=torch.float32, device=pos_t.device)
            return out[name].data_ptr()
        return None

    v = (_P * 3)(*[cube(nm, want_v) for nm in ("vx", "vy", "vz")])
    p = (_P * 3)(*[cube(nm, on) for nm, on in zip(("px", "py", "pz"), want_p)])
    e = cube("e", want_e)
    m = cube("m", want_m)
    nn_idx = torch.empty(shape, dtype=torch.int32, device=pos_t.device) if want_idx else None
    _check(load_library().vp_nn_grid_fields(
        ctx(), _P(pos_t.data_ptr()), _P(vel_t.data_ptr()), _P(rho_t.data_ptr()) if rho_t is not None else None,
        _dtype_code(pos_t), pos_t.shape[0], qx_p, shape[0], qy_p, shape[1], qz_p, shape[2], float(lcell3), v, p,
        _P(e) if e else None, _P(m) if m else None, _P(nn_idx.data_ptr()) if want_idx else None,
        C.byref(opts) if opts is not None else None, stream_ptr()))
    return out, nn_idx


def fields_sorted(nn_pos_t, spay_t, want_v=True, want_p=(False, False, False), want_e=False, want_m=False):
    """-> dict of float32 CUDA cubes (vx,vy,vz,px,py,pz,e,m as requested) from the sorted payload."""
    torch = _torch()
    shape = tuple(nn_pos_t.shape)
    out = {}

    def cube(name, on):
        if on:
            out[name] = torch.empty(shape, dtype=torch.float32, device=nn_pos_t.device)
            return out[name].data_ptr()
        return None

    v = (_P * 3)(*[cube(nm, want_v) for nm in ("vx", "vy", "vz")])
    p = (_P * 3)(*[cube(nm, on) for nm, on in zip(("px", "py", "pz"), want_p)])
    e = cube("e", want_e)
    m = cube("m", want_m)
    _check(load_library().vp_fields_sorted(ctx(), _P(nn_pos_t.data_ptr()), nn_pos_t.numel(), _P(spay_t.data_ptr()), v, p,
                                           _P(e) if e else None, _P(m) if m else None, stream_ptr()))
    return out


def nn_grid_plan(np_particles, qx, qy, qz, opts: NNOpts | None = None):
    """Host-only: the cell list vp_nn_grid would build (no device needed).  -> dict, see include/vpower_b200.h."""
    qx_a, qx_p = _as_dp(qx)
    qy_a, qy_p = _as_dp(qy)
    qz_a, qz_p = _as_dp(qz)
    out = (_L * 10)()
    _check(load_library().vp_nn_grid_plan(int(np_particles), qx_p, len(qx_a), qy_p, len(qy_a), qz_p, len(qz_a),
                                          C.byref(opts) if opts is not None else None, out))
    names = ("cells_x", "cells_y", "cells_z", "bucket_shift", "n_buckets", "r5", "r6", "r7", "scratch_MiB", "corner_aligned")
    return dict(zip(names, (int(v) for v in out)))


def fft_x_layout(N, kz_columns=None):
    """Host-only view of the blocked half-spectrum layout and the x pass's tensor map (vp_fft_x_layout)."""
    out = (_L * 24)()
    _check(load_library().vp_fft_x_layout(int(N), int(N // 2 if kz_columns is None else kz_columns), out))
    v = [int(x) for x in out]
    return {"C": v[0], "kyb": v[1], "tiles": v[2], "box_x": v[3], "boxes": v[4], "tma": bool(v[5]), "dims": v[6:11],
            "strides": v[11:15], "box": v[15:20], "box_slot_bytes": v[20], "ring_items": v[21], "smem_bytes": v[22], "threads": v[23]}


def nn_grid_stats():
    o = (_L * 5)()
    _check(load_library().vp_nn_grid_stats_ex(ctx(), o, stream_ptr()))
    return {"n_wide": int(o[0]), "n_unresolved": int(o[1]), "n_kept": int(o[2]), "n_stage_b": int(o[3]), "n_far": int(o[4])}


def gather_rows(idx_t, src_t):
    torch = _torch()
    src_t = src_t.contiguous()
    row = src_t[0].numel() * src_t.element_size() if src_t.dim() > 1 else src_t.element_size()
    n = idx_t.numel()
    out = torch.empty((n,) + tuple(src_t.shape[1:]), dtype=src_t.dtype, device=src_t.device)
    _check(load_library().vp_gather_rows(ctx(), _P(idx_t.data_ptr()), n, _P(src_t.data_ptr()), row, _P(out.data_ptr()),
                                         stream_ptr()))
    return out


def snapshot_preamble(pos_t, vel_t, mass_t, shift=True, bulk=True):
    """In place on CUDA tensors: pos -= min(pos) per axis, vel -= mass-weighted mean velocity.  -> (min[3], bulk[3])."""
    mn, bk = (C.c_double * 3)(), (C.c_double * 3)()
    ref = pos_t if pos_t is not None else vel_t
    _check(load_library().vp_snapshot_preamble(
        ctx(), _P(pos_t.data_ptr()) if pos_t is not None else None, _P(vel_t.data_ptr()) if vel_t is not None else None,
        _P(mass_t.data_ptr()) if mass_t is not None else None, _dtype_code(ref), ref.shape[0], 1 if shift else 0, 1 if bulk else 0,
        mn, bk, stream_ptr()))
    return np.array(list(mn)), np.array(list(bk))


# --------------------------------------------------------------------------- K3
def build_fields(nn_t, vel_t, rho_t, lcell3, want_v=True, want_p=(False, False, False), want_e=False, want_m=False):
    """-> dict of float32 CUDA cubes with the shape of nn_t: vx,vy,vz,px,py,pz,e,m (only the requested ones)."""
    torch = _torch()
    shape = tuple(nn_t.shape)
    n = nn_t.numel()
    out = {}

    def cube(name, on):
        if on:
            out[name] = torch.empty(shape, dtype=torch.float32, device=nn_t.device)
            return out[name].data_ptr()
        return None

    v = (_P * 3)(*[cube(nm, want_v) for nm in ("vx", "vy", "vz")])
    p = (_P * 3)(*[cube(nm, on) for nm, on in zip(("px", "py", "pz"), want_p)])
    e = cube("e", want_e)
    m = cube("m", want_m)
    if rho_t is not None:
        assert rho_t.dtype == vel_t.dtype
    _check(load_library().vp_build_fields(ctx(), _P(nn_t.data_ptr()), n, _P(vel_t.data_ptr()),
                                          _P(rho_t.data_ptr()) if rho_t is not None else None, _dtype_code(vel_t),
                                          float(lcell3), v, p, _P(e) if e else None, _P(m) if m else None, stream_ptr()))
    return out


# --------------------------------------------------------------------------- K2
def deposit_ngp(pos_t, w_t, N, Lbox):
    torch = _torch()
    w2 = w_t.to(torch.float64).contiguous()
    ncomp = 1 if w2.dim() == 1 else w2.shape[1]
    shape = (N, N, N) if w2.dim() == 1 else (N, N, N, ncomp)
    grid = torch.empty(shape, dtype=torch.float64, device=pos_t.device)
    _check(load_library().vp_deposit_ngp(ctx(), _P(pos_t.data_ptr()), _dtype_code(pos_t), pos_t.shape[0],
                                         _P(w2.data_ptr()), ncomp, N, float(Lbox), _P(grid.data_ptr()), stream_ptr()))
    return grid


# --------------------------------------------------------------------------- K4/K5
class PkPlan:
    """Geometry of one transform+binning: N, k axis table (2 pi fftfreq), bin edges."""

    def __init__(self, N, k_axis, edges, nranks=1, rank=0):
        self.N = int(N)
        self.k_axis, kp = _as_dp(k_axis)
        self.edges, ep = _as_dp(edges)
        self.nbins = len(self.edges) - 1
        self.nranks, self.rank = int(nranks), int(rank)
        assert len(self.k_axis) == self.N
        self._h = _P()
        _check(load_library().vp_pk_plan_create_dist(ctx(), self.N, self.nranks, self.rank, kp, ep, self.nbins,
                                                     C.byref(self._h)))

    def dist_local(self, slabs):
        """slabs: 1..3 float32 CUDA tensors [N/nranks, N, N] (overwritten) -> list of send buffers
        [nranks, N/nranks, N, N/2/nranks] complex64 (block d goes to rank d)."""
        torch = _torch()
        nx, kzc = self.N // self.nranks, self.N // 2 // self.nranks
        send = [torch.empty((self.nranks, nx, self.N, kzc), dtype=torch.complex64, device=slabs[0].device) for _ in slabs]
        fp = (_P * len(slabs))(*[s.data_ptr() for s in slabs])
        sp = (_P * len(slabs))(*[s.data_ptr() for s in send])
        _check(load_library().vp_pk_dist_local(self._h, fp, len(slabs), sp, stream_ptr()))
        return send

    def p2p_setup(self, group=None, ncomp_max=3):
        """Allocate the receive buffers in the plan, exchange their CUDA IPC handles, map every peer's buffers."""
        torch = _torch()
        import torch.distributed as dist
        buf = C.create_string_buffer(64 * ncomp_max)
        _check(load_library().vp_pk_dist_p2p_alloc(self._h, ncomp_max, buf))
        mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(mine) for _ in range(self.nranks)]
        dist.all_gather(allh, mine, group=group)
        raw = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        _check(load_library().vp_pk_dist_p2p_open(self._h, raw))
        self.p2p = True

    def dist_local_p2p(self, slabs):
        fp = (_P * len(slabs))(*[s.data_ptr() for s in slabs])
        _check(load_library().vp_pk_dist_local_p2p(self._h, fp, len(slabs), stream_ptr()))

    def dist_final_p2p(self, ncomp, device):
        torch = _torch()
        psum = torch.empty(self.nbins, dtype=torch.float64, device=device)
        ns = torch.empty(self.nbins, dtype=torch.int64, device=device)
        _check(load_library().vp_pk_dist_final_p2p(self._h, ncomp, _P(psum.data_ptr()), _P(ns.data_ptr()), stream_ptr()))
        return psum, ns

    def dist_final(self, recv):
        """recv: 1..3 complex64 CUDA tensors [N, N, N/2/nranks] -> partial (psum, nsample) CUDA tensors."""
        torch = _torch()
        rp = (_P * len(recv))(*[r.data_ptr() for r in recv])
        psum = torch.empty(self.nbins, dtype=torch.float64, device=recv[0].device)
        ns = torch.empty(self.nbins, dtype=torch.int64, device=recv[0].device)
        _check(load_library().vp_pk_dist_final(self._h, rp, len(recv), _P(psum.data_ptr()), _P(ns.data_ptr()), stream_ptr()))
        return psum, ns

    def close(self, group=None):
        """Collective teardown of a plan with peer-mapped receive buffers: unmap the peers on every rank, barrier, free."""
        if not self._h:
            return
        if getattr(self, "p2p", False) and self.nranks > 1:
            import torch.distributed as dist
            _check(load_library().vp_pk_dist_p2p_close(self._h))
            if dist.is_initialized():
                dist.barrier(group=group)
        load_library().vp_pk_plan_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            if self._h:
                load_library().vp_pk_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def fields(self, cubes):
        """cubes: list of 1..3 float32 CUDA tensors [N,N,N]; they are OVERWRITTEN.
        -> (sum_c |FFT(f_c)|^2 per shell as numpy f64 [nbins], Nsample numpy int64 [nbins])."""
        torch = _torch()
        for c in cubes:
            assert c.is_cuda and c.dtype == torch.float32 and c.is_contiguous() and tuple(c.shape) == (self.N,) * 3
        ptrs = (_P * len(cubes))(*[c.data_ptr() for c in cubes])
        psum = torch.empty(self.nbins, dtype=torch.float64, device=cubes[0].device)
        ns = torch.empty(self.nbins, dtype=torch.int64, device=cubes[0].device)
        _check(load_library().vp_pk_fields(self._h, ptrs, len(cubes), _P(psum.data_ptr()), _P(ns.data_ptr()), stream_ptr()))
        return psum.cpu().numpy(), ns.cpu().numpy()

    def power_cube(self, cubes):
        """-> CUDA f64 tensor [N,N,N] = sum_c |FFT(f_c)|^2 over the FULL spectrum (cubes are overwritten)."""
        torch = _torch()
        ptrs = (_P * len(cubes))(*[c.data_ptr() for c in cubes])
        P = torch.empty((self.N,) * 3, dtype=torch.float64, device=cubes[0].device)
        _check(load_library().vp_power_cube(self._h, ptrs, len(cubes), _P(P.data_ptr()), stream_ptr()))
        return P

    def bin_full(self, P_t):
        torch = _torch()
        P_t = P_t.to(torch.float64).contiguous()
        psum = torch.empty(self.nbins, dtype=torch.float64, device=P_t.device)
        ns = torch.empty(self.nbins, dtype=torch.int64, device=P_t.device)
        _check(load_library().vp_power_bin_full(self._h, _P(P_t.data_ptr()), _P(psum.data_ptr()), _P(ns.data_ptr()), stream_ptr()))
        return psum.cpu().numpy(), ns.cpu().numpy()

    def fft_half(self, cube):
        """Diagnostic: in-place r2c of one cube, returned as complex64 CUDA tensor [N,N,N/2+1]."""
        torch = _torch()
        _check(load_library().vp_fft_r2c_inplace(self._h, _P(cube.data_ptr()), stream_ptr()))
        half = torch.empty((self.N, self.N, self.N // 2 + 1, 2), dtype=torch.float32, device=cube.device)
        _check(load_library().vp_fft_unpack_half(self._h, _P(cube.data_ptr()), _P(half.data_ptr()), stream_ptr()))
        return torch.view_as_complex(half)


def fold_field(cubes, m, beta):
    """BoxField.fold on the device: cubes = 1..3 float32 CUDA tensors [N,N,N] (read only) ->
    complex128 CUDA tensor [N/m, N/m, N/m, ncomp] = fold_field(v * phase_beta, m) / m**1.5."""
    torch = _torch()
    N = int(cubes[0].shape[0])
    for c in cubes:
        assert c.is_cuda and c.dtype == torch.float32 and c.is_contiguous() and tuple(c.shape) == (N,) * 3
    if N % m:
        raise VPowerError(f"fold: Nsize={N} is not a multiple of the folding factor m={m}")
    n = N // m
    out = torch.empty((n, n, n, len(cubes), 2), dtype=torch.float64, device=cubes[0].device)
    ptrs = (_P * len(cubes))(*[c.data_ptr() for c in cubes])
    b = (_I * 3)(*[int(t) for t in beta])
    _check(load_library().vp_fold_field(ctx(), ptrs, len(cubes), N, int(m), b, _P(out.data_ptr()), stream_ptr()))
    return torch.view_as_complex(out)


def fold_power(folded_t):
    """sum_c |FFT_n(folded_c)|^2 of a complex128 CUDA tensor [n,n,n,ncomp] (or [n,n,n]) -> f64 CUDA tensor [n,n,n]."""
    torch = _torch()
    f = folded_t if folded_t.dim() == 4 else folded_t[..., None]
    f = torch.view_as_real(f.to(torch.complex128).contiguous())
    n, ncomp = int(f.shape[0]), int(f.shape[3])
    P = torch.empty((n, n, n), dtype=torch.float64, device=f.device)
    _check(load_library().vp_fold_power(ctx(), _P(f.data_ptr()), ncomp, n, _P(P.data_ptr()), stream_ptr()))
    return P


def k_magnitude(kx, ky, kz):
    """|k| = sqrt((kx*kx + ky*ky) + kz*kz) for the separable table, flattened C order (CUDA f64 tensor)."""
    torch = _torch()
    kx_a, kxp = _as_dp(kx)
    ky_a, kyp = _as_dp(ky)
    kz_a, kzp = _as_dp(kz)
    assert len(kx_a) == len(ky_a) == len(kz_a)
    n = len(kx_a)
    out = torch.empty(n ** 3, dtype=torch.float64, device="cuda")
    _check(load_library().vp_k_magnitude(ctx(), kxp, kyp, kzp, n, _P(out.data_ptr()), stream_ptr()))
    return out


def hist_weighted(k_t, w_t, edges):
    """numpy.histogram(k, bins=edges, weights=w) and the unweighted counts, on the device."""
    torch = _torch()
    k_t = k_t.to(torch.float64).contiguous()
    w_t = w_t.to(torch.float64).contiguous()
    e_a, ep = _as_dp(edges)
    nb = len(e_a) - 1
    psum = torch.empty(nb, dtype=torch.float64, device=k_t.device)
    ns = torch.empty(nb, dtype=torch.int64, device=k_t.device)
    _check(load_library().vp_hist_weighted(ctx(), _P(k_t.data_ptr()), _P(w_t.data_ptr()), k_t.numel(), ep, nb,
                                           _P(psum.data_ptr()), _P(ns.data_ptr()), stream_ptr()))
    return psum.cpu().numpy(), ns.cpu().numpy()


def sort_pairs(keys_t, vals_t, bits=32):
    _check(load_library().vp_sort_pairs(ctx(), _P(keys_t.data_ptr()), _P(vals_t.data_ptr()), keys_t.numel(), bits, stream_ptr()))


# --------------------------------------------------------------------------- whole path
def particles_to_pk(pos, vel, rho, qx, qy, qz, N, lcell3, norm, k_axis, edges, quantities=("velocity",),
                    momentum_strict=True):
    """Whole path in one C call.  pos/vel/rho: numpy arrays (HOST buffers; copied inside the call) or CUDA
    tensors (device resident).  -> dict quantity -> Psum[nbins] (normalised by `norm`), and Nsample."""
    lib = load_library()
    mask = sum({"velocity": 1, "momentum": 2, "energy": 4}[q] for q in quantities)
    qx_a, qxp = _as_dp(qx)
    qy_a, qyp = _as_dp(qy)
    qz_a, qzp = _as_dp(qz)
    k_a, kp = _as_dp(k_axis)
    e_a, ep = _as_dp(edges)
    nb = len(e_a) - 1
    psum = np.zeros((3, nb), dtype=np.float64)
    ns = np.zeros(nb, dtype=np.uint64)
    on_host = isinstance(pos, np.ndarray)
    if on_host:
        _torch()
        dt = VP_F64 if pos.dtype == np.float64 else VP_F32
        want = np.float64 if dt == VP_F64 else np.float32
        pos = np.ascontiguousarray(pos, dtype=want)
        vel = np.ascontiguousarray(vel, dtype=want)
        rho = None if rho is None else np.ascontiguousarray(rho, dtype=want)
        ptr = lambda a: _P(a.ctypes.data) if a is not None else None  # noqa: E731
        fn = lib.vp_host_particles_to_pk
    else:
        dt = _dtype_code(pos)
        ptr = lambda a: _P(a.data_ptr()) if a is not None else None  # noqa: E731
        fn = lib.vp_dev_particles_to_pk
    _check(fn(ctx(), ptr(pos), ptr(vel), ptr(rho), dt, pos.shape[0], qxp, qyp, qzp, int(N), float(lcell3), float(norm),
              kp, ep, nb, mask, 1 if momentum_strict else 0, _P(psum.ctypes.data), _P(ns.ctypes.data), stream_ptr()))
    out = {q: psum[i] for i, q in enumerate(("velocity", "momentum", "energy")) if q in quantities}
    return out, ns.astype(np.int64)
