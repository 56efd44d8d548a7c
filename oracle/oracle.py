"""TEST INFRASTRUCTURE -- ctypes wrappers for oracle/liboracle.so (the plain-C
restatement of the reference CPU hot path) and oracle/_ref/libdvren_ref.so
(the unmodified reference compiled from /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing here is product code.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
sys.path.insert(0, os.path.join(_REPO, "diff-volume-renderer_b200", "python"))
import hp_abi as A  # noqa: E402

ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libdvren_ref.so")


class orc_grid(C.Structure):
    _fields_ = [("data", C.c_void_p), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("channels", C.c_int32), ("interp", C.c_uint32), ("oob", C.c_uint32),
                ("wmin", C.c_float * 3), ("wmax", C.c_float * 3)]


class orc_render_out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("radiance", "transmittance", "opacity", "depth", "image",
                                          "trans", "opac", "depth_img", "hitmask", "sigma_grad",
                                          "color_grad")] + [("sample_count", C.c_uint64),
                                                            ("live_sample_count", C.c_uint64)]


class orc_render_shadow(C.Structure):
    _fields_ = [("box_o", C.c_int32 * 3), ("box_n", C.c_int32 * 3), ("sigma_sum", C.c_void_p),
                ("sigma_abs", C.c_void_p), ("color_sum", C.c_void_p), ("color_abs", C.c_void_p),
                ("misses", C.c_uint64)]


def build_oracle(force: bool = False) -> str:
    """Compile liboracle.so if missing (gcc, seconds)."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "dvren_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        f32p, u32p = C.c_void_p, C.c_void_p
        _lib.orc_plan_resolve.restype = C.c_int
        _lib.orc_plan_resolve.argtypes = [C.POINTER(A.hp_plan_desc)]
        _lib.orc_rays.restype = C.c_int
        _lib.orc_rays.argtypes = [C.POINTER(A.hp_plan_desc), f32p, f32p, f32p, f32p, u32p]
        _lib.orc_ray_sample_count.restype = C.c_uint32
        _lib.orc_ray_sample_count.argtypes = [C.POINTER(A.hp_plan_desc), C.c_float, C.c_float]
        _lib.orc_jitter.restype = C.c_float
        _lib.orc_jitter.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        _lib.orc_alpha.restype = C.c_float
        _lib.orc_alpha.argtypes = [C.c_float, C.c_float]
        _lib.orc_alpha_fast.restype = C.c_float
        _lib.orc_alpha_fast.argtypes = [C.c_float, C.c_float]
        _lib.orc_alpha_fast_mismatches.restype = C.c_uint64
        _lib.orc_alpha_fast_mismatches.argtypes = [C.c_float, C.c_float, C.c_uint32, C.POINTER(C.c_uint64)]
        _lib.orc_sample.restype = C.c_int
        _lib.orc_sample.argtypes = [C.POINTER(A.hp_plan_desc), C.POINTER(orc_grid), C.POINTER(orc_grid),
                                    C.c_size_t, f32p, f32p, f32p, f32p, C.c_uint64, C.c_size_t,
                                    f32p, f32p, f32p, f32p, u32p, C.POINTER(C.c_size_t)]
        _lib.orc_integrate.restype = C.c_int
        _lib.orc_integrate.argtypes = [C.POINTER(A.hp_plan_desc), C.c_size_t, C.c_size_t, f32p, f32p,
                                       f32p, u32p, f32p, f32p, f32p, f32p, f32p]
        _lib.orc_diff.restype = C.c_int
        _lib.orc_diff.argtypes = [C.c_size_t, C.c_size_t, f32p, C.c_int64, C.c_int64, f32p, f32p, u32p,
                                  f32p, f32p, f32p]
        _lib.orc_scatter.restype = C.c_int
        _lib.orc_scatter.argtypes = [C.POINTER(C.c_int32 * 3), C.POINTER(C.c_float * 3),
                                     C.POINTER(C.c_float * 3), C.c_uint32, C.c_uint32, C.c_size_t,
                                     f32p, f32p, f32p, f32p, f32p]
        _lib.orc_image.restype = C.c_int
        _lib.orc_image.argtypes = [C.POINTER(A.hp_plan_desc), C.c_size_t, u32p, f32p, f32p, f32p, f32p,
                                   f32p, f32p, f32p, f32p, u32p]
        _lib.orc_render.restype = C.c_int
        _lib.orc_render.argtypes = [C.POINTER(A.hp_plan_desc), C.POINTER(orc_grid), C.POINTER(orc_grid),
                                    C.c_uint64, f32p, C.POINTER(C.c_int32 * 3), C.POINTER(C.c_float * 3),
                                    C.POINTER(C.c_float * 3), C.POINTER(orc_render_out)]
        _lib.orc_render_shadowed.restype = C.c_int
        _lib.orc_render_shadowed.argtypes = _lib.orc_render.argtypes + [C.POINTER(orc_render_shadow)]
        _lib.orc_camera_grad.restype = C.c_int
        _lib.orc_camera_grad.argtypes = [C.POINTER(A.hp_plan_desc), C.POINTER(orc_grid),
                                         C.POINTER(orc_grid), C.c_uint64, f32p,
                                         C.POINTER(C.c_double * 16)]
        _lib.orc_camera_grad_mag.restype = C.c_int
        _lib.orc_camera_grad_mag.argtypes = _lib.orc_camera_grad.argtypes + [C.POINTER(C.c_double * 16)]
    return _lib


def _p(a: Optional[np.ndarray]):
    return a.ctypes.data if a is not None else None


def make_grid(data: np.ndarray, channels: int, interp=A.HP_INTERP_LINEAR, oob=A.HP_OOB_ZERO) -> orc_grid:
    """data: (nz,ny,nx) for sigma (channels=1) or (nz,ny,nx,3) for colour.  Keep `data` alive."""
    assert data.dtype == np.float32 and data.flags.c_contiguous
    g = orc_grid()
    g.data = data.ctypes.data
    g.nz, g.ny, g.nx = (int(v) for v in data.shape[:3])
    g.channels = channels
    g.interp, g.oob = int(interp), int(oob)
    for i in range(3):
        g.wmin[i], g.wmax[i] = 0.0, 1.0  # hp_runtime.cpp:289-294 hard-codes [0,1]^3
    return g


def plan_resolve(desc: A.hp_plan_desc):
    d = A.copy_desc(desc)
    st = lib().orc_plan_resolve(C.byref(d))
    return st, d


def rays(desc: A.hp_plan_desc) -> Dict[str, np.ndarray]:
    n = desc.roi.width * desc.roi.height
    r = {"origins": np.zeros((n, 3), np.float32), "directions": np.zeros((n, 3), np.float32),
         "t_near": np.zeros(n, np.float32), "t_far": np.zeros(n, np.float32),
         "pixel_ids": np.zeros(n, np.uint32)}
    st = lib().orc_rays(C.byref(desc), _p(r["origins"]), _p(r["directions"]), _p(r["t_near"]),
                        _p(r["t_far"]), _p(r["pixel_ids"]))
    assert st == 0, st
    return r


def sample(desc, gs, gc, r, capacity, ray_index_base=0):
    n = r["t_near"].shape[0]
    b = {"positions": np.zeros((capacity, 3), np.float32), "dt": np.zeros(capacity, np.float32),
         "ray_offset": np.zeros(n + 1, np.uint32), "sigma": np.zeros(capacity, np.float32),
         "color": np.zeros((capacity, 3), np.float32)}
    cnt = C.c_size_t(0)
    st = lib().orc_sample(C.byref(desc), C.byref(gs) if gs else None, C.byref(gc) if gc else None, n,
                          _p(r["origins"]), _p(r["directions"]), _p(r["t_near"]), _p(r["t_far"]),
                          ray_index_base, capacity, _p(b["positions"]), _p(b["dt"]), _p(b["sigma"]),
                          _p(b["color"]), _p(b["ray_offset"]), C.byref(cnt))
    if st != 0:
        return st, None
    m = cnt.value
    out = {k: (v if k == "ray_offset" else v[:m]) for k, v in b.items()}
    out["count"] = m
    return 0, out


def integrate(desc, samp):
    n = samp["ray_offset"].shape[0] - 1
    m = samp["count"]
    b = {"radiance": np.zeros((n, 3), np.float32), "transmittance": np.zeros(n, np.float32),
         "opacity": np.zeros(n, np.float32), "depth": np.zeros(n, np.float32),
         "aux": np.zeros((m, 4), np.float32)}
    st = lib().orc_integrate(C.byref(desc), n, m, _p(samp["dt"]), _p(samp["sigma"]), _p(samp["color"]),
                             _p(samp["ray_offset"]), _p(b["radiance"]), _p(b["transmittance"]),
                             _p(b["opacity"]), _p(b["depth"]), _p(b["aux"]))
    assert st == 0, st
    return b


def diff(dL_dI: np.ndarray, samp, intl):
    n = samp["ray_offset"].shape[0] - 1
    m = samp["count"]
    dl = np.ascontiguousarray(dL_dI, np.float32)
    g = {"sigma": np.zeros(m, np.float32), "color": np.zeros((m, 3), np.float32)}
    st = lib().orc_diff(n, m, _p(dl), dl.shape[1], 1, _p(samp["dt"]), _p(samp["color"]),
                        _p(samp["ray_offset"]), _p(intl["aux"]), _p(g["sigma"]), _p(g["color"]))
    assert st == 0, st
    return g


def scatter(res, bmin, bmax, interp, oob, positions, gsig, gcol):
    v = int(res[0]) * int(res[1]) * int(res[2])
    sg, cg = np.zeros(v, np.float32), np.zeros(v * 3, np.float32)
    r3 = (C.c_int32 * 3)(*[int(x) for x in res])
    lo = (C.c_float * 3)(*[float(x) for x in bmin])
    hi = (C.c_float * 3)(*[float(x) for x in bmax])
    pos = np.ascontiguousarray(positions, np.float32)
    a = np.ascontiguousarray(gsig, np.float32)
    b = np.ascontiguousarray(gcol, np.float32)
    st = lib().orc_scatter(C.byref(r3), C.byref(lo), C.byref(hi), interp, oob, pos.shape[0], _p(pos),
                           _p(a), _p(b), _p(sg), _p(cg))
    assert st == 0, st
    return sg, cg


def image(desc, r, intl):
    h, w = desc.height, desc.width
    b = {"image": np.zeros((h, w, 3), np.float32), "trans": np.zeros((h, w), np.float32),
         "opacity": np.zeros((h, w), np.float32), "depth": np.zeros((h, w), np.float32),
         "hitmask": np.zeros((h, w), np.uint32)}
    n = intl["transmittance"].shape[0]
    st = lib().orc_image(C.byref(desc), n, _p(r["pixel_ids"]), _p(intl["radiance"]),
                         _p(intl["transmittance"]), _p(intl["opacity"]), _p(intl["depth"]),
                         _p(b["image"]), _p(b["trans"]), _p(b["opacity"]), _p(b["depth"]), _p(b["hitmask"]))
    return st, b


def render(desc, gs, gc, dL_dI=None, res=None, bmin=(0, 0, 0), bmax=(1, 1, 1), ray_index_base=0,
           per_ray=True, frames=True, shadow=False, shadow_box=None):
    """Whole path; returns dict with per-ray, per-pixel and (if dL_dI) grid-gradient arrays.

    shadow=True (needs dL_dI) adds the float64 shadow of the scatter (orc_render_shadowed): `sigma_sum`,
    `color_sum` = the reference's own float32 terms accumulated in double, `sigma_abs`, `color_abs` = the
    sums of their magnitudes, over `shadow_box` = (x0, y0, z0, nx, ny, nz) (default: the whole grid),
    flattened x-fastest like the gradient arrays; `shadow_misses` counts contributions outside the box."""
    n = desc.roi.width * desc.roi.height
    h, w = desc.height, desc.width
    out = orc_render_out()
    o: Dict[str, np.ndarray] = {}
    if per_ray:
        o.update(radiance=np.zeros((n, 3), np.float32), transmittance=np.zeros(n, np.float32),
                 ray_opacity=np.zeros(n, np.float32), ray_depth=np.zeros(n, np.float32))
        out.radiance, out.transmittance = _p(o["radiance"]), _p(o["transmittance"])
        out.opacity, out.depth = _p(o["ray_opacity"]), _p(o["ray_depth"])
    if frames:
        o.update(image=np.zeros((h, w, 3), np.float32), trans=np.zeros((h, w), np.float32),
                 opacity=np.zeros((h, w), np.float32), depth=np.zeros((h, w), np.float32),
                 hitmask=np.zeros((h, w), np.uint32))
        out.image, out.trans, out.opac = _p(o["image"]), _p(o["trans"]), _p(o["opacity"])
        out.depth_img, out.hitmask = _p(o["depth"]), _p(o["hitmask"])
    dl = None
    r3 = lo = hi = None
    if dL_dI is not None:
        dl = np.ascontiguousarray(dL_dI, np.float32).reshape(n, 3)
        if res is None:
            g = gs if gs else gc
            res = (g.nx, g.ny, g.nz)
        v = int(res[0]) * int(res[1]) * int(res[2])
        o.update(sigma_grad=np.zeros(v, np.float32), color_grad=np.zeros(v * 3, np.float32))
        out.sigma_grad, out.color_grad = _p(o["sigma_grad"]), _p(o["color_grad"])
    if res is None:
        res = (1, 1, 1)
    r3 = (C.c_int32 * 3)(*[int(x) for x in res])
    lo = (C.c_float * 3)(*[float(x) for x in bmin])
    hi = (C.c_float * 3)(*[float(x) for x in bmax])
    sh = None
    if shadow and dl is not None:
        box = tuple(int(v) for v in (shadow_box if shadow_box is not None else (0, 0, 0) + tuple(res)))
        bv = box[3] * box[4] * box[5]
        sh = orc_render_shadow()
        for i in range(3):
            sh.box_o[i], sh.box_n[i] = box[i], box[3 + i]
        o.update(sigma_sum=np.zeros(bv, np.float64), sigma_abs=np.zeros(bv, np.float64),
                 color_sum=np.zeros(bv * 3, np.float64), color_abs=np.zeros(bv * 3, np.float64))
        sh.sigma_sum, sh.sigma_abs = _p(o["sigma_sum"]), _p(o["sigma_abs"])
        sh.color_sum, sh.color_abs = _p(o["color_sum"]), _p(o["color_abs"])
        o["shadow_box"] = box
    st = lib().orc_render_shadowed(C.byref(desc), C.byref(gs) if gs else None, C.byref(gc) if gc else None,
                                   ray_index_base, _p(dl), C.byref(r3), C.byref(lo), C.byref(hi), C.byref(out),
                                   C.byref(sh) if sh is not None else None)
    if sh is not None:
        o["shadow_misses"] = int(sh.misses)
    o["status"] = st
    o["sample_count"] = int(out.sample_count)
    o["live_sample_count"] = int(out.live_sample_count)
    return o


def camera_grad(desc, gs, gc, dL_dI, ray_index_base=0, with_mag=False):
    """Analytic camera adjoint d/d c2w[12], d/d {fx,fy,cx,cy}; with_mag=True also returns the per-output upper
    bound of the sum of |terms| (orc_camera_grad_mag)."""
    dl = np.ascontiguousarray(dL_dI, np.float32)
    out = (C.c_double * 16)()
    mag = (C.c_double * 16)()
    st = lib().orc_camera_grad_mag(C.byref(desc), C.byref(gs), C.byref(gc), ray_index_base, _p(dl), C.byref(out),
                                   C.byref(mag) if with_mag else None)
    assert st == 0, st
    if with_mag:
        return np.array(list(out), dtype=np.float64), np.array(list(mag), dtype=np.float64)
    return np.array(list(out), dtype=np.float64)


# ---------------------------------------------------------------------------
# the compiled, unmodified reference (oracle/_ref)
# ---------------------------------------------------------------------------
_ref = None


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def ref_lib() -> C.CDLL:
    """The reference through its own hp.h ABI (+ the ref_shim door to its C++ classes)."""
    global _ref
    if _ref is None:
        _ref = A.bind(C.CDLL(REF_SO))
        f32p = C.c_void_p
        _ref.ref_scatter.restype = C.c_int
        _ref.ref_scatter.argtypes = [C.POINTER(C.c_int32 * 3), C.POINTER(C.c_float * 3),
                                     C.POINTER(C.c_float * 3), C.c_uint32, C.c_uint32, C.c_size_t,
                                     f32p, f32p, f32p, f32p, f32p]
        _ref.ref_render.restype = C.c_int
        _ref.ref_render.argtypes = [C.POINTER(A.hp_plan_desc), C.POINTER(C.c_int32 * 3), f32p, f32p,
                                    C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.c_uint32,
                                    C.c_uint32, C.c_int, f32p, f32p, f32p, f32p, f32p, C.c_void_p, f32p,
                                    f32p, f32p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _ref.ref_worker_create.restype = C.c_void_p
        _ref.ref_worker_create.argtypes = [C.POINTER(C.c_int32 * 3), f32p, f32p, C.c_uint32, C.c_uint32]
        _ref.ref_worker_destroy.restype = None
        _ref.ref_worker_destroy.argtypes = [C.c_void_p]
        _ref.ref_worker_run.restype = C.c_int
        _ref.ref_worker_run.argtypes = [C.c_void_p, C.POINTER(A.hp_plan_desc), f32p, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return _ref


class RefWorker:
    """One host thread's handle on the unmodified reference's hot-path calls (oracle/ref_shim.cpp: ref_worker_*):
    hp_ray -> hp_samp_int_fused -> hp_img, then hp_diff -> DenseGridField::AccumulateSampleGradients, for ROI bands of a
    plan.  sigma / color are SHARED between workers (the reference's fields view them); each worker owns the gradient
    vectors its scatter writes (4 floats per voxel, plus the reference field's own zero-filled value copy)."""

    def __init__(self, sigma: np.ndarray, color: np.ndarray, interp=A.HP_INTERP_LINEAR, oob=A.HP_OOB_ZERO):
        assert sigma.dtype == np.float32 and color.dtype == np.float32 and sigma.flags.c_contiguous and color.flags.c_contiguous
        self._keep = (sigma, color)
        nz, ny, nx = sigma.shape
        r3 = (C.c_int32 * 3)(nx, ny, nz)
        self.handle = ref_lib().ref_worker_create(C.byref(r3), _p(sigma), _p(color), interp, oob)
        if not self.handle:
            raise RuntimeError("ref_worker_create failed")

    def run(self, desc, dL_dI=None):
        """Returns (sample_count, forward_ms, backward_ms) of one band (`desc` = unresolved plan descriptor with its ROI)."""
        dl = np.ascontiguousarray(dL_dI, np.float32) if dL_dI is not None else None
        n, f, b = C.c_uint64(0), C.c_double(0), C.c_double(0)
        rc = ref_lib().ref_worker_run(self.handle, C.byref(desc), _p(dl), C.byref(n), C.byref(f), C.byref(b))
        if rc != 0:
            raise RuntimeError(f"ref_worker_run -> {rc}")
        return n.value, f.value, b.value

    def close(self):
        if self.handle:
            ref_lib().ref_worker_destroy(self.handle)
            self.handle = None


def ref_scatter(res, bmin, bmax, interp, oob, positions, gsig, gcol):
    v = int(res[0]) * int(res[1]) * int(res[2])
    sg, cg = np.zeros(v, np.float32), np.zeros(v * 3, np.float32)
    r3 = (C.c_int32 * 3)(*[int(x) for x in res])
    lo = (C.c_float * 3)(*[float(x) for x in bmin])
    hi = (C.c_float * 3)(*[float(x) for x in bmax])
    pos = np.ascontiguousarray(positions, np.float32)
    a = np.ascontiguousarray(gsig, np.float32)
    b = np.ascontiguousarray(gcol, np.float32)
    st = ref_lib().ref_scatter(C.byref(r3), C.byref(lo), C.byref(hi), interp, oob, pos.shape[0],
                               _p(pos), _p(a), _p(b), _p(sg), _p(cg))
    assert st == 0, st
    return sg, cg


def ref_render(desc, sigma: np.ndarray, color: np.ndarray, dL_dI=None, interp=A.HP_INTERP_LINEAR,
               oob=A.HP_OOB_ZERO, bmin=(0, 0, 0), bmax=(1, 1, 1), use_fused=True):
    """dvren::Renderer::Forward/Backward of the unmodified reference.
    sigma: (nz,ny,nx), color: (nz,ny,nx,3).  `desc` is the UNRESOLVED plan descriptor."""
    nz, ny, nx = sigma.shape
    h, w = desc.height, desc.width
    v = nx * ny * nz
    o = {"image": np.zeros((h, w, 3), np.float32), "trans": np.zeros((h, w), np.float32),
         "opacity": np.zeros((h, w), np.float32), "depth": np.zeros((h, w), np.float32),
         "hitmask": np.zeros((h, w), np.uint32)}
    dl = None
    if dL_dI is not None:
        dl = np.ascontiguousarray(dL_dI, np.float32)
        o.update(sigma_grad=np.zeros(v, np.float32), color_grad=np.zeros(v * 3, np.float32),
                 camera_grad=np.zeros(12, np.float32))
    r3 = (C.c_int32 * 3)(nx, ny, nz)
    lo = (C.c_float * 3)(*[float(x) for x in bmin])
    hi = (C.c_float * 3)(*[float(x) for x in bmax])
    rays_n, samp_n = C.c_uint64(0), C.c_uint64(0)
    fms, bms = C.c_double(0), C.c_double(0)
    sig = np.ascontiguousarray(sigma, np.float32)
    col = np.ascontiguousarray(color, np.float32)
    st = ref_lib().ref_render(C.byref(desc), C.byref(r3), _p(sig), _p(col), C.byref(lo), C.byref(hi),
                              interp, oob, 1 if use_fused else 0, _p(dl), _p(o["image"]), _p(o["trans"]),
                              _p(o["opacity"]), _p(o["depth"]), _p(o["hitmask"]), _p(o.get("sigma_grad")),
                              _p(o.get("color_grad")), _p(o.get("camera_grad")), C.byref(rays_n),
                              C.byref(samp_n), C.byref(fms), C.byref(bms))
    o.update(status=st, ray_count=rays_n.value, sample_count=samp_n.value, forward_ms=fms.value,
             backward_ms=bms.value)
    return o
