"""numpy-level driver for any shared library that speaks include/hotpath/hp.h
with HOST tensors (the product library stages them to the GPU; the compiled
reference under oracle/_ref runs its CPU code).  The calls mirror the staged
pipeline of the reference's Renderer (src/render/renderer.cpp:259-365,415):
hp_ray -> hp_samp -> hp_int -> hp_img -> hp_diff, plus hp_samp_int_fused.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

import hp_abi as A


class HpError(RuntimeError):
    def __init__(self, what: str, status: int):
        super().__init__(f"{what} -> {A.STATUS_NAMES[status] if 0 <= status < 6 else status}")
        self.status = status


def _check(what: str, status: int):
    if status != A.HP_STATUS_SUCCESS:
        raise HpError(what, status)


class HpHostPipeline:
    """Owns ctx/plan/field handles of one library and runs the ABI on numpy buffers."""

    def __init__(self, lib: C.CDLL, ctx_desc: Optional[A.hp_ctx_desc] = None):
        self.lib = lib
        self._keep = []
        self.ctx = C.c_void_p()
        _check("hp_ctx_create", lib.hp_ctx_create(C.byref(ctx_desc) if ctx_desc else None, C.byref(self.ctx)))
        self._plans, self._fields = [], []

    def close(self):
        for f in self._fields:
            self.lib.hp_field_release(f)
        for p in self._plans:
            self.lib.hp_plan_release(p)
        if self.ctx:
            self.lib.hp_ctx_release(self.ctx)
        self._fields, self._plans, self.ctx = [], [], C.c_void_p()

    # -- handles ---------------------------------------------------------
    def plan(self, desc: A.hp_plan_desc):
        h = C.c_void_p()
        _check("hp_plan_create", self.lib.hp_plan_create(self.ctx, C.byref(desc), C.byref(h)))
        self._plans.append(h)
        out = A.hp_plan_desc()
        _check("hp_plan_get_desc", self.lib.hp_plan_get_desc(h, C.byref(out)))
        return h, out

    def sigma_field(self, grid: np.ndarray, interp=A.HP_INTERP_LINEAR, oob=A.HP_OOB_ZERO):
        grid = np.ascontiguousarray(grid, dtype=np.float32)
        self._keep.append(grid)
        t = A.host_tensor(grid)
        h = C.c_void_p()
        _check("hp_field_create_grid_sigma",
               self.lib.hp_field_create_grid_sigma(self.ctx, C.byref(t), interp, oob, C.byref(h)))
        self._fields.append(h)
        return h

    def color_field(self, grid: np.ndarray, interp=A.HP_INTERP_LINEAR, oob=A.HP_OOB_ZERO):
        grid = np.ascontiguousarray(grid, dtype=np.float32)
        self._keep.append(grid)
        t = A.host_tensor(grid)
        h = C.c_void_p()
        _check("hp_field_create_grid_color",
               self.lib.hp_field_create_grid_color(self.ctx, C.byref(t), interp, oob, C.byref(h)))
        self._fields.append(h)
        return h

    # -- stages ----------------------------------------------------------
    @staticmethod
    def _rays_struct(r: Dict[str, np.ndarray]) -> A.hp_rays_t:
        s = A.hp_rays_t()
        s.origins = A.host_tensor(r["origins"])
        s.directions = A.host_tensor(r["directions"])
        s.t_near = A.host_tensor(r["t_near"])
        s.t_far = A.host_tensor(r["t_far"])
        s.pixel_ids = A.host_tensor(r["pixel_ids"])
        return s

    def ray(self, plan, n_rays: int, override: Optional[Dict[str, np.ndarray]] = None):
        r = {"origins": np.zeros((n_rays, 3), np.float32), "directions": np.zeros((n_rays, 3), np.float32),
             "t_near": np.zeros(n_rays, np.float32), "t_far": np.zeros(n_rays, np.float32),
             "pixel_ids": np.zeros(n_rays, np.uint32)}
        s = self._rays_struct(r)
        ov = self._rays_struct(override) if override is not None else None
        _check("hp_ray", self.lib.hp_ray(plan, C.byref(ov) if ov is not None else None, C.byref(s), None, 0))
        n = int(s.t_near.shape[0])
        return {k: v[:n] for k, v in r.items()}

    @staticmethod
    def _samp_buffers(capacity: int, n_rays: int):
        return {"positions": np.zeros((capacity, 3), np.float32), "dt": np.zeros(capacity, np.float32),
                "ray_offset": np.zeros(n_rays + 1, np.uint32), "sigma": np.zeros(capacity, np.float32),
                "color": np.zeros((capacity, 3), np.float32)}

    @staticmethod
    def _samp_struct(b) -> A.hp_samp_t:
        s = A.hp_samp_t()
        for k in ("positions", "dt", "ray_offset", "sigma", "color"):
            setattr(s, k, A.host_tensor(b[k]))
        return s

    @staticmethod
    def _trim_samp(b, s: A.hp_samp_t):
        m = int(s.dt.shape[0])
        out = {k: (v if k == "ray_offset" else v[:m]) for k, v in b.items()}
        out["count"] = m
        return out

    def samp(self, plan, fs, fc, rays, capacity: int):
        n = rays["t_near"].shape[0]
        b = self._samp_buffers(capacity, n)
        s = self._samp_struct(b)
        rs = self._rays_struct(rays)
        st = self.lib.hp_samp(plan, fs, fc, C.byref(rs), C.byref(s), None, 0)
        _check("hp_samp", st)
        return self._trim_samp(b, s)

    @staticmethod
    def _intl_buffers(n_rays: int, m: int):
        return {"radiance": np.zeros((n_rays, 3), np.float32), "transmittance": np.zeros(n_rays, np.float32),
                "opacity": np.zeros(n_rays, np.float32), "depth": np.zeros(n_rays, np.float32),
                "aux": np.zeros((m, 4), np.float32)}

    @staticmethod
    def _intl_struct(b) -> A.hp_intl_t:
        s = A.hp_intl_t()
        for k in ("radiance", "transmittance", "opacity", "depth", "aux"):
            setattr(s, k, A.host_tensor(b[k]))
        return s

    def integrate(self, plan, samp):
        n = samp["ray_offset"].shape[0] - 1
        m = samp["count"]
        sb = {k: samp[k] for k in ("positions", "dt", "ray_offset", "sigma", "color")}
        ss = self._samp_struct(sb)
        b = self._intl_buffers(n, m)
        s = self._intl_struct(b)
        _check("hp_int", self.lib.hp_int(plan, C.byref(ss), C.byref(s), None, 0))
        return b

    def fused(self, plan, fs, fc, rays, capacity: int):
        """hp_samp_int_fused the way the reference's Renderer calls it (renderer.cpp:276-311): every
        output pointer NULL, one workspace sized for samples (capacity) + integrator outputs."""
        n = rays["t_near"].shape[0]
        need = capacity * 32 + (n + 1) * 4 + n * 24 + capacity * 16
        ws = np.zeros(need + 16, np.uint8)
        ss, istr = A.hp_samp_t(), A.hp_intl_t()
        rs = self._rays_struct(rays)
        _check("hp_samp_int_fused",
               self.lib.hp_samp_int_fused(plan, fs, fc, C.byref(rs), C.byref(ss), C.byref(istr),
                                          ws.ctypes.data, need))
        self._keep_ws = ws
        samp = {k: A.host_array(getattr(ss, k)) for k in ("positions", "dt", "ray_offset", "sigma", "color")}
        samp["count"] = int(ss.dt.shape[0])
        intl = {k: A.host_array(getattr(istr, k)) for k in ("radiance", "transmittance", "opacity", "depth", "aux")}
        return samp, intl

    def img(self, plan, desc: A.hp_plan_desc, intl, rays):
        h, w = desc.height, desc.width
        b = {"image": np.zeros((h, w, 3), np.float32), "trans": np.zeros((h, w), np.float32),
             "opacity": np.zeros((h, w), np.float32), "depth": np.zeros((h, w), np.float32),
             "hitmask": np.zeros((h, w), np.uint32)}
        s = A.hp_img_t()
        for k in b:
            setattr(s, k, A.host_tensor(b[k]))
        istr = self._intl_struct(intl)
        rs = self._rays_struct(rays)
        _check("hp_img", self.lib.hp_img(plan, C.byref(istr), C.byref(rs), C.byref(s), None, 0))
        return b

    def diff(self, plan, dL_dI: np.ndarray, samp, intl):
        m = samp["count"]
        sb = {k: samp[k] for k in ("positions", "dt", "ray_offset", "sigma", "color")}
        ss = self._samp_struct(sb)
        istr = self._intl_struct(intl)
        g = {"sigma": np.zeros(m, np.float32), "color": np.zeros((m, 3), np.float32),
             "camera": np.zeros((3, 4), np.float32)}
        gs = A.hp_grads_t()
        for k in g:
            setattr(gs, k, A.host_tensor(g[k]))
        dl = np.ascontiguousarray(dL_dI, dtype=np.float32)
        t = A.host_tensor(dl)
        _check("hp_diff", self.lib.hp_diff(plan, C.byref(t), C.byref(ss), C.byref(istr), C.byref(gs), None, 0))
        return g
