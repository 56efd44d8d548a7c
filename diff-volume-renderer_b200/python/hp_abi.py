"""ctypes mirror of include/hotpath/hp.h (the drop-in C ABI; reference
hotpath/include/hotpath/hp.h:24-216).  Pure declarations: no compute, no
fallback.  Used by the tests, bench.py and the oracle/reference wrappers, all
of which speak this same ABI to different shared libraries.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

HP_STATUS_SUCCESS = 0
HP_STATUS_INVALID_ARGUMENT = 1
HP_STATUS_OUT_OF_MEMORY = 2
HP_STATUS_NOT_IMPLEMENTED = 3
HP_STATUS_UNSUPPORTED = 4
HP_STATUS_INTERNAL_ERROR = 5
STATUS_NAMES = ["success", "invalid_argument", "out_of_memory", "not_implemented", "unsupported",
                "internal_error"]

HP_MEMSPACE_HOST, HP_MEMSPACE_DEVICE = 0, 1
HP_DTYPE_F16, HP_DTYPE_BF16, HP_DTYPE_F32, HP_DTYPE_I32, HP_DTYPE_U32 = range(5)
HP_CAMERA_PINHOLE, HP_CAMERA_ORTHOGRAPHIC = 0, 1
HP_SAMPLING_FIXED, HP_SAMPLING_STRATIFIED = 0, 1
HP_INTERP_NEAREST, HP_INTERP_LINEAR = 0, 1
HP_OOB_ZERO, HP_OOB_CLAMP = 0, 1


class hp_version(C.Structure):
    _fields_ = [("major", C.c_uint32), ("minor", C.c_uint32), ("patch", C.c_uint32)]


class hp_sampling_desc(C.Structure):
    _fields_ = [("dt", C.c_float), ("max_steps", C.c_uint32), ("mode", C.c_int)]


class hp_tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int), ("memspace", C.c_int),
                ("rank", C.c_uint32), ("shape", C.c_int64 * 8), ("stride", C.c_int64 * 8)]


class hp_ctx_desc(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("preferred_device", C.c_char_p), ("reserved", C.c_void_p)]


class hp_camera_desc(C.Structure):
    _fields_ = [("model", C.c_int), ("K", C.c_float * 9), ("c2w", C.c_float * 12),
                ("ortho_scale", C.c_float)]


class hp_roi_desc(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32)]


class hp_plan_desc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("t_near", C.c_float),
                ("t_far", C.c_float), ("max_rays", C.c_uint32), ("max_samples", C.c_uint32),
                ("seed", C.c_uint64), ("camera", hp_camera_desc), ("roi", hp_roi_desc),
                ("sampling", hp_sampling_desc)]


class hp_rays_t(C.Structure):
    _fields_ = [(n, hp_tensor) for n in ("origins", "directions", "t_near", "t_far", "pixel_ids")]


class hp_samp_t(C.Structure):
    _fields_ = [(n, hp_tensor) for n in ("positions", "dt", "ray_offset", "sigma", "color")]


class hp_intl_t(C.Structure):
    _fields_ = [(n, hp_tensor) for n in ("radiance", "transmittance", "opacity", "depth", "aux")]


class hp_img_t(C.Structure):
    _fields_ = [(n, hp_tensor) for n in ("image", "trans", "opacity", "depth", "hitmask")]


class hp_grads_t(C.Structure):
    _fields_ = [(n, hp_tensor) for n in ("sigma", "color", "camera")]


ABI_STRUCTS = [hp_version, hp_sampling_desc, hp_tensor, hp_ctx_desc, hp_camera_desc, hp_roi_desc,
               hp_plan_desc, hp_rays_t, hp_samp_t, hp_intl_t, hp_img_t, hp_grads_t]

# name -> (restype, argtypes); the 21 symbols of hp.h
P = C.POINTER
ABI_FUNCTIONS = {
    "hp_get_version": (hp_version, []),
    "hp_ctx_create": (C.c_int, [P(hp_ctx_desc), P(C.c_void_p)]),
    "hp_ctx_release": (None, [C.c_void_p]),
    "hp_ctx_get_desc": (C.c_int, [C.c_void_p, P(hp_ctx_desc)]),
    "hp_plan_create": (C.c_int, [C.c_void_p, P(hp_plan_desc), P(C.c_void_p)]),
    "hp_plan_release": (None, [C.c_void_p]),
    "hp_plan_get_desc": (C.c_int, [C.c_void_p, P(hp_plan_desc)]),
    "hp_ray": (C.c_int, [C.c_void_p, P(hp_rays_t), P(hp_rays_t), C.c_void_p, C.c_size_t]),
    "hp_samp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, P(hp_rays_t), P(hp_samp_t),
                          C.c_void_p, C.c_size_t]),
    "hp_int": (C.c_int, [C.c_void_p, P(hp_samp_t), P(hp_intl_t), C.c_void_p, C.c_size_t]),
    "hp_img": (C.c_int, [C.c_void_p, P(hp_intl_t), P(hp_rays_t), P(hp_img_t), C.c_void_p,
                         C.c_size_t]),
    "hp_diff": (C.c_int, [C.c_void_p, P(hp_tensor), P(hp_samp_t), P(hp_intl_t), P(hp_grads_t),
                          C.c_void_p, C.c_size_t]),
    "hp_field_create_grid_sigma": (C.c_int, [C.c_void_p, P(hp_tensor), C.c_uint32, C.c_uint32,
                                             P(C.c_void_p)]),
    "hp_field_create_grid_color": (C.c_int, [C.c_void_p, P(hp_tensor), C.c_uint32, C.c_uint32,
                                             P(C.c_void_p)]),
    "hp_field_create_hash_mlp": (C.c_int, [C.c_void_p, P(hp_tensor), P(C.c_void_p)]),
    "hp_field_release": (None, [C.c_void_p]),
    "hp_samp_int_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, P(hp_rays_t),
                                    P(hp_samp_t), P(hp_intl_t), C.c_void_p, C.c_size_t]),
    "hp_graph_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                  C.c_size_t, C.c_size_t, P(C.c_void_p)]),
    "hp_graph_capture": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, P(hp_tensor)]),
    "hp_graph_execute": (C.c_int, [C.c_void_p, P(hp_rays_t), P(hp_samp_t), P(hp_intl_t),
                                   P(hp_img_t), P(hp_grads_t)]),
    "hp_graph_release": (None, [C.c_void_p]),
}


def bind(lib: C.CDLL, table=None) -> C.CDLL:
    """Attach restype/argtypes for every ABI symbol; raises if one is missing."""
    for name, (res, args) in (table or ABI_FUNCTIONS).items():
        fn = getattr(lib, name)  # AttributeError => missing export, fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


_NP_DTYPE = {np.dtype(np.float32): HP_DTYPE_F32, np.dtype(np.uint32): HP_DTYPE_U32,
             np.dtype(np.int32): HP_DTYPE_I32, np.dtype(np.float16): HP_DTYPE_F16}


def contiguous_strides(shape: Sequence[int]):
    strides, acc = [], 1
    for dim in reversed(shape):
        strides.append(acc)
        acc *= int(dim)
    return list(reversed(strides))


def make_tensor(ptr: Optional[int], dtype: int, memspace: int, shape: Sequence[int],
                strides: Optional[Sequence[int]] = None) -> hp_tensor:
    t = hp_tensor()
    t.data = ptr
    t.dtype = dtype
    t.memspace = memspace
    t.rank = len(shape)
    st = list(strides) if strides is not None else contiguous_strides(shape)
    for i, (s, k) in enumerate(zip(shape, st)):
        t.shape[i] = int(s)
        t.stride[i] = int(k)
    return t


def host_tensor(arr: np.ndarray) -> hp_tensor:
    """View of a numpy array (element strides).  The caller keeps `arr` alive."""
    itemsize = arr.dtype.itemsize
    return make_tensor(arr.ctypes.data, _NP_DTYPE[arr.dtype], HP_MEMSPACE_HOST, arr.shape,
                       [s // itemsize for s in arr.strides])


def empty_tensor(memspace: int) -> hp_tensor:
    """Output slot with .data == NULL: the callee allocates it from the workspace."""
    t = hp_tensor()
    t.memspace = memspace
    return t


def tensor_shape(t: hp_tensor):
    return tuple(int(t.shape[i]) for i in range(t.rank))


def host_array(t: hp_tensor) -> np.ndarray:
    """Copy a HOST tensor view (contiguous) into a numpy array."""
    shape = tensor_shape(t)
    dt = {HP_DTYPE_F32: np.float32, HP_DTYPE_U32: np.uint32, HP_DTYPE_I32: np.int32}[t.dtype]
    n = int(np.prod(shape)) if shape else 0
    if n == 0 or not t.data:
        return np.zeros(shape, dtype=dt)
    buf = (C.c_byte * (n * np.dtype(dt).itemsize)).from_address(t.data)
    return np.frombuffer(buf, dtype=dt).reshape(shape).copy()


def make_plan_desc(width, height, t_near, t_far, dt=0.0, max_steps=0, mode=HP_SAMPLING_FIXED,
                   K=None, c2w=None, model=HP_CAMERA_PINHOLE, ortho_scale=0.0, roi=None,
                   max_rays=0, max_samples=0, seed=0) -> hp_plan_desc:
    d = hp_plan_desc()
    d.width, d.height = int(width), int(height)
    d.t_near, d.t_far = float(t_near), float(t_far)
    d.max_rays, d.max_samples, d.seed = int(max_rays), int(max_samples), int(seed)
    d.camera.model = int(model)
    d.camera.ortho_scale = float(ortho_scale)
    if K is not None:
        for i, v in enumerate(np.asarray(K, dtype=np.float32).reshape(9)):
            d.camera.K[i] = float(v)
    if c2w is not None:
        for i, v in enumerate(np.asarray(c2w, dtype=np.float32).reshape(12)):
            d.camera.c2w[i] = float(v)
    if roi is not None:
        d.roi.x, d.roi.y, d.roi.width, d.roi.height = (int(v) for v in roi)
    d.sampling.dt = float(dt)
    d.sampling.max_steps = int(max_steps)
    d.sampling.mode = int(mode)
    return d


def copy_desc(d: hp_plan_desc) -> hp_plan_desc:
    out = hp_plan_desc()
    C.memmove(C.byref(out), C.byref(d), C.sizeof(hp_plan_desc))
    return out
