"""ctypes front end of the product library libdvren_hp.so (hp.h + hp_b200.h).

No fallback of any kind lives here: if the library is missing, fails to load or
reports no CUDA device, the call raises.  Reference-side counterparts of the
classes below: dvren::Context / Plan / DenseGridField / Renderer
(reference include/dvren/**), which libdvren.so re-implements in C++ on the
same C ABI; this module is the Python door used by tests and bench.py.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

import hp_abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.dirname(_HERE)
LIB_PATH = os.path.join(PKG_DIR, "libdvren_hp.so")

HPX_CTX_EXT_MAGIC = 0x42323030
HPX_CTX_EXT2_MAGIC = 0x42323031
HPX_COMM_ID_BYTES = 128
HPX_SHARD_RESULT_OWNED, HPX_SHARD_RESULT_REPLICATED = 1, 2
HPX_ORDER_COLUMNS = 4
HPX_BACKWARD_GRID, HPX_BACKWARD_CAMERA, HPX_BACKWARD_ZERO = 1, 2, 4
HPX_BACKWARD_SCATTER_PER_RAY, HPX_BACKWARD_SCATTER_MERGED, HPX_BACKWARD_DETERMINISTIC = 0x10, 0x20, 0x40


class hpx_ctx_ext(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("device_ordinal", C.c_int32), ("stream", C.c_void_p)]


class hpx_ctx_ext2(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("device_ordinal", C.c_int32), ("stream", C.c_void_p),
                ("reserve_sms", C.c_uint32), ("flags", C.c_uint32)]


class hpx_counts(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("samples", C.c_uint64), ("live_samples", C.c_uint64)]


P = C.POINTER
f3 = C.c_float * 3
HPX_FUNCTIONS = {
    "hpx_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "hpx_ctx_device": (C.c_int, [C.c_void_p, P(C.c_int32), P(C.c_void_p)]),
    "hpx_last_error": (C.c_char_p, []),
    "hpx_ctx_mark": (C.c_int, [C.c_void_p, C.c_uint32]),
    "hpx_ctx_elapsed_ms": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, P(C.c_float)]),
    "hpx_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "hpx_device_alloc": (C.c_int, [C.c_void_p, C.c_size_t, P(C.c_void_p)]),
    "hpx_device_free": (None, [C.c_void_p, C.c_void_p]),
    "hpx_copy_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "hpx_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "hpx_host_unregister": (None, [C.c_void_p, C.c_void_p]),
    "hpx_grid_accumulate_samples": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "hpx_grid_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, P(f3), P(f3), P(C.c_void_p)]),
    "hpx_grid_create_raw": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_uint32, C.c_uint32, P(f3), P(f3), P(C.c_void_p)]),
    "hpx_grid_adopt_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hpx_grid_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "hpx_grid_zero_grad": (C.c_int, [C.c_void_p]),
    "hpx_grid_set_storage": (C.c_int, [C.c_void_p, C.c_uint32]),
    "hpx_grid_build_occupancy": (C.c_int, [C.c_void_p, C.c_int32, P(C.c_float), P(C.c_float)]),
    "hpx_grid_set_occupancy": (C.c_int, [C.c_void_p, C.c_int32]),
    "hpx_grid_grad_buffer": (C.c_int, [C.c_void_p, P(C.c_void_p), P(C.c_size_t)]),
    "hpx_grid_read_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "hpx_grid_read_grad_range": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "hpx_grid_release": (None, [C.c_void_p]),
    "hpx_frame_create": (C.c_int, [C.c_void_p, P(C.c_void_p)]),
    "hpx_frame_bytes": (C.c_size_t, [C.c_void_p]),
    "hpx_plan_step_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, P(C.c_uint32)]),
    "hpx_frame_set_view": (C.c_int, [C.c_void_p, P(A.hp_camera_desc), C.c_uint64, C.c_uint64]),
    "hpx_forward": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpx_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32]),
    "hpx_backward_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, P(C.c_uint32)]),
    "hpx_backward_streamed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hpx_frame_set_interleave": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "hpx_frame_set_row_order": (C.c_int, [C.c_void_p, C.c_int32]),
    "hpx_tile_row_order": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int32, P(C.c_uint32)]),
    "hpx_tile_order": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, P(C.c_uint32), P(C.c_uint32)]),
    "hpx_frame_bounds": (C.c_int, [C.c_void_p, C.c_void_p, P(C.c_int32 * 6)]),
    "hpx_backward_box": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, P(C.c_int32 * 6)]),
    "hpx_frame_box_misses": (C.c_int, [C.c_void_p, P(C.c_uint32)]),
    "hpx_backward_signalled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, P(C.c_uint32), C.c_uint32,
                                         P(C.c_void_p), P(C.c_uint32)]),
    "hpx_stream_wait_counter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "hpx_frame_reset_group_counters": (C.c_int, [C.c_void_p, P(C.c_void_p)]),
    "hpx_grid_add_box": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, P(C.c_int32 * 6)]),
    "hpx_grid_set_grad_layout": (C.c_int, [C.c_void_p, C.c_int32, P(C.c_size_t), P(C.c_int32)]),
    "hpx_frame_image": (C.c_int, [C.c_void_p, P(A.hp_img_t)]),
    "hpx_frame_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hpx_frame_counts": (C.c_int, [C.c_void_p, P(hpx_counts)]),
    "hpx_frame_cube_samples": (C.c_int, [C.c_void_p, C.c_void_p, P(C.c_uint64)]),
    "hpx_grid_touched_voxels": (C.c_int, [C.c_void_p, P(C.c_uint64)]),
    "hpx_frame_capture": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "hpx_frame_replay": (C.c_int, [C.c_void_p]),
    "hpx_frame_grad_input": (C.c_int, [C.c_void_p, P(C.c_void_p)]),
    "hpx_frame_release": (None, [C.c_void_p]),
    "hpx_ctx_sm_counts": (C.c_int, [C.c_void_p, P(C.c_uint32), P(C.c_uint32)]),
    "hpx_comm_unique_id": (C.c_int, [C.c_void_p]),
    "hpx_comm_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, P(C.c_void_p)]),
    "hpx_comm_release": (None, [C.c_void_p]),
    "hpx_comm_info": (C.c_int, [C.c_void_p, P(C.c_int32), P(C.c_int32), P(C.c_int32)]),
    "hpx_comm_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "hpx_grid_allreduce_grad": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hpx_shard_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, P(C.c_float), C.c_uint32, P(C.c_void_p)]),
    "hpx_shard_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "hpx_shard_frame": (C.c_int, [C.c_void_p, P(C.c_void_p)]),
    "hpx_shard_set_reduce": (C.c_int, [C.c_void_p, C.c_int32]),
    "hpx_shard_layout": (C.c_int, [C.c_void_p, P(C.c_int32), P(C.c_uint32), P(C.c_uint32), P(C.c_int32)]),
    "hpx_shard_release": (None, [C.c_void_p]),
    "hpx_shard_create_bands": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, P(C.c_void_p)]),
    "hpx_shard_set_result": (C.c_int, [C.c_void_p, C.c_uint32]),
    "hpx_shard_rebalance": (C.c_int, [C.c_void_p, P(C.c_int32)]),
    "hpx_shard_exchange_is_direct": (C.c_int, [C.c_void_p, P(C.c_int32)]),
    "hpx_shard_tile_order": (C.c_int, [C.c_void_p, P(C.c_int32)]),
    "hpx_shard_tune_order": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, P(C.c_int32)]),
    "hpx_plan_best_tile_order": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, P(C.c_int32), P(C.c_double)]),
    "hpx_plan_owner_cuts": (C.c_int, [C.c_uint32, C.c_int32, P(C.c_int32), C.c_uint32, P(C.c_int32)]),
    "hpx_plan_balanced_bands": (C.c_int, [C.c_void_p, C.c_uint32, P(C.c_uint32), P(C.c_uint32), P(C.c_double)]),
    "hpx_shard_bands": (C.c_int, [C.c_void_p, P(C.c_uint32), P(C.c_uint32), P(C.c_int32), P(C.c_int32), P(C.c_size_t), P(C.c_size_t)]),
    "hpx_shard_owned": (C.c_int, [C.c_void_p, P(C.c_void_p), P(C.c_int32), P(C.c_int32), P(C.c_size_t), P(C.c_int32)]),
}

_lib = None


def load(path: Optional[str] = None) -> C.CDLL:
    """Load and bind the product library.  Raises if it is not built."""
    global _lib
    if _lib is None or path is not None:
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise RuntimeError(f"{p} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(p)
        A.bind(lib)
        A.bind(lib, HPX_FUNCTIONS)
        if path is not None:
            return lib
        _lib = lib
    return _lib


class DvrenError(RuntimeError):
    def __init__(self, what: str, status: int, detail: str = ""):
        name = A.STATUS_NAMES[status] if 0 <= status < len(A.STATUS_NAMES) else str(status)
        super().__init__(f"{what} -> {name}" + (f" ({detail})" if detail else ""))
        self.status = status


def check(what: str, status: int):
    if status != A.HP_STATUS_SUCCESS:
        detail = load().hpx_last_error()
        raise DvrenError(what, status, detail.decode() if detail else "")


def _vec3(v):
    return None if v is None else f3(*[float(x) for x in v])


class Context:
    """hp_ctx bound to a device ordinal and (optionally) a caller stream."""

    def __init__(self, device: int = -1, stream: int = 0, reserve_sms: int = 0):
        """reserve_sms > 0: the context's kernels stay off that many SMs (green context; the library then owns the
        stream -- `stream` must be 0 -- and `self.stream` returns it)."""
        self.lib = load()
        if reserve_sms:
            self._ext = hpx_ctx_ext2(HPX_CTX_EXT2_MAGIC, device, None, reserve_sms, 0)
        else:
            self._ext = hpx_ctx_ext(HPX_CTX_EXT_MAGIC, device, stream or None)
        desc = A.hp_ctx_desc(0, None, C.cast(C.pointer(self._ext), C.c_void_p))
        self.handle = C.c_void_p()
        check("hp_ctx_create", self.lib.hp_ctx_create(C.byref(desc), C.byref(self.handle)))

    def synchronize(self):
        check("hpx_ctx_synchronize", self.lib.hpx_ctx_synchronize(self.handle))

    @property
    def stream(self) -> int:
        """cudaStream_t the context enqueues on (creates the device state on first use)."""
        dev, st = C.c_int32(), C.c_void_p()
        check("hpx_ctx_device", self.lib.hpx_ctx_device(self.handle, C.byref(dev), C.byref(st)))
        return st.value or 0

    def sm_counts(self):
        """(SMs this context's kernels may use, SMs of the GPU)."""
        a, b = C.c_uint32(), C.c_uint32()
        check("hpx_ctx_sm_counts", self.lib.hpx_ctx_sm_counts(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def mark(self, slot: int):
        check("hpx_ctx_mark", self.lib.hpx_ctx_mark(self.handle, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        check("hpx_ctx_elapsed_ms", self.lib.hpx_ctx_elapsed_ms(self.handle, a, b, C.byref(ms)))
        return ms.value

    def wait_counter(self, device_counter_ptr: int, value: int):
        """This context's stream waits (without occupying an SM) until *device_counter_ptr >= value."""
        check("hpx_stream_wait_counter", self.lib.hpx_stream_wait_counter(self.handle, int(device_counter_ptr), int(value)))

    def close(self):
        if self.handle:
            self.lib.hp_ctx_release(self.handle)
            self.handle = C.c_void_p()


class Plan:
    def __init__(self, ctx: Context, desc: A.hp_plan_desc):
        self.ctx, self.lib = ctx, ctx.lib
        self.handle = C.c_void_p()
        check("hp_plan_create", self.lib.hp_plan_create(ctx.handle, C.byref(desc), C.byref(self.handle)))
        self.desc = A.hp_plan_desc()
        check("hp_plan_get_desc", self.lib.hp_plan_get_desc(self.handle, C.byref(self.desc)))

    @property
    def n_rays(self) -> int:
        return self.desc.roi.width * self.desc.roi.height

    def close(self):
        if self.handle:
            self.lib.hp_plan_release(self.handle)
            self.handle = C.c_void_p()


class Grid:
    """hpx_grid: packed {r,g,b,sigma} device grid + gradient block."""

    def __init__(self, ctx: Context, sigma, color, interp=A.HP_INTERP_LINEAR, oob=A.HP_OOB_ZERO,
                 bbox_min=None, bbox_max=None, device_shape=None):
        """sigma (nz,ny,nx) / color (nz,ny,nx,3): numpy arrays (HOST), or raw device pointers
        (ints) together with device_shape=(nx,ny,nz)."""
        self.ctx, self.lib = ctx, ctx.lib
        self.handle = C.c_void_p()
        lo, hi = _vec3(bbox_min), _vec3(bbox_max)
        if device_shape is None:
            sig = np.ascontiguousarray(sigma, np.float32)
            col = np.ascontiguousarray(color, np.float32)
            nz, ny, nx = sig.shape
            assert col.shape == (nz, ny, nx, 3)
            ms, ps, pc = A.HP_MEMSPACE_HOST, sig.ctypes.data, col.ctypes.data
        else:
            nx, ny, nz = device_shape
            ms, ps, pc = A.HP_MEMSPACE_DEVICE, int(sigma), int(color)
        self.shape = (nx, ny, nz)
        self.voxels = nx * ny * nz
        check("hpx_grid_create_raw",
              self.lib.hpx_grid_create_raw(ctx.handle, nx, ny, nz, ps, pc, ms, interp, oob,
                                           C.byref(lo) if lo else None, C.byref(hi) if hi else None,
                                           C.byref(self.handle)))

    def zero_grad(self):
        check("hpx_grid_zero_grad", self.lib.hpx_grid_zero_grad(self.handle))

    def set_storage(self, half: bool):
        """Store the packed values as four halfs per voxel (8 B) instead of four floats (16 B); arithmetic stays fp32."""
        check("hpx_grid_set_storage", self.lib.hpx_grid_set_storage(self.handle, 1 if half else 0))

    def build_occupancy(self, enable: bool = True):
        """Empty-space skipping: (fraction of bricks the forward can skip, fraction the backward can skip)."""
        a, b = C.c_float(), C.c_float()
        check("hpx_grid_build_occupancy", self.lib.hpx_grid_build_occupancy(self.handle, 1 if enable else 0, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_occupancy(self, enable: bool):
        check("hpx_grid_set_occupancy", self.lib.hpx_grid_set_occupancy(self.handle, 1 if enable else 0))

    def set_grad_layout(self, slow_axis: int):
        """Make axis 0 = x / 1 = y / 2 = z the slowest one of the gradient block; returns (floats per slab, slabs)."""
        f, n = C.c_size_t(), C.c_int32()
        check("hpx_grid_set_grad_layout", self.lib.hpx_grid_set_grad_layout(self.handle, slow_axis, C.byref(f), C.byref(n)))
        return f.value, n.value

    def add_box(self, stream_ctx: "Context", box_device_ptr: int, box):
        """gradient[box] += box buffer, box buffer = 0, on stream_ctx's stream."""
        b = (C.c_int32 * 6)(*box)
        check("hpx_grid_add_box", self.lib.hpx_grid_add_box(stream_ctx.handle, self.handle, int(box_device_ptr), C.byref(b)))

    def grad_buffer(self):
        ptr, n = C.c_void_p(), C.c_size_t()
        check("hpx_grid_grad_buffer", self.lib.hpx_grid_grad_buffer(self.handle, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def read_grad(self):
        sg = np.zeros(self.voxels, np.float32)
        cg = np.zeros(self.voxels * 3, np.float32)
        cam = np.zeros(16, np.float32)
        check("hpx_grid_read_grad", self.lib.hpx_grid_read_grad(self.handle, sg.ctypes.data, cg.ctypes.data,
                                                                cam.ctypes.data, A.HP_MEMSPACE_HOST))
        return sg, cg, cam

    def touched_voxels(self) -> int:
        """Voxels whose gradient is non-zero (what the backward passes since the last zero have written)."""
        n = C.c_uint64()
        check("hpx_grid_touched_voxels", self.lib.hpx_grid_touched_voxels(self.handle, C.byref(n)))
        return n.value

    def update(self, sigma=None, color=None):
        s = np.ascontiguousarray(sigma, np.float32) if sigma is not None else None
        c = np.ascontiguousarray(color, np.float32) if color is not None else None
        check("hpx_grid_update", self.lib.hpx_grid_update(self.handle, s.ctypes.data if s is not None else None,
                                                          c.ctypes.data if c is not None else None,
                                                          A.HP_MEMSPACE_HOST))

    def close(self):
        if self.handle:
            self.lib.hpx_grid_release(self.handle)
            self.handle = C.c_void_p()


class Frame:
    """hpx_frame: per-plan device workspace; forward / backward / graph replay."""

    def __init__(self, plan: Plan):
        self.plan, self.lib = plan, plan.lib
        self.handle = C.c_void_p()
        check("hpx_frame_create", self.lib.hpx_frame_create(plan.handle, C.byref(self.handle)))

    def set_view(self, camera: Optional[A.hp_camera_desc], seed: int, ray_index_base: int = 0):
        check("hpx_frame_set_view", self.lib.hpx_frame_set_view(
            self.handle, C.byref(camera) if camera is not None else None, seed, ray_index_base))

    def forward(self, grid: Grid):
        check("hpx_forward", self.lib.hpx_forward(self.handle, grid.handle))

    def backward(self, grid: Grid, dL_dI, flags=HPX_BACKWARD_GRID | HPX_BACKWARD_ZERO, device=False):
        if device:
            ptr, ms = int(dL_dI), A.HP_MEMSPACE_DEVICE
        else:
            self._g = np.ascontiguousarray(dL_dI, np.float32)
            assert self._g.size == self.plan.n_rays * 3
            ptr, ms = self._g.ctypes.data, A.HP_MEMSPACE_HOST
        check("hpx_backward", self.lib.hpx_backward(self.handle, grid.handle, ptr, ms, flags))

    def set_interleave(self, stride: int, phase: int):
        check("hpx_frame_set_interleave", self.lib.hpx_frame_set_interleave(self.handle, stride, phase))

    def bounds(self, grid: Grid):
        """(x0, y0, z0, nx, ny, nz): box of voxels this frame's backward can touch."""
        box = (C.c_int32 * 6)()
        check("hpx_frame_bounds", self.lib.hpx_frame_bounds(self.handle, grid.handle, C.byref(box)))
        return tuple(int(v) for v in box)

    def backward_box(self, grid: Grid, dL_dI_device_ptr: int, box, box_device_ptr: int,
                     flags=HPX_BACKWARD_GRID | HPX_BACKWARD_ZERO):
        b = (C.c_int32 * 6)(*box)
        check("hpx_backward_box", self.lib.hpx_backward_box(self.handle, grid.handle, int(dL_dI_device_ptr),
                                                           A.HP_MEMSPACE_DEVICE, flags, int(box_device_ptr), C.byref(b)))

    def reset_group_counters(self) -> int:
        p = C.c_void_p()
        check("hpx_frame_reset_group_counters", self.lib.hpx_frame_reset_group_counters(self.handle, C.byref(p)))
        return p.value

    def backward_signalled(self, grid: Grid, dL_dI_device_ptr: int, group_end_rows, flags=HPX_BACKWARD_GRID):
        """One backward launch with a completion counter per row group; returns (device pointer of the counters,
        expected counts)."""
        n = len(group_end_rows)
        ends = (C.c_uint32 * n)(*[int(v) for v in group_end_rows])
        expected = (C.c_uint32 * n)()
        counters = C.c_void_p()
        check("hpx_backward_signalled", self.lib.hpx_backward_signalled(
            self.handle, grid.handle, int(dL_dI_device_ptr), A.HP_MEMSPACE_DEVICE, flags, ends, n, C.byref(counters), expected))
        return counters.value, [int(v) for v in expected]

    def box_misses(self) -> int:
        n = C.c_uint32()
        check("hpx_frame_box_misses", self.lib.hpx_frame_box_misses(self.handle, C.byref(n)))
        return n.value

    def scatter_mode(self, grid: Grid, flags: int = HPX_BACKWARD_GRID) -> str:
        out = C.c_uint32()
        check("hpx_backward_scatter", self.lib.hpx_backward_scatter(self.handle, grid.handle, flags, C.byref(out)))
        return "merged" if out.value == HPX_BACKWARD_SCATTER_MERGED else "per_ray"

    def read(self):
        d = self.plan.desc
        h, w = d.height, d.width
        o = {"image": np.zeros((h, w, 3), np.float32), "trans": np.zeros((h, w), np.float32),
             "opacity": np.zeros((h, w), np.float32), "depth": np.zeros((h, w), np.float32),
             "hitmask": np.zeros((h, w), np.uint32)}
        check("hpx_frame_read", self.lib.hpx_frame_read(self.handle, o["image"].ctypes.data, o["trans"].ctypes.data,
                                                        o["opacity"].ctypes.data, o["depth"].ctypes.data,
                                                        o["hitmask"].ctypes.data))
        return o

    def counts(self):
        c = hpx_counts()
        check("hpx_frame_counts", self.lib.hpx_frame_counts(self.handle, C.byref(c)))
        return {"rays": c.rays, "samples": c.samples, "live_samples": c.live_samples}

    def cube_samples(self, grid: Grid) -> int:
        """Live samples of the last forward that lie inside the unit cube (gather 8 corners / scatter 8 reds)."""
        n = C.c_uint64()
        check("hpx_frame_cube_samples", self.lib.hpx_frame_cube_samples(self.handle, grid.handle, C.byref(n)))
        return n.value

    def capture(self, grid: Grid, backward_flags: int = 0):
        check("hpx_frame_capture", self.lib.hpx_frame_capture(self.handle, grid.handle, backward_flags))

    def replay(self):
        check("hpx_frame_replay", self.lib.hpx_frame_replay(self.handle))

    def grad_input_ptr(self) -> int:
        p = C.c_void_p()
        check("hpx_frame_grad_input", self.lib.hpx_frame_grad_input(self.handle, C.byref(p)))
        return p.value

    def image_ptrs(self):
        v = A.hp_img_t()
        check("hpx_frame_image", self.lib.hpx_frame_image(self.handle, C.byref(v)))
        return v

    def bytes(self) -> int:
        return self.lib.hpx_frame_bytes(self.handle)

    def close(self):
        if self.handle:
            self.lib.hpx_frame_release(self.handle)
            self.handle = C.c_void_p()


def comm_unique_id() -> bytes:
    """128-byte NCCL rendezvous id (rank 0 makes it, the others receive it by any means)."""
    buf = (C.c_uint8 * HPX_COMM_ID_BYTES)()
    check("hpx_comm_unique_id", load().hpx_comm_unique_id(buf))
    return bytes(buf)


class Comm:
    """hpx_comm: one rank of the NCCL communicator behind the C ABI (world 1 = no NCCL)."""

    def __init__(self, ctx: Context, unique_id: Optional[bytes], rank: int, world: int, max_ctas: int = 0):
        self.ctx, self.lib, self.rank, self.world = ctx, ctx.lib, rank, world
        self.handle = C.c_void_p()
        ident = (C.c_uint8 * HPX_COMM_ID_BYTES).from_buffer_copy(unique_id) if unique_id else None
        check("hpx_comm_create", self.lib.hpx_comm_create(ctx.handle, ident, rank, world, max_ctas, C.byref(self.handle)))

    def nccl_version(self) -> int:
        v = C.c_int32()
        check("hpx_comm_info", self.lib.hpx_comm_info(self.handle, None, None, C.byref(v)))
        return v.value

    def allreduce_grad(self, grid: Grid):
        check("hpx_grid_allreduce_grad", self.lib.hpx_grid_allreduce_grad(self.handle, grid.handle))

    def allreduce(self, device_ptr: int, floats: int):
        check("hpx_comm_allreduce", self.lib.hpx_comm_allreduce(self.handle, int(device_ptr), floats))

    def close(self):
        if self.handle:
            self.lib.hpx_comm_release(self.handle)
            self.handle = C.c_void_p()


class _BorrowedFrame(Frame):
    """The frame a shard owns (never released from Python)."""

    def __init__(self, plan: Plan, handle):
        self.plan, self.lib, self.handle = plan, plan.lib, handle

    def close(self):
        self.handle = C.c_void_p()


class Shard:
    """hpx_shard: ONE frame rendered by all ranks of a communicator.

    bands=None   interleaved tile rows, slab all-reduces hidden behind a device-signalled backward (hpx_shard_create)
    bands="replicated" | "owned"   contiguous work-balanced bands with a sparse point-to-point exchange
                 (hpx_shard_create_bands); "owned": every rank ends with the finished sum of the slabs it owns."""

    RESULTS = {"owned": HPX_SHARD_RESULT_OWNED, "replicated": HPX_SHARD_RESULT_REPLICATED}

    def __init__(self, comm: Comm, plan: Plan, grid: Grid, group_weights=(1.0,), bands: Optional[str] = None):
        self.comm, self.lib, self.plan, self.grid, self.bands = comm, comm.lib, plan, grid, bands
        self.handle = C.c_void_p()
        if bands is None:
            w = (C.c_float * len(group_weights))(*[float(v) for v in group_weights])
            check("hpx_shard_create", self.lib.hpx_shard_create(comm.handle, plan.handle, grid.handle, w, len(group_weights),
                                                                C.byref(self.handle)))
        else:
            check("hpx_shard_create_bands", self.lib.hpx_shard_create_bands(comm.handle, plan.handle, grid.handle,
                                                                            self.RESULTS[bands], C.byref(self.handle)))
        fh = C.c_void_p()
        check("hpx_shard_frame", self.lib.hpx_shard_frame(self.handle, C.byref(fh)))
        self.frame = _BorrowedFrame(plan, fh) if fh.value else None

    def step(self, dL_dI_device_ptr: int, flags: int = HPX_BACKWARD_GRID | HPX_BACKWARD_ZERO):
        check("hpx_shard_step", self.lib.hpx_shard_step(self.handle, int(dL_dI_device_ptr), flags))

    def tune_order(self, dL_dI_device_ptr: int, flags: int = HPX_BACKWARD_GRID) -> int:
        """Measured choice of the band's tile dispatch order (rank-local; the gradient block is left dirty)."""
        order = C.c_int32()
        check("hpx_shard_tune_order", self.lib.hpx_shard_tune_order(self.handle, int(dL_dI_device_ptr), flags, C.byref(order)))
        return order.value

    def rebalance(self) -> bool:
        """Collective: re-cut the bands from the measured time of the last step; True when they moved (self.frame is then
        a new frame)."""
        changed = C.c_int32()
        check("hpx_shard_rebalance", self.lib.hpx_shard_rebalance(self.handle, C.byref(changed)))
        if changed.value:
            fh = C.c_void_p()
            check("hpx_shard_frame", self.lib.hpx_shard_frame(self.handle, C.byref(fh)))
            self.frame = _BorrowedFrame(self.plan, fh) if fh.value else None
        return bool(changed.value)

    def set_reduce(self, enabled: bool):
        check("hpx_shard_set_reduce", self.lib.hpx_shard_set_reduce(self.handle, 1 if enabled else 0))

    def set_result(self, result: str):
        check("hpx_shard_set_result", self.lib.hpx_shard_set_result(self.handle, self.RESULTS[result]))

    def layout(self):
        if self.bands is not None:
            n = self.comm.world
            row0, rows = (C.c_uint32 * n)(), (C.c_uint32 * n)()
            wedges, cuts = (C.c_int32 * (2 * n))(), (C.c_int32 * (n + 1))()
            out, inn = C.c_size_t(), C.c_size_t()
            check("hpx_shard_bands", self.lib.hpx_shard_bands(self.handle, row0, rows, wedges, cuts, C.byref(out), C.byref(inn)))
            axis = C.c_int32()
            check("hpx_shard_owned", self.lib.hpx_shard_owned(self.handle, None, None, None, None, C.byref(axis)))
            direct = C.c_int32()
            check("hpx_shard_exchange_is_direct", self.lib.hpx_shard_exchange_is_direct(self.handle, C.byref(direct)))
            order = C.c_int32()
            check("hpx_shard_tile_order", self.lib.hpx_shard_tile_order(self.handle, C.byref(order)))
            return {"exchange": "own kernels over mapped peer memory (NVLink)" if direct.value else "NCCL send/recv + broadcast",
                    "slow_axis": "xyz"[axis.value], "band_row0": list(row0), "band_rows": list(rows), "tile_order": order.value,
                    "wedges": [(int(wedges[2 * i]), int(wedges[2 * i + 1])) for i in range(n)], "owner_cuts": list(cuts),
                    "send_bytes": out.value * 4, "recv_bytes": inn.value * 4}
        axis, n = C.c_int32(), C.c_uint32()
        rows, ranges = (C.c_uint32 * 16)(), (C.c_int32 * 32)()
        check("hpx_shard_layout", self.lib.hpx_shard_layout(self.handle, C.byref(axis), C.byref(n), rows, ranges))
        return {"slow_axis": "xyz"[axis.value], "group_rows": [int(rows[i]) for i in range(n.value)],
                "slab_ranges": [(int(ranges[2 * i]), int(ranges[2 * i + 1])) for i in range(n.value)]}

    def owned(self):
        """(device pointer, first slab, slabs, floats per slab) of the slabs this rank owns inside the gradient block."""
        ptr, first, count, sf = C.c_void_p(), C.c_int32(), C.c_int32(), C.c_size_t()
        check("hpx_shard_owned", self.lib.hpx_shard_owned(self.handle, C.byref(ptr), C.byref(first), C.byref(count), C.byref(sf), None))
        return ptr.value or 0, first.value, count.value, sf.value

    def close(self):
        if self.handle:
            self.lib.hpx_shard_release(self.handle)
            self.handle = C.c_void_p()
