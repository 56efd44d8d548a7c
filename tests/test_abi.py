"""The drop-in boundary without a GPU: struct layout, exported symbols, host-side logic.

No compute entry point is exercised here beyond checking that, without a CUDA device,
it fails loudly (HP_STATUS_UNSUPPORTED) instead of falling back to anything.
"""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import dvren_b200 as D
import hp_abi as A
import oracle as O
import util as U

REPO = U.REPO
LAYOUT_JSON = os.path.join(U.GOLDEN_DIR, "hp_abi_layout.json")

STRUCT_FIELDS = {
    "hp_version": ["major", "minor", "patch"],
    "hp_sampling_desc": ["dt", "max_steps", "mode"],
    "hp_tensor": ["data", "dtype", "memspace", "rank", "shape", "stride"],
    "hp_ctx_desc": ["flags", "preferred_device", "reserved"],
    "hp_camera_desc": ["model", "K", "c2w", "ortho_scale"],
    "hp_roi_desc": ["x", "y", "width", "height"],
    "hp_plan_desc": ["width", "height", "t_near", "t_far", "max_rays", "max_samples", "seed", "camera", "roi",
                     "sampling"],
    "hp_rays_t": ["origins", "directions", "t_near", "t_far", "pixel_ids"],
    "hp_samp_t": ["positions", "dt", "ray_offset", "sigma", "color"],
    "hp_intl_t": ["radiance", "transmittance", "opacity", "depth", "aux"],
    "hp_img_t": ["image", "trans", "opacity", "depth", "hitmask"],
    "hp_grads_t": ["sigma", "color", "camera"],
}


def _layout_program():
    lines = ["#include <stdio.h>", "#include <stddef.h>", "#define HP_WITH_CUDA 1", '#include "hotpath/hp.h"',
             "int main(void){", 'printf("{");']
    first = True
    for name, fields in STRUCT_FIELDS.items():
        sep = "" if first else ","
        first = False
        lines.append(f'printf("{sep}\\"{name}\\":{{\\"size\\":%zu", sizeof({name}));')
        for f in fields:
            lines.append(f'printf(",\\"{f}\\":%zu", offsetof({name}, {f}));')
        lines.append('printf("}");')
    enums = ["HP_STATUS_SUCCESS", "HP_STATUS_INVALID_ARGUMENT", "HP_STATUS_OUT_OF_MEMORY", "HP_STATUS_NOT_IMPLEMENTED",
             "HP_STATUS_UNSUPPORTED", "HP_STATUS_INTERNAL_ERROR", "HP_MEMSPACE_HOST", "HP_MEMSPACE_DEVICE",
             "HP_DTYPE_F16", "HP_DTYPE_BF16", "HP_DTYPE_F32", "HP_DTYPE_I32", "HP_DTYPE_U32", "HP_CAMERA_PINHOLE",
             "HP_CAMERA_ORTHOGRAPHIC", "HP_SAMPLING_FIXED", "HP_SAMPLING_STRATIFIED", "HP_INTERP_NEAREST",
             "HP_INTERP_LINEAR", "HP_OOB_ZERO", "HP_OOB_CLAMP", "HP_VERSION_MAJOR", "HP_VERSION_MINOR",
             "HP_VERSION_PATCH"]
    lines.append('printf(",\\"enums\\":{");')
    for i, e in enumerate(enums):
        lines.append(f'printf("{"," if i else ""}\\"{e}\\":%d", (int){e});')
    lines += ['printf("}}\\n");', "return 0;}"]
    return "\n".join(lines)


def _compile_layout(include_dir):
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "layout.c"), os.path.join(td, "layout")
        with open(src, "w") as f:
            f.write(_layout_program())
        subprocess.check_call(["gcc", "-std=c11", "-I", include_dir, src, "-o", exe])
        return json.loads(subprocess.check_output([exe]).decode())


def test_header_layout_matches_committed_reference_layout():
    """include/hotpath/hp.h must be layout-identical to the reference header.  The committed table
    was produced by compiling this same program against reference hotpath/include/hotpath/hp.h."""
    mine = _compile_layout(os.path.join(REPO, "include"))
    ref_inc = "/root/reference/hotpath/include"
    if os.path.exists(os.path.join(ref_inc, "hotpath", "hp.h")):
        ref = _compile_layout(ref_inc)
        if not os.path.exists(LAYOUT_JSON) or json.load(open(LAYOUT_JSON)) != ref:
            with open(LAYOUT_JSON, "w") as f:
                json.dump(ref, f, indent=1, sort_keys=True)
        assert mine == ref
    assert mine == json.load(open(LAYOUT_JSON))


def test_ctypes_mirror_matches_header():
    layout = json.load(open(LAYOUT_JSON))
    for st in A.ABI_STRUCTS:
        ref = layout[st.__name__]
        assert C.sizeof(st) == ref["size"], st.__name__
        for fname, _ in st._fields_:
            assert getattr(st, fname).offset == ref[fname], (st.__name__, fname)


def test_library_exports_every_declared_symbol():
    lib = D.load()   # bind() raises AttributeError for any missing symbol
    for name in list(A.ABI_FUNCTIONS) + list(D.HPX_FUNCTIONS):
        assert hasattr(lib, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", D.LIB_PATH]).decode()
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    # every prototype in the two public headers is exported, nothing under oracle/ is linked in
    import re
    for header in ("hp.h", "hp_b200.h"):
        text = open(os.path.join(REPO, "include", "hotpath", header)).read()
        for name in re.findall(r"HP_API[^;(]*?\b(hpx?_[a-z0-9_]+)\s*\(", text):
            assert name in exported, f"{name} declared in {header} but not exported"
    assert not any(s.startswith("orc_") or s.startswith("ref_") for s in exported)
    v = lib.hp_get_version()
    assert (v.major, v.minor, v.patch) == (0, 1, 0)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_plan_defaults_match_oracle_and_reference_rules():
    """hp_plan_create is pure host logic (reference hp_runtime.cpp:45-146): same status, same
    resolved descriptor as the pinned oracle, for valid and invalid inputs."""
    lib = D.load()
    ctx = C.c_void_p()
    assert lib.hp_ctx_create(None, C.byref(ctx)) == 0
    rng = np.random.default_rng(5)
    descs = [c["desc"] for c in U.random_cases(12, seed=3)]
    descs += [U.load_golden(p)["desc_in"] for p in U.golden_cases()]
    bad = [A.make_plan_desc(0, 4, 0, 1), A.make_plan_desc(4, 4, 1.0, 1.0), A.make_plan_desc(4, 4, 0, 1, roi=(2, 2, 3, 1)),
           A.make_plan_desc(8, 8, 0, 1, max_rays=10), A.make_plan_desc(8, 8, 0, 1, max_rays=64, max_samples=10),
           A.make_plan_desc(4, 4, 0, 1, roi=(0xFFFFFFFF, 0, 2, 2)), A.make_plan_desc(4, 4, 0, 1, dt=-1.0, mode=7, model=9)]
    for d in descs + bad:
        for _ in range(2):
            plan = C.c_void_p()
            st = lib.hp_plan_create(ctx, C.byref(d), C.byref(plan))
            ost, odesc = O.plan_resolve(d)
            assert st == ost, (st, ost)
            if st == 0:
                got = A.hp_plan_desc()
                assert lib.hp_plan_get_desc(plan, C.byref(got)) == 0
                assert bytes(got) == bytes(odesc)
                lib.hp_plan_release(plan)
            # perturb and retry once
            d = A.copy_desc(d)
            d.max_samples = int(rng.integers(0, 2)) * d.max_samples
    assert lib.hp_plan_create(None, C.byref(descs[0]), C.byref(plan)) == A.HP_STATUS_INVALID_ARGUMENT
    assert lib.hp_plan_get_desc(None, None) == A.HP_STATUS_INVALID_ARGUMENT
    lib.hp_plan_release(None)
    lib.hp_ctx_release(ctx)


def test_ctx_desc_roundtrip_and_null_safety():
    lib = D.load()
    assert lib.hp_ctx_create(None, None) == A.HP_STATUS_INVALID_ARGUMENT
    desc = A.hp_ctx_desc(7, b"cuda:0", None)
    ctx = C.c_void_p()
    assert lib.hp_ctx_create(C.byref(desc), C.byref(ctx)) == 0
    got = A.hp_ctx_desc()
    assert lib.hp_ctx_get_desc(ctx, C.byref(got)) == 0
    assert got.flags == 7 and got.preferred_device == b"cuda:0"
    lib.hp_ctx_release(ctx)
    lib.hp_ctx_release(None)
    lib.hp_field_release(None)
    lib.hp_graph_release(None)
    lib.hpx_grid_release(None)
    lib.hpx_frame_release(None)


def test_field_validation_order_matches_reference():
    """reference hp_runtime.cpp:259-339: dtype, then rank, then non-positive shape."""
    lib = D.load()
    ctx = C.c_void_p()
    assert lib.hp_ctx_create(None, C.byref(ctx)) == 0
    field = C.c_void_p()
    grid = np.zeros((2, 2, 2), np.float32)
    t = A.host_tensor(grid)
    assert lib.hp_field_create_grid_sigma(None, C.byref(t), 1, 0, C.byref(field)) == A.HP_STATUS_INVALID_ARGUMENT
    t.dtype = A.HP_DTYPE_F16
    assert lib.hp_field_create_grid_sigma(ctx, C.byref(t), 1, 0, C.byref(field)) == A.HP_STATUS_INVALID_ARGUMENT
    t = A.host_tensor(np.zeros((2, 2), np.float32))
    assert lib.hp_field_create_grid_sigma(ctx, C.byref(t), 1, 0, C.byref(field)) == A.HP_STATUS_INVALID_ARGUMENT
    t = A.host_tensor(np.zeros((2, 2, 2, 2), np.float32))  # colour needs a trailing 3
    assert lib.hp_field_create_grid_color(ctx, C.byref(t), 1, 0, C.byref(field)) == A.HP_STATUS_INVALID_ARGUMENT
    t = A.host_tensor(grid)
    t.shape[1] = 0
    assert lib.hp_field_create_grid_sigma(ctx, C.byref(t), 1, 0, C.byref(field)) == A.HP_STATUS_INVALID_ARGUMENT
    params = A.host_tensor(np.zeros(16, np.float32))
    assert lib.hp_field_create_hash_mlp(ctx, C.byref(params), C.byref(field)) == A.HP_STATUS_UNSUPPORTED
    lib.hp_ctx_release(ctx)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_unsupported_not_fallback():
    """The product has no CPU compute path: without a CUDA device every compute entry point
    fails with HP_STATUS_UNSUPPORTED and says why."""
    lib = D.load()
    ctx = C.c_void_p()
    assert lib.hp_ctx_create(None, C.byref(ctx)) == 0
    grid = np.ones((2, 2, 2), np.float32)
    t = A.host_tensor(grid)
    field = C.c_void_p()
    assert lib.hp_field_create_grid_sigma(ctx, C.byref(t), 1, 0, C.byref(field)) == A.HP_STATUS_UNSUPPORTED
    assert b"no CPU compute path" in lib.hpx_last_error()
    desc = A.make_plan_desc(4, 4, 0.0, 1.0, dt=0.1, max_steps=16)
    plan = C.c_void_p()
    assert lib.hp_plan_create(ctx, C.byref(desc), C.byref(plan)) == 0
    rays = A.hp_rays_t()
    buf = {k: np.zeros((16, 3) if k in ("origins", "directions") else 16,
                       np.uint32 if k == "pixel_ids" else np.float32)
           for k in ("origins", "directions", "t_near", "t_far", "pixel_ids")}
    for k, v in buf.items():
        setattr(rays, k, A.host_tensor(v))
    assert lib.hp_ray(plan, None, C.byref(rays), None, 0) == A.HP_STATUS_UNSUPPORTED
    frame = C.c_void_p()
    assert lib.hpx_frame_create(plan, C.byref(frame)) == A.HP_STATUS_UNSUPPORTED
    with pytest.raises(D.DvrenError):
        D.Grid(D.Context(), grid, np.ones((2, 2, 2, 3), np.float32))
    lib.hp_plan_release(plan)
    lib.hp_ctx_release(ctx)


@pytest.mark.parametrize("t_near,t_far,dt,steps", [(0.9, 4.0, 1.5 / 512, 512), (0.1, 2.0, 0.1, 64), (0.0, 1.0, 0.1, 16),
                                                   (0.5, 1.6, 0.04, 40), (0.3, 0.95, 0.0371, 100), (2.0, 2.5, 0.3, 7)])
def test_step_table_matches_the_oracle_march(t_near, t_far, dt, steps):
    """The host-built per-step table of the lean kernels (hpx_plan_step_table; no GPU needed) against the oracle's march of
    one ray: sample count, dt_actual of every step and the fixed-mode sample position bit for bit; base and depth cursor
    against their float32 definitions (reference samp_cpu.cpp:227-241, int_cpu.cpp:170,211)."""
    import ctypes as C
    import dvren_b200 as D
    import oracle as O
    lib = D.load()
    desc = A.make_plan_desc(1, 1, t_near, t_far, dt=dt, max_steps=steps)
    ctx, plan = C.c_void_p(), C.c_void_p()
    assert lib.hp_ctx_create(None, C.byref(ctx)) == 0
    assert lib.hp_plan_create(ctx, C.byref(desc), C.byref(plan)) == 0
    count = C.c_uint32()
    assert lib.hpx_plan_step_table(plan, None, 0, C.byref(count)) == 0
    table = np.zeros((count.value, 4), np.float32)
    assert lib.hpx_plan_step_table(plan, table.ctypes.data, count.value, C.byref(count)) == 0
    st, odesc = O.plan_resolve(desc)
    rays = O.rays(odesc)
    sig = np.ones((2, 2, 2), np.float32)
    gs, gc = O.make_grid(sig, 1), O.make_grid(np.ones((2, 2, 2, 3), np.float32), 3)
    st, samp = O.sample(odesc, gs, gc, rays, odesc.max_samples)
    assert st == 0 and samp["count"] == count.value
    assert table[:, 2].tobytes() == samp["dt"].tobytes()                       # dt_actual
    o, d = rays["origins"][0], rays["directions"][0]
    pos = np.stack([o + d * np.float32(t) for t in table[:, 1]]).astype(np.float32)
    assert pos.tobytes() == samp["positions"].tobytes()                         # fixed-mode sample time -> position
    k = np.arange(count.value, dtype=np.float32)
    base = (np.float32(odesc.t_near) + k * np.float32(odesc.sampling.dt)).astype(np.float32)
    assert table[:, 0].tobytes() == base.tobytes()
    cursor = np.float32(odesc.t_near)
    for i in range(count.value):
        assert table[i, 3] == cursor
        cursor = np.float32(cursor + table[i, 2])
    lib.hp_plan_release(plan)
    lib.hp_ctx_release(ctx)


def test_extension_entry_points_validate_arguments_without_a_gpu():
    """hp_b200.h: null / out-of-range arguments are rejected before any device call (so this runs on a CPU-only box)."""
    import ctypes as C
    import dvren_b200 as D
    lib = D.load()
    INV = A.HP_STATUS_INVALID_ARGUMENT
    box = (C.c_int32 * 6)()
    n = C.c_uint32()
    ptr = C.c_void_p()
    assert lib.hpx_frame_set_interleave(None, 2, 0) == INV
    assert lib.hpx_frame_bounds(None, None, C.byref(box)) == INV
    assert lib.hpx_backward_box(None, None, None, A.HP_MEMSPACE_DEVICE, 1, None, C.byref(box)) == INV
    assert lib.hpx_frame_box_misses(None, C.byref(n)) == INV
    assert lib.hpx_grid_add_box(None, None, None, C.byref(box)) == INV
    assert lib.hpx_grid_set_grad_layout(None, 1, None, None) == INV
    assert lib.hpx_backward_signalled(None, None, None, A.HP_MEMSPACE_DEVICE, 1, None, 0, C.byref(ptr), C.byref(n)) == INV
    assert lib.hpx_frame_reset_group_counters(None, C.byref(ptr)) == INV
    assert lib.hpx_stream_wait_counter(None, None, 1) == INV
    assert lib.hpx_backward_scatter(None, None, 0, C.byref(n)) == INV
    assert lib.hpx_plan_step_table(None, None, 0, C.byref(n)) == INV


def test_round2_entry_points_validate_arguments_without_a_gpu():
    """hp_b200.h, multi-GPU / storage / read-back additions: null handles and out-of-range enumerators are rejected before
    any device or NCCL call; the releases accept null (reference hp_*_release behaviour, hp_runtime.cpp:150-155)."""
    import ctypes as C
    import dvren_b200 as D
    lib = D.load()
    INV = A.HP_STATUS_INVALID_ARGUMENT
    i32, u32, sz, ptr = C.c_int32(), C.c_uint32(), C.c_size_t(), C.c_void_p()
    f32 = C.c_float()
    assert lib.hpx_grid_set_storage(None, 1) == INV
    assert lib.hpx_grid_build_occupancy(None, 1, C.byref(f32), C.byref(f32)) == INV
    assert lib.hpx_grid_set_occupancy(None, 1) == INV
    assert lib.hpx_grid_read_grad_range(None, 0, 1, None, None, None, A.HP_MEMSPACE_HOST) == INV
    assert lib.hpx_backward_streamed(None, None, None, A.HP_MEMSPACE_DEVICE, 1, None, None, None) == INV
    assert lib.hpx_frame_set_row_order(None, 1) == INV
    assert lib.hpx_tile_row_order(0, 0, 0, None) == INV
    assert lib.hpx_tile_row_order(0, 8, 3, C.byref(u32)) == INV        # order is 0, 1 or 2
    assert lib.hpx_tile_order(0, 0, 0, 0, None, None) == INV
    assert lib.hpx_shard_tile_order(None, C.byref(i32)) == INV
    assert lib.hpx_shard_tune_order(None, None, 1, C.byref(i32)) == INV
    assert lib.hpx_plan_best_tile_order(None, 0, 8, 740, C.byref(i32), None) == INV
    assert lib.hpx_ctx_sm_counts(None, C.byref(u32), C.byref(u32)) == INV
    assert lib.hpx_comm_unique_id(None) == INV
    assert lib.hpx_comm_create(None, None, 0, 1, 0, C.byref(ptr)) == INV
    ctx = C.c_void_p()
    assert lib.hp_ctx_create(None, C.byref(ctx)) == 0
    assert lib.hpx_comm_create(ctx, None, 2, 2, 0, C.byref(ptr)) == INV     # rank outside [0, world)
    assert lib.hpx_comm_create(ctx, None, 0, 2, 0, C.byref(ptr)) == INV     # world > 1 needs a rendezvous id
    assert lib.hpx_comm_create(ctx, None, 0, 0, 0, C.byref(ptr)) == INV
    lib.hp_ctx_release(ctx)
    assert lib.hpx_comm_info(None, C.byref(i32), C.byref(i32), C.byref(i32)) == INV
    assert lib.hpx_comm_allreduce(None, None, 0) == INV
    assert lib.hpx_grid_allreduce_grad(None, None) == INV
    assert lib.hpx_shard_create(None, None, None, None, 1, C.byref(ptr)) == INV
    assert lib.hpx_shard_create_bands(None, None, None, 1, C.byref(ptr)) == INV
    assert lib.hpx_shard_step(None, None, 0) == INV
    assert lib.hpx_shard_frame(None, C.byref(ptr)) == INV
    assert lib.hpx_shard_set_reduce(None, 1) == INV
    assert lib.hpx_shard_set_result(None, 1) == INV
    assert lib.hpx_shard_rebalance(None, C.byref(i32)) == INV
    assert lib.hpx_shard_exchange_is_direct(None, C.byref(i32)) == INV
    assert lib.hpx_shard_layout(None, C.byref(i32), C.byref(u32), None, None) == INV
    assert lib.hpx_shard_bands(None, None, None, None, None, C.byref(sz), C.byref(sz)) == INV
    assert lib.hpx_shard_owned(None, C.byref(ptr), C.byref(i32), C.byref(i32), C.byref(sz), C.byref(i32)) == INV
    wedges = (C.c_int32 * 4)(0, 4, 2, 8)
    cuts = (C.c_int32 * 3)()
    assert lib.hpx_plan_owner_cuts(0, 8, wedges, 1, cuts) == INV
    assert lib.hpx_plan_owner_cuts(2, 8, wedges, 7, cuts) == INV             # result is OWNED (1) or REPLICATED (2)
    assert lib.hpx_plan_owner_cuts(2, 4, wedges, 1, cuts) == INV             # a wedge reaches past the last slab
    assert lib.hpx_plan_balanced_bands(None, 2, None, None, None) == INV
    lib.hpx_comm_release(None)
    lib.hpx_shard_release(None)
