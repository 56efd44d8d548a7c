"""GPU tests of the device-resident training surface and of object lifetimes (hp_b200.h):
parameter updates (hpx_grid_update, hpx_grid_adopt_fields), graph capture with the deterministic
backward, context / stream rules, release order, and the measurement counters bench.py relies on."""
import ctypes as C

import numpy as np
import pytest

import dvren_b200 as D
import hp_abi as A
import hp_host as H
import oracle as O
import synth as S
import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = D.Context()
    yield c
    c.close()


def _render(ctx, grid, desc, dl=None, flags=D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO):
    plan = D.Plan(ctx, desc)
    frame = D.Frame(plan)
    frame.forward(grid)
    out = frame.read()
    out.update(frame.counts())
    if dl is not None:
        frame.backward(grid, dl, flags)
        out["sigma_grad"], out["color_grad"], out["camera_grad"] = grid.read_grad()
    frame.close(); plan.close()
    return out


def test_grid_update_equals_a_fresh_grid_bit_for_bit(ctx):
    """hpx_grid_update (the optimiser hand-off): after replacing the values -- both arrays, sigma only, colour only --
    forward AND backward equal those of a grid created from the new values, bit for bit on the images and on the
    deterministic gradients."""
    desc = S.bench_plan(70, 52, 96, stratified=True, view=2, views=9)
    dl = S.hashed_image_grad(70 * 52)
    sig0, col0 = S.hashed_volume(24, "dense", seed=11)
    sig1, col1 = S.hashed_volume(24, "thin", seed=99)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_DETERMINISTIC
    for new_sig, new_col in ((sig1, col1), (sig1, None), (None, col1)):
        grid = D.Grid(ctx, sig0, col0)
        before = _render(ctx, grid, desc, dl, flags)                     # also primes the deterministic |value| maximum
        grid.update(new_sig, new_col)
        got = _render(ctx, grid, desc, dl, flags)
        grid.close()
        fresh_grid = D.Grid(ctx, new_sig if new_sig is not None else sig0, new_col if new_col is not None else col0)
        fresh = _render(ctx, fresh_grid, desc, dl, flags)
        fresh_grid.close()
        for k in ("image", "trans", "opacity", "depth", "hitmask"):
            U.assert_bits(got[k], fresh[k], f"updated grid {k}")
        assert got["live_samples"] == fresh["live_samples"]
        U.assert_bits(got["sigma_grad"], fresh["sigma_grad"], "updated grid sigma_grad (deterministic)")
        U.assert_bits(got["color_grad"], fresh["color_grad"], "updated grid color_grad (deterministic)")
        assert not U.bits_equal(before["image"], got["image"])           # the update really changed the result


def test_adopted_fields_follow_grid_updates_in_the_staged_path():
    """DenseGridField::UpdateValues refreshed only the packed grid; the staged path (hp_samp on the two
    hp_fields) kept rendering the old values.  With hpx_grid_adopt_fields the fields are views of the packed voxels:
    after an update, staged hp_samp output equals the oracle on the NEW values bit for bit, and equals the fused path."""
    lib = D.load()
    pipe = H.HpHostPipeline(lib)
    sig0, col0 = S.hashed_volume((7, 9, 11), "dense", seed=5)
    sig1, col1 = S.hashed_volume((7, 9, 11), "thin", seed=6)
    desc = S.bench_plan(21, 17, 40, stratified=True, view=1, views=7)
    plan, rdesc = pipe.plan(desc)
    n = rdesc.roi.width * rdesc.roi.height
    rays = pipe.ray(plan, n)
    fs, fc = pipe.sigma_field(sig0), pipe.color_field(col0)
    grid = C.c_void_p()
    D.check("hpx_grid_create", lib.hpx_grid_create(pipe.ctx, fs, fc, None, None, C.byref(grid)))
    D.check("hpx_grid_adopt_fields", lib.hpx_grid_adopt_fields(grid, fs, fc))
    st, odesc = O.plan_resolve(desc)
    orays = O.rays(odesc)
    for sig, col, what in ((sig0, col0, "adopted"), (sig1, col1, "updated")):
        if what == "updated":
            D.check("hpx_grid_update", lib.hpx_grid_update(grid, sig.ctypes.data, col.ctypes.data, A.HP_MEMSPACE_HOST))
        samp = pipe.samp(plan, fs, fc, rays, rdesc.max_samples)
        gs, gc = U.oracle_grids(sig, col, 1, 0)
        st, osamp = O.sample(odesc, gs, gc, orays, odesc.max_samples)
        assert st == 0 and samp["count"] == osamp["count"]
        for k in ("positions", "dt", "ray_offset", "sigma", "color"):
            U.assert_bits(samp[k], osamp[k], f"{what} fields samp.{k}")
        fsamp, fintl = pipe.fused(plan, fs, fc, rays, rdesc.max_samples)
        U.assert_bits(fsamp["sigma"], osamp["sigma"], f"{what} fused sigma")
        # a second grid packed FROM the views (stride-4 sources) holds the same voxels
        grid2 = C.c_void_p()
        D.check("hpx_grid_create", lib.hpx_grid_create(pipe.ctx, fs, fc, None, None, C.byref(grid2)))
        frame = C.c_void_p()
        D.check("hpx_frame_create", lib.hpx_frame_create(plan, C.byref(frame)))
        imgs = []
        for g in (grid, grid2):
            D.check("hpx_forward", lib.hpx_forward(frame, g))
            img = np.zeros((rdesc.height, rdesc.width, 3), np.float32)
            D.check("hpx_frame_read", lib.hpx_frame_read(frame, img.ctypes.data, None, None, None, None))
            imgs.append(img)
        U.assert_bits(imgs[0], imgs[1], f"{what}: grid packed from the views")
        ref = O.render(odesc, gs, gc)
        U.assert_close(imgs[0], ref["image"], U.IMAGE_RTOL, f"{what} lean image")
        lib.hpx_frame_release(frame)
        lib.hpx_grid_release(grid2)
    # releasing the grid first leaves the views empty, not dangling: the next query fails cleanly
    lib.hpx_grid_release(grid)
    with pytest.raises(H.HpError):
        pipe.samp(plan, fs, fc, rays, rdesc.max_samples)
    pipe.close()


def test_capture_with_deterministic_backward_survives_value_updates(ctx):
    """capturing HPX_BACKWARD_DETERMINISTIC used to cudaMalloc inside the capture (first use) and to freeze the
    fixed-point quantum of the values at capture time.  Capture FIRST (no eager deterministic backward before), replay,
    grow the values 1000x with hpx_grid_update, replay again: each replay must equal an eager deterministic backward on
    the same values bit for bit."""
    desc = S.bench_plan(64, 40, 64, stratified=True, view=1, views=5)
    n = 64 * 40
    dl = S.hashed_image_grad(n)
    sig, col = S.hashed_volume(16, "thin", seed=3)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_DETERMINISTIC
    plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
    d_dl = frame.grad_input_ptr()
    D.check("h2d", ctx.lib.hpx_copy_to_device(ctx.handle, C.c_void_p(d_dl), np.ascontiguousarray(dl).ctypes.data, dl.nbytes))
    frame.capture(grid, flags)                       # nothing deterministic has run eagerly on this grid yet
    for scale in (1.0, 1000.0, 0.001):
        grid.update(None, col * np.float32(scale))   # colour magnitudes set the quantum (fixed_scale_kernel)
        frame.replay()
        sg_r, cg_r, _ = grid.read_grad()
        fresh = D.Grid(ctx, sig, col * np.float32(scale))
        eager = _render(ctx, fresh, desc, dl, flags)
        fresh.close()
        U.assert_bits(sg_r, eager["sigma_grad"], f"replayed deterministic sigma_grad, colours x{scale}")
        U.assert_bits(cg_r, eager["color_grad"], f"replayed deterministic color_grad, colours x{scale}")
        assert np.isfinite(cg_r).all() and np.abs(cg_r).max() > 0
    # an eager deterministic backward after the capture still works (meta block initialised, staleness honoured)
    frame.backward(grid, dl, flags)
    sg_e, cg_e, _ = grid.read_grad()
    U.assert_bits(sg_e, sg_r, "eager after capture")
    frame.close(); grid.close(); plan.close()


def test_frame_and_grid_must_share_a_context(ctx):
    """a frame and a grid from different contexts (= different streams) are rejected instead of racing."""
    other = D.Context(device=0)
    sig, col = S.hashed_volume(8, "thin")
    desc = S.bench_plan(16, 16, 16, stratified=False)
    plan = D.Plan(ctx, desc); frame = D.Frame(plan)
    foreign = D.Grid(other, sig, col)
    with pytest.raises(D.DvrenError) as e:
        frame.forward(foreign)
    assert e.value.status == A.HP_STATUS_INVALID_ARGUMENT
    own = D.Grid(ctx, sig, col)
    frame.forward(own)
    with pytest.raises(D.DvrenError):
        frame.backward(foreign, np.zeros((256, 3), np.float32))
    frame.close(); plan.close(); own.close(); foreign.close(); other.close()


def test_set_interleave_drops_a_captured_graph(ctx):
    """a captured graph has the CTA count of the old interleave baked in; replay must fail until recapture."""
    sig, col = S.hashed_volume(12, "thin")
    desc = S.bench_plan(48, 40, 32, stratified=True)
    plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
    frame.capture(grid, 0)
    frame.replay()
    full = frame.read()
    frame.set_interleave(2, 1)
    with pytest.raises(D.DvrenError):
        frame.replay()
    frame.capture(grid, 0)
    frame.replay()
    half = frame.read()
    own = half["hitmask"] == 1
    assert own.sum() == frame.counts()["rays"] and 0 < own.sum() < full["hitmask"].sum()
    U.assert_bits(half["image"][own], full["image"][own], "interleaved replay")
    frame.close(); grid.close(); plan.close()


def test_context_may_be_released_before_its_objects():
    """the reference's fields never touch their context, so releasing the context first is legal there
    (hp_runtime.cpp:33-36).  Here every object keeps its context alive: use after hp_ctx_release still works and the
    releases in any order neither crash nor leak the stream."""
    lib = D.load()
    c = D.Context(device=0)
    sig, col = S.hashed_volume(10, "thin")
    plan = D.Plan(c, S.bench_plan(24, 20, 24, stratified=False))
    grid = D.Grid(c, sig, col)
    frame = D.Frame(plan)
    pipe_field = C.c_void_p()
    t = A.host_tensor(np.ascontiguousarray(sig))
    D.check("field", lib.hp_field_create_grid_sigma(c.handle, C.byref(t), 1, 0, C.byref(pipe_field)))
    c.close()                                   # the caller's reference goes first
    plan_handle = plan.handle
    frame.forward(grid)                         # still usable: the objects hold the context
    assert frame.read()["hitmask"].all()
    plan.close()                                # the frame keeps its plan alive too
    frame.forward(grid)
    frame.close()
    lib.hp_field_release(pipe_field)
    grid.close()                                # last object: the context goes away here
    assert plan_handle


def test_entry_points_restore_the_callers_device():
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    c = D.Context(device=1)
    sig, col = S.hashed_volume(8, "thin")
    grid = D.Grid(c, sig, col)
    assert torch.cuda.current_device() == 0
    grid.close(); c.close()


def test_cube_sample_and_touched_voxel_counters_match_the_oracle(ctx):
    """The counters behind bench.py's roofline line: in-cube live samples (the samples that gather and scatter) and touched
    voxels, against counts derived from the oracle's materialised samples."""
    for kind, strat in (("thin", True), ("dense", False)):
        sig, col = S.hashed_volume(20, kind)
        desc = S.bench_plan(57, 43, 80, stratified=strat, view=2, views=9)
        st, odesc = O.plan_resolve(desc)
        gs, gc = U.oracle_grids(sig, col, 1, 0)
        orays = O.rays(odesc)
        st, osamp = O.sample(odesc, gs, gc, orays, odesc.max_samples)
        ointl = O.integrate(odesc, osamp)
        live = ointl["aux"][:, 2] != 0                      # T_before > 0 on every integrated sample, rows after the stop stay 0
        pos = osamp["positions"]
        inside = np.all((pos >= 0) & (pos <= 1), axis=1)
        plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
        frame.forward(grid)
        assert frame.counts()["live_samples"] == int(live.sum())
        assert frame.cube_samples(grid) == int((live & inside).sum())
        dl = np.ones((57 * 43, 3), np.float32)
        frame.backward(grid, dl)
        sg, cg, _ = grid.read_grad()
        touched = (sg != 0) | (cg.reshape(-1, 3) != 0).any(axis=1)
        assert grid.touched_voxels() == int(touched.sum()) > 0
        ref = O.render(odesc, gs, gc, dl)
        ref_touched = (ref["sigma_grad"] != 0) | (ref["color_grad"].reshape(-1, 3) != 0).any(axis=1)
        assert abs(int(touched.sum()) - int(ref_touched.sum())) <= 0.001 * ref_touched.sum() + 2   # exact zeros may differ by rounding
        frame.close(); grid.close(); plan.close()


def test_cxx_surface_selftest():
    """diff-volume-renderer_b200/apps/dvren_bench.cpp `selftest`: dvren::DenseGridField::UpdateValues reaches the staged
    path (fused == staged == fresh field), RenderStats carry CUDA-event stage times, Backward stats exist, the host
    gradient mirrors are lazy but correct, and a Renderer survives the release of its Context."""
    import os
    import subprocess
    exe = os.path.join(U.REPO, "diff-volume-renderer_b200", "dvren_bench")
    assert os.path.exists(exe), f"{exe} is not built (__graft_entry__.build())"
    r = subprocess.run([exe, "selftest"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "selftest ok" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_balanced_bands_sum_to_the_whole_frame(ctx):
    """What hpx_shard_create_bands gives each rank, replayed on ONE GPU: the bands of hpx_plan_balanced_bands rendered one
    after the other (ROI sub-plan, global ray-index base for the stratified jitter, every band with its tiles taken in
    another dispatch order: rows last-to-first, centre-out, column by column with each row order, first-to-last; the last tile
    column is a partial one) reproduce the whole frame's image planes bit for bit and sum to its gradient; the voxel wedges of
    hpx_frame_bounds contain everything a band's backward writes."""
    W, Hh, steps, world = 104, 80, 64, 6
    orders = [1, 2, D.HPX_ORDER_COLUMNS, D.HPX_ORDER_COLUMNS | 1, D.HPX_ORDER_COLUMNS | 2, 0]
    desc = S.bench_plan(W, Hh, steps, stratified=True, view=1, views=7)
    sig, col = S.hashed_volume(20, "dense", seed=5)
    dl = S.hashed_image_grad(W * Hh)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_DETERMINISTIC
    grid = D.Grid(ctx, sig, col)
    whole = _render(ctx, grid, desc, dl, flags | D.HPX_BACKWARD_ZERO)

    plan = D.Plan(ctx, desc)
    row0, rows = (C.c_uint32 * world)(), (C.c_uint32 * world)()
    D.check("hpx_plan_balanced_bands", ctx.lib.hpx_plan_balanced_bands(plan.handle, world, row0, rows, None))
    assert sum(rows) == Hh and all(r > 0 for r in rows)
    grid.zero_grad()
    image = np.zeros((Hh, W, 3), np.float32)
    samples = 0
    for r in range(world):
        bd = S.bench_plan(W, Hh, steps, stratified=True, view=1, views=7, roi=(0, int(row0[r]), W, int(rows[r])))
        bplan = D.Plan(ctx, bd)
        frame = D.Frame(bplan)
        frame.set_view(None, desc.seed, int(row0[r]) * W)
        D.check("hpx_frame_set_row_order", ctx.lib.hpx_frame_set_row_order(frame.handle, orders[r]))
        box = (C.c_int32 * 6)()
        D.check("hpx_frame_bounds", ctx.lib.hpx_frame_bounds(frame.handle, grid.handle, box))
        frame.forward(grid)
        out = frame.read()
        image[row0[r]:row0[r] + rows[r]] = out["image"][row0[r]:row0[r] + rows[r]]
        samples += frame.counts()["samples"]
        before = grid.read_grad()[0].reshape(20, 20, 20).copy()
        frame.backward(grid, dl[int(row0[r]) * W:(int(row0[r]) + int(rows[r])) * W], flags)
        delta = grid.read_grad()[0].reshape(20, 20, 20) - before           # [z][y][x]
        zs, ys, xs = np.nonzero(delta)
        if len(zs):
            assert xs.min() >= box[0] and xs.max() < box[0] + box[3]
            assert ys.min() >= box[1] and ys.max() < box[1] + box[4]
            assert zs.min() >= box[2] and zs.max() < box[2] + box[5]
        frame.close(); bplan.close()
    assert samples == whole["samples"]
    U.assert_bits(image, whole["image"], "bands vs whole frame: image")
    sg, cg, _ = grid.read_grad()
    # deterministic (fixed-point) accumulation: the band sums differ from the whole frame's only by the float conversion of
    # each band's integer total
    U.assert_close(sg, whole["sigma_grad"], 1e-5, "bands vs whole frame: sigma gradient", floor_frac=1e-3)
    U.assert_close(cg, whole["color_grad"], 1e-5, "bands vs whole frame: colour gradient", floor_frac=1e-3)
    plan.close(); grid.close()


def test_sharded_frame_on_two_gpus():
    """hpx_comm + hpx_shard_* with NO Python in the loop (apps/dvren_bench.cpp `shard`: one host thread per GPU): interleaved
    tile rows with slab all-reduces, balanced bands with the replicated and with the owned result -- each verified inside
    the tool against a single-GPU backward of the whole frame."""
    import os
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = os.path.join(U.REPO, "diff-volume-renderer_b200", "dvren_bench")
    assert os.path.exists(exe), f"{exe} is not built (__graft_entry__.build())"
    # (the tool settles the bands' tile dispatch order by measurement, hpx_shard_tune_order; the last two runs pin it instead)
    for mode, env in ((0, {}), (1, {}), (2, {}), (1, {"DVREN_SHARD_EXCHANGE": "nccl"}), (2, {"DVREN_SHARD_TILE_ORDER": "columns"}),
                      (1, {"DVREN_BENCH_NO_TUNE": "1", "DVREN_SHARD_TILE_ORDER": "rows"})):
        r = subprocess.run([exe, "shard", "96", "384", "192", "1", "2", "1", "2", "3", "0", "0", str(mode)], capture_output=True,
                           text=True, timeout=300, env={**os.environ, **env})
        assert r.returncode == 0 and '"verify_max_rel_err_vs_single_gpu"' in r.stdout, (mode, env, r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.parametrize("stratified", [False, True])
def test_empty_space_skipping_changes_nothing(ctx, stratified):
    """hpx_grid_build_occupancy: on a sparse volume -- a shell of density, colour that extends a little
    further than the density, zeros elsewhere -- forward images, counts, live counts and the deterministic gradients
    (grid and fused camera) are IDENTICAL with skipping on and off; and the oracle agrees as it does without it."""
    n, W, Hh, steps = 40, 88, 72, 160
    z, y, x = np.meshgrid(*(np.linspace(0, 1, n, dtype=np.float32),) * 3, indexing="ij")
    r = np.sqrt((x - 0.5) ** 2 + (y - 0.45) ** 2 + (z - 0.55) ** 2)
    base_s, base_c = S.hashed_volume(n, "dense", seed=3)
    sigma = np.where((r > 0.18) & (r < 0.3), base_s, 0).astype(np.float32)
    color = np.where(((r > 0.15) & (r < 0.33))[..., None], base_c, 0).astype(np.float32)
    desc = S.bench_plan(W, Hh, steps, stratified=stratified, view=1, views=11)
    dl = S.hashed_image_grad(W * Hh)
    grid = D.Grid(ctx, sigma, color)
    empty_sigma, empty_all = grid.build_occupancy(enable=True)
    assert 0.2 < empty_all <= empty_sigma < 1.0
    results = {}
    for on in (True, False):
        grid.set_occupancy(on)
        for scatter in (D.HPX_BACKWARD_SCATTER_MERGED, D.HPX_BACKWARD_SCATTER_PER_RAY):
            flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_DETERMINISTIC | scatter
            if scatter == D.HPX_BACKWARD_SCATTER_MERGED:
                flags |= D.HPX_BACKWARD_CAMERA
            results[(on, scatter)] = _render(ctx, grid, desc, dl, flags)
    for scatter in (D.HPX_BACKWARD_SCATTER_MERGED, D.HPX_BACKWARD_SCATTER_PER_RAY):
        a, b = results[(True, scatter)], results[(False, scatter)]
        for k in ("image", "trans", "opacity", "depth", "hitmask", "sigma_grad", "color_grad"):
            U.assert_bits(a[k], b[k], f"occupancy on vs off: {k}")
        assert a["samples"] == b["samples"] and a["live_samples"] == b["live_samples"]
    a, b = results[(True, D.HPX_BACKWARD_SCATTER_MERGED)], results[(False, D.HPX_BACKWARD_SCATTER_MERGED)]
    np.testing.assert_allclose(a["camera_grad"], b["camera_grad"], rtol=1e-6, atol=1e-9)
    # and against the oracle, as every other case
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sigma, color, A.HP_INTERP_LINEAR, A.HP_OOB_ZERO)
    ref = O.render(odesc, gs, gc, dl, shadow=True)
    got = results[(True, D.HPX_BACKWARD_SCATTER_MERGED)]
    U.assert_close(got["image"].reshape(-1, 3), ref["image"].reshape(-1, 3), U.IMAGE_RTOL, "occupancy image vs oracle")
    U.assert_grads(got["sigma_grad"], got["color_grad"], ref, "occupancy vs oracle", res=(n, n, n))
    # new values invalidate the bits: everything renders as occupied, still equal to a fresh grid
    grid.set_occupancy(True)
    grid.update(sigma=base_s)
    fresh = D.Grid(ctx, base_s, color)
    U.assert_bits(_render(ctx, grid, desc)["image"], _render(ctx, fresh, desc)["image"], "after update: image")
    fresh.close(); grid.close()


@pytest.mark.parametrize("stratified,scatter", [(True, D.HPX_BACKWARD_SCATTER_MERGED), (False, D.HPX_BACKWARD_SCATTER_PER_RAY)])
def test_half_storage_equals_the_oracle_on_the_rounded_grid(ctx, stratified, scatter):
    """hpx_grid_set_storage(HPX_STORAGE_F16): values are rounded to IEEE half once, widened exactly on
    load, and all arithmetic stays fp32 -- so the parity twin is the ORACLE run on the rounded values, at the usual gates
    (counts and hit mask bit-exact, image 1e-5, grid gradients 1e-4 with the float64 adjudication, camera 1e-4).  Against
    the unrounded grid the image differs by the storage rounding only (<= 2^-10 relative, checked loosely)."""
    n, W, Hh, steps = 28, 72, 60, 128
    sig, col = S.hashed_volume(n, "dense", seed=21)
    sig16, col16 = sig.astype(np.float16).astype(np.float32), col.astype(np.float16).astype(np.float32)
    desc = S.bench_plan(W, Hh, steps, stratified=stratified, view=2, views=13)
    dl = S.hashed_image_grad(W * Hh)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | scatter
    if scatter == D.HPX_BACKWARD_SCATTER_MERGED:
        flags |= D.HPX_BACKWARD_CAMERA
    grid = D.Grid(ctx, sig, col)
    full = _render(ctx, grid, desc)
    grid.set_storage(half=True)
    got = _render(ctx, grid, desc, dl, flags)
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig16, col16, A.HP_INTERP_LINEAR, A.HP_OOB_ZERO)
    ref = O.render(odesc, gs, gc, dl, shadow=True)
    assert got["samples"] == ref["sample_count"]
    U.assert_bits(got["hitmask"], ref["hitmask"], "half storage: hitmask")
    for k in ("image", "trans", "opacity", "depth"):
        U.assert_close(got[k], ref[k], U.IMAGE_RTOL, f"half storage vs oracle on the rounded grid: {k}")
    U.assert_grads(got["sigma_grad"], got["color_grad"], ref, "half storage vs oracle on the rounded grid", res=(n, n, n))
    np.testing.assert_allclose(got["image"], full["image"], rtol=0, atol=2e-3 * float(np.abs(full["image"]).max()))
    # the same through a grid that was created from the rounded values and never converted: bit for bit
    twin = D.Grid(ctx, sig16, col16)
    same = _render(ctx, twin, desc, dl, flags | D.HPX_BACKWARD_DETERMINISTIC)
    det = _render(ctx, grid, desc, dl, flags | D.HPX_BACKWARD_DETERMINISTIC)
    for k in ("image", "trans", "opacity", "depth", "sigma_grad", "color_grad"):
        U.assert_bits(det[k], same[k], f"half storage vs fp32 grid of the rounded values: {k}")
    # parameter update on a half grid (one channel set kept), occupancy on a half grid, and the way back
    grid.update(sigma=sig * np.float32(0.5))
    twin.update(sigma=(sig * np.float32(0.5)).astype(np.float16).astype(np.float32))
    U.assert_bits(_render(ctx, grid, desc)["image"], _render(ctx, twin, desc)["image"], "half storage after update")
    grid.build_occupancy(enable=True)
    U.assert_bits(_render(ctx, grid, desc)["image"], _render(ctx, twin, desc)["image"], "half storage + occupancy")
    grid.set_storage(half=False)
    U.assert_bits(_render(ctx, grid, desc)["image"], _render(ctx, twin, desc)["image"], "back to fp32 storage")
    twin.close(); grid.close()


def test_half_storage_is_refused_where_it_cannot_hold(ctx):
    sig, col = S.hashed_volume(8, "thin")
    for kw in (dict(oob=A.HP_OOB_CLAMP), dict(interp=A.HP_INTERP_NEAREST), dict(bbox_min=(-1, -1, -1), bbox_max=(2, 2, 2))):
        g = D.Grid(ctx, sig, col, **kw)
        assert ctx.lib.hpx_grid_set_storage(g.handle, 1) == A.HP_STATUS_UNSUPPORTED
        g.close()
    g = D.Grid(ctx, sig, col)
    assert ctx.lib.hpx_grid_set_storage(g.handle, 7) == A.HP_STATUS_INVALID_ARGUMENT
    g.close()


@pytest.mark.parametrize("view,views", [(0, 1), (2, 7), (1, 4)])
def test_streamed_backward_equals_backward_plus_read(ctx, view, views):
    """hpx_backward_streamed: per-row-group signals + slab copies under the kernel deliver exactly what hpx_backward followed
    by hpx_grid_read_grad delivers (same contributions, float reds in another order), for cameras whose image rows advance
    along world y (view 0 and the orbit views) -- and fall back to the two plain calls where streaming cannot apply."""
    n, W, Hh, steps = 36, 120, 96, 128
    sig, col = S.hashed_volume(n, "dense", seed=8)
    desc = S.bench_plan(W, Hh, steps, stratified=True, view=view, views=views)
    dl = S.hashed_image_grad(W * Hh)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_CAMERA | D.HPX_BACKWARD_SCATTER_MERGED
    for oob in (A.HP_OOB_ZERO, A.HP_OOB_CLAMP):          # clamp: the fallback path
        grid = D.Grid(ctx, sig, col, oob=oob)
        plan = D.Plan(ctx, desc)
        frame = D.Frame(plan)
        frame.forward(grid)
        frame.backward(grid, dl, flags)
        ref_s, ref_c, ref_cam = grid.read_grad()
        for _ in range(2):                                # second call reuses the cached row groups
            sg, cg, cam = np.full(n ** 3, 7.0, np.float32), np.full(3 * n ** 3, 7.0, np.float32), np.zeros(16, np.float32)
            D.check("hpx_backward_streamed", ctx.lib.hpx_backward_streamed(frame.handle, grid.handle, dl.ctypes.data, A.HP_MEMSPACE_HOST,
                                                                           flags, sg.ctypes.data, cg.ctypes.data, cam.ctypes.data))
            ctx.synchronize()
            U.assert_close(sg, ref_s, U.GRAD_RTOL, f"streamed sigma gradient (oob {oob})", floor_frac=1e-2)   # two float32 red ORDERS of one sum:
            # the noise scales with sum|terms|, not with the entry (tests/util.py); the oracle comparisons adjudicate that in float64
            U.assert_close(cg, ref_c, U.GRAD_RTOL, f"streamed colour gradient (oob {oob})", floor_frac=1e-2)
            np.testing.assert_allclose(cam, ref_cam, rtol=1e-4, atol=1e-5 * float(np.abs(ref_cam).max()))
            # and the block itself still reads back the same way
            again_s, again_c, _ = grid.read_grad()
            U.assert_bits(again_s, sg, "streamed vs read_grad of the same block: sigma")
            U.assert_bits(again_c, cg, "streamed vs read_grad of the same block: colour")
        # one array only
        only = np.zeros(n ** 3, np.float32)
        D.check("hpx_backward_streamed", ctx.lib.hpx_backward_streamed(frame.handle, grid.handle, dl.ctypes.data, A.HP_MEMSPACE_HOST,
                                                                       flags, only.ctypes.data, None, None))
        ctx.synchronize()
        U.assert_close(only, ref_s, U.GRAD_RTOL, "streamed, sigma only", floor_frac=1e-2)
        frame.close(); plan.close(); grid.close()
