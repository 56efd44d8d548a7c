"""GPU parity of the non-materialising path (hpx_forward / hpx_backward, hp_b200.h) against the
pinned oracle: counts bit-exact, images <= 1e-5 relative, grid gradients <= 1e-4 relative,
camera gradients <= 1e-4 relative to the pinned analytic adjoint (north_star tolerances)."""
import ctypes as C

import numpy as np
import pytest

import dvren_b200 as D
import hp_abi as A
import oracle as O
import synth as S
import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = D.Context()
    yield c
    c.close()


def run_lean(ctx, desc, sigma, color, interp, oob, bmin, bmax, dl=None, flags=None, ray_index_base=0):
    plan = D.Plan(ctx, desc)
    grid = D.Grid(ctx, sigma, color, interp, oob, bmin, bmax)
    frame = D.Frame(plan)
    if ray_index_base:
        frame.set_view(None, plan.desc.seed, ray_index_base)
    frame.forward(grid)
    out = frame.read()
    out.update(frame.counts())
    out["box"] = tuple(frame.bounds(grid))
    if dl is not None:
        frame.backward(grid, dl, flags if flags is not None else D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO)
        sg, cg, cam = grid.read_grad()
        out.update(sigma_grad=sg, color_grad=cg, camera_grad=cam)
    out["desc"] = plan.desc
    frame.close(); grid.close(); plan.close()
    return out


def check_forward(got, ref, what):
    U.assert_bits(got["hitmask"], ref["hitmask"], what + " hitmask")
    assert got["samples"] == ref["sample_count"], what
    assert got["live_samples"] == ref["live_sample_count"], f"{what}: live {got['live_samples']} vs {ref['live_sample_count']}"
    U.assert_close(got["image"], ref["image"], U.IMAGE_RTOL, what + " image")
    U.assert_close(got["trans"], ref["trans"], U.IMAGE_RTOL, what + " trans")
    U.assert_close(got["opacity"], ref["opacity"], U.IMAGE_RTOL, what + " opacity")
    U.assert_close(got["depth"], ref["depth"], U.IMAGE_RTOL, what + " depth")


@pytest.mark.parametrize("case", list(U.random_cases(12, seed=0)), ids=lambda c: f"case{c['case']}")
def test_lean_matches_oracle_random(ctx, case):
    desc = case["desc"]
    st, odesc = O.plan_resolve(desc)
    assert st == 0
    gs, gc = U.oracle_grids(case["sigma"], case["color"], case["interp"], case["oob"])
    n = odesc.roi.width * odesc.roi.height
    dl = S.hashed_image_grad(n)
    ref = O.render(odesc, gs, gc, dl, case["res"], case["bmin"], case["bmax"], shadow=True)
    got = run_lean(ctx, desc, case["sigma"], case["color"], case["interp"], case["oob"], case["bmin"], case["bmax"], dl)
    assert bytes(got["desc"]) == bytes(odesc)
    check_forward(got, ref, f"case{case['case']}")
    U.assert_grads(got["sigma_grad"], got["color_grad"], ref, f"case{case['case']}")


SCATTER_MODES = {"per_ray": D.HPX_BACKWARD_SCATTER_PER_RAY, "merged": D.HPX_BACKWARD_SCATTER_MERGED}


@pytest.mark.parametrize("mode", sorted(SCATTER_MODES))
@pytest.mark.parametrize("case", list(U.random_cases(12, seed=3)), ids=lambda c: f"case{c['case']}")
def test_lean_backward_scatter_modes_match_oracle(ctx, case, mode):
    """Both scatter strategies of the grid backward (one lane per ray; 2x2 quads x 2 steps merged in
    registers) against the oracle, forced through the hpx_backward flags so that the automatic choice
    cannot hide either kernel.  The merged kernel falls back to per-ray for nearest / 1-voxel axes."""
    desc = case["desc"]
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(case["sigma"], case["color"], case["interp"], case["oob"])
    dl = S.hashed_image_grad(odesc.roi.width * odesc.roi.height)
    ref = O.render(odesc, gs, gc, dl, case["res"], case["bmin"], case["bmax"], shadow=True)
    got = run_lean(ctx, desc, case["sigma"], case["color"], case["interp"], case["oob"], case["bmin"], case["bmax"], dl,
                   flags=D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
    U.assert_grads(got["sigma_grad"], got["color_grad"], ref, mode)


@pytest.mark.parametrize("mode", sorted(SCATTER_MODES))
@pytest.mark.parametrize("kind,strat,oob", [("thin", True, 0), ("dense", False, 0), ("dense", True, 1)])
def test_lean_backward_scatter_modes_dense_pixels(ctx, kind, strat, oob, mode):
    """Pixels much denser than voxels (the regime the merged scatter is built for): 128x96 rays over a 20^3 grid,
    ragged image size so that partially filled tiles and quads are exercised."""
    sig, col = S.hashed_volume(20, kind)
    desc = S.bench_plan(125, 91, 120, stratified=strat, view=3, views=11)
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, oob)
    dl = S.hashed_image_grad(125 * 91)
    ref = O.render(odesc, gs, gc, dl, shadow=True)
    got = run_lean(ctx, desc, sig, col, 1, oob, None, None, dl,
                   flags=D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
    U.assert_grads(got["sigma_grad"], got["color_grad"], ref, mode)


@pytest.mark.parametrize("mode", sorted(SCATTER_MODES))
def test_lean_axis_limits_of_the_merged_kernel(ctx, mode):
    """Edge sizes: the merged kernel packs a cell into 10 bits per axis, so 1024 voxels is its largest axis and
    1025 must fall back to the per-ray kernel; both against the oracle (thin slabs keep the grids small)."""
    rng = np.random.default_rng(9)
    for shape in [(2, 3, 1024), (2, 1024, 3), (1024, 2, 3), (3, 2, 1025)]:      # (nz, ny, nx)
        sig = (rng.random(shape, dtype=np.float32) * 4).astype(np.float32)
        col = rng.random(shape + (3,), dtype=np.float32)
        desc = S.bench_plan(33, 21, 48, stratified=True, view=1, views=5)
        st, odesc = O.plan_resolve(desc)
        gs, gc = U.oracle_grids(sig, col, 1, 0)
        dl = S.hashed_image_grad(33 * 21)
        ref = O.render(odesc, gs, gc, dl, shadow=True)
        got = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl,
                       flags=D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
        check_forward(got, ref, str(shape))
        U.assert_grads(got["sigma_grad"], got["color_grad"], ref, str(shape))


@pytest.mark.parametrize("path", U.golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_lean_matches_reference_golden(ctx, path):
    g = U.load_golden(path)
    got = run_lean(ctx, g["desc_in"], g["sigma"], g["color"], g["interp"], g["oob"], g["bmin"], g["bmax"], g["dL_dI"])
    assert bytes(got["desc"]) == bytes(g["desc_resolved"])
    assert got["samples"] == int(g["sample_count"])
    U.assert_bits(got["hitmask"], g["img_hitmask"], "hitmask")
    for k in ("image", "trans", "opacity", "depth"):
        U.assert_close(got[k], g[f"img_{k}"], U.IMAGE_RTOL, k)
    # the float64 shadow for the adjudication comes from the oracle, which reproduces the golden gradients bit for bit
    gs, gc = U.oracle_grids(g["sigma"], g["color"], g["interp"], g["oob"])
    nz, ny, nx = g["sigma"].shape
    ref = O.render(g["desc_resolved"], gs, gc, g["dL_dI"], (nx, ny, nz), g["bmin"], g["bmax"], shadow=True)
    U.assert_bits(ref["sigma_grad"], g["sigma_grad"], "oracle vs golden sigma_grad")
    U.assert_bits(ref["color_grad"], g["color_grad"], "oracle vs golden color_grad")
    U.assert_grads(got["sigma_grad"], got["color_grad"], ref, "golden")


@pytest.mark.parametrize("kind,strat", [("thin", False), ("dense", True), ("dense", False)])
def test_lean_hashed_volumes_medium(ctx, kind, strat):
    """BASELINE config shapes scaled to what the oracle finishes in seconds: 96x80, 48^3, 160 steps."""
    sig, col = S.hashed_volume(48, kind)
    desc = S.bench_plan(96, 80, 160, stratified=strat, view=2, views=9)
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    dl = S.hashed_image_grad(96 * 80)
    ref = O.render(odesc, gs, gc, dl, shadow=True)
    got = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl)
    check_forward(got, ref, kind)
    if kind == "dense":
        assert ref["live_sample_count"] < ref["sample_count"]   # early termination is exercised
    U.assert_grads(got["sigma_grad"], got["color_grad"], ref, kind)


def test_lean_backward_accumulates_and_is_linear(ctx):
    sig, col = S.hashed_volume(24, "thin")
    desc = S.bench_plan(40, 36, 64, stratified=True)
    n = 40 * 36
    dl = S.hashed_image_grad(n)
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    ref = O.render(odesc, gs, gc, dl, shadow=True)
    a = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl)
    b = run_lean(ctx, desc, sig, col, 1, 0, None, None, 2.0 * dl)
    # scaling dL/dI by 2 is exact on every product of the adjoint (so twice the oracle's gradient IS the oracle's
    # gradient of 2 dL/dI); what remains is the order of the float reds in the scatter
    U.assert_grads(a["sigma_grad"], a["color_grad"], ref, "linearity x1")
    U.assert_grads(b["sigma_grad"], b["color_grad"], ref, "linearity x2", scale_by=2.0)
    # without HPX_BACKWARD_ZERO gradients accumulate (DenseGridField semantics, dense_grid.cpp:166-169)
    plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
    frame.forward(grid)
    frame.backward(grid, dl, D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO)
    frame.backward(grid, dl, D.HPX_BACKWARD_GRID)
    sg, cg, _ = grid.read_grad()
    U.assert_grads(sg, cg, ref, "accumulate", scale_by=2.0)
    frame.close(); grid.close(); plan.close()


@pytest.mark.parametrize("mode", sorted(SCATTER_MODES))
def test_deterministic_backward_is_bitwise_reproducible(ctx, mode):
    """HPX_BACKWARD_DETERMINISTIC: fixed-point integer accumulation -> the same bits on every run, within the gradient
    tolerance of the oracle, and accumulation across calls (no ZERO flag) still works."""
    sig, col = S.hashed_volume(24, "dense")
    desc = S.bench_plan(101, 67, 96, stratified=True, view=2, views=9)
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    dl = S.hashed_image_grad(101 * 67) * np.float32(3.7)
    ref = O.render(odesc, gs, gc, dl, shadow=True)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_DETERMINISTIC | SCATTER_MODES[mode]
    runs = [run_lean(ctx, desc, sig, col, 1, 0, None, None, dl, flags=flags) for _ in range(3)]
    for r in runs[1:]:
        U.assert_bits(r["sigma_grad"], runs[0]["sigma_grad"], "deterministic sigma_grad")
        U.assert_bits(r["color_grad"], runs[0]["color_grad"], "deterministic color_grad")
    U.assert_grads(runs[0]["sigma_grad"], runs[0]["color_grad"], ref, "deterministic")
    plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
    frame.forward(grid)
    frame.backward(grid, dl, flags)
    frame.backward(grid, dl, flags & ~D.HPX_BACKWARD_ZERO)
    sg, cg, _ = grid.read_grad()
    U.assert_grads(sg, cg, ref, "deterministic accumulated", scale_by=2.0)
    frame.close(); grid.close(); plan.close()


def test_lean_roi_tiles_reproduce_full_frame(ctx):
    """Sharding contract: ROI sub-plans with a global ray-index base render the same
    pixels bit for bit as the unsharded stratified plan, and their gradients add up to it."""
    sig, col = S.hashed_volume(32, "dense")
    W, Hh, steps = 64, 48, 96
    full_desc = S.bench_plan(W, Hh, steps, stratified=True)
    dl_full = S.hashed_image_grad(W * Hh)
    full = run_lean(ctx, full_desc, sig, col, 1, 0, None, None, dl_full)
    image = np.zeros_like(full["image"]); depth = np.zeros_like(full["depth"])
    sg = np.zeros_like(full["sigma_grad"]); cg = np.zeros_like(full["color_grad"])
    bands = [(0, 16), (16, 8), (24, 24)]
    for y0, h in bands:
        desc = S.bench_plan(W, Hh, steps, stratified=True, roi=(0, y0, W, h))
        part = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl_full[y0 * W:(y0 + h) * W], ray_index_base=y0 * W)
        assert part["hitmask"][y0:y0 + h].all() and part["hitmask"].sum() == W * h
        image[y0:y0 + h] = part["image"][y0:y0 + h]
        depth[y0:y0 + h] = part["depth"][y0:y0 + h]
        sg += part["sigma_grad"]; cg += part["color_grad"]
    U.assert_bits(image, full["image"], "tiled image")
    U.assert_bits(depth, full["depth"], "tiled depth")
    st, odesc = O.plan_resolve(full_desc)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    ref = O.render(odesc, gs, gc, dl_full, shadow=True)
    U.assert_grads(full["sigma_grad"], full["color_grad"], ref, "full frame")
    U.assert_grads(sg, cg, ref, "sum of the tiles")


@pytest.mark.parametrize("slow_axis", [0, 1, 2])
@pytest.mark.parametrize("mode", sorted(SCATTER_MODES))
def test_gradient_block_axis_orders(ctx, slow_axis, mode):
    """hpx_grid_set_grad_layout: any axis may be the slowest one of the gradient block (contiguous slabs for the in-place
    pipelined all-reduce); hpx_grid_read_grad always returns the reference order."""
    sig, col = S.hashed_volume((9, 14, 11), "dense")
    desc = S.bench_plan(70, 45, 64, stratified=True, view=2, views=11)
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    dl = S.hashed_image_grad(70 * 45)
    ref = O.render(odesc, gs, gc, dl, shadow=True)
    plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
    slab_floats, slabs = grid.set_grad_layout(slow_axis)
    assert slabs == (9, 14, 11)[slow_axis] and slab_floats * slabs == 9 * 14 * 11 * 4
    frame.forward(grid)
    frame.backward(grid, dl, D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
    sg, cg, _ = grid.read_grad()
    U.assert_grads(sg, cg, ref, f"slow axis {slow_axis}")
    frame.close(); grid.close(); plan.close()


def test_interleaved_rows_and_box_backward_reproduce_full_frame(ctx):
    """Strong-scaling building blocks (hp_b200.h): 2 row groups x 2 'ranks' that own interleaved CTA tile rows, each
    scattering into the dense voxel box hpx_frame_bounds reports.  Run one after the other on one GPU, the pieces must
    tile the plain full-frame result: images bit for bit, gradients within the tolerance, no contribution dropped."""
    import sharding as SH
    sig, col = S.hashed_volume(28, "dense")
    W, Hh, steps, world, groups = 72, 61, 96, 2, 2
    full_desc = S.bench_plan(W, Hh, steps, stratified=True, view=1, views=7)
    dl_full = S.hashed_image_grad(W * Hh)
    full = run_lean(ctx, full_desc, sig, col, 1, 0, None, None, dl_full)
    st, odesc = O.plan_resolve(full_desc)
    ogs, ogc = U.oracle_grids(sig, col, 1, 0)
    ref = O.render(odesc, ogs, ogc, dl_full, shadow=True)
    lib = ctx.lib
    grid = D.Grid(ctx, sig, col)
    grid.zero_grad()
    nx = ny = nz = 28
    G = np.zeros((nz, ny, nx, 4), np.float64)
    image = np.zeros_like(full["image"]); hit = np.zeros_like(full["hitmask"])
    total_samples = total_live = 0
    for band in SH.row_bands(full_desc, groups, align=16):
        n_rays = band.rows * W
        dl = np.ascontiguousarray(dl_full[band.ray_index_base: band.ray_index_base + n_rays])
        d_dl = C.c_void_p()
        D.check("alloc", lib.hpx_device_alloc(ctx.handle, dl.nbytes, C.byref(d_dl)))
        D.check("h2d", lib.hpx_copy_to_device(ctx.handle, d_dl, dl.ctypes.data, dl.nbytes))
        for rank in range(world):
            plan = D.Plan(ctx, SH.band_desc(full_desc, band)); frame = D.Frame(plan)
            frame.set_view(None, plan.desc.seed, band.ray_index_base)
            frame.set_interleave(world, rank)
            frame.forward(grid)
            part = frame.read(); cnt = frame.counts()
            total_samples += cnt["samples"]; total_live += cnt["live_samples"]
            own = part["hitmask"] == 1
            assert not (hit.astype(bool) & own).any()          # ranks own disjoint pixels
            image[own] = part["image"][own]; hit[own] = 1
            box = frame.bounds(grid)
            x0, y0, z0, bx, by, bz = box
            assert bx > 0 and by > 0 and bz > 0 and x0 + bx <= nx and y0 + by <= ny and z0 + bz <= nz
            d_box = C.c_void_p()
            D.check("alloc", lib.hpx_device_alloc(ctx.handle, bx * by * bz * 16, C.byref(d_box)))
            frame.backward_box(grid, d_dl.value, box, d_box.value)
            assert frame.box_misses() == 0
            host = np.zeros((bz, by, bx, 4), np.float32)
            D.check("d2h", lib.hpx_copy_to_host(ctx.handle, host.ctypes.data, d_box, host.nbytes))
            G[z0:z0 + bz, y0:y0 + by, x0:x0 + bx] += host
            grid.add_box(ctx, d_box.value, box)            # device hand-over: gradient block += box, box = 0
            D.check("d2h", lib.hpx_copy_to_host(ctx.handle, host.ctypes.data, d_box, host.nbytes))
            assert not host.any()
            lib.hpx_device_free(ctx.handle, d_box)
            frame.close(); plan.close()
        lib.hpx_device_free(ctx.handle, d_dl)
    sg_dev, cg_dev, _ = grid.read_grad()                   # what hpx_grid_add_box accumulated on the device
    U.assert_grads(sg_dev, cg_dev, ref, "add_box")
    grid.close()
    assert hit.all() and total_samples == full["samples"] and total_live == full["live_samples"]
    U.assert_bits(image, full["image"], "interleaved image")
    U.assert_grads(G[..., 3].reshape(-1), G[..., :3].reshape(-1), ref, "boxed")


def test_lean_forward_is_deterministic_and_graph_replay_matches(ctx):
    sig, col = S.hashed_volume(32, "dense")
    desc = S.bench_plan(72, 40, 128, stratified=True)
    n = 72 * 40
    dl = S.hashed_image_grad(n)
    plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
    frame.forward(grid); a = frame.read(); ca = frame.counts()
    frame.forward(grid); b = frame.read()
    for k in a:
        U.assert_bits(a[k], b[k], "rerun " + k)
    frame.backward(grid, dl); sg_a, cg_a, _ = grid.read_grad()
    # real CUDA graph: forward + backward captured once; dL/dI is read from the frame-owned device
    # buffer, which the HOST-memspace hpx_backward above has already filled with `dl`
    frame.capture(grid, D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO)
    for _ in range(2):
        frame.replay()
    c = frame.read(); cc = frame.counts()
    sg_c, cg_c, _ = grid.read_grad()
    for k in a:
        U.assert_bits(a[k], c[k], "graph " + k)
    assert ca == cc
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    ref = O.render(odesc, gs, gc, dl, shadow=True)
    U.assert_grads(sg_a, cg_a, ref, "eager")
    U.assert_grads(sg_c, cg_c, ref, "graph replay")
    frame.close(); grid.close(); plan.close()


@pytest.mark.parametrize("scatter", [0, D.HPX_BACKWARD_SCATTER_PER_RAY, D.HPX_BACKWARD_SCATTER_MERGED],
                         ids=["auto", "separate_kernel", "fused_in_merged_backward"])
@pytest.mark.parametrize("strat,oob", [(False, A.HP_OOB_ZERO), (True, A.HP_OOB_ZERO), (True, A.HP_OOB_CLAMP)])
def test_camera_gradient_matches_pinned_adjoint(ctx, strat, oob, scatter):
    """Both evaluations of the camera adjoint -- the stand-alone forward-order kernel and the one fused into the
    merged backward (reverse order, from the corners that kernel has loaded anyway) -- against the oracle."""
    sig, col = S.smooth_volume(40)
    W = Hh = 36
    desc = S.bench_plan(W, Hh, 128, stratified=strat, view=1, views=9)
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, oob)
    dl = S.hashed_image_grad(W * Hh) + np.float32(0.25)
    ref, mag = O.camera_grad(odesc, gs, gc, dl, with_mag=True)
    got = run_lean(ctx, desc, sig, col, 1, oob, None, None, dl,
                   flags=D.HPX_BACKWARD_CAMERA | D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | scatter)
    U.assert_camera_close(got["camera_grad"], ref, mag, "camera gradient")
    # the grid gradient is unaffected by asking for the camera gradient too
    ref_grid = O.render(odesc, gs, gc, dl, shadow=True)
    U.assert_grads(got["sigma_grad"], got["color_grad"], ref_grid, "grid gradient next to the camera adjoint")


def test_full_size_config2_properties(ctx):
    """BASELINE config 2 at full size (1024x1024, 256^3, 512 stratified steps, 537 M samples):
    too big for the CPU oracle, so check size-independent properties plus an oracle band."""
    n_grid, W, steps = 256, 1024, 512
    sig, col = S.hashed_volume(n_grid, "dense")
    desc = S.bench_plan(W, W, steps, stratified=True)
    plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
    frame.forward(grid)
    full = frame.read(); counts = frame.counts()
    assert counts["samples"] == W * W * steps and counts["rays"] == W * W
    assert 0 < counts["live_samples"] < counts["samples"]
    assert full["hitmask"].all()
    assert np.isfinite(full["image"]).all() and (full["trans"] >= 0).all() and (full["trans"] <= 1).all()
    np.testing.assert_allclose(full["opacity"], 1.0 - full["trans"], atol=1e-6)
    # an 8-row band through the middle against the oracle (8*1024 rays * 512 steps = 4.2 M samples)
    y0, h = 508, 8
    band_desc = S.bench_plan(W, W, steps, stratified=True, roi=(0, y0, W, h))
    st, oband = O.plan_resolve(band_desc)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    ref = O.render(oband, gs, gc, ray_index_base=y0 * W, per_ray=False, frames=True)
    U.assert_close(full["image"][y0:y0 + h], ref["image"][y0:y0 + h], U.IMAGE_RTOL, "band image")
    U.assert_close(full["depth"][y0:y0 + h], ref["depth"][y0:y0 + h], U.IMAGE_RTOL, "band depth")
    # backward: gradients of a constant dL/dI = 1: colour-gradient mass equals the sum of weights,
    # i.e. sum over voxels of d/dc_r == sum over pixels of opacity (trilinear weights sum to 1)
    dl = np.ones((W * W, 3), np.float32)
    frame.backward(grid, dl)
    sg, cg, _ = grid.read_grad()
    assert np.isfinite(sg).all() and np.isfinite(cg).all()
    mass = cg.reshape(-1, 3).astype(np.float64).sum(axis=0)
    expect = full["opacity"].astype(np.float64).sum()
    np.testing.assert_allclose(mass, [expect] * 3, rtol=2e-4)
    frame.close(); grid.close(); plan.close()


# ---- gradient parity at the FULL BASELINE grid / image sizes, through ROI bands the oracle renders in seconds ----------
_VOLUMES = {}


def _volume(n, kind):
    """Hashed volumes are expensive at 512^3 (host hashing): build each one once per session."""
    if (n, kind) not in _VOLUMES:
        _VOLUMES.clear()                       # one big volume at a time
        _VOLUMES[(n, kind)] = S.hashed_volume(n, kind)
    return _VOLUMES[(n, kind)]


def _band_case(ctx, n_grid, W, steps, strat, kind, y0, h, flags, view=0, views=1, camera=False):
    sig, col = _volume(n_grid, kind)
    desc = S.bench_plan(W, W, steps, stratified=strat, view=view, views=views, roi=(0, y0, W, h))
    st, oband = O.plan_resolve(desc)
    assert st == 0
    dl = S.hashed_image_grad(W * h)
    got = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl, flags=flags, ray_index_base=y0 * W)
    assert bytes(got["desc"]) == bytes(oband)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    ref = O.render(oband, gs, gc, dl, ray_index_base=y0 * W, per_ray=False, frames=True, shadow=True, shadow_box=got["box"])
    what = f"{n_grid}^3 / {W} px / {steps} steps / rows {y0}..{y0 + h} / {kind}"
    assert got["samples"] == ref["sample_count"] == W * h * steps
    assert got["live_samples"] == ref["live_sample_count"], what
    if kind == "dense":
        assert ref["live_sample_count"] < ref["sample_count"]
    U.assert_bits(got["hitmask"], ref["hitmask"], what + " hitmask")
    for k in ("image", "trans", "opacity", "depth"):
        U.assert_close(got[k][y0:y0 + h], ref[k][y0:y0 + h], U.IMAGE_RTOL, f"{what} {k}")
    n_adj = U.assert_grads(got["sigma_grad"], got["color_grad"], ref, what, res=(n_grid, n_grid, n_grid))
    if camera:
        cref, cmag = O.camera_grad(oband, gs, gc, dl, ray_index_base=y0 * W, with_mag=True)
        U.assert_camera_close(got["camera_grad"], cref, cmag, what + " camera")
    return n_adj


@pytest.mark.parametrize("mode", sorted(SCATTER_MODES))
@pytest.mark.parametrize("kind,y0", [("thin", 508), ("dense", 96)])
def test_config2_full_size_band_gradients(ctx, kind, y0, mode):
    """BASELINE config 2 (256^3 grid, 1024 px wide, 512 stratified steps): an 8-row band (4.2 M samples) of the
    full-size plan, forward AND grid gradients against the oracle, both scatter kernels, a band through the centre
    (thin volume, no early stop) and one near the top edge (dense volume, early termination, oblique rays)."""
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode]
    _band_case(ctx, 256, 1024, 512, True, kind, y0, 8, flags)


@pytest.mark.parametrize("y0", [1022, 40])
def test_config3_full_size_band_gradients(ctx, y0):
    """BASELINE config 3 (512^3 grid, 2048 px wide, 1024 fixed steps): a 4-row band (8.4 M samples) of the full-size plan
    through the merged backward -- the kernel the 32 768-CTA frame runs -- against the oracle."""
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_SCATTER_MERGED
    _band_case(ctx, 512, 2048, 1024, False, "thin", y0, 4, flags)


@pytest.mark.parametrize("view,y0", [(0, 396), (13, 120), (29, 700)])
def test_config4_full_size_band_camera_and_grid_gradients(ctx, view, y0):
    """BASELINE config 4 (64 views at 800x800 over a 256^3 grid, stratified, 512 steps): an 8-row band of three of the
    views through the merged backward with the FUSED camera adjoint (kCamera): grid gradients against the oracle, camera
    gradients (c2w and intrinsics) against the oracle's analytic adjoint."""
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_CAMERA | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_SCATTER_MERGED
    _band_case(ctx, 256, 800, 512, True, "thin", y0, 8, flags, view=view, views=64, camera=True)


def test_camera_gradient_clamp_policy_full_width(ctx):
    """The fused and the stand-alone camera adjoint with the CLAMP policy at config-4 width (800 px, 128^3 smooth volume,
    8-row band): both evaluations against the oracle's analytic adjoint."""
    sig, col = S.smooth_volume(128)
    W, y0, h = 800, 300, 8
    desc = S.bench_plan(W, W, 256, stratified=True, view=5, views=64, roi=(0, y0, W, h))
    st, oband = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, A.HP_OOB_CLAMP)
    dl = S.hashed_image_grad(W * h) + np.float32(0.25)
    cref, cmag = O.camera_grad(oband, gs, gc, dl, ray_index_base=y0 * W, with_mag=True)
    for scatter in (D.HPX_BACKWARD_SCATTER_PER_RAY, D.HPX_BACKWARD_SCATTER_MERGED):
        got = run_lean(ctx, desc, sig, col, 1, A.HP_OOB_CLAMP, None, None, dl,
                       flags=D.HPX_BACKWARD_CAMERA | D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | scatter, ray_index_base=y0 * W)
        U.assert_camera_close(got["camera_grad"], cref, cmag, f"clamp camera gradient (scatter flag {scatter:#x})")


@pytest.mark.parametrize("mode", sorted(SCATTER_MODES))
@pytest.mark.parametrize("strat", [False, True])
def test_scatter_modes_sparse_pixels(ctx, mode, strat):
    """Pixels much SPARSER than voxels (3 voxels between neighbouring rays, 0.5 voxel between steps): the regime
    where the merged kernel merges only along a ray; both kernels must still be exact."""
    sig, col = S.hashed_volume(64, "thin")
    desc = S.bench_plan(27, 22, 128, stratified=strat, view=1, views=9)
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    dl = np.ones((27 * 22, 3), np.float32)
    ref = O.render(odesc, gs, gc, dl, shadow=True)
    got = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl,
                   flags=D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
    check_forward(got, ref, "sparse")
    np.testing.assert_allclose(got["color_grad"].reshape(-1, 3).astype(np.float64).sum(axis=0),
                               [ref["opacity"].astype(np.float64).sum()] * 3, rtol=1e-4)
    U.assert_grads(got["sigma_grad"], got["color_grad"], ref, mode)


def test_signalled_backward_counts_every_cta_and_matches_plain(ctx):
    """hpx_backward_signalled: one launch, a device counter per group of tile rows.  After the stream drains every counter
    equals the expected CTA count, a second context's stream can wait on them (cuStreamWaitValue32), and the gradient
    equals the plain backward's."""
    sig, col = S.hashed_volume(24, "dense")
    W, Hh, steps = 150, 131, 64
    desc = S.bench_plan(W, Hh, steps, stratified=True, view=1, views=6)
    dl = S.hashed_image_grad(W * Hh)
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    ref = O.render(odesc, gs, gc, dl, shadow=True)
    lib = ctx.lib
    for world, rank in ((1, 0), (3, 1)):
        plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
        frame.set_interleave(world, rank)
        frame.forward(grid)
        d_dl = C.c_void_p()
        D.check("alloc", lib.hpx_device_alloc(ctx.handle, dl.nbytes, C.byref(d_dl)))
        D.check("h2d", lib.hpx_copy_to_device(ctx.handle, d_dl, np.ascontiguousarray(dl).ctypes.data, dl.nbytes))
        grid.zero_grad()
        counters = frame.reset_group_counters()
        tile_rows = (Hh + 7) // 8
        owned = (tile_rows - rank + world - 1) // world
        ends = [owned // 3, 2 * owned // 3, owned]
        ptr, expected = frame.backward_signalled(grid, d_dl.value, ends, D.HPX_BACKWARD_GRID)
        assert ptr == counters and sum(expected) == owned * ((W + 15) // 16)
        ctx.synchronize()
        got = np.zeros(8, np.uint32)
        D.check("d2h", lib.hpx_copy_to_host(ctx.handle, got.ctypes.data, C.c_void_p(counters), 32))
        assert list(got[:3]) == expected and not got[3:].any()
        # only now (the counts are known to be there, so this cannot hang): a second context's stream waits on a counter
        other = D.Context(device=0)
        other.wait_counter(counters + 8, expected[2])
        other.synchronize()
        if world == 1:
            sg, cg, _ = grid.read_grad()
            U.assert_grads(sg, cg, ref, "signalled")
        other.close()
        lib.hpx_device_free(ctx.handle, d_dl)
        frame.close(); grid.close(); plan.close()


def test_grid_1024_cubed_maximum_size_properties(ctx):
    """BASELINE config 5's grid: 1024^3 voxels (17.2 GB packed + 17.2 GB gradient), the largest axis the merged kernel's
    10-bit cell keys allow and 2^30 voxels for the 32-bit voxel indices.  Far beyond the CPU oracle, so the check is by
    size-independent properties; the volume is generated on the device and handed over as DEVICE arrays."""
    torch = pytest.importorskip("torch")
    free, _total = torch.cuda.mem_get_info()
    if free < 80 * (1 << 30):
        pytest.skip("needs ~80 GB of free HBM")
    n, W, steps = 1024, 256, 1024
    gen = torch.Generator(device="cuda").manual_seed(5)
    sigma = torch.rand((n, n, n), generator=gen, device="cuda", dtype=torch.float32) * 3.0
    color = torch.rand((n, n, n, 3), generator=gen, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    grid = D.Grid(ctx, sigma.data_ptr(), color.data_ptr(), device_shape=(n, n, n))
    ctx.synchronize()
    corner = (float(sigma[-1, -1, -1]), [float(v) for v in color[-1, -1, -1]])
    del sigma, color
    torch.cuda.empty_cache()
    desc = S.bench_plan(W, W, steps, stratified=False)
    plan = D.Plan(ctx, desc); frame = D.Frame(plan)
    assert frame.scatter_mode(grid) == "per_ray"          # 4 voxels between neighbouring rays: nothing to merge
    frame.forward(grid)
    out = frame.read(); counts = frame.counts()
    assert counts["samples"] == W * W * steps and counts["live_samples"] == counts["samples"]
    assert np.isfinite(out["image"]).all() and (out["trans"] > 0).all() and (out["trans"] < 1).all()
    np.testing.assert_allclose(out["opacity"], 1.0 - out["trans"], atol=1e-6)
    dl = np.ones((W * W, 3), np.float32)
    for mode in ("per_ray", "merged"):
        frame.backward(grid, dl, D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
        ctx.synchronize()   # the library has its own stream: torch must not read the block before the kernel is done
        # colour-gradient mass = sum of the weights = sum of opacity (trilinear weights sum to 1), through a device reduction
        ptr, floats = grid.grad_buffer()
        class _V:  # zero-copy view of the packed gradient block
            __cuda_array_interface__ = {"shape": (floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        block = torch.as_tensor(_V(), device="cuda")
        g4 = block[: n * n * n * 4].view(-1, 4)
        mass = g4[:, :3].sum(dim=0, dtype=torch.float64).cpu().numpy()
        assert torch.isfinite(g4).all()
        np.testing.assert_allclose(mass, [out["opacity"].astype(np.float64).sum()] * 3, rtol=2e-4)
        # the far corner voxel (index 2^30 - 1) is reachable: its gradient slot exists and is finite
        assert np.isfinite(g4[-1].cpu().numpy()).all()
    assert np.isfinite(corner[0]) and len(corner[1]) == 3
    frame.close(); grid.close(); plan.close()
