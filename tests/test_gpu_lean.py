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
    ref = O.render(odesc, gs, gc, dl, case["res"], case["bmin"], case["bmax"])
    got = run_lean(ctx, desc, case["sigma"], case["color"], case["interp"], case["oob"], case["bmin"], case["bmax"], dl)
    assert bytes(got["desc"]) == bytes(odesc)
    check_forward(got, ref, f"case{case['case']}")
    U.assert_close(got["sigma_grad"], ref["sigma_grad"], U.GRAD_RTOL, "sigma_grad")
    U.assert_close(got["color_grad"], ref["color_grad"], U.GRAD_RTOL, "color_grad")


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
    ref = O.render(odesc, gs, gc, dl, case["res"], case["bmin"], case["bmax"])
    got = run_lean(ctx, desc, case["sigma"], case["color"], case["interp"], case["oob"], case["bmin"], case["bmax"], dl,
                   flags=D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
    U.assert_close(got["sigma_grad"], ref["sigma_grad"], U.GRAD_RTOL, f"{mode} sigma_grad")
    U.assert_close(got["color_grad"], ref["color_grad"], U.GRAD_RTOL, f"{mode} color_grad")


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
    ref = O.render(odesc, gs, gc, dl)
    got = run_lean(ctx, desc, sig, col, 1, oob, None, None, dl,
                   flags=D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
    U.assert_close(got["sigma_grad"], ref["sigma_grad"], U.GRAD_RTOL, f"{mode} sigma_grad")
    U.assert_close(got["color_grad"], ref["color_grad"], U.GRAD_RTOL, f"{mode} color_grad")


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
        ref = O.render(odesc, gs, gc, dl)
        got = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl,
                       flags=D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
        check_forward(got, ref, str(shape))
        U.assert_close(got["sigma_grad"], ref["sigma_grad"], U.GRAD_RTOL, f"{shape} sigma_grad")
        U.assert_close(got["color_grad"], ref["color_grad"], U.GRAD_RTOL, f"{shape} color_grad")


@pytest.mark.parametrize("path", U.golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_lean_matches_reference_golden(ctx, path):
    g = U.load_golden(path)
    got = run_lean(ctx, g["desc_in"], g["sigma"], g["color"], g["interp"], g["oob"], g["bmin"], g["bmax"], g["dL_dI"])
    assert bytes(got["desc"]) == bytes(g["desc_resolved"])
    assert got["samples"] == int(g["sample_count"])
    U.assert_bits(got["hitmask"], g["img_hitmask"], "hitmask")
    for k in ("image", "trans", "opacity", "depth"):
        U.assert_close(got[k], g[f"img_{k}"], U.IMAGE_RTOL, k)
    U.assert_close(got["sigma_grad"], g["sigma_grad"], U.GRAD_RTOL, "sigma_grad")
    U.assert_close(got["color_grad"], g["color_grad"], U.GRAD_RTOL, "color_grad")


@pytest.mark.parametrize("kind,strat", [("thin", False), ("dense", True), ("dense", False)])
def test_lean_hashed_volumes_medium(ctx, kind, strat):
    """BASELINE config shapes scaled to what the oracle finishes in seconds: 96x80, 48^3, 160 steps."""
    sig, col = S.hashed_volume(48, kind)
    desc = S.bench_plan(96, 80, 160, stratified=strat, view=2, views=9)
    st, odesc = O.plan_resolve(desc)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    dl = S.hashed_image_grad(96 * 80)
    ref = O.render(odesc, gs, gc, dl)
    got = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl)
    check_forward(got, ref, kind)
    if kind == "dense":
        assert ref["live_sample_count"] < ref["sample_count"]   # early termination is exercised
    U.assert_close(got["sigma_grad"], ref["sigma_grad"], U.GRAD_RTOL, "sigma_grad")
    U.assert_close(got["color_grad"], ref["color_grad"], U.GRAD_RTOL, "color_grad")


def test_lean_backward_accumulates_and_is_linear(ctx):
    sig, col = S.hashed_volume(24, "thin")
    desc = S.bench_plan(40, 36, 64, stratified=True)
    n = 40 * 36
    dl = S.hashed_image_grad(n)
    a = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl)
    b = run_lean(ctx, desc, sig, col, 1, 0, None, None, 2.0 * dl)
    # scaling dL/dI by 2 is exact on every product of the adjoint; what remains is the order of the
    # float atomics in the scatter, which differs from run to run -> the gradient tolerance applies
    U.assert_close(b["sigma_grad"], 2.0 * a["sigma_grad"], U.GRAD_RTOL, "linearity sigma")
    U.assert_close(b["color_grad"], 2.0 * a["color_grad"], U.GRAD_RTOL, "linearity color")
    # without HPX_BACKWARD_ZERO gradients accumulate (DenseGridField semantics, dense_grid.cpp:166-169)
    plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
    frame.forward(grid)
    frame.backward(grid, dl, D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO)
    frame.backward(grid, dl, D.HPX_BACKWARD_GRID)
    sg, cg, _ = grid.read_grad()
    U.assert_close(sg, 2.0 * a["sigma_grad"], U.GRAD_RTOL, "accumulate sigma")
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
    ref = O.render(odesc, gs, gc, dl)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_DETERMINISTIC | SCATTER_MODES[mode]
    runs = [run_lean(ctx, desc, sig, col, 1, 0, None, None, dl, flags=flags) for _ in range(3)]
    for r in runs[1:]:
        U.assert_bits(r["sigma_grad"], runs[0]["sigma_grad"], "deterministic sigma_grad")
        U.assert_bits(r["color_grad"], runs[0]["color_grad"], "deterministic color_grad")
    U.assert_close(runs[0]["sigma_grad"], ref["sigma_grad"], U.GRAD_RTOL, "sigma_grad")
    U.assert_close(runs[0]["color_grad"], ref["color_grad"], U.GRAD_RTOL, "color_grad")
    plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
    frame.forward(grid)
    frame.backward(grid, dl, flags)
    frame.backward(grid, dl, flags & ~D.HPX_BACKWARD_ZERO)
    sg, cg, _ = grid.read_grad()
    U.assert_close(sg, 2.0 * ref["sigma_grad"], U.GRAD_RTOL, "accumulated sigma_grad")
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
    U.assert_close(sg, full["sigma_grad"], U.GRAD_RTOL, "tiled sigma_grad")
    U.assert_close(cg, full["color_grad"], U.GRAD_RTOL, "tiled color_grad")


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
    ref = O.render(odesc, gs, gc, dl)
    plan = D.Plan(ctx, desc); grid = D.Grid(ctx, sig, col); frame = D.Frame(plan)
    slab_floats, slabs = grid.set_grad_layout(slow_axis)
    assert slabs == (9, 14, 11)[slow_axis] and slab_floats * slabs == 9 * 14 * 11 * 4
    frame.forward(grid)
    frame.backward(grid, dl, D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
    sg, cg, _ = grid.read_grad()
    U.assert_close(sg, ref["sigma_grad"], U.GRAD_RTOL, "sigma_grad")
    U.assert_close(cg, ref["color_grad"], U.GRAD_RTOL, "color_grad")
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
    U.assert_close(sg_dev, full["sigma_grad"], U.GRAD_RTOL, "add_box sigma_grad")
    U.assert_close(cg_dev, full["color_grad"], U.GRAD_RTOL, "add_box color_grad")
    grid.close()
    assert hit.all() and total_samples == full["samples"] and total_live == full["live_samples"]
    U.assert_bits(image, full["image"], "interleaved image")
    U.assert_close(G[..., 3].reshape(-1), full["sigma_grad"], U.GRAD_RTOL, "boxed sigma_grad")
    U.assert_close(G[..., :3].reshape(-1), full["color_grad"], U.GRAD_RTOL, "boxed color_grad")


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
    U.assert_close(sg_c, sg_a, U.GRAD_RTOL, "graph sigma_grad")
    U.assert_close(cg_c, cg_a, U.GRAD_RTOL, "graph color_grad")
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
    ref = O.camera_grad(odesc, gs, gc, dl)
    got = run_lean(ctx, desc, sig, col, 1, oob, None, None, dl,
                   flags=D.HPX_BACKWARD_CAMERA | D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | scatter)
    cam = got["camera_grad"].astype(np.float64)
    scale = np.abs(ref[:12]).max()
    assert np.all(np.abs(cam[:12] - ref[:12]) <= 1e-4 * np.maximum(np.abs(ref[:12]), 0.05 * scale)), (cam[:12], ref[:12])
    kscale = np.abs(ref[12:]).max()
    assert np.all(np.abs(cam[12:] - ref[12:]) <= 1e-4 * np.maximum(np.abs(ref[12:]), 0.05 * kscale)), (cam[12:], ref[12:])
    # the grid gradient is unaffected by asking for the camera gradient too
    ref_grid = O.render(odesc, gs, gc, dl)
    U.assert_close(got["sigma_grad"], ref_grid["sigma_grad"], U.GRAD_RTOL, "sigma_grad")


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
    ref = O.render(odesc, gs, gc, dl)
    got = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl,
                   flags=D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | SCATTER_MODES[mode])
    check_forward(got, ref, "sparse")
    np.testing.assert_allclose(got["color_grad"].reshape(-1, 3).astype(np.float64).sum(axis=0),
                               [ref["opacity"].astype(np.float64).sum()] * 3, rtol=1e-4)
    U.assert_close(got["sigma_grad"], ref["sigma_grad"], U.GRAD_RTOL, f"{mode} sigma_grad")
    U.assert_close(got["color_grad"], ref["color_grad"], U.GRAD_RTOL, f"{mode} color_grad")


def test_signalled_backward_counts_every_cta_and_matches_plain(ctx):
    """hpx_backward_signalled: one launch, a device counter per group of tile rows.  After the stream drains every counter
    equals the expected CTA count, a second context's stream can wait on them (cuStreamWaitValue32), and the gradient
    equals the plain backward's."""
    sig, col = S.hashed_volume(24, "dense")
    W, Hh, steps = 150, 131, 64
    desc = S.bench_plan(W, Hh, steps, stratified=True, view=1, views=6)
    dl = S.hashed_image_grad(W * Hh)
    plain = run_lean(ctx, desc, sig, col, 1, 0, None, None, dl)
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
            U.assert_close(sg, plain["sigma_grad"], U.GRAD_RTOL, "signalled sigma_grad")
            U.assert_close(cg, plain["color_grad"], U.GRAD_RTOL, "signalled color_grad")
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
