"""N > 1 path on CPU: world_size-2 `gloo` processes run the host-side sharding logic
(diff-volume-renderer_b200/python/sharding.py) end to end.  Each rank renders its row band with
the CPU oracle standing in for the GPU (tests may use the oracle as the checker AND as the per-rank
worker here: what is under test is the partition, the global ray-index base, the packed-gradient
all-reduce and the image gather -- not the kernels, which tests/test_gpu_lean.py covers)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import hp_abi as A
import oracle as O
import sharding as SH
import synth as S
import util as U


# ---- pure host logic ----------------------------------------------------------------------------
@pytest.mark.parametrize("height,world", [(48, 2), (1024, 8), (20, 8), (7, 2), (2048, 4), (100, 3)])
def test_row_bands_partition(height, world):
    desc = S.bench_plan(64, height, 16, stratified=True)
    bands = SH.row_bands(desc, world)
    assert len(bands) == world
    assert sum(b.rows for b in bands) == height
    y = 0
    for b in bands:
        assert b.y0 == y or b.empty
        assert b.ray_index_base == (b.y0 if not b.empty else min(b.y0, height)) * 64 or b.empty
        if not b.empty:
            assert b.y0 % SH.TILE_ROWS == 0           # cut on CTA tile rows
        y += b.rows
    sizes = [b.rows for b in bands if not b.empty]
    assert max(sizes) - min(sizes) < 2 * SH.TILE_ROWS  # one tile row of imbalance + the ragged last row


def test_row_bands_respect_parent_roi():
    desc = S.bench_plan(64, 64, 16, stratified=True, roi=(8, 16, 40, 24))
    bands = SH.row_bands(desc, 2)
    assert [(b.y0, b.rows, b.ray_index_base) for b in bands] == [(16, 16, 0), (32, 8, 16 * 40)]
    d = SH.band_desc(desc, bands[1])
    assert (d.roi.x, d.roi.y, d.roi.width, d.roi.height) == (8, 32, 40, 8)
    assert d.width == 64 and d.height == 64 and d.seed == desc.seed


def test_views_of_rank_and_u32_bands():
    assert [SH.views_of_rank(64, 8, r) for r in (0, 7)] == [list(range(0, 8)), list(range(56, 64))]
    got = sum((SH.views_of_rank(10, 4, r) for r in range(4)), [])
    assert got == list(range(10))
    c3 = S.bench_plan(2048, 2048, 1024, stratified=False)
    assert SH.u32_safe_bands(c3, 1024) == 2            # 2^32 samples do not fit one u32 plan (SURVEY finding 9)
    c2 = S.bench_plan(1024, 1024, 512, stratified=True)
    assert SH.u32_safe_bands(c2, 512) == 1


def test_weighted_row_bands():
    desc = S.bench_plan(64, 2048, 16, stratified=False)
    bands = SH.weighted_row_bands(desc, (0.75, 0.25), align=64)
    assert [(b.y0, b.rows) for b in bands] == [(0, 1536), (1536, 512)]
    assert bands[1].ray_index_base == 1536 * 64
    bands = SH.weighted_row_bands(desc, (1, 1, 1), align=8)
    assert sum(b.rows for b in bands) == 2048 and all(b.y0 % 8 == 0 for b in bands)
    ragged = SH.weighted_row_bands(S.bench_plan(64, 100, 16, stratified=False), (0.9, 0.1), align=16)
    assert sum(b.rows for b in ragged) == 100 and ragged[0].rows % 16 == 0


def test_final_slab_runs_cover_every_touched_slab_once():
    # monotone ranges with overlap (the usual case), a group that touches nothing, and a reversed order
    for ranges in ([(0, 10), (7, 18), (15, 30), (27, 40)], [(0, 10), None, (8, 20)], [(30, 40), (18, 33), (5, 20), (0, 8)],
                   [(3, 9), (3, 9)], [None, None]):
        runs = SH.final_slab_runs(ranges)
        assert len(runs) == len(ranges)
        seen = []
        for g, rs in enumerate(runs):
            later = set()
            for r in ranges[g + 1:]:
                if r:
                    later.update(range(*r))
            for a, b in rs:
                assert a < b
                assert not (set(range(a, b)) & later), "a slab was declared final while a later group still writes it"
                seen += list(range(a, b))
        touched = set()
        for r in ranges:
            if r:
                touched.update(range(*r))
        assert sorted(seen) == sorted(touched)          # each touched slab exactly once
    assert SH.final_slab_runs([(0, 10), (7, 18)]) == [[(0, 7)], [(7, 18)]]


# ---- two gloo ranks -----------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, Hh, steps, n = 40, 24, 48, 20
        sig, col = S.hashed_volume(n, "dense")
        gs, gc = U.oracle_grids(sig, col, A.HP_INTERP_LINEAR, A.HP_OOB_ZERO)
        full = S.bench_plan(W, Hh, steps, stratified=True, view=1, views=5)
        st, full_res = O.plan_resolve(full)
        assert st == 0
        dl_full = S.hashed_image_grad(W * Hh)
        band = SH.row_bands(full, world)[rank]
        st, bdesc = O.plan_resolve(SH.band_desc(full, band))
        assert st == 0
        dl = dl_full[band.ray_index_base: band.ray_index_base + band.rows * W]
        part = O.render(bdesc, gs, gc, dl, ray_index_base=band.ray_index_base)
        cam = O.camera_grad(bdesc, gs, gc, dl, ray_index_base=band.ray_index_base)
        V = n ** 3
        # packed gradient block exactly as hpx_grid_grad_buffer lays it out: [V x {dr,dg,db,dsigma} | 16 camera]
        block = np.zeros(4 * V + 16, np.float32)
        block[:4 * V].reshape(V, 4)[:, :3] = part["color_grad"].reshape(V, 3)
        block[:4 * V].reshape(V, 4)[:, 3] = part["sigma_grad"]
        block[4 * V:] = cam.astype(np.float32)
        t = torch.from_numpy(block)
        red = SH.GradientAllReduce(t)
        red()
        assert red.bytes_on_wire_per_rank == block.nbytes      # 2 (N-1)/N at N = 2
        # image: every rank contributes its rows (disjoint pixels, no reduction needed)
        img = torch.from_numpy(part["image"].copy())
        dist.all_reduce(img)                                   # zeros outside the band
        samples = torch.tensor([part["sample_count"], part["live_sample_count"]], dtype=torch.int64)
        dist.all_reduce(samples)
        if rank == 0:
            ref = O.render(full_res, gs, gc, dl_full, shadow=True)
            ref_cam, cam_mag = O.camera_grad(full_res, gs, gc, dl_full, with_mag=True)
            ok_img = U.bits_equal(img.numpy(), ref["image"])
            sg = block[:4 * V].reshape(V, 4)[:, 3]
            cg = block[:4 * V].reshape(V, 4)[:, :3].reshape(-1)
            U.assert_grads(sg, cg, ref, "sharded")
            U.assert_camera_close(block[4 * V:], ref_cam, cam_mag, "sharded camera gradient")
            q.put(("ok", ok_img, int(samples[0]) == ref["sample_count"], int(samples[1]) == ref["live_sample_count"]))
    except Exception as e:  # surface the failure in the parent
        if rank == 0:
            q.put(("fail", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_step_matches_unsharded():
    import torch.multiprocessing as mp
    O.build_oracle()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = q.get(timeout=5)
    assert res[0] == "ok", res
    assert res[1], "sharded image differs from the unsharded one (must be bit-exact)"
    assert res[2] and res[3], "sample counts do not add up"


# ---- balanced bands of the C library (host-only entry point, no GPU needed) ---------------------
@pytest.mark.parametrize("width,height,steps,world,roi", [(2048, 2048, 1024, 8, None), (1024, 1024, 512, 4, None), (1024, 1024, 512, 2, None),
                                                          (320, 200, 64, 3, (16, 24, 256, 150)), (64, 20, 16, 8, None)])
def test_library_balanced_bands(width, height, steps, world, roi):
    """hpx_plan_balanced_bands (what hpx_shard_create_bands cuts): contiguous, complete, on CTA tile rows, and equal
    in marching work -- in-cube steps of the band's rays, recounted here from the oracle's ray generator."""
    import dvren_b200 as D
    ctx = D.Context(device=0)
    desc = S.bench_plan(width, height, steps, stratified=False, roi=roi)
    plan = D.Plan(ctx, desc)
    row0, rows, work = (C.c_uint32 * world)(), (C.c_uint32 * world)(), (C.c_double * world)()
    D.check("hpx_plan_balanced_bands", ctx.lib.hpx_plan_balanced_bands(plan.handle, world, row0, rows, work))
    h = plan.desc.roi.height
    y = 0
    for r in range(world):
        assert row0[r] == y
        assert rows[r] % 8 == 0 or row0[r] + rows[r] == h
        y += rows[r]
    assert y == h
    # independent recount on a coarse pixel lattice: slab-method cube interval of pinhole rays from the same camera
    d = plan.desc
    K, c2w = np.array(d.camera.K[:], np.float64), np.array(d.camera.c2w[:], np.float64).reshape(3, 4)
    ys = np.arange(d.roi.y, d.roi.y + h, dtype=np.float64)
    xs = np.arange(d.roi.x, d.roi.x + d.roi.width, 8, dtype=np.float64)
    qx, qy = np.meshgrid((xs + 0.5 - K[2]) / K[0], (ys + 0.5 - K[5]) / K[4])
    v = np.stack([qx, qy, np.ones_like(qx)], -1) @ c2w[:, :3].T
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    o = c2w[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        a, b = (0.0 - o) / v, (1.0 - o) / v
    t_in = np.maximum(np.minimum(a, b).max(-1), d.t_near)
    t_out = np.minimum(np.maximum(a, b).min(-1), min(d.t_far, d.t_near + d.sampling.max_steps * d.sampling.dt))
    per_row = (np.clip(t_out - t_in, 0, None) / d.sampling.dt + 8.0).sum(-1)
    mine = np.array([per_row[row0[r]:row0[r] + rows[r]].sum() for r in range(world)])
    lib = np.array(work[:])
    busy = mine > 0
    np.testing.assert_allclose(lib[busy] / lib[busy].sum(), mine[busy] / mine[busy].sum(), rtol=0.03, atol=1e-3)
    if h >= 64 * world:   # enough tile rows to balance: nobody more than 6 % above the mean
        assert mine.max() <= 1.06 * mine.mean()
    plan.close(); ctx.close()


def _exchange_volumes(wedges, cuts, replicated):
    world = len(wedges)
    hull = (min(w[0] for w in wedges if w[0] < w[1]), max(w[1] for w in wedges if w[0] < w[1]))
    ov = lambda r, lo, hi: max(0, min(wedges[r][1], hi) - max(wedges[r][0], lo))
    out, inn = [], []
    for r in range(world):
        lo, hi = cuts[r], cuts[r + 1]
        o = max(0, (wedges[r][1] - wedges[r][0]) - ov(r, lo, hi))
        i = sum(ov(q, lo, hi) for q in range(world) if q != r)
        if replicated:
            own = max(0, min(hi, hull[1]) - max(lo, hull[0]))
            o, i = o + own * (world - 1), i + (hull[1] - hull[0]) - own
        out.append(o); inn.append(i)
    return out, inn


@pytest.mark.parametrize("replicated", [False, True])
@pytest.mark.parametrize("wedges,n", [
    ([(0, 127), (0, 170), (78, 213), (164, 258), (254, 348), (299, 434), (342, 512), (385, 512)], 512),   # c3 at 8 GPUs (measured)
    ([(0, 258), (254, 512)], 512), ([(0, 168), (74, 258), (254, 441), (345, 512)], 512),
    ([(10, 40), (0, 0), (35, 90)], 100), ([(0, 64)], 64)])
def test_library_owner_cuts(wedges, n, replicated):
    """hpx_plan_owner_cuts (host-only): the cuts partition all slabs in rank order, and the busiest port (max over ranks of
    slabs sent / received, both exchange phases in the replicated mode) is never worse than with the plain mid-overlap cuts."""
    import dvren_b200 as D
    lib = D.load()
    world = len(wedges)
    flat = (C.c_int32 * (2 * world))(*[v for w in wedges for v in w])
    cuts = (C.c_int32 * (world + 1))()
    D.check("hpx_plan_owner_cuts", lib.hpx_plan_owner_cuts(world, n, flat, 2 if replicated else 1, cuts))
    cuts = list(cuts)
    assert cuts[0] == 0 and cuts[-1] == n and all(a <= b for a, b in zip(cuts, cuts[1:]))
    mid, prev = [0] * (world + 1), 0
    mid[world] = n
    for r, (lo, hi) in enumerate(wedges):
        if lo >= hi:
            lo = hi = prev
        if r:
            mid[r] = min(n, max(mid[r - 1], (lo + prev) // 2))
        prev = max(prev, hi)
    worst = lambda c: max(max(v) for v in _exchange_volumes(wedges, c, replicated))
    assert worst(cuts) <= worst(mid)
    assert lib.hpx_plan_owner_cuts(world, n, flat, 0, (C.c_int32 * (world + 1))()) == A.HP_STATUS_INVALID_ARGUMENT


@pytest.mark.parametrize("rows", [1, 2, 7, 8, 27, 256])
def test_tile_row_orders_are_permutations(rows):
    """hpx_frame_set_row_order: every order visits every tile row once; centre-out starts at the middle row and alternates
    below / above it, so that the dispatched prefix is always one contiguous ring around the centre (hpx_backward_streamed
    cuts its row groups from that)."""
    import dvren_b200 as D
    lib = D.load()
    for order in (0, 1, 2):
        seq = []
        for i in range(rows):
            out = C.c_uint32()
            D.check("hpx_tile_row_order", lib.hpx_tile_row_order(i, rows, order, C.byref(out)))
            seq.append(out.value)
        assert sorted(seq) == list(range(rows)), (order, seq)
        if order == 0:
            assert seq == list(range(rows))
        if order == 1:
            assert seq == list(range(rows))[::-1]
        if order == 2:
            assert seq[0] == rows // 2
            for k in range(1, rows + 1):          # every prefix is a contiguous run of rows containing the centre
                pre = sorted(seq[:k])
                assert pre == list(range(pre[0], pre[0] + k)) and pre[0] <= rows // 2 <= pre[-1]
    assert lib.hpx_tile_row_order(rows, rows, 0, C.byref(C.c_uint32())) == A.HP_STATUS_INVALID_ARGUMENT


@pytest.mark.parametrize("tiles_x,rows", [(1, 1), (1, 9), (7, 1), (6, 10), (128, 27), (13, 4)])
def test_tile_orders_are_permutations(tiles_x, rows):
    """hpx_frame_set_row_order with HPX_ORDER_COLUMNS: every CTA of a launch takes another tile; without the flag the tiles go
    row by row, with it column by column from the middle column outwards, the rows of a column in the chosen row order."""
    import dvren_b200 as D
    lib = D.load()
    def walk(order):
        seq = []
        for b in range(tiles_x * rows):
            col, row = C.c_uint32(), C.c_uint32()
            D.check("hpx_tile_order", lib.hpx_tile_order(b, tiles_x, rows, order, C.byref(col), C.byref(row)))
            seq.append((col.value, row.value))
        return seq
    everything = sorted((c, r) for c in range(tiles_x) for r in range(rows))
    for row_order in (0, 1, 2):
        row_seq = []
        for i in range(rows):
            out = C.c_uint32()
            D.check("hpx_tile_row_order", lib.hpx_tile_row_order(i, rows, row_order, C.byref(out)))
            row_seq.append(out.value)
        plain = walk(row_order)
        assert sorted(plain) == everything
        assert plain == [(c, r) for r in row_seq for c in range(tiles_x)]
        cols = walk(D.HPX_ORDER_COLUMNS | row_order)
        assert sorted(cols) == everything
        col_seq = [cols[i * rows][0] for i in range(tiles_x)]
        assert cols == [(c, r) for c in col_seq for r in row_seq]
        assert col_seq[0] == tiles_x // 2
        for k in range(1, tiles_x + 1):          # the dispatched columns are always one contiguous run around the middle
            pre = sorted(col_seq[:k])
            assert pre == list(range(pre[0], pre[0] + k)) and pre[0] <= tiles_x // 2 <= pre[-1]
    bad = C.c_uint32()
    assert lib.hpx_tile_order(tiles_x * rows, tiles_x, rows, 0, C.byref(bad), C.byref(bad)) == A.HP_STATUS_INVALID_ARGUMENT
    assert lib.hpx_tile_order(0, tiles_x, rows, 3, C.byref(bad), C.byref(bad)) == A.HP_STATUS_INVALID_ARGUMENT
    assert lib.hpx_tile_order(0, tiles_x, rows, 8, C.byref(bad), C.byref(bad)) == A.HP_STATUS_INVALID_ARGUMENT


def test_best_tile_order_prefers_columns_for_the_middle_bands_of_a_sharded_frame():
    """hpx_plan_best_tile_order (what hpx_shard_create_bands applies to its band): on BASELINE configs[2] cut for 8 GPUs the
    middle bands -- every row equally expensive, cheap tiles at the left and right edge -- end sooner column by column,
    the outermost bands (rows get cheaper towards the image border) row by row with the cheap rows last; the estimate of the
    chosen order is never worse than the row orders'."""
    import dvren_b200 as D
    import synth as S
    lib = D.load()
    ctx, plan = C.c_void_p(), C.c_void_p()
    assert lib.hp_ctx_create(None, C.byref(ctx)) == 0
    desc = S.bench_plan(2048, 2048, 1024, stratified=False)
    assert lib.hp_plan_create(ctx, C.byref(desc), C.byref(plan)) == 0
    world = 8
    row0, rows = (C.c_uint32 * world)(), (C.c_uint32 * world)()
    D.check("hpx_plan_balanced_bands", lib.hpx_plan_balanced_bands(plan, world, row0, rows, None))
    chosen = []
    for r in range(world):
        order, ends = C.c_int32(), (C.c_double * 4)()
        D.check("hpx_plan_best_tile_order", lib.hpx_plan_best_tile_order(plan, row0[r], rows[r], 148 * 5, C.byref(order), ends))
        table = dict(zip((0, 1, D.HPX_ORDER_COLUMNS, D.HPX_ORDER_COLUMNS | 1), ends))
        assert table[order.value] <= min(table[0], table[1]) * 1.0001
        if order.value & D.HPX_ORDER_COLUMNS:
            assert table[order.value] < 0.995 * min(table[0], table[1])
        chosen.append(order.value)
    assert chosen[0] == 1 and chosen[-1] == 0                     # rays get longer towards the middle of the image
    assert all(o & D.HPX_ORDER_COLUMNS for o in chosen[2:6])
    one = C.c_int32()
    assert lib.hpx_plan_best_tile_order(plan, 2040, 16, 740, C.byref(one), None) == A.HP_STATUS_INVALID_ARGUMENT   # past the ROI
    assert lib.hpx_plan_best_tile_order(plan, 0, 8, 0, C.byref(one), None) == A.HP_STATUS_INVALID_ARGUMENT
    lib.hp_plan_release(plan)
    lib.hp_ctx_release(ctx)
