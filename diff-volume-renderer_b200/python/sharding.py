"""Ray-tile / view data parallelism of the hot path across the GPUs of one box.

The reference has no multi-device support at all (single process, default stream).
Rays are independent, so the path shards with NO forward collective: every rank owns a replica of
the grid and renders either a band of image rows of one view or whole views of a batch.  The only
exchange is the sum of the packed gradient block [4*V grid floats | 16 camera floats]
(hpx_grid_grad_buffer, include/hotpath/hp_b200.h), one all-reduce per step.

Two facts make a shard reproduce the unsharded plan:
  * pixel ids are global: a sub-plan is the full-size plan with a smaller ROI, and the reference
    computes pixel_id = py * W + px from full-frame coordinates (hotpath/src/cpu/ray_cpu.cpp:224);
  * the stratified jitter hashes the ray index WITHIN the plan's ray list
    (hotpath/src/cpu/samp_cpu.cpp:28-35), so a band starting `r` rays into the parent ROI passes
    ray_index_base = r (hpx_frame_set_view) and draws the same jitter as the unsharded plan.

Everything here is host logic on plain Python / ctypes structs; the collective goes through
torch.distributed (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import copy
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import hp_abi as A

# CTA tile of the lean kernels (csrc/dv_types.h: kTileH * kWarpsY rows): bands are cut on tile rows so
# that no CTA straddles two ranks and every rank launches whole tiles.
TILE_ROWS = 8
U32_MAX = 0xFFFFFFFF


@dataclass(frozen=True)
class Band:
    """Rows [y0, y0 + rows) of the parent plan's ROI, owned by `rank`."""
    rank: int
    y0: int             # absolute image row (parent roi.y + offset)
    rows: int
    ray_index_base: int  # rays of the parent ROI that precede this band

    @property
    def empty(self) -> bool:
        return self.rows == 0


def resolved_roi(desc: A.hp_plan_desc) -> Tuple[int, int, int, int]:
    """ROI as hp_plan_create resolves it (reference hp_runtime.cpp:100-118): zero size = full frame."""
    r = desc.roi
    if r.width == 0 or r.height == 0:
        return 0, 0, desc.width, desc.height
    return r.x, r.y, r.width, r.height


def row_bands(desc: A.hp_plan_desc, world: int, align: int = TILE_ROWS) -> List[Band]:
    """Cut the plan's ROI into `world` contiguous row bands of near-equal size, cut on multiples of
    `align` rows (relative to the ROI top).  Ranks beyond the number of tile rows get empty bands."""
    if world < 1:
        raise ValueError("world must be >= 1")
    x, y, w, h = resolved_roi(desc)
    units = (h + align - 1) // align          # tile rows
    bands, start = [], 0
    for rank in range(world):
        n_units = units // world + (1 if rank < units % world else 0)
        r0 = min(start * align, h)
        r1 = min((start + n_units) * align, h)
        bands.append(Band(rank, y + r0, r1 - r0, r0 * w))
        start += n_units
    assert sum(b.rows for b in bands) == h
    return bands


def weighted_row_bands(desc: A.hp_plan_desc, weights: Sequence[float], align: int = TILE_ROWS) -> List[Band]:
    """Contiguous row bands whose heights follow `weights` (cut on multiples of `align` rows).  Used by the pipelined
    strong scaling to make the LAST group small: only its share of the all-reduce is exposed."""
    x, y, w, h = resolved_roi(desc)
    units = (h + align - 1) // align
    total = float(sum(weights))
    cuts, acc = [0], 0.0
    for wgt in weights[:-1]:
        acc += wgt
        cuts.append(min(units, max(cuts[-1], int(round(units * acc / total)))))
    cuts.append(units)
    bands = []
    for i in range(len(weights)):
        r0, r1 = min(cuts[i] * align, h), min(cuts[i + 1] * align, h)
        bands.append(Band(i, y + r0, r1 - r0, r0 * w))
    assert sum(b.rows for b in bands) == h
    return bands


def band_desc(desc: A.hp_plan_desc, band: Band) -> A.hp_plan_desc:
    """Sub-plan of `desc` restricted to `band` (same frame size, camera, sampling and seed)."""
    if band.empty:
        raise ValueError("empty band has no plan")
    x, _, w, _ = resolved_roi(desc)
    out = A.hp_plan_desc()
    C.memmove(C.byref(out), C.byref(desc), C.sizeof(A.hp_plan_desc))
    out.roi.x, out.roi.y, out.roi.width, out.roi.height = x, band.y0, w, band.rows
    # capacities are re-derived by hp_plan_create from the new ROI
    out.max_rays = 0
    out.max_samples = 0
    return out


def views_of_rank(n_views: int, world: int, rank: int) -> List[int]:
    """Whole-view partition of a view batch: contiguous blocks, sizes differ by at most one."""
    base, extra = divmod(n_views, world)
    first = rank * base + min(rank, extra)
    return list(range(first, first + base + (1 if rank < extra else 0)))


def u32_safe_bands(desc: A.hp_plan_desc, steps: int) -> int:
    """Number of row bands a MATERIALISING plan needs so that rays * steps fits hp_plan_desc's u32
    max_samples (SURVEY finding 9: 2048^2 x 1024 steps = 2^32 does not fit one plan).  The lean
    path never materialises samples and is not subject to this limit."""
    _, _, w, h = resolved_roi(desc)
    rows_per_plan = max(1, U32_MAX // max(1, w * steps))
    return (h + rows_per_plan - 1) // rows_per_plan


class GradientAllReduce:
    """Sum of the packed gradient block across ranks.  `tensor` is a flat float32 torch tensor that
    aliases hpx_grid_grad_buffer on the GPU (or any CPU tensor under gloo)."""

    def __init__(self, tensor, group=None):
        import torch.distributed as dist
        self.dist, self.tensor, self.group = dist, tensor, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self):
        if self.world > 1:
            self.dist.all_reduce(self.tensor, op=self.dist.ReduceOp.SUM, group=self.group)
        return self.tensor

    @property
    def bytes_on_wire_per_rank(self) -> int:
        """Ring all-reduce volume: 2 (N-1)/N of the buffer leaves and enters every rank."""
        n = self.world
        return 0 if n == 1 else int(2 * (n - 1) / n * self.tensor.numel() * self.tensor.element_size())


def final_slab_runs(ranges: Sequence[Optional[Tuple[int, int]]]) -> List[List[Tuple[int, int]]]:
    """ranges[g] = [lo, hi) of gradient slabs group g can touch (None: touches nothing), groups processed in order.
    Returns, per group, the contiguous runs of slabs that are FINAL once that group is done: touched by it or an earlier
    group and by no later one.  Every touched slab appears in exactly one run, so reducing the runs reduces the whole
    gradient once."""
    out: List[List[Tuple[int, int]]] = []
    n = len(ranges)
    done: set = set()
    for g in range(n):
        touched = set()
        for r in ranges[: g + 1]:
            if r is not None:
                touched.update(range(r[0], r[1]))
        later = set()
        for r in ranges[g + 1:]:
            if r is not None:
                later.update(range(r[0], r[1]))
        final = sorted(touched - later - done)
        done.update(final)
        runs: List[Tuple[int, int]] = []
        for s_ in final:
            if runs and runs[-1][1] == s_:
                runs[-1] = (runs[-1][0], s_ + 1)
            else:
                runs.append((s_, s_ + 1))
        out.append(runs)
    return out


class PipelinedFrame:
    """Strong scaling of ONE frame over the ranks of a box with the gradient all-reduce hidden behind the rendering.

    The frame is cut into `groups` row groups, rendered one after the other.  Inside a group every rank owns the CTA
    tile rows t with t % world == rank (hpx_frame_set_interleave): all ranks carry the same mix of short and long rays
    and all of them touch the same region of the grid.  The gradient block is laid out with the world axis the image
    rows run along as its SLOWEST axis (hpx_grid_set_grad_layout), so the region a group touches is a contiguous range
    of slabs (hpx_frame_bounds, unioned over ranks).  When a group is done, the slabs no later group will touch are
    final on every rank: a side stream all-reduces exactly those, IN PLACE, while the compute stream is already
    rendering the next group.  Every slab is reduced once (no extra copies, buffers or passes); only the last group's
    share is exposed.

    Everything here is host orchestration (torch streams / events / torch.distributed); kernels are the library's.
    """

    def __init__(self, D, ctx, grid, full_desc, groups, world: int, rank: int, device, compute_stream):
        """groups: number of equal row groups, or a sequence of relative heights (e.g. (0.75, 0.25))."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.D = torch, dist, D
        self.ctx, self.grid, self.world, self.rank = ctx, grid, world, rank
        self.compute = compute_stream
        self.side = torch.cuda.Stream(device=device)
        self.reduce = True   # False: skip the collectives (timing experiments)
        weights = [1.0] * groups if isinstance(groups, int) else [float(v) for v in groups]
        # world axis the image rows advance along = the camera's "down" vector (second column of c2w's rotation)
        st_desc = full_desc
        c2w = [st_desc.camera.c2w[i] for i in range(12)]
        down = [abs(c2w[1]), abs(c2w[5]), abs(c2w[9])]
        if not any(down):
            down = [0.0, 1.0, 0.0]           # all-zero pose = identity (hp_plan_create default)
        self.slow_axis = max(range(3), key=lambda i: down[i])
        self.slab_floats, self.n_slabs = grid.set_grad_layout(self.slow_axis)
        ptr, floats = grid.grad_buffer()
        self.block = torch.as_tensor(_CudaView(ptr, floats), device=device)
        self.parts = []
        ranges: List[Optional[Tuple[int, int]]] = []
        for band in weighted_row_bands(full_desc, weights, align=TILE_ROWS * world):
            if band.empty:
                continue
            plan = D.Plan(ctx, band_desc(full_desc, band))
            frame = D.Frame(plan)
            frame.set_view(None, plan.desc.seed, band.ray_index_base)
            frame.set_interleave(world, rank)
            box = frame.bounds(grid)
            lo = box[self.slow_axis] if box[3 + self.slow_axis] > 0 else 1 << 40
            hi = box[self.slow_axis] + box[3 + self.slow_axis] if box[3 + self.slow_axis] > 0 else -1
            if world > 1:   # every rank must reduce the same slabs: union of the ranks' ranges
                t_lo = torch.tensor([lo], dtype=torch.int64, device=device)
                t_hi = torch.tensor([hi], dtype=torch.int64, device=device)
                dist.all_reduce(t_lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(t_hi, op=dist.ReduceOp.MAX)
                lo, hi = int(t_lo.item()), int(t_hi.item())
            ranges.append((lo, hi) if hi > lo else None)
            self.parts.append(dict(band=band, plan=plan, frame=frame, done=torch.cuda.Event()))
        for p, runs in zip(self.parts, final_slab_runs(ranges)):
            p["runs"] = runs
        self.ranges = ranges

    @property
    def samples(self) -> int:
        return sum(p["frame"].counts()["samples"] for p in self.parts)

    def step(self, dL_dI_ptr: int, flags: int):
        """dL_dI_ptr: DEVICE pointer of the WHOLE frame's (rays, 3) gradient.  Leaves the summed gradient of all
        ranks in the grid's gradient block (hpx_grid_grad_buffer, slab order of hpx_grid_set_grad_layout)."""
        torch = self.torch
        self.grid.zero_grad()
        last = len(self.parts) - 1
        for i, p in enumerate(self.parts):
            band, frame = p["band"], p["frame"]
            frame.forward(self.grid)
            frame.backward(self.grid, dL_dI_ptr + band.ray_index_base * 12, flags & ~self.D.HPX_BACKWARD_ZERO, device=True)
            if self.world == 1 or not self.reduce:
                continue
            p["done"].record(self.compute)
            with torch.cuda.stream(self.side):
                self.side.wait_event(p["done"])
                for a, b in p["runs"]:
                    self.dist.all_reduce(self.block[a * self.slab_floats: b * self.slab_floats], op=self.dist.ReduceOp.SUM)
                if i == last:   # camera gradients ride with the last group
                    self.dist.all_reduce(self.block[-16:], op=self.dist.ReduceOp.SUM)
        self.compute.wait_stream(self.side)

    def close(self):
        for p in self.parts:
            p["frame"].close()
            p["plan"].close()
        self.parts = []
        self.grid.set_grad_layout(2)


class SignalledFrame:
    """PipelinedFrame without per-group launches: ONE forward and ONE backward launch over the rank's interleaved tile
    rows.  The backward's CTAs run in tile-row order and count themselves into a per-group device counter
    (hpx_backward_signalled); the side stream waits on those counters (hpx_stream_wait_counter -> cuStreamWaitValue32) and
    all-reduces, in place, the gradient slabs each finished group leaves behind while later rows are still running.
    The GPU stays full the whole time (no launch tails); only the last group's slabs are reduced after the kernel."""

    def __init__(self, D, ctx, grid, full_desc, groups, world: int, rank: int, device, compute_stream, interleave: bool = True):
        """interleave=True: ONE frame shared by the ranks (strong scaling, tile rows t % world == rank).
        interleave=False: every rank renders its OWN full frame `full_desc` (weak scaling over views); the slab ranges are
        still unioned over the ranks, so the overlap works whenever the views' image rows advance along the same world
        axis in the same order (views of one orbit about that axis, neighbouring views of a batch)."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.D = torch, dist, D
        self.ctx, self.grid, self.world, self.rank = ctx, grid, world, rank
        self.compute = compute_stream
        self.side = torch.cuda.Stream(device=device, priority=-1)   # waits + collectives go ahead of queued rendering CTAs
        stride, phase = (world, rank) if interleave else (1, 0)
        self.side_ctx = D.Context(device=device.index, stream=self.side.cuda_stream)
        self.reduce = True
        weights = [1.0] * groups if isinstance(groups, int) else [float(v) for v in groups]
        c2w = [full_desc.camera.c2w[i] for i in range(12)]
        down = [abs(c2w[1]), abs(c2w[5]), abs(c2w[9])]
        if not any(down):
            down = [0.0, 1.0, 0.0]
        self.slow_axis = max(range(3), key=lambda i: down[i])
        self.slab_floats, self.n_slabs = grid.set_grad_layout(self.slow_axis)
        ptr, floats = grid.grad_buffer()
        self.block = torch.as_tensor(_CudaView(ptr, floats), device=device)
        self.plan = D.Plan(ctx, full_desc)
        self.frame = D.Frame(self.plan)
        self.frame.set_interleave(stride, phase)
        # groups: row bands cut on multiples of (tile rows x stride) so that every rank owns the same number of tile rows
        ranges: List[Optional[Tuple[int, int]]] = []
        ends, owned = [], 0
        bands = [b for b in weighted_row_bands(full_desc, weights, align=TILE_ROWS * stride) if not b.empty]
        for band in bands:
            plan = D.Plan(ctx, band_desc(full_desc, band))
            probe = D.Frame(plan)
            probe.set_interleave(stride, phase)
            box = probe.bounds(grid)
            probe.close(); plan.close()
            lo = box[self.slow_axis] if box[3 + self.slow_axis] > 0 else 1 << 40
            hi = box[self.slow_axis] + box[3 + self.slow_axis] if box[3 + self.slow_axis] > 0 else -1
            if world > 1:
                t_lo = torch.tensor([lo], dtype=torch.int64, device=device)
                t_hi = torch.tensor([hi], dtype=torch.int64, device=device)
                dist.all_reduce(t_lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(t_hi, op=dist.ReduceOp.MAX)
                lo, hi = int(t_lo.item()), int(t_hi.item())
            ranges.append((lo, hi) if hi > lo else None)
            tile_rows = (band.rows + TILE_ROWS - 1) // TILE_ROWS
            owned += (tile_rows - phase + stride - 1) // stride if tile_rows > phase else 0
            ends.append(owned)
        self.group_end_rows, self.ranges, self.bands = ends, ranges, bands
        self.runs = final_slab_runs(ranges)
        self.start = torch.cuda.Event()

    @property
    def samples(self) -> int:
        return self.frame.counts()["samples"]

    def step(self, dL_dI_ptr: int, flags: int):
        torch = self.torch
        self.grid.zero_grad()
        self.frame.forward(self.grid)
        counters = self.frame.reset_group_counters()
        self.start.record(self.compute)
        _, expected = self.frame.backward_signalled(self.grid, dL_dI_ptr, self.group_end_rows,
                                                    flags & ~self.D.HPX_BACKWARD_ZERO)
        if self.world > 1 and self.reduce:
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.start)     # counters cleared: the previous step's counts cannot satisfy the waits
                for g, runs in enumerate(self.runs):
                    self.side_ctx.wait_counter(counters + 4 * g, expected[g])
                    for a, b in runs:
                        self.dist.all_reduce(self.block[a * self.slab_floats: b * self.slab_floats], op=self.dist.ReduceOp.SUM)
                self.dist.all_reduce(self.block[-16:], op=self.dist.ReduceOp.SUM)
            self.compute.wait_stream(self.side)

    def close(self):
        self.frame.close()
        self.plan.close()
        self.side_ctx.close()
        self.grid.set_grad_layout(2)


class _CudaView:
    """Raw device pointer as a torch-importable array (zero copy)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}
