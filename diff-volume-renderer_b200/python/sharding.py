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
