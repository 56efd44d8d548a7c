#!/usr/bin/env python
"""bench.py -- Msamples/s of the dvren hot path (fused forward + backward to the dense grid).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3]

A "step" is one pass of the hot path over one batch of synthetic input: ray generation ->
fused march / integrate / compose -> reverse-march backward with grid scatter (-> NCCL all-reduce
of the packed gradient block when N > 1).  Default workload = BASELINE.json configs[1]:
256^3 grid, 1024x1024, stratified sampling, 512 steps (537 M samples), on one B200.  For N > 1
every rank renders its own view of an N-view batch against a replicated grid (weak scaling) and
the ranks all-reduce the gradients; `value` counts the samples of all ranks.

Prints ONE JSON line (see README / DESIGN.md section "Measurement" for every key).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly ONE JSON line.  Native libraries write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there), so descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved one.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


REPO = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(REPO, "diff-volume-renderer_b200", "python")]

CONFIGS = {
    # name: grid n, image W, steps, stratified, description
    "c1": dict(grid=64, width=512, steps=256, stratified=False,
               workload="BASELINE configs[0] shape: 64^3 grid, 512x512, fixed, 256 steps"),
    "c2": dict(grid=256, width=1024, steps=512, stratified=True,
               workload="BASELINE configs[1]: 256^3 dense grid, 1024x1024, stratified, 512 steps, fused fwd + sigma/color bwd"),
    "c3": dict(grid=512, width=2048, steps=1024, stratified=False,
               workload="BASELINE configs[2]: 512^3 dense grid, 2048x2048, fixed, 1024 steps, fwd+bwd"),
    "c4": dict(grid=256, width=800, steps=512, stratified=True, views=64,
               workload="BASELINE configs[3]: 64 views at 800x800 over a 256^3 grid, stratified, 512 steps, grid + camera "
                        "gradients, one captured CUDA graph replayed per view"),
    "c5": dict(grid=1024, width=1024, steps=1024, stratified=True, views=128, camera=False, device_volume=True,
               workload="BASELINE configs[4]: 1024^3 dense grid (17.2 GB packed, replicated per GPU), 128 views at 1024x1024, "
                        "stratified, 1024 steps, fwd+bwd, views sharded across the GPUs, one captured CUDA graph replayed per view"),
}
METRIC = "Msamples/s fwd+bwd (fused forward + backward adjoint to the dense sigma/color grid)"
UNIT = "Msamples/s"
# algorithmic bytes per live sample: 8 corners x 16 B gathered; + 8 x 16 B gradient reds
BYTES_FWD_PER_SAMPLE = 128
BYTES_BWD_PER_SAMPLE = 256


def read_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region.  The thread starts (NVML init
    included) before the warm-up and samples every 5 ms; only samples between mark_begin() and mark_end() count."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.stop_flag, self.samples, self.max_mhz, self.err = index, False, [], None, None
        self.t0 = self.t1 = None
        self.thread = None
        self.ready = threading.Event()

    def _run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            self.ready.set()
            while not self.stop_flag:
                self.samples.append((time.perf_counter(), float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)),
                                     int(N.nvmlDeviceGetCurrentClocksEventReasons(h))))
                time.sleep(0.005)
        except Exception as e:  # NVML missing: report it, never fake a clock
            self.err = repr(e)
            self.ready.set()

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        self.ready.wait(timeout=10.0)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.t1 is None:
            self.mark_end()
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2.0)
        t0 = self.t0 if self.t0 is not None else 0.0
        inside = [s for s in self.samples if t0 <= s[0] <= self.t1]
        sm = sorted(s[1] for s in inside)
        reasons = set()
        for s in inside:
            for bit, name in self.REASONS.items():
                if s[2] & bit:
                    reasons.add(name)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
               "samples": len(sm)}
        if self.err:
            out["error"] = self.err
        return out


class CudaArrayView:
    """Exposes a raw device pointer to torch (zero copy) via __cuda_array_interface__."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import dvren_b200 as D
    import synth as S

    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun (python -m torch.distributed.run --nproc-per-node {args.gpus} ...)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL's kernels on a HIGH-PRIORITY stream: when a collective is issued while a long rendering kernel still has
        # CTAs queued (overlapped strong scaling), its few CTAs get the next free slots instead of waiting for the tail
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    import sharding as SH

    n, W, steps = cfg["grid"], cfg["width"], cfg["steps"]
    sigma, color = S.hashed_volume(n, "thin")       # no early termination: live samples == samples
    ctx = D.Context(device=local_rank, stream=stream.cuda_stream)
    rows_mode = args.sharding == "rows" and world > 1
    if args.sharding in ("pipeline", "signalled", "views-overlap"):
        return run_pipelined(args, cfg, ctx, sigma, color, stream, dev, world, rank)
    if rows_mode:
        # strong scaling: ONE frame cut into row bands, global pixel ids + global ray-index base
        full = S.bench_plan(W, W, steps, stratified=cfg["stratified"])
        band = SH.row_bands(full, world)[rank]
        plan = D.Plan(ctx, SH.band_desc(full, band))
    else:
        # weak scaling: every rank renders its own view of a `world`-view batch.  The views sit 2.5 degrees apart on an arc
        # centred on the canonical camera, so that per-rank work is (nearly) the same: kernel time depends on the view
        # direction relative to the grid's x-fastest layout (profiles/README.md: 2.25 ms at 0 deg, 2.83 ms at 90 deg)
        # and a full orbit would measure that spread, not the scaling.
        plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=cfg["stratified"], view=rank - (world - 1) / 2.0, views=144))
    grid = D.Grid(ctx, sigma, color)
    del sigma, color
    frame = D.Frame(plan)
    if rows_mode:
        frame.set_view(None, plan.desc.seed, band.ray_index_base)
    n_rays = plan.n_rays
    g_host = torch.from_numpy(S.hashed_image_grad(n_rays)).pin_memory()
    g_dev = g_host.to(dev, non_blocking=True)
    grad_ptr, grad_floats = grid.grad_buffer()
    grad_view = torch.as_tensor(CudaArrayView(grad_ptr, grad_floats), device=dev)
    img = frame.image_ptrs()
    pixels = W * W
    reducer = SH.GradientAllReduce(grad_view) if world > 1 else None
    planes = [torch.as_tensor(CudaArrayView(img.image.data, pixels * 3), device=dev),
              torch.as_tensor(CudaArrayView(img.trans.data, pixels), device=dev),
              torch.as_tensor(CudaArrayView(img.opacity.data, pixels), device=dev),
              torch.as_tensor(CudaArrayView(img.depth.data, pixels), device=dev)]
    planes_host = [torch.empty(p.shape, dtype=p.dtype).pin_memory() for p in planes]
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO

    def step_resident():
        frame.forward(grid)
        frame.backward(grid, g_dev.data_ptr(), flags, device=True)
        if reducer is not None:
            reducer()

    # e2e: the same step through the C ABI with HOST buffers (pinned), i.e. what dvren::Renderer
    # Forward/Backward move per step (reference renderer.hpp:50-66): dL/dI host->device; image planes
    # and the un-interleaved sigma / colour gradient grids device->host.  The image planes are read from
    # the frame's device views (hpx_frame_image) on a side stream while the backward runs; the gradient
    # read blocks.
    import ctypes as C
    import hp_abi as A
    lib = ctx.lib
    sg_host = torch.empty(grid.voxels, dtype=torch.float32).pin_memory()
    cg_host = torch.empty(grid.voxels * 3, dtype=torch.float32).pin_memory()
    cam_host = torch.empty(16, dtype=torch.float32).pin_memory()
    mask_host = torch.empty(pixels, dtype=torch.int32).pin_memory()

    side = torch.cuda.Stream(device=dev)
    fwd_done = torch.cuda.Event()
    mask_dev = torch.as_tensor(CudaArrayView(img.hitmask.data, pixels), device=dev).view(torch.int32)

    def read_planes_async():
        """Image planes device -> pinned host on a side stream, so that the copy runs under the backward kernel."""
        fwd_done.record(stream)
        with torch.cuda.stream(side):
            side.wait_event(fwd_done)
            for h, d in zip(planes_host, planes):
                h.copy_(d, non_blocking=True)
            mask_host.copy_(mask_dev, non_blocking=True)

    def step_e2e():
        D.check("hpx_forward", lib.hpx_forward(frame.handle, grid.handle))
        read_planes_async()
        D.check("hpx_backward", lib.hpx_backward(frame.handle, grid.handle, g_host.data_ptr(), A.HP_MEMSPACE_HOST, flags))
        if reducer is not None:
            reducer()
        D.check("hpx_grid_read_grad", lib.hpx_grid_read_grad(grid.handle, sg_host.data_ptr(), cg_host.data_ptr(),
                                                             cam_host.data_ptr(), A.HP_MEMSPACE_HOST))
        stream.wait_stream(side)

    def step_e2e_device_grads():
        """Same, but the gradient block stays in HBM for a device-side optimiser (hpx_grid_grad_buffer): host
        traffic is dL/dI in, the five image planes out."""
        D.check("hpx_forward", lib.hpx_forward(frame.handle, grid.handle))
        read_planes_async()
        D.check("hpx_backward", lib.hpx_backward(frame.handle, grid.handle, g_host.data_ptr(), A.HP_MEMSPACE_HOST, flags))
        if reducer is not None:
            reducer()
        stream.wait_stream(side)
        ctx.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, warm, sampler=None):
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.mark_begin()
        a.record(stream)
        for _ in range(k):
            fn()
        b.record(stream)
        barrier()
        if sampler is not None:
            sampler.mark_end()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(step_resident, args.steps, args.warmup, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    counts = frame.counts()
    samples, live = counts["samples"], counts["live_samples"]
    e2e_ms = timed(step_e2e, args.steps, max(args.warmup, 1))
    e2e_dev_ms = timed(step_e2e_device_grads, args.steps, 1)
    # kernel-level timing for the roofline lines (same stream, CUDA events, after the runs above)
    fwd_ms = timed(lambda: frame.forward(grid), args.steps, 1)
    bwd_ms = timed(lambda: frame.backward(grid, g_dev.data_ptr(), D.HPX_BACKWARD_GRID, device=True), args.steps, 1)

    ms_per_step = total_ms / args.steps
    total_samples = samples
    if world > 1:
        ts = torch.tensor([samples], dtype=torch.int64, device=dev)
        dist.all_reduce(ts)
        total_samples = int(ts.item())
    value = total_samples / (ms_per_step * 1e-3) / 1e6
    e2e_value = total_samples / (e2e_ms / args.steps * 1e-3) / 1e6
    peak, peak_src = read_peaks()
    bwd_bytes = BYTES_BWD_PER_SAMPLE * live + 12 * n_rays
    fwd_bytes = BYTES_FWD_PER_SAMPLE * live + 24 * n_rays
    bwd_gbs = bwd_bytes / (bwd_ms / args.steps * 1e-3) / 1e9
    fwd_gbs = fwd_bytes / (fwd_ms / args.steps * 1e-3) / 1e9
    bwd_kernel = "lean_backward_merge_kernel" if frame.scatter_mode(grid, flags) == "merged" else "lean_backward_kernel"
    traffic = fwd_traffic = None
    l1_pct = {}
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f).get(args.config, {})
        traffic, fwd_traffic = tj.get(bwd_kernel), tj.get("lean_forward_kernel")
        l1_pct = tj.get("_l1_wavefront_pct", {})

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if rows_mode else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "volume": "hashed thin (sigma = 2u, no early termination)",
                   "rays_per_gpu": n_rays, "samples_per_gpu_step": samples, "live_samples_per_gpu_step": live,
                   "parallelism": (f"one frame in {world} row bands" if rows_mode else f"{world} views (2.5 deg apart), one per GPU") +
                                  ", grid replicated, one NCCL all-reduce of the packed gradient block per step"
                                  if world > 1 else "single GPU",
                   "allreduce_bytes": int(grad_floats * 4) if world > 1 else 0,
                   "l2": "inputs larger than L2 (packed grid %d MB + gradient grid %d MB vs 126 MB)" % (n ** 3 * 16 >> 20, n ** 3 * 16 >> 20)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(g_host.numel() * 4),
                "d2h_bytes_per_step": int(sum(p.numel() for p in planes_host) * 4 + pixels * 4 + grad_floats * 4),
                "ms_per_step": e2e_ms / args.steps,
                "note": "full dvren::Renderer result contract: image planes AND un-interleaved sigma/colour gradient grids "
                        "copied to host every step (PCIe-bound)",
                "device_resident_gradients": {
                    "value": total_samples / (e2e_dev_ms / args.steps * 1e-3) / 1e6, "ms_per_step": e2e_dev_ms / args.steps,
                    "d2h_bytes_per_step": int(sum(p.numel() for p in planes_host) * 4 + pixels * 4)}},
        "gpu_launches": 2 * args.steps,
        "clocks": clocks,
        "fwd": {"ms": fwd_ms / args.steps, "msamples_s": samples / (fwd_ms / args.steps * 1e-3) / 1e6},
        "bwd": {"ms": bwd_ms / args.steps, "msamples_s": samples / (bwd_ms / args.steps * 1e-3) / 1e6},
        "roofline": {"bound": "hbm", "kernel": bwd_kernel, "achieved": bwd_gbs, "peak": peak,
                     "unit": "GB/s", "frac": bwd_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bwd_bytes,
                     "note": "algorithmic bytes (SURVEY 8d) are gather/scatter bytes at the L1/L2 level: 256 B per live "
                             "sample for the backward, 128 B for the forward; a pixel tile re-touches the same voxels, "
                             "so caches absorb them and frac > 1.  Measured DRAM bytes per launch are in `traffic`; "
                             "the binding resource is the SM's L1/LSU data pipe (l1_wavefront_pct_of_peak, from the ncu "
                             "capture under profiles/), see DESIGN.md section 5",
                     "l1_wavefront_pct_of_peak": l1_pct.get(bwd_kernel),
                     "hbm_frac_of_peak": (traffic / (bwd_ms / args.steps * 1e-3) / 1e9 / peak) if traffic else None,
                     "forward_kernel": {"kernel": "lean_forward_kernel", "achieved": fwd_gbs, "frac": fwd_gbs / peak,
                                        "algorithmic_bytes_per_launch": fwd_bytes, "traffic": fwd_traffic,
                                        "l1_wavefront_pct_of_peak": l1_pct.get("lean_forward_kernel")}},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, rows=args.cpu_rows, threads=1)
    if rank == 0:
        emit(line)
    frame.close(); grid.close(); plan.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_pipelined(args, cfg, ctx, sigma, color, stream, dev, world, rank):
    """Strong scaling of ONE frame with the gradient all-reduce hidden behind the rendering (sharding.PipelinedFrame):
    row groups, interleaved tile rows inside a group, per-group voxel boxes reduced on a side stream."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import dvren_b200 as D
    import sharding as SH
    import synth as S

    n, W, steps = cfg["grid"], cfg["width"], cfg["steps"]
    weak = args.sharding == "views-overlap"
    if weak:   # every rank its own view (2.5 degrees apart, like the default weak-scaling run), all-reduce hidden behind the backward
        full = S.bench_plan(W, W, steps, stratified=cfg["stratified"], view=rank - (world - 1) / 2.0, views=144)
    else:
        full = S.bench_plan(W, W, steps, stratified=cfg["stratified"])
    grid = D.Grid(ctx, sigma, color)
    del sigma, color
    g_host = torch.from_numpy(S.hashed_image_grad(W * W)).pin_memory()
    g_dev = g_host.to(dev, non_blocking=True)
    groups = [float(v) for v in args.group_split.split(",")] if args.group_split else args.groups
    if weak:
        pf = SH.SignalledFrame(D, ctx, grid, full, groups, world, rank, dev, stream, interleave=False)
    else:
        cls = SH.SignalledFrame if args.sharding == "signalled" else SH.PipelinedFrame
        pf = cls(D, ctx, grid, full, groups, world, rank, dev, stream)
    flags = D.HPX_BACKWARD_GRID
    cam_host = torch.empty(16, dtype=torch.float32).pin_memory()

    def step():
        pf.step(g_dev.data_ptr(), flags)

    def step_e2e():
        g_dev.copy_(g_host, non_blocking=True)
        pf.step(g_dev.data_ptr(), flags)
        cam_host.copy_(pf.block[-16:], non_blocking=True)
        (pf.frame if hasattr(pf, "frame") else pf.parts[0]["frame"]).read()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, warm):
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(k):
            fn()
        b.record(stream)
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # correctness first: the pipelined, all-reduced gradient against a plain single-GPU full-frame backward (rank 0)
    step()
    barrier()
    verify = None
    if rank == 0 and not weak:   # (weak mode sums DIFFERENT views: no single-GPU frame to compare with)
        got = pf.block.clone()
        plan = D.Plan(ctx, full)
        frame = D.Frame(plan)
        frame.forward(grid)
        frame.backward(grid, g_dev.data_ptr(), D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO, device=True)
        torch.cuda.synchronize()
        ref = pf.block
        scale = torch.maximum(ref.abs(), 1e-2 * ref.abs().max())
        verify = float(((got - ref).abs() / scale).max().item())
        frame.close(); plan.close()
    barrier()

    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    total_ms = timed(step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = timed(step_e2e, args.steps, 1)
    pf.reduce = False                      # the same step without the collectives: what the all-reduce still costs
    no_reduce_ms = timed(step, args.steps, 1)
    pf.reduce = True
    samples = torch.tensor([pf.samples], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(samples)
    total = int(samples.item())
    ms_per_step = total_ms / args.steps
    line = {"metric": METRIC, "value": total / (ms_per_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if weak else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "volume": "hashed thin (sigma = 2u, no early termination)",
                       "parallelism": f"one frame, {len(pf.ranges)} row groups ({args.sharding}), tile rows interleaved over {world} GPUs; gradient "
                                      f"block laid out with axis {'xyz'[pf.slow_axis]} slowest, the slabs a finished group leaves "
                                      "behind all-reduced in place on a side stream while the next group renders",
                       "slab_ranges": pf.ranges,
                       "group_rows": [b.rows for b in pf.bands] if hasattr(pf, "bands") else [p["band"].rows for p in pf.parts], "allreduce_bytes": grid.voxels * 16,
                       "verify_max_rel_err_vs_single_gpu": verify,
                       "ms_per_step_without_collectives": no_reduce_ms / args.steps,
                       "l2": "inputs larger than L2"},
            "e2e": {"value": total / (e2e_ms / args.steps * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(g_host.numel() * 4), "d2h_bytes_per_step": 64 + W * W * 28,
                    "note": "gradient block stays in HBM (device-side optimiser)"},
            "gpu_launches": (2 if hasattr(pf, "frame") else 2 * len(pf.parts)) * args.steps, "clocks": clocks}
    if rank == 0:
        emit(line)
    pf.close(); grid.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_view_batch(args):
    """Config 4: a batch of views through ONE captured CUDA graph (forward + backward to the grid AND the camera);
    the view changes between replays through the frame's device parameter block.  N > 1: views split across ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import dvren_b200 as D
    import sharding as SH
    import synth as S

    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    n, W, steps, views = cfg["grid"], cfg["width"], cfg["steps"], cfg["views"]
    ctx = D.Context(device=local_rank, stream=stream.cuda_stream)
    my_views = SH.views_of_rank(views, world, rank)
    descs = [S.bench_plan(W, W, steps, stratified=cfg["stratified"], view=v, views=views) for v in my_views]
    plan = D.Plan(ctx, descs[0])
    if cfg.get("device_volume"):
        # too large to hash on the host in reasonable time: a seeded device generator (same seed on every rank = replicas)
        gen = torch.Generator(device=dev).manual_seed(1234)
        sigma = torch.rand((n, n, n), generator=gen, device=dev, dtype=torch.float32) * 2.0
        color = torch.rand((n, n, n, 3), generator=gen, device=dev, dtype=torch.float32)
        torch.cuda.synchronize()
        grid = D.Grid(ctx, sigma.data_ptr(), color.data_ptr(), device_shape=(n, n, n))
        ctx.synchronize()
        del sigma, color
        torch.cuda.empty_cache()
    else:
        sigma, color = S.hashed_volume(n, "thin")
        grid = D.Grid(ctx, sigma, color)
        del sigma, color
    frame = D.Frame(plan)
    g_host = torch.from_numpy(S.hashed_image_grad(plan.n_rays)).pin_memory()
    g_frame = torch.as_tensor(CudaArrayView(frame.grad_input_ptr(), plan.n_rays * 3), device=dev)
    g_frame.copy_(g_host.reshape(-1), non_blocking=True)
    grad_ptr, grad_floats = grid.grad_buffer()
    grad_view = torch.as_tensor(CudaArrayView(grad_ptr, grad_floats), device=dev)
    reducer = SH.GradientAllReduce(grad_view) if world > 1 else None
    flags = D.HPX_BACKWARD_GRID | (D.HPX_BACKWARD_CAMERA if cfg.get("camera", True) else 0)
    frame.capture(grid, flags)
    cams = [d.camera for d in descs]
    cam_host = torch.empty(16, dtype=torch.float32).pin_memory()

    def step():
        grid.zero_grad()
        for v, cam in zip(my_views, cams):
            frame.set_view(cam, 42 + v, 0)
            frame.replay()
        if reducer is not None:
            reducer()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(args.steps):
        step()
    b.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    samples = views * plan.n_rays * steps
    # e2e: per step the cameras go in (tiny) and the camera gradients + the last image come out; the grid gradient stays in HBM
    a.record(stream)
    for _ in range(args.steps):
        step()
        cam_host.copy_(grad_view[-16:], non_blocking=True)
        frame.read()
    b.record(stream)
    barrier()
    e2e_ms = a.elapsed_time(b) / args.steps
    line = {"metric": METRIC + (" + camera adjoint" if cfg.get("camera", True) else ""), "value": samples / (ms_per_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "views": views, "views_per_gpu": len(my_views),
                       "volume": "device-generated uniform (sigma = 2u)" if cfg.get("device_volume") else "hashed thin (sigma = 2u)",
                       "samples_per_step": samples, "allreduce_bytes": int(grad_floats * 4) if world > 1 else 0,
                       "backward_kernel": ("lean_backward_merge_kernel" if frame.scatter_mode(grid, flags) == "merged"
                                           else "lean_backward_kernel") + (" (+ camera adjoint)" if cfg.get("camera", True) else ""),
                       "l2": "inputs larger than L2 (grid %d MB + gradient %d MB)" % (n ** 3 * 16 >> 20, n ** 3 * 16 >> 20)},
            "e2e": {"value": samples / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": len(my_views) * 128, "d2h_bytes_per_step": 64 + W * W * 28},
            "gpu_launches": 4 * len(my_views) * args.steps, "clocks": clocks,
            "ms_per_view": ms_per_step / len(my_views)}
    if rank == 0:
        emit(line)
    frame.close(); grid.close(); plan.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(cfg, rows: int, threads: int):
    """The reference's own CPU implementation (oracle/_ref, unmodified, via dvren::Renderer) -- or the
    oracle port when that library is absent -- timed on a band of `rows` image rows per thread."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import numpy as np

    import oracle as O
    import synth as S

    n, W, steps = cfg["grid"], cfg["width"], cfg["steps"]
    sigma, color = S.hashed_volume(n, "thin")
    use_ref = O.ref_available()
    if not use_ref:
        O.build_oracle()
    results = [None] * threads
    y_start = (W - rows * threads) // 2

    def work(t):
        y0 = y_start + t * rows
        desc = S.bench_plan(W, W, steps, stratified=cfg["stratified"], roi=(0, y0, W, rows))
        dl = S.hashed_image_grad(W * rows)
        t0 = time.perf_counter()
        if use_ref:
            r = O.ref_render(desc, sigma, color, dl)
            assert r["status"] == 0, r["status"]
            ms = r["forward_ms"] + r["backward_ms"]
            cnt = r["sample_count"]
        else:
            st, rd = O.plan_resolve(desc)
            gs, gc = O.make_grid(sigma, 1), O.make_grid(color, 3)
            r = O.render(rd, gs, gc, dl, per_ray=False, frames=True)
            ms = (time.perf_counter() - t0) * 1e3
            cnt = r["sample_count"]
        results[t] = (cnt, ms, (time.perf_counter() - t0) * 1e3)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in results)
    # throughput of the hot path itself: samples / (forward + backward time), slowest thread
    busy_ms = max(r[1] for r in results)
    return {"value": total / (busy_ms * 1e-3) / 1e6, "unit": UNIT, "cores": threads,
            "kind": "reference" if use_ref else "port",
            "sample": f"{threads} band(s) of {rows} rows x {W} px x {steps} steps = {total} samples, "
                      f"Renderer::Forward+Backward time {busy_ms:.0f} ms (wall incl. setup {wall:.1f} s)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation on all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    threads = max(1, min(os.cpu_count() or 1, args.cpu_threads))
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(cfg, rows=1, threads=threads)
    t0 = time.perf_counter()
    vals = [cpu_baseline(cfg, rows=args.cpu_rows_ref, threads=threads) for _ in range(args.steps)]
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    v = sum(x["value"] for x in vals) / len(vals)
    base = vals[-1]
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "volume": "hashed thin (sigma = 2u, no early termination)"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-rows", type=int, default=32, help="image rows of the cpu_baseline sample")
    ap.add_argument("--cpu-rows-ref", type=int, default=8, help="rows per thread per step for --impl reference")
    ap.add_argument("--cpu-threads", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sharding", default="views", choices=["views", "rows", "pipeline", "signalled", "views-overlap"],
                    help="N > 1: one view per GPU (weak scaling, default); one frame cut into row bands (strong); or one "
                         "frame in row groups with interleaved tile rows and the all-reduce overlapped (strong, pipelined)")
    ap.add_argument("--groups", type=int, default=2, help="equal row groups of --sharding pipeline")
    ap.add_argument("--group-split", default="", help="relative heights of the row groups instead, e.g. 0.75,0.25")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    elif "views" in CONFIGS[args.config]:
        run_view_batch(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
