#!/usr/bin/env python
"""bench.py -- Msamples/s of the dvren hot path (fused forward + backward to the dense grid).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5]

A "step" is one pass of the hot path over one batch of synthetic input: ray generation -> fused
march / integrate / compose -> reverse-march backward with the grid scatter (-> exchange of the
gradient sums when N > 1).

Default workload for EVERY N = BASELINE.json configs[2], the configuration the scaling target is
quoted on: ONE 2048x2048 frame over a 512^3 grid, fixed sampling, 1024 steps (2^32 samples per
step).  N = 1 runs it on one B200 (it fits: 2.1 GB grid + 2.1 GB gradient + 2.1 GB checkpoints);
N > 1 is STRONG scaling of that same frame through the library's band sharding (hpx_shard_*,
include/hotpath/hp_b200.h): work-balanced row bands, grid replicated, sparse slab exchange.  The
N = 1 line also carries the numbers of BASELINE configs[1] (256^3 / 1024^2 / stratified, the round-1
default) under "c2".

Prints ONE JSON line (README / DESIGN.md section "Measurement" explain every key).
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


def _finite(x):
    """JSON has no NaN / Infinity: they become null."""
    if isinstance(x, float) and (x != x or x in (float("inf"), float("-inf"))):
        return None
    if isinstance(x, dict):
        return {k: _finite(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_finite(v) for v in x]
    return x


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(_finite(line), allow_nan=False) + "\n").encode())


REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "diff-volume-renderer_b200")
sys.path[:0] = [os.path.join(PKG, "python")]

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


def read_peaks():
    """Roofline denominators: HBM from the driver's MEASURED_PEAKS.json (fallback: B200_PROFILING.md), the SM-side ones
    from this repo's own probe (tools/mem_probe.cu, results committed as profiles/r01_mem_probe.json)."""
    out = {"hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)"}
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            out["hbm_gbs"], out["hbm_source"] = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    with open(os.path.join(REPO, "profiles", "r01_mem_probe.json")) as f:
        probe = json.load(f)
    out["l2_gather_gbs"] = next(r["gbs"] for r in probe["gather_random_16B"] if r["footprint_mb"] == 64)
    out["l1_gather_gbs"] = max(r["gbs"] for r in probe["gather_warp_local_16B"])
    out["red_lane_gops"] = max(r["gops"] for r in probe["red_v4_f32"])
    out["sm_side_source"] = "profiles/r01_mem_probe.json (tools/mem_probe.cu on a B200: 16-B gathers, L2-resident 64 MB random / warp-local window)"
    return out


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

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_hashed_volume(torch, n, kind, dev, seed=1234):
    """synth.hashed_volume (u(i, s) = top 24 bits of mix64((s ^ i) + GOLDEN) / 2^24) evaluated on the GPU with
    wrapping int64 arithmetic -- identical bytes, seconds instead of minutes of host hashing at 512^3."""
    def lsr(x, k):
        return (x >> k) & ((1 << (64 - k)) - 1)

    def s64(v):
        return v - (1 << 64) if v >= (1 << 63) else v

    def unit(idx, s):
        x = (idx ^ s) + s64(0x9E3779B97F4A7C15)
        x = (x ^ lsr(x, 30)) * s64(0xBF58476D1CE4E5B9)
        x = (x ^ lsr(x, 27)) * s64(0x94D049BB133111EB)
        x = x ^ lsr(x, 31)
        return lsr(x, 40).to(torch.float32) * (1.0 / 16777216.0)

    idx = torch.arange(n * n * n, dtype=torch.int64, device=dev)
    scale = {"thin": 2.0, "dense": 40.0}[kind]
    sigma = (unit(idx, seed) * scale).reshape(n, n, n).contiguous()
    color = torch.stack([unit(idx, seed + 1 + c) for c in range(3)], dim=-1).reshape(n, n, n, 3).contiguous()
    return sigma, color


def make_grid(D, S, torch, ctx, n, kind, dev):
    import numpy as np
    if not getattr(make_grid, "checked", False):   # the device generator against the numpy one, once
        a, b = device_hashed_volume(torch, 8, kind, dev)
        ha, hb = S.hashed_volume(8, kind)
        assert np.array_equal(a.cpu().numpy(), ha) and np.array_equal(b.cpu().numpy(), hb), "device volume generator drifted"
        make_grid.checked = True
    sigma, color = device_hashed_volume(torch, n, kind, dev)
    torch.cuda.synchronize()
    grid = D.Grid(ctx, sigma.data_ptr(), color.data_ptr(), device_shape=(n, n, n))
    ctx.synchronize()
    del sigma, color
    torch.cuda.empty_cache()
    return grid


def roofline(cfg_name, bwd_kernel, fwd_ms, bwd_ms, cube, live, rays, touched, peaks):
    """Three levels, each a fraction that stays <= ~1 and can be recomputed from this object + profiles/:
      hbm  compulsory HBM bytes (the touched voxels once per pass) / time / measured HBM copy bandwidth
      l2   SURVEY 8d two-level model: algorithmic gather/red bytes / the measured L2-resident gather bandwidth -> T_roof;
           frac = T_roof / T_measured.  > 1 means L1 absorbs the re-touches of a pixel tile (it does).
      l1   the same algorithmic bytes, counted for IN-CUBE samples only (samples outside the cube gather nothing),
           / time / the warp-local 16-B gather ceiling of tools/mem_probe -- the level that binds (ncu: L1/LSU data pipe)
    """
    tj, l1_pct = {}, {}
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f).get(cfg_name, {})
        l1_pct = tj.get("_l1_wavefront_pct", {})
    wf = tj.get("_wavefronts_per_request", {})

    def level(ms, alg_per_sample, per_ray, hbm_bytes, kernel):
        t = ms * 1e-3
        alg_cube = alg_per_sample * cube + per_ray * rays
        traffic = tj.get(kernel)
        o = {"kernel": kernel, "ms": ms,
             "algorithmic_bytes_per_launch": alg_cube,
             "l1": {"achieved_gbs": alg_cube / t / 1e9, "peak_gbs": peaks["l1_gather_gbs"], "frac": alg_cube / t / 1e9 / peaks["l1_gather_gbs"],
                    "lsu_data_pipe_pct_ncu": l1_pct.get(kernel), "wavefronts_per_request_ncu": wf.get(kernel)},
             "l2": {"t_roof_ms": alg_cube / (peaks["l2_gather_gbs"] * 1e9) * 1e3, "peak_gbs": peaks["l2_gather_gbs"],
                    "frac": alg_cube / (peaks["l2_gather_gbs"] * 1e9) / t},
             "hbm": {"compulsory_bytes": hbm_bytes, "achieved_gbs": hbm_bytes / t / 1e9, "peak_gbs": peaks["hbm_gbs"],
                     "frac": hbm_bytes / t / 1e9 / peaks["hbm_gbs"], "traffic_ncu": traffic,
                     "traffic_over_compulsory": (traffic / hbm_bytes) if traffic else None}}
        return o

    fwd = level(fwd_ms, 128, 24, 16 * touched + 24 * rays, "lean_forward_kernel")
    bwd = level(bwd_ms, 256, 12, 48 * touched + 12 * rays, bwd_kernel)
    return {"bound": "l1_lsu", "kernel": bwd_kernel, "achieved": bwd["l1"]["achieved_gbs"], "peak": peaks["l1_gather_gbs"],
            "unit": "GB/s", "frac": bwd["l1"]["frac"], "traffic": bwd["hbm"]["traffic_ncu"],
            "peak_source": "warp-local 16-B gather ceiling measured by tools/mem_probe.cu (" + peaks["sm_side_source"] + "); HBM: " + peaks["hbm_source"],
            "units_per_launch": {"in_cube_live_samples": cube, "live_samples": live, "rays": rays, "touched_voxels": touched},
            "algorithmic_bytes_per_unit": "SURVEY 8(d): forward 128 B per in-cube live sample (8 corners x 16 B) + 24 B per ray; "
                                          "backward 256 B per in-cube live sample (128 B re-gather + 8 x 16 B reds) + 12 B per ray; "
                                          "compulsory HBM: 16 B per touched voxel forward, 48 B backward",
            "note": "neither kernel is HBM-bound (hbm.frac of a few percent): a pixel tile re-touches the same voxels and L1 "
                    "absorbs it, so the binding level is the SM's L1/LSU data pipe (ncu: lsu_data_pipe_pct_ncu)",
            "backward": bwd, "forward": fwd}


def renderer_e2e(cfg, iters=5, warmup=2):
    """The drop-in path a reference caller links: dvren::Renderer::Forward + Backward through libdvren.so with std::vector
    results (apps/dvren_bench.cpp), pageable as the reference's results are, and with the opt-in result pinning."""
    exe = os.path.join(PKG, "dvren_bench")
    if not os.path.exists(exe):
        return {"unavailable": "dvren_bench not built"}
    out = {}
    for key, pin, staged in (("pageable", 0, 0), ("pinned_opt_in", 1, 0), ("staged_path_pinned", 1, 1)):
        try:
            r = subprocess.run([exe, "bench", str(cfg["grid"]), str(cfg["width"]), str(cfg["steps"]), "1" if cfg["stratified"] else "0",
                                str(iters), str(warmup), str(pin), str(staged)], capture_output=True, text=True, timeout=600)
            j = json.loads(r.stdout.strip().splitlines()[-1])
            out[key] = {"value": j["samples"] / (j["ms_per_step"] * 1e-3) / 1e6, "ms_per_step": j["ms_per_step"],
                        "forward_kernel_ms": j["forward_kernel_ms"], "forward_readback_ms": j["forward_readback_ms"],
                        "backward_kernel_ms": j["backward_kernel_ms"], "backward_readback_ms": j["backward_readback_ms"],
                        "d2h_bytes_per_step": j["d2h_bytes_per_step"], "h2d_bytes_per_step": j["h2d_bytes_per_step"]}
        except Exception as e:   # never let the side measurement kill the bench line
            out[key] = {"error": repr(e)[:200]}
    out["what"] = ("dvren::Renderer::Forward + Backward (C++ surface, host clock, every host<->device copy inside), " + cfg["workload"] +
                   "; staged_path_pinned = RenderOptions::use_fused_path = false: the hp.h calls one by one on materialised samples "
                   "(48 B per sample in HBM), the compatibility path")
    return out


class Env:
    """Process-wide plumbing: torch.distributed for rendezvous / barriers / max-over-ranks, one CUDA stream, one hp_ctx."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import dvren_b200 as D
        self.torch, self.dist, self.D = torch, dist, D
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if args.gpus != self.world and self.world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun (python -m torch.distributed.run --nproc-per-node {args.gpus} ...)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.ctx = D.Context(device=self.local_rank, stream=self.stream.cuda_stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, k, warm, sampler=None):
        """W untimed steps, then EXACTLY k steps between CUDA events on the launching stream, barrier + synchronize on both
        sides, max over ranks.  Returns total milliseconds."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.mark_begin()
        a.record(self.stream)
        for _ in range(k):
            fn()
        b.record(self.stream)
        self.barrier()
        if sampler is not None:
            sampler.mark_end()
        ms = torch.tensor([a.elapsed_time(b)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def close(self):
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def measure_single(env, cfg, cfg_name, args, sampler=None, with_renderer=False):
    """One GPU, one frame: resident step, end-to-end step through the C ABI with HOST buffers, per-kernel times, roofline."""
    import hp_abi as A
    import synth as S
    torch, D, ctx, dev, stream = env.torch, env.D, env.ctx, env.dev, env.stream
    n, W, steps = cfg["grid"], cfg["width"], cfg["steps"]
    plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=cfg["stratified"]))
    grid = make_grid(D, S, torch, ctx, n, "thin", dev)      # thin: no early termination, live samples == samples
    frame = D.Frame(plan)
    n_rays = plan.n_rays
    g_host = torch.from_numpy(S.hashed_image_grad(n_rays)).pin_memory()
    g_dev = g_host.to(dev, non_blocking=True)
    grad_ptr, grad_floats = grid.grad_buffer()
    img = frame.image_ptrs()
    pixels = W * W
    planes = [torch.as_tensor(CudaArrayView(img.image.data, pixels * 3), device=dev),
              torch.as_tensor(CudaArrayView(img.trans.data, pixels), device=dev),
              torch.as_tensor(CudaArrayView(img.opacity.data, pixels), device=dev),
              torch.as_tensor(CudaArrayView(img.depth.data, pixels), device=dev)]
    planes_host = [torch.empty(p.shape, dtype=p.dtype).pin_memory() for p in planes]
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO
    lib = ctx.lib

    def step_resident():
        frame.forward(grid)
        frame.backward(grid, g_dev.data_ptr(), flags, device=True)

    # e2e: the same step through the C ABI with HOST buffers (pinned) -- what dvren::Renderer Forward/Backward move per
    # step (reference renderer.hpp:50-66): dL/dI host->device; image planes and the un-interleaved sigma / colour gradient
    # grids device->host.  The image planes are read on a side stream while the backward runs; the gradient read blocks.
    sg_host = torch.empty(grid.voxels, dtype=torch.float32).pin_memory()
    cg_host = torch.empty(grid.voxels * 3, dtype=torch.float32).pin_memory()
    cam_host = torch.empty(16, dtype=torch.float32).pin_memory()
    mask_host = torch.empty(pixels, dtype=torch.int32).pin_memory()
    side = torch.cuda.Stream(device=dev)
    fwd_done = torch.cuda.Event()
    mask_dev = torch.as_tensor(CudaArrayView(img.hitmask.data, pixels), device=dev).view(torch.int32)

    def read_planes_async():
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
        D.check("hpx_grid_read_grad", lib.hpx_grid_read_grad(grid.handle, sg_host.data_ptr(), cg_host.data_ptr(),
                                                             cam_host.data_ptr(), A.HP_MEMSPACE_HOST))
        stream.wait_stream(side)

    def step_e2e_device_grads():
        """Same, but the gradient block stays in HBM for a device-side optimiser (hpx_grid_grad_buffer)."""
        D.check("hpx_forward", lib.hpx_forward(frame.handle, grid.handle))
        read_planes_async()
        D.check("hpx_backward", lib.hpx_backward(frame.handle, grid.handle, g_host.data_ptr(), A.HP_MEMSPACE_HOST, flags))
        stream.wait_stream(side)
        ctx.synchronize()

    k = args.steps
    total_ms = env.timed(step_resident, k, args.warmup, sampler)
    counts = frame.counts()
    samples, live = counts["samples"], counts["live_samples"]
    cube = frame.cube_samples(grid)
    touched = grid.touched_voxels()
    e2e_plain_ms = env.timed(step_e2e, k, max(1, min(args.warmup, 2)))
    e2e_dev_ms = env.timed(step_e2e_device_grads, k, 1)
    fwd_ms = env.timed(lambda: frame.forward(grid), k, 1)
    bwd_ms = env.timed(lambda: frame.backward(grid, g_dev.data_ptr(), D.HPX_BACKWARD_GRID, device=True), k, 1)
    bwd_kernel = "lean_backward_merge_kernel" if frame.scatter_mode(grid, flags) == "merged" else "lean_backward_kernel"
    # e2e, streamed: the SAME host traffic, but the gradient read-back runs UNDER the backward kernel (hpx_backward_streamed:
    # the backward signals per group of image rows, a copy stream un-interleaves and copies the slabs a finished group leaves
    # behind while later rows still render).  Checked once against hpx_grid_read_grad (bitwise).
    def step_e2e_streamed():
        D.check("hpx_forward", lib.hpx_forward(frame.handle, grid.handle))
        read_planes_async()
        D.check("hpx_backward_streamed", lib.hpx_backward_streamed(frame.handle, grid.handle, g_host.data_ptr(), A.HP_MEMSPACE_HOST, flags,
                                                                   sg_host.data_ptr(), cg_host.data_ptr(), cam_host.data_ptr()))
        stream.wait_stream(side)
        stream.synchronize()

    step_e2e_streamed()
    check_sg = torch.empty_like(sg_host)
    check_cg = torch.empty_like(cg_host)
    D.check("hpx_grid_read_grad", lib.hpx_grid_read_grad(grid.handle, check_sg.data_ptr(), check_cg.data_ptr(), None, A.HP_MEMSPACE_HOST))
    assert torch.equal(check_sg, sg_host) and torch.equal(check_cg, cg_host), "streamed gradient read-back differs from hpx_grid_read_grad"
    del check_sg, check_cg
    e2e_ms = env.timed(step_e2e_streamed, k, 1)
    peaks = read_peaks()
    d2h = int(sum(p.numel() for p in planes_host) * 4 + pixels * 4 + grad_floats * 4)
    out = {
        "workload": cfg["workload"], "value": samples / (total_ms / k * 1e-3) / 1e6, "ms_per_step": total_ms / k,
        "rays": n_rays, "samples": samples, "live_samples": live, "in_cube_live_samples": cube, "touched_voxels": touched,
        "fwd": {"ms": fwd_ms / k, "msamples_s": samples / (fwd_ms / k * 1e-3) / 1e6},
        "bwd": {"ms": bwd_ms / k, "msamples_s": samples / (bwd_ms / k * 1e-3) / 1e6, "kernel": bwd_kernel},
        "e2e": {"value": samples / (e2e_ms / k * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(g_host.numel() * 4),
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / k,
                "note": "C ABI with pinned HOST buffers, full dvren::Renderer result contract: image planes AND un-interleaved "
                        "sigma/colour gradient grids copied to host every step (PCIe-bound); the gradient read-back is streamed "
                        "under the backward kernel (hpx_backward_streamed)",
                "without_streaming": {"value": samples / (e2e_plain_ms / k * 1e-3) / 1e6, "ms_per_step": e2e_plain_ms / k,
                                      "what": "hpx_forward, hpx_backward, then hpx_grid_read_grad (round 1's e2e)"},
                "device_resident_gradients": {"value": samples / (e2e_dev_ms / k * 1e-3) / 1e6, "ms_per_step": e2e_dev_ms / k,
                                              "d2h_bytes_per_step": int(sum(p.numel() for p in planes_host) * 4 + pixels * 4)}},
        "roofline": roofline(cfg_name, bwd_kernel, fwd_ms / k, bwd_ms / k, cube, live, n_rays, touched, peaks),
        "l2": "inputs larger than L2 (packed grid %d MB + gradient grid %d MB vs 126 MB)" % (n ** 3 * 16 >> 20, n ** 3 * 16 >> 20),
    }
    frame.close(); grid.close(); plan.close()
    del sg_host, cg_host, planes_host, g_host, g_dev
    torch.cuda.empty_cache()
    if with_renderer:
        out["e2e"]["renderer"] = renderer_e2e(cfg)
    return out


def measure_sparse(env, cfg, args):
    """Empty-space skipping (hpx_grid_build_occupancy) on a SPARSE volume of the same shape: a spherical shell of the dense
    hashed volume (about a fifth of the bricks occupied), same frame, skipping on vs off.  Results are identical by
    construction (tests/test_gpu_runtime.py::test_empty_space_skipping_changes_nothing); this leg records what it buys."""
    import synth as S
    torch, D, ctx, dev = env.torch, env.D, env.ctx, env.dev
    n, W, steps = cfg["grid"], cfg["width"], cfg["steps"]
    sigma, color = device_hashed_volume(torch, n, "dense", dev)
    ax = torch.linspace(0, 1, n, device=dev)
    r = ((ax[None, None, :] - 0.5) ** 2 + (ax[None, :, None] - 0.5) ** 2 + (ax[:, None, None] - 0.5) ** 2).sqrt()
    shell = (r > 0.25) & (r < 0.4)
    sigma = torch.where(shell, sigma, torch.zeros_like(sigma)).contiguous()
    color = torch.where(shell[..., None], color, torch.zeros_like(color)).contiguous()
    torch.cuda.synchronize()
    grid = D.Grid(ctx, sigma.data_ptr(), color.data_ptr(), device_shape=(n, n, n))
    ctx.synchronize()
    del sigma, color, r, shell
    plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=cfg["stratified"]))
    frame = D.Frame(plan)
    g_dev = torch.from_numpy(S.hashed_image_grad(plan.n_rays)).to(dev)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO
    empty_fwd, empty_bwd = grid.build_occupancy(enable=True)
    out = {"volume": "spherical shell 0.25 < r < 0.4 of the dense hashed volume (sigma = 40u), zeros elsewhere",
           "empty_brick_fraction_forward": empty_fwd, "empty_brick_fraction_backward": empty_bwd}
    k = args.steps
    for key, on in (("skipping_on", True), ("skipping_off", False)):
        grid.set_occupancy(on)
        f = env.timed(lambda: frame.forward(grid), k, 2) / k
        b = env.timed(lambda: frame.backward(grid, g_dev.data_ptr(), flags, device=True), k, 2) / k
        c = frame.counts()
        out[key] = {"fwd_ms": f, "bwd_ms": b, "value": c["samples"] / ((f + b) * 1e-3) / 1e6, "live_samples": c["live_samples"]}
    out["speedup"] = out["skipping_on"]["value"] / out["skipping_off"]["value"]
    frame.close(); plan.close(); grid.close()
    torch.cuda.empty_cache()
    return out


def measure_half(env, cfg, args):
    """Half-precision grid storage (hpx_grid_set_storage): the same thin volume and frame with the values kept as four
    halfs per voxel (8 B gathers instead of 16 B; arithmetic stays fp32, results = an fp32 grid of the rounded values)."""
    import synth as S
    torch, D, ctx, dev = env.torch, env.D, env.ctx, env.dev
    n, W, steps = cfg["grid"], cfg["width"], cfg["steps"]
    grid = make_grid(D, S, torch, ctx, n, "thin", dev)
    grid.set_storage(half=True)
    plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=cfg["stratified"]))
    frame = D.Frame(plan)
    g_dev = torch.from_numpy(S.hashed_image_grad(plan.n_rays)).to(dev)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO
    k = args.steps
    f = env.timed(lambda: frame.forward(grid), k, 2) / k
    b = env.timed(lambda: frame.backward(grid, g_dev.data_ptr(), flags, device=True), k, 2) / k
    c = frame.counts()
    out = {"fwd_ms": f, "bwd_ms": b, "value": c["samples"] / ((f + b) * 1e-3) / 1e6, "grid_bytes": n ** 3 * 8,
           "what": "values stored as IEEE half {r,g,b,sigma} (8 B per voxel), widened exactly on load, fp32 arithmetic; gradient "
                   "block stays fp32; parity twin: the oracle on the rounded grid (tests/test_gpu_runtime.py)"}
    frame.close(); plan.close(); grid.close()
    torch.cuda.empty_cache()
    return out


def measure_dense(env, cfg, cfg_name, args):
    """The same frame over the DENSE hashed volume (sigma = 40u): every ray saturates and stops early (T <= 1e-4,
    int_cpu.cpp:209-215), so the work per ray is a fraction of the thin volume's and it varies inside a warp.  `value` keeps the
    metric's definition (samples of the plan per second); the roofline fractions count the LIVE in-cube samples only."""
    import synth as S
    torch, D, ctx, dev = env.torch, env.D, env.ctx, env.dev
    n, W, steps = cfg["grid"], cfg["width"], cfg["steps"]
    grid = make_grid(D, S, torch, ctx, n, "dense", dev)
    plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=cfg["stratified"]))
    frame = D.Frame(plan)
    g_dev = torch.from_numpy(S.hashed_image_grad(plan.n_rays)).to(dev)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO
    k = args.steps
    f = env.timed(lambda: frame.forward(grid), k, 2) / k
    b = env.timed(lambda: frame.backward(grid, g_dev.data_ptr(), flags, device=True), k, 2) / k
    c = frame.counts()
    cube = frame.cube_samples(grid)
    touched = grid.touched_voxels()
    bwd_kernel = "lean_backward_merge_kernel" if frame.scatter_mode(grid, flags) == "merged" else "lean_backward_kernel"
    out = {"volume": "hashed dense (sigma = 40u): early termination on every ray", "fwd_ms": f, "bwd_ms": b,
           "value": c["samples"] / ((f + b) * 1e-3) / 1e6, "samples": c["samples"], "live_samples": c["live_samples"],
           "in_cube_live_samples": cube, "touched_voxels": touched,
           "live_msamples_s": c["live_samples"] / ((f + b) * 1e-3) / 1e6,
           "roofline": roofline(cfg_name + "_dense", bwd_kernel, f, b, cube, c["live_samples"], plan.n_rays, touched, read_peaks())}
    frame.close(); plan.close(); grid.close()
    torch.cuda.empty_cache()
    return out


def base_line(env, args, cfg, value, ms_per_step, scaling):
    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}


def run_single(env, args):
    cfg = CONFIGS[args.config]
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    m = measure_single(env, cfg, args.config, args, sampler, with_renderer=args.config == "c2")
    clocks = sampler.stop()
    line = base_line(env, args, cfg, m["value"], m["ms_per_step"], "strong")
    line["config"] = {"workload": m["workload"], "volume": "hashed thin (sigma = 2u, no early termination)", "rays": m["rays"],
                      "samples_per_step": m["samples"], "live_samples_per_step": m["live_samples"],
                      "in_cube_live_samples_per_step": m["in_cube_live_samples"], "parallelism": "single GPU", "l2": m["l2"]}
    line.update({"e2e": m["e2e"], "gpu_launches": 2 * args.steps, "clocks": clocks, "fwd": m["fwd"], "bwd": m["bwd"],
                 "roofline": m["roofline"]})
    if args.config == "c3" and not args.no_c2:
        c2 = measure_single(env, CONFIGS["c2"], "c2", args, None, with_renderer=True)
        line["c2"] = {k: c2[k] for k in ("workload", "value", "ms_per_step", "fwd", "bwd", "e2e", "roofline", "samples",
                                         "in_cube_live_samples", "touched_voxels")}
        line["c2"]["empty_space_skipping"] = measure_sparse(env, CONFIGS["c2"], args)
        line["c2"]["half_storage"] = measure_half(env, CONFIGS["c2"], args)
        try:   # (a side measurement never takes the line down)
            line["c2"]["dense_volume"] = measure_dense(env, CONFIGS["c2"], "c2", args)
        except Exception as e:
            line["c2"]["dense_volume"] = {"error": repr(e)[:200]}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, rows=args.cpu_rows, threads=1)
    emit(line)


def run_sharded(env, args):
    """N > 1: ONE frame of the configuration, strong scaling, through the library's sharding (hp_b200.h hpx_comm / hpx_shard):
    work-balanced contiguous row bands (default) or interleaved tile rows; grid replicated."""
    import numpy as np
    import synth as S
    torch, dist, D, ctx, dev, stream = env.torch, env.dist, env.D, env.ctx, env.dev, env.stream
    world, rank = env.world, env.rank
    cfg = CONFIGS[args.config]
    n, W, steps = cfg["grid"], cfg["width"], cfg["steps"]
    # rendezvous: rank 0's NCCL id travels over torch.distributed
    uid = torch.zeros(D.HPX_COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(D.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    comm = D.Comm(ctx, bytes(uid.cpu().numpy().tobytes()), rank, world)
    full = S.bench_plan(W, W, steps, stratified=cfg["stratified"])
    plan = D.Plan(ctx, full)
    grid = make_grid(D, S, torch, ctx, n, "thin", dev)
    bands = args.sharding == "bands"
    shard = D.Shard(comm, plan, grid, bands="replicated") if bands else \
        D.Shard(comm, plan, grid, [0.72 ** g for g in range(args.groups)])
    g_host = torch.from_numpy(S.hashed_image_grad(W * W)).pin_memory()
    g_dev = g_host.to(dev, non_blocking=True)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO
    grad_ptr, grad_floats = grid.grad_buffer()
    block = torch.as_tensor(CudaArrayView(grad_ptr, grad_floats), device=dev)

    def step():
        shard.step(g_dev.data_ptr(), flags)

    # untimed: the library re-cuts the bands from the measured per-rank time of a step (hpx_shard_rebalance) until they
    # stop moving -- start-up work of a training loop, like its warm-up
    rebalances = 0
    if bands:
        step()
        shard.tune_order(g_dev.data_ptr(), flags)   # rank-local: the band's tile dispatch order by measurement
        for _ in range(4):
            step(); step()
            if not shard.rebalance():
                break
            rebalances += 1
    layout = shard.layout()
    layout["rebalance_rounds"] = rebalances

    # correctness first: every rank must hold the gradient a single GPU computes for the whole frame (rank 0 checks)
    step()
    env.barrier()
    verify = None
    if rank == 0:
        got = block.clone()
        fplan = D.Plan(ctx, full)
        frame = D.Frame(fplan)
        gref = make_grid(D, S, torch, ctx, n, "thin", dev)
        gref.set_grad_layout("xyz".index(layout["slow_axis"]))
        frame.forward(gref)
        frame.backward(gref, g_dev.data_ptr(), flags, device=True)
        torch.cuda.synchronize()
        rptr, rfloats = gref.grad_buffer()
        ref = torch.as_tensor(CudaArrayView(rptr, rfloats), device=dev)
        # both are float32 red accumulations of the same contributions in different orders (each kernel has its own oracle
        # parity tests, tests/test_gpu_lean.py); this catches a slab summed twice or not at all, hence the 1e-2 floor
        scale = torch.maximum(ref.abs(), 1e-2 * ref.abs().max())
        verify = float(((got - ref).abs() / scale).max().item())
        del got, ref
        frame.close(); gref.close(); fplan.close()
        torch.cuda.empty_cache()
    env.barrier()

    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    k = args.steps
    results = {}
    # headline (bands): the reduce-scatter contract -- every slab's finished sum on its owner; the all-reduce contract
    # (every rank holds everything) is timed right after it
    if bands:
        shard.set_result("owned")
        rounds = layout["rebalance_rounds"]
        layout = shard.layout()        # the owner cuts of this result mode
        layout["rebalance_rounds"] = rounds
        mine = torch.tensor([layout["tile_order"]], dtype=torch.int32, device=dev)   # hpx_frame_set_row_order value per rank
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        layout["tile_order"] = [int(t.item()) for t in every]
    total_ms = env.timed(step, k, args.warmup, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    results["owned" if bands else "replicated"] = total_ms / k
    total = W * W * steps
    frame = shard.frame
    my_samples = frame.counts()["samples"] if frame is not None else 0
    ts = torch.tensor([my_samples], dtype=torch.int64, device=dev)
    dist.all_reduce(ts)
    assert int(ts.item()) == total, (int(ts.item()), total)
    if bands:
        shard.set_result("replicated")
        results["replicated"] = env.timed(step, k, 2) / k
        shard.set_result("owned")
    shard.set_reduce(False)
    results["without_exchange"] = env.timed(step, k, 1) / k
    shard.set_reduce(True)

    # e2e (owned result, what a slab-sharded optimiser on the host side would consume): per step every rank copies ITS
    # rows of dL/dI host->device, runs the step, and reads ITS band of the image planes and ITS owned slabs device->host
    e2e = None
    if bands:
        row0, rows = layout["band_row0"][rank], layout["band_rows"][rank]
        optr, first, count, slab_floats = shard.owned()
        owned_dev = torch.as_tensor(CudaArrayView(optr, max(count * slab_floats, 1)), device=dev)[:count * slab_floats]
        owned_host = torch.empty(count * slab_floats, dtype=torch.float32).pin_memory()
        img = frame.image_ptrs() if frame is not None else None
        lo, hi = row0 * W, (row0 + rows) * W
        planes, hosts = [], []
        if img is not None and rows:
            for ptr, ch, ts_ in ((img.image.data, 3, "<f4"), (img.trans.data, 1, "<f4"), (img.opacity.data, 1, "<f4"),
                                 (img.depth.data, 1, "<f4"), (img.hitmask.data, 1, "<u4")):
                t = torch.as_tensor(CudaArrayView(ptr, W * W * ch, "<f4"), device=dev)[lo * ch:hi * ch]
                planes.append(t)
                hosts.append(torch.empty(t.shape, dtype=t.dtype).pin_memory())
        g_rows_host = g_host.reshape(-1)[lo * 3:hi * 3]
        g_rows_dev = g_dev.reshape(-1)[lo * 3:hi * 3]

        def step_e2e():
            g_rows_dev.copy_(g_rows_host, non_blocking=True)
            shard.step(g_dev.data_ptr(), flags)
            owned_host.copy_(owned_dev, non_blocking=True)
            for h, d in zip(hosts, planes):
                h.copy_(d, non_blocking=True)
            stream.synchronize()

        e2e_ms = env.timed(step_e2e, k, 1) / k
        h2d = torch.tensor([g_rows_host.numel() * 4], dtype=torch.int64, device=dev)
        d2h = torch.tensor([owned_host.numel() * 4 + sum(h.numel() for h in hosts) * 4], dtype=torch.int64, device=dev)
        dist.all_reduce(h2d)
        dist.all_reduce(d2h)
        e2e = {"value": total / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": int(d2h.item()),
               "note": "all ranks together: every rank moves its own rows of dL/dI in and its band of the image planes + the "
                       "finished gradient sums of the slabs it owns out (packed {dr,dg,db,dsigma} slabs), over its own PCIe link"}

    # continuity with round 1: weak scaling of configs[1] (one view of SURVEY 8d's orbit per GPU, whole-block all-reduce in
    # the library)
    weak = None
    if not args.no_c2:
        weak = run_weak_views(env, args, comm)

    ms = results["owned"] if bands else results["replicated"]
    line = base_line(env, args, cfg, total / (ms * 1e-3) / 1e6, ms, "strong")
    line["config"] = {
        "workload": cfg["workload"], "volume": "hashed thin (sigma = 2u, no early termination)",
        "samples_per_step": total,
        "parallelism": (f"ONE frame in {world} contiguous row bands cut for equal marching work (hpx_shard_create_bands), grid "
                        "replicated; the gradient block is laid out in slabs along world axis " + layout["slow_axis"] +
                        ", a band's backward touches one slab wedge; the wedge parts a rank does not own are added by their "
                        "owners in rank order (" + layout.get("exchange", "") + ").  RESULT of the timed step: a reduce-scatter -- every "
                        "slab's finished sum (the gradient of ALL rays of the frame) lives on the rank that owns the slab, the "
                        "hand-over for a slab-sharded optimiser (hpx_shard_owned; e2e reads exactly that to the host).  The "
                        "all-reduce contract -- every rank additionally fetches the other owners' sums, so that all ranks hold the "
                        "whole gradient -- is timed as well: replicated_result") if bands else
                       (f"ONE frame, tile rows interleaved over {world} GPUs, {args.groups} row groups, slab all-reduces behind a "
                        "device-signalled backward (hpx_shard_create)"),
        "layout": layout, "verify_max_rel_err_vs_single_gpu": verify,
        "ms_per_step_without_exchange": results["without_exchange"],
        "exchange_bytes_sent_by_rank0": layout.get("send_bytes"), "l2": "inputs larger than L2"}
    if bands:
        line["replicated_result"] = {"value": total / (results["replicated"] * 1e-3) / 1e6, "ms_per_step": results["replicated"],
                                     "what": "same step followed by the all-gather of the owned sums: every rank holds the whole "
                                             "summed gradient, as after an all-reduce (SURVEY 8e's collective); verified against a "
                                             "single-GPU backward of the frame: config.verify_max_rel_err_vs_single_gpu"}
    if e2e is not None:
        line["e2e"] = e2e
    # per rank and step: lean_forward_kernel, lean_backward_merge_kernel, peer_reduce_kernel (bands; the two barriers of the
    # exchange are 4-byte NCCL all-reduces and not counted)
    line["gpu_launches"] = (3 if bands else 2) * args.steps
    line["clocks"] = clocks
    if weak is not None:
        line["c2_weak"] = weak
    if rank == 0:
        emit(line)
    shard.close(); comm.close(); grid.close(); plan.close()


def run_weak_views(env, args, comm):
    """configs[1] weak scaling: rank j renders view j of SURVEY 8(d)'s orbit (360 deg * j / N about the cube centre), then
    hpx_grid_allreduce_grad sums the whole 268 MB gradient block in the library."""
    import synth as S
    torch, dist, D, ctx, dev = env.torch, env.dist, env.D, env.ctx, env.dev
    cfg = CONFIGS["c2"]
    n, W, steps = cfg["grid"], cfg["width"], cfg["steps"]
    plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=True, view=env.rank, views=env.world))
    grid = make_grid(D, S, torch, ctx, n, "thin", dev)
    frame = D.Frame(plan)
    g_dev = torch.from_numpy(S.hashed_image_grad(W * W)).to(dev)
    flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO
    ptr, floats = grid.grad_buffer()
    block = torch.as_tensor(CudaArrayView(ptr, floats), device=dev)

    def local():
        frame.forward(grid)
        frame.backward(grid, g_dev.data_ptr(), flags, device=True)

    def step():
        local()
        comm.allreduce_grad(grid)

    # the all-reduced block must equal the sum of the per-view gradients (checked through two independent reductions:
    # the library's NCCL call against torch.distributed's on a copy)
    local()
    mine = block.clone()
    dist.all_reduce(mine)
    step()
    env.barrier()
    scale = torch.maximum(mine.abs(), 1e-2 * mine.abs().max())   # two float32 reductions in different orders
    err = ((block - mine).abs() / scale).max()
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    assert float(err.item()) < 1e-3, f"all-reduced gradient block differs from the sum of the per-view gradients: {float(err.item()):.3e}"
    ms = env.timed(step, args.steps, 2) / args.steps
    ms_local = env.timed(local, args.steps, 1) / args.steps
    total = env.world * W * W * steps
    out = {"workload": cfg["workload"] + f"; {env.world} views on the orbit of SURVEY 8(d), one per GPU, whole-block all-reduce "
                                         "(hpx_grid_allreduce_grad)",
           "scaling": "weak", "value": total / (ms * 1e-3) / 1e6, "ms_per_step": ms, "ms_per_step_without_allreduce": ms_local,
           "allreduce_bytes": floats * 4, "allreduced_block_vs_sum_of_views_max_rel_err": float(err.item())}
    del mine, block
    frame.close(); grid.close(); plan.close()
    torch.cuda.empty_cache()
    return out


def run_view_batch(env, args):
    """Configs 4 / 5: a batch of views through ONE captured CUDA graph (forward + backward to the grid [+ camera]); the
    view changes between replays through the frame's device parameter block.  N > 1: views split across ranks, gradient
    block all-reduced by the library."""
    import numpy as np
    import synth as S
    torch, dist, D, ctx, dev, stream = env.torch, env.dist, env.D, env.ctx, env.dev, env.stream
    world, rank = env.world, env.rank
    cfg = CONFIGS[args.config]
    n, W, steps, views = cfg["grid"], cfg["width"], cfg["steps"], cfg["views"]
    comm = None
    if world > 1:
        uid = torch.zeros(D.HPX_COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(D.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = D.Comm(ctx, bytes(uid.cpu().numpy().tobytes()), rank, world)
    per = (views + world - 1) // world
    my_views = list(range(rank * per, min(views, (rank + 1) * per)))
    descs = [S.bench_plan(W, W, steps, stratified=cfg["stratified"], view=v, views=views) for v in my_views]
    plan = D.Plan(ctx, descs[0])
    if cfg.get("device_volume"):
        gen = torch.Generator(device=dev).manual_seed(1234)   # same seed on every rank = replicas
        sigma = torch.rand((n, n, n), generator=gen, device=dev, dtype=torch.float32) * 2.0
        color = torch.rand((n, n, n, 3), generator=gen, device=dev, dtype=torch.float32)
        torch.cuda.synchronize()
        grid = D.Grid(ctx, sigma.data_ptr(), color.data_ptr(), device_shape=(n, n, n))
        ctx.synchronize()
        del sigma, color
        torch.cuda.empty_cache()
    else:
        grid = make_grid(D, S, torch, ctx, n, "thin", dev)
    frame = D.Frame(plan)
    g_host = torch.from_numpy(S.hashed_image_grad(plan.n_rays)).pin_memory()
    g_frame = torch.as_tensor(CudaArrayView(frame.grad_input_ptr(), plan.n_rays * 3), device=dev)
    g_frame.copy_(g_host.reshape(-1), non_blocking=True)
    grad_ptr, grad_floats = grid.grad_buffer()
    grad_view = torch.as_tensor(CudaArrayView(grad_ptr, grad_floats), device=dev)
    flags = D.HPX_BACKWARD_GRID | (D.HPX_BACKWARD_CAMERA if cfg.get("camera", True) else 0)
    frame.capture(grid, flags)
    cams = [d.camera for d in descs]
    cam_host = torch.empty(16, dtype=torch.float32).pin_memory()

    def step():
        grid.zero_grad()
        for v, cam in zip(my_views, cams):
            frame.set_view(cam, 42 + v, 0)
            frame.replay()
        if comm is not None:
            comm.allreduce_grad(grid)

    def step_e2e():
        step()
        cam_host.copy_(grad_view[-16:], non_blocking=True)
        frame.read()

    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    ms_per_step = env.timed(step, args.steps, args.warmup, sampler if rank == 0 else None) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = env.timed(step_e2e, args.steps, 1) / args.steps
    samples = views * plan.n_rays * steps
    line = base_line(env, args, cfg, samples / (ms_per_step * 1e-3) / 1e6, ms_per_step, "strong")
    line["metric"] = METRIC + (" + camera adjoint" if cfg.get("camera", True) else "")
    line["config"] = {"workload": cfg["workload"], "views": views, "views_per_gpu": len(my_views),
                      "volume": "device-generated uniform (sigma = 2u)" if cfg.get("device_volume") else "hashed thin (sigma = 2u)",
                      "samples_per_step": samples, "allreduce_bytes": int(grad_floats * 4) if world > 1 else 0,
                      "backward_kernel": ("lean_backward_merge_kernel" if frame.scatter_mode(grid, flags) == "merged"
                                          else "lean_backward_kernel") + (" (+ camera adjoint)" if cfg.get("camera", True) else ""),
                      "l2": "inputs larger than L2 (grid %d MB + gradient %d MB)" % (n ** 3 * 16 >> 20, n ** 3 * 16 >> 20)}
    line["e2e"] = {"value": samples / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": len(my_views) * 128, "d2h_bytes_per_step": 64 + W * W * 28}
    line.update({"gpu_launches": 4 * len(my_views) * args.steps, "clocks": clocks, "ms_per_view": ms_per_step / max(len(my_views), 1)})
    if rank == 0:
        emit(line)
    frame.close(); grid.close(); plan.close()
    if comm is not None:
        comm.close()


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


_CPU_STATE = {}


def cpu_baseline(cfg, rows: int, threads: int, repeats: int = 1, keep: bool = False):
    """The reference's own CPU implementation of the path (oracle/_ref: the UNMODIFIED reference compiled here; its hp_ray ->
    hp_samp_int_fused -> hp_img, hp_diff -> DenseGridField::AccumulateSampleGradients call sequence, oracle/ref_shim.cpp
    ref_worker_*) -- or the oracle port when that library is absent -- on `threads` host threads, each rendering a band of
    `rows` image rows of the same workload around the image centre.  keep: the volume and the per-thread reference objects
    (fields + a DenseGridField as scatter target, 4.3 GB each at 512^3) stay alive for the next call."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import numpy as np

    import oracle as O
    import synth as S

    n, W, steps = cfg["grid"], cfg["width"], cfg["steps"]
    use_ref = O.ref_available()
    if not use_ref:
        O.build_oracle()
    key = (n, threads, use_ref)
    state = _CPU_STATE.get(key)
    if state is None:
        sigma, color = S.hashed_volume(n, "thin")
        state = {"sigma": sigma, "color": color,
                 "workers": [O.RefWorker(sigma, color) for _ in range(threads)] if use_ref else None}
        if keep:
            _CPU_STATE[key] = state
    sigma, color, workers = state["sigma"], state["color"], state["workers"]
    results = [None] * threads
    y_start = max(0, (W - rows * threads) // 2)

    def work(t):
        y0 = y_start + t * rows
        desc = S.bench_plan(W, W, steps, stratified=cfg["stratified"], roi=(0, y0, W, rows))
        dl = S.hashed_image_grad(W * rows)
        cnt, ms = 0, 0.0
        for _ in range(repeats):
            if use_ref:
                c, f, b = workers[t].run(desc, dl)
                cnt += c
                ms += f + b
            else:
                t0 = time.perf_counter()
                st, rd = O.plan_resolve(desc)
                gs, gc = O.make_grid(sigma, 1), O.make_grid(color, 3)
                r = O.render(rd, gs, gc, dl, per_ray=False, frames=True)
                ms += (time.perf_counter() - t0) * 1e3
                cnt += r["sample_count"]
        results[t] = (cnt, ms)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    wall = time.perf_counter() - t0
    if workers and not keep:
        for w in workers:
            w.close()
    total = sum(r[0] for r in results)
    busy_ms = max(r[1] for r in results)   # throughput of the hot path itself: samples / (forward + backward time), slowest thread
    return {"value": total / (busy_ms * 1e-3) / 1e6, "unit": UNIT, "cores": threads,
            "kind": "reference" if use_ref else "port",
            "sample": f"{threads} band(s) of {rows} rows x {W} px x {steps} steps x {repeats} = {total} samples, reference hp_ray -> "
                      f"hp_samp_int_fused -> hp_img -> hp_diff -> AccumulateSampleGradients time {busy_ms:.0f} ms (wall {wall:.1f} s)",
            "busy_ms": busy_ms}


def run_reference(args):
    """--impl reference: the reference's CPU implementation on all host threads the process may use, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = CONFIGS[args.config]
    threads = host_threads() if args.cpu_threads <= 0 else min(host_threads(), args.cpu_threads)
    if cfg["grid"] >= 512:
        threads = min(threads, 24)   # every worker owns a 4.3 GB gradient target at 512^3 (reference DenseGridField)
    for _ in range(max(1, min(args.warmup, 2))):   # builds the volume and the per-thread reference objects once
        cpu_baseline(cfg, rows=1, threads=threads, keep=True)
    t0 = time.perf_counter()
    vals = [cpu_baseline(cfg, rows=args.cpu_rows_ref, threads=threads, keep=True) for _ in range(args.steps)]
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    v = sum(x["value"] for x in vals) / len(vals)
    base = vals[-1]
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "volume": "hashed thin (sigma = 2u, no early termination)"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-rows", type=int, default=16, help="image rows of the cpu_baseline sample")
    ap.add_argument("--cpu-rows-ref", type=int, default=4, help="rows per thread per step for --impl reference")
    ap.add_argument("--cpu-threads", type=int, default=0, help="host threads of --impl reference (0 = all the process may use)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c2", action="store_true", help="skip the configs[1] continuity numbers")
    ap.add_argument("--sharding", default="bands", choices=["bands", "interleaved"],
                    help="N > 1: work-balanced row bands with a sparse slab exchange (default), or interleaved tile rows with slab "
                         "all-reduces behind a device-signalled backward")
    ap.add_argument("--groups", type=int, default=4, help="row groups of --sharding interleaved")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3   # timing rule: at least 3 warm-up steps
    env = Env(args)
    if "views" in CONFIGS[args.config]:
        run_view_batch(env, args)
    elif env.world == 1:
        run_single(env, args)
    else:
        run_sharded(env, args)
    env.close()


if __name__ == "__main__":
    main()
