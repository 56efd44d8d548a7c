"""Scratch timing of the lean kernels (CUDA events on the launching stream)."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "diff-volume-renderer_b200", "python")]
import numpy as np, torch
import dvren_b200 as D, hp_abi as A, synth as S

def run(n_grid, W, steps, strat, kind, iters=5):
    sig, col = S.hashed_volume(n_grid, kind)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = D.Context(device=0, stream=stream)
    plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=strat))
    grid = D.Grid(ctx, sig, col)
    frame = D.Frame(plan)
    dl = torch.from_numpy(S.hashed_image_grad(W * W)).cuda()
    grid.zero_grad()
    def timeit(fn):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        return float(np.median(ts))
    f = timeit(lambda: frame.forward(grid))
    c = frame.counts()
    b = timeit(lambda: frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_SCATTER_PER_RAY, device=True))
    bm = timeit(lambda: frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_SCATTER_MERGED, device=True))
    z = timeit(lambda: frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_ZERO, device=True))
    cam = timeit(lambda: frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_CAMERA, device=True))
    fused = timeit(lambda: frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_CAMERA | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_SCATTER_MERGED, device=True))
    det = timeit(lambda: frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_DETERMINISTIC, device=True))
    M = c["samples"]
    print(json.dumps(dict(grid=n_grid, W=W, steps=steps, strat=strat, kind=kind, samples=M, live=c["live_samples"],
          fwd_ms=f, bwd_ms=b, bwd_merged_ms=bm, zero_ms=z, cam_ms=cam, bwd_cam_fused_ms=fused, bwd_deterministic_ms=det, fwd_gsamp=M / f / 1e6, fwdbwd_gsamp=M / (f + b) / 1e6,
          live_fwd_gsamp=c["live_samples"] / f / 1e6)), flush=True)
    frame.close(); grid.close(); plan.close(); ctx.close()

if __name__ == "__main__":
  if os.environ.get("DVREN_HP_LIB"):
    D._lib = D.load(os.environ["DVREN_HP_LIB"])
  with torch.cuda.stream(torch.cuda.Stream()):
    run(64, 512, 256, False, "thin")
    run(64, 512, 256, False, "dense")
    run(256, 1024, 512, True, "thin")
    run(256, 1024, 512, True, "dense")
    run(256, 1024, 512, False, "thin")
