"""Per-ray vs merged backward as a function of the pixel / voxel spacing ratio (tunes merge_scatter_pays)."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "diff-volume-renderer_b200", "python")]
import numpy as np, torch
import dvren_b200 as D, synth as S
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = D.Context(device=0, stream=stream.cuda_stream)
def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for n, W, steps in ((256, 640, 512), (256, 512, 512), (256, 448, 512), (256, 384, 512), (512, 1024, 512)):
    gen = torch.Generator(device="cuda").manual_seed(5)
    sigma = torch.rand((n, n, n), generator=gen, device="cuda") * 2.0
    color = torch.rand((n, n, n, 3), generator=gen, device="cuda")
    torch.cuda.synchronize()
    grid = D.Grid(ctx, sigma.data_ptr(), color.data_ptr(), device_shape=(n, n, n)); ctx.synchronize()
    del sigma, color; torch.cuda.empty_cache()
    plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=True)); frame = D.Frame(plan)
    dl = torch.from_numpy(S.hashed_image_grad(W * W)).cuda()
    f = timeit(lambda: frame.forward(grid))
    res = dict(n=n, W=W, steps=steps, spacing_voxels=1.5 / (1.2 * W) * (n - 1), auto=frame.scatter_mode(grid), fwd_ms=f)
    for name, flag in (("per_ray", D.HPX_BACKWARD_SCATTER_PER_RAY), ("merged", D.HPX_BACKWARD_SCATTER_MERGED)):
        res[name + "_ms"] = timeit(lambda: frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | flag, device=True))
    print(json.dumps(res), flush=True)
    frame.close(); plan.close(); grid.close()
