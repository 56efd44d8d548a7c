"""Kernel time of config 2 as a function of the orbit view (layout / orientation sensitivity)."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "diff-volume-renderer_b200", "python")]
import numpy as np, torch
import dvren_b200 as D, synth as S

n, W, steps = 256, 1024, 512
sig, col = S.hashed_volume(n, "thin")
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)   # a non-default stream: the library enqueues on it and the events below see it
ctx = D.Context(device=0, stream=_stream.cuda_stream)
grid = D.Grid(ctx, sig, col)
dl = torch.from_numpy(S.hashed_image_grad(W * W)).cuda()
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for view, views in [(0, 8), (1, 8), (2, 8), (3, 8), (1, 16), (1, 3)]:
    plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=True, view=view, views=views))
    frame = D.Frame(plan)
    f = timeit(lambda: frame.forward(grid))
    b = timeit(lambda: frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO, device=True))
    c = frame.counts()
    print(json.dumps(dict(view=view, views=views, deg=360.0 * view / views, fwd_ms=f, bwd_ms=b, live=c["live_samples"], mode=frame.scatter_mode(grid))), flush=True)
    frame.close(); plan.close()
