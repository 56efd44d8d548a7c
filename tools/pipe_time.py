"""Single-GPU cost of the strong-scaling pipeline ingredients on config 2: gradient layout, row groups, interleave."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "diff-volume-renderer_b200", "python")]
import numpy as np, torch
import dvren_b200 as D, synth as S, sharding as SH

n, W, steps = 256, 1024, 512
sig, col = S.hashed_volume(n, "thin")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
dev = torch.device("cuda", 0)
ctx = D.Context(device=0, stream=stream.cuda_stream)
grid = D.Grid(ctx, sig, col)
dl = torch.from_numpy(S.hashed_image_grad(W * W)).cuda()
full = S.bench_plan(W, W, steps, stratified=True)
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
plan = D.Plan(ctx, full); frame = D.Frame(plan)
for axis in (2, 1, 0):
    grid.set_grad_layout(axis)
    f = timeit(lambda: frame.forward(grid))
    b = timeit(lambda: frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO, device=True))
    print(json.dumps(dict(what="plain", slow_axis=axis, fwd_ms=f, bwd_ms=b)), flush=True)
frame.close(); plan.close()
for groups in (1, 2, 4, 8):
    pf = SH.PipelinedFrame(D, ctx, grid, full, groups, 1, 0, dev, stream)
    t = timeit(lambda: pf.step(dl.data_ptr(), D.HPX_BACKWARD_GRID))
    print(json.dumps(dict(what="pipeline world=1", groups=groups, step_ms=t)), flush=True)
    pf.close()
for world in (2, 8):   # one rank's share of an interleaved frame (no collectives): should be step / world
    pf = SH.PipelinedFrame(D, ctx, grid, full, 4, 1, 0, dev, stream)
    for p in pf.parts: p["frame"].set_interleave(world, 0)
    t = timeit(lambda: pf.step(dl.data_ptr(), D.HPX_BACKWARD_GRID))
    print(json.dumps(dict(what="one rank of an interleaved frame", world=world, groups=4, step_ms=t, ideal_ms=8.26 / world)), flush=True)
    pf.close()
