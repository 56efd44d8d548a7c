"""One forward + backward (both scatter strategies) of a bench config, for ncu captures."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "diff-volume-renderer_b200", "python")]
import numpy as np, torch
import dvren_b200 as D, synth as S

cfg = {"c1": (64, 512, 256, False), "c2": (256, 1024, 512, True), "c3": (512, 2048, 1024, False)}[sys.argv[1] if len(sys.argv) > 1 else "c2"]
kind = sys.argv[2] if len(sys.argv) > 2 else "thin"
n, W, steps, strat = cfg
sig, col = S.hashed_volume(n, kind)
ctx = D.Context(device=0)
plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=strat))
grid = D.Grid(ctx, sig, col)
frame = D.Frame(plan)
dl = torch.from_numpy(S.hashed_image_grad(W * W)).cuda()
for _ in range(2):
    frame.forward(grid)
    frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_SCATTER_PER_RAY, device=True)
    frame.backward(grid, dl.data_ptr(), D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | D.HPX_BACKWARD_SCATTER_MERGED, device=True)
ctx.synchronize()
print("ok", frame.counts())
