"""Scratch timing of hpx_backward_streamed pieces (DVREN_STREAM_DEBUG=0 all, 1 no D2H copies, 2 no un-interleave either)."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "diff-volume-renderer_b200", "python"), REPO]
import numpy as np, torch
import dvren_b200 as D, hp_abi as A, synth as S
import bench as B

cfg = {"c2": (256, 1024, 512, True), "c3": (512, 2048, 1024, False)}[sys.argv[1] if len(sys.argv) > 1 else "c3"]
n, W, steps, strat = cfg
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = D.Context(device=0, stream=stream.cuda_stream)
plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=strat))
grid = B.make_grid(D, S, torch, ctx, n, "thin", dev)
frame = D.Frame(plan)
g_host = torch.from_numpy(S.hashed_image_grad(W * W)).pin_memory()
g_dev = g_host.to(dev)
sg = torch.empty(n ** 3, dtype=torch.float32).pin_memory()
cg = torch.empty(3 * n ** 3, dtype=torch.float32).pin_memory()
flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO
lib = ctx.lib

def timeit(fn, k=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(k): fn()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / k

frame.forward(grid)
out = {"debug": os.environ.get("DVREN_STREAM_DEBUG", "0")}
out["plain_backward_default_layout"] = timeit(lambda: frame.backward(grid, g_dev.data_ptr(), flags, device=True))
out["streamed"] = timeit(lambda: D.check("s", lib.hpx_backward_streamed(frame.handle, grid.handle, g_dev.data_ptr(), A.HP_MEMSPACE_DEVICE, flags, sg.data_ptr(), cg.data_ptr(), None)))
out["plain_backward_slab_layout"] = timeit(lambda: frame.backward(grid, g_dev.data_ptr(), flags, device=True))
out["read_grad_host"] = timeit(lambda: D.check("r", lib.hpx_grid_read_grad(grid.handle, sg.data_ptr(), cg.data_ptr(), None, A.HP_MEMSPACE_HOST)), 3)
os.write(B._REAL_STDOUT, (json.dumps(out) + "\n").encode())   # (importing bench points descriptor 1 at stderr)
