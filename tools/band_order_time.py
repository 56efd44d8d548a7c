"""One GPU, the bands hpx_shard_create_bands cuts for `world` ranks rendered one after the other under every tile dispatch
order (hpx_frame_set_row_order): forward + backward ms per band and order, and the order the library would choose.

    python tools/band_order_time.py [c3] [world=8]
"""
import os, sys, json, ctypes as C
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "diff-volume-renderer_b200", "python"), REPO]
import numpy as np, torch
import dvren_b200 as D, synth as S
import bench as B

cfg = {"c2": (256, 1024, 512, True), "c3": (512, 2048, 1024, False)}[sys.argv[1] if len(sys.argv) > 1 else "c3"]
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n, W, steps, strat = cfg
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = D.Context(device=0, stream=stream.cuda_stream)
lib = ctx.lib
full = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=strat))
grid = B.make_grid(D, S, torch, ctx, n, "thin", dev)
g_dev = torch.from_numpy(S.hashed_image_grad(W * W)).to(dev)
flags = D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO
row0, rows = (C.c_uint32 * world)(), (C.c_uint32 * world)()
D.check("bands", lib.hpx_plan_balanced_bands(full.handle, world, row0, rows, None))
usable, total = C.c_uint32(), C.c_uint32()
D.check("sms", lib.hpx_ctx_sm_counts(ctx.handle, C.byref(usable), C.byref(total)))

def timeit(fn, k=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(k): fn()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / k

orders = [0, 1, 2, D.HPX_ORDER_COLUMNS, D.HPX_ORDER_COLUMNS | 1]
for r in range(world):
    plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=strat, roi=(0, int(row0[r]), W, int(rows[r]))))
    frame = D.Frame(plan)
    frame.set_view(None, 42, int(row0[r]) * W)
    dl = g_dev.data_ptr() + int(row0[r]) * W * 3 * 4
    chosen, ends = C.c_int32(), (C.c_double * 4)()
    D.check("best", lib.hpx_plan_best_tile_order(full.handle, row0[r], rows[r], max(1, usable.value) * 5, C.byref(chosen), ends))
    line = {"rank": r, "row0": int(row0[r]), "rows": int(rows[r]), "library_choice": chosen.value,
            "estimated_end_steps": dict(zip(("0", "1", "4", "5"), [round(e) for e in ends]))}
    for o in orders:
        D.check("order", lib.hpx_frame_set_row_order(frame.handle, o))
        f = timeit(lambda: frame.forward(grid))
        b = timeit(lambda: frame.backward(grid, dl, flags, device=True))
        line[f"order_{o}_ms"] = {"forward": round(f, 4), "backward": round(b, 4), "step": round(f + b, 4)}
    os.write(B._REAL_STDOUT, (json.dumps(line) + "\n").encode())   # (importing bench points descriptor 1 at stderr)
    frame.close(); plan.close()
