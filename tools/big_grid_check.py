import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "diff-volume-renderer_b200", "python")]
import numpy as np, torch
import dvren_b200 as D, synth as S
ctx = D.Context(device=0)
for n, W, steps in ((256, 128, 256), (512, 1024, 1024), (512, 128, 1024), (1024, 256, 1024)):
    gen = torch.Generator(device="cuda").manual_seed(5)
    sigma = torch.rand((n, n, n), generator=gen, device="cuda") * 3.0
    color = torch.rand((n, n, n, 3), generator=gen, device="cuda")
    torch.cuda.synchronize()
    grid = D.Grid(ctx, sigma.data_ptr(), color.data_ptr(), device_shape=(n, n, n)); ctx.synchronize()
    del sigma, color; torch.cuda.empty_cache()
    plan = D.Plan(ctx, S.bench_plan(W, W, steps, stratified=False)); frame = D.Frame(plan)
    frame.forward(grid); out = frame.read()
    want = float(out["opacity"].astype(np.float64).sum())
    dl = np.ones((W * W, 3), np.float32)
    res = dict(n=n, W=W, steps=steps, want=want)
    for name, flag in (("per_ray", D.HPX_BACKWARD_SCATTER_PER_RAY), ("merged", D.HPX_BACKWARD_SCATTER_MERGED)):
        frame.backward(grid, dl, D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | flag)
        ctx.synchronize()   # the library runs on its own stream: torch must not read the block before it is done
        ptr, floats = grid.grad_buffer()
        class V: __cuda_array_interface__ = {"shape": (floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        block = torch.as_tensor(V(), device="cuda")
        tot = 0.0; V4 = n * n * n * 4; chunk = 1 << 28
        nz_hi = 0
        for a in range(0, V4, chunk):
            part = block[a:min(a + chunk, V4)].view(-1, 4)
            tot += float(part[:, 0].sum(dtype=torch.float64))
            if float(part.abs().sum()) > 0: nz_hi = a
        res[name] = tot; res[name + "_last_nonzero_chunk"] = nz_hi // chunk
    print(json.dumps(res), flush=True)
    frame.close(); plan.close(); grid.close()
