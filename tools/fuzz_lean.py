"""One-off differential sweep: many random small configurations, lean forward + both backward kernels (+ deterministic,
+ fused camera) against the oracle.  Prints failures; exit code 1 if any."""
import os, sys, traceback
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("diff-volume-renderer_b200/python", "oracle", "tests", "tools"):
    sys.path.insert(0, os.path.join(REPO, p))
import numpy as np
import dvren_b200 as D, hp_abi as A, oracle as O, synth as S, util as U

O.build_oracle()
ctx = D.Context(device=0)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
fails = 0
from fuzz_lean_cases import cases


ran = 0
adjudicated = 0
for case in cases(n_cases, seed0, 70):
    desc = case["desc"]
    try:
        st, odesc = O.plan_resolve(desc)
        if st != 0:
            continue
        ran += 1
        gs, gc = U.oracle_grids(case["sigma"], case["color"], case["interp"], case["oob"])
        n = odesc.roi.width * odesc.roi.height
        dl = S.hashed_image_grad(n)
        ref = O.render(odesc, gs, gc, dl, case["res"], case["bmin"], case["bmax"], shadow=True)
        plan = D.Plan(ctx, desc)
        grid = D.Grid(ctx, case["sigma"], case["color"], case["interp"], case["oob"], case["bmin"], case["bmax"])
        frame = D.Frame(plan)
        frame.forward(grid)
        got = frame.read(); cnt = frame.counts()
        assert cnt["samples"] == ref["sample_count"] and cnt["live_samples"] == ref["live_sample_count"], "counts"
        U.assert_bits(got["hitmask"], ref["hitmask"], "hitmask")
        for k in ("image", "trans", "opacity", "depth"):
            U.assert_close(got[k], ref[k], U.IMAGE_RTOL, k)
        for name, extra in (("per_ray", D.HPX_BACKWARD_SCATTER_PER_RAY), ("merged", D.HPX_BACKWARD_SCATTER_MERGED),
                            ("merged+det", D.HPX_BACKWARD_SCATTER_MERGED | D.HPX_BACKWARD_DETERMINISTIC),
                            ("merged+cam", D.HPX_BACKWARD_SCATTER_MERGED | D.HPX_BACKWARD_CAMERA)):
            frame.backward(grid, dl, D.HPX_BACKWARD_GRID | D.HPX_BACKWARD_ZERO | extra)
            sg, cg, cam = grid.read_grad()
            adjudicated += U.assert_grads(sg, cg, ref, name)   # contract gate + float64 adjudication (tests/util.py)
            assert np.isfinite(cam).all()
        frame.close(); grid.close(); plan.close()
    except Exception as e:
        fails += 1
        print("FAIL case", case["case"], "interp", case["interp"], "oob", case["oob"], "res", case["res"], "bbox", case["bmin"], case["bmax"],
              "wh", desc.width, desc.height, "roi", (desc.roi.x, desc.roi.y, desc.roi.width, desc.roi.height), "mode", desc.sampling.mode,
              "model", desc.camera.model, "->", repr(e)[:300], flush=True)
print(f"{n_cases} cases generated, {ran} valid plans run, {fails} failures, {adjudicated} gradient entries adjudicated against the float64 shadow", flush=True)
sys.exit(1 if fails else 0)
