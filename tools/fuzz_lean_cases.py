"""Random small configurations shared by tools/fuzz_lean.py and tools/fuzz_staged.py."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("diff-volume-renderer_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(REPO, p))
import numpy as np
import hp_abi as A, synth as S


def cases(count, seed, max_dim):
    """tests/util.random_cases with the case number folded so that every plan is valid (t_near < t_far)."""
    rng = np.random.default_rng(seed)
    for case in range(count):
        W, Hh = int(rng.integers(3, max_dim)), int(rng.integers(3, max_dim))
        n = tuple(int(v) for v in rng.integers(1, 24, 3))
        strat = int(rng.integers(0, 2))
        interp = A.HP_INTERP_NEAREST if rng.random() < 0.2 else A.HP_INTERP_LINEAR
        oob = A.HP_OOB_CLAMP if rng.random() < 0.3 else A.HP_OOB_ZERO
        sig = (rng.random((n[2], n[1], n[0]), dtype=np.float32) * (30 if rng.random() < 0.3 else 3)).astype(np.float32)
        col = rng.random((n[2], n[1], n[0], 3), dtype=np.float32)
        steps = int(rng.integers(5, 140))
        K = [float(rng.uniform(0.6, 2.5)) * W, 0, W / 2 + float(rng.uniform(-2, 2)), 0, float(rng.uniform(0.6, 2.5)) * W,
             Hh / 2 + float(rng.uniform(-2, 2)), 0, 0, 1]
        roi = None
        if rng.random() < 0.3 and W > 6 and Hh > 6:
            rx, ry = int(rng.integers(0, W // 2)), int(rng.integers(0, Hh // 2))
            roi = (rx, ry, int(rng.integers(1, W - rx + 1)), int(rng.integers(1, Hh - ry + 1)))
        t_near = float(rng.uniform(0.0, 1.2))
        desc = A.make_plan_desc(W, Hh, t_near, t_near + float(rng.uniform(0.5, 3.0)), dt=float(np.float32(rng.uniform(1.0, 3.0) / steps)),
                                max_steps=steps, mode=strat, K=K, c2w=S.orbit_c2w(int(rng.integers(0, 16)), 16, radius=float(rng.uniform(0.8, 2.0))),
                                roi=roi, seed=int(rng.integers(0, 1 << 40)),
                                model=A.HP_CAMERA_ORTHOGRAPHIC if rng.random() < 0.05 else A.HP_CAMERA_PINHOLE)
        bbox = ((0, 0, 0), (1, 1, 1)) if rng.random() < 0.6 else ((-0.1, 0.05, 0.0), (1.2, 0.9, 1.0))
        yield dict(case=case, desc=desc, sigma=sig, color=col, interp=interp, oob=oob, res=n, bmin=bbox[0], bmax=bbox[1])


