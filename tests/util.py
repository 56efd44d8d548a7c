"""Helpers shared by the test modules: golden loading, tolerances, case tables."""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

import hp_abi as A
import oracle as O
import synth as S

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")

# north_star tolerances (SURVEY 8d "parity gates")
IMAGE_RTOL = 1e-5
GRAD_RTOL = 1e-4
FLOOR_FRAC = 1e-3
# Grid gradients are float32 sums of signed terms; the GPU adds them with float atomics in an
# arbitrary order, the reference in sample order.  Two orders of the same sum differ by about
# eps * sum|terms|, so entries far below the largest gradient are compared against a floor of 1e-2
# of the largest magnitude (absolute 1e-6 * max|ref|; the reference's own determinism gate is an
# absolute 1e-6, hp_runner.cpp:2580-2594, and its CPU<->CUDA gradient gate 1e-3 relative).
GRAD_FLOOR_FRAC = 1e-2


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    for key in ("desc_in", "desc_resolved"):
        d = A.hp_plan_desc()
        C.memmove(C.byref(d), g[key].tobytes(), C.sizeof(A.hp_plan_desc))
        g[key] = d
    g["interp"], g["oob"] = int(g["interp"]), int(g["oob"])
    return g


def bits_equal(a, b) -> bool:
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return a.tobytes() == b.tobytes()


def assert_bits(a, b, what):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.tobytes() != b.tobytes():
        diff = np.flatnonzero(a.reshape(-1) != b.reshape(-1))
        raise AssertionError(f"{what}: {diff.size} of {a.size} elements differ, first at {diff[:5]}: "
                             f"{a.reshape(-1)[diff[:5]]} vs {b.reshape(-1)[diff[:5]]}")


def assert_close(got, ref, rtol, what, floor_frac=None):
    """|got - ref| <= rtol * max(|ref|, floor_frac * max|ref|)  (SURVEY 8d)."""
    if floor_frac is None:
        floor_frac = GRAD_FLOOR_FRAC if rtol >= GRAD_RTOL else FLOOR_FRAC
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    scale = np.maximum(np.abs(ref), floor_frac * (np.abs(ref).max() if ref.size else 0.0))
    err = np.abs(got - ref)
    bad = err > rtol * scale + 1e-30
    if bad.any():
        i = np.argmax(err / (scale + 1e-300))
        raise AssertionError(f"{what}: {bad.sum()} of {ref.size} outside rtol={rtol}; worst rel "
                             f"{(err / (scale + 1e-300)).reshape(-1)[i]:.3e} (got {got.reshape(-1)[i]!r}, ref {ref.reshape(-1)[i]!r})")


def oracle_grids(sigma, color, interp, oob):
    return O.make_grid(sigma, 1, interp, oob), O.make_grid(color, 3, interp, oob)


def random_cases(count=10, seed=0, max_dim=40):
    """Small random configurations covering fixed/stratified, linear/nearest, zero/clamp,
    ROI, orthographic cameras, dense (early stop) and thin volumes."""
    rng = np.random.default_rng(seed)
    for case in range(count):
        W, Hh = int(rng.integers(3, max_dim)), int(rng.integers(3, max_dim))
        n = tuple(int(v) for v in rng.integers(1, 20, 3))
        strat = case % 2
        interp = A.HP_INTERP_NEAREST if case % 3 == 0 else A.HP_INTERP_LINEAR
        oob = A.HP_OOB_CLAMP if case % 4 == 1 else A.HP_OOB_ZERO
        sig = (rng.random((n[2], n[1], n[0]), dtype=np.float32) * (30 if case % 5 == 0 else 3)).astype(np.float32)
        col = rng.random((n[2], n[1], n[0], 3), dtype=np.float32)
        steps = int(rng.integers(5, 90))
        K = [1.1 * W, 0, W / 2, 0, 1.3 * W, Hh / 2, 0, 0, 1]
        roi = (1, 1, W - 2, Hh - 2) if case % 3 == 1 else None
        desc = A.make_plan_desc(W, Hh, 0.3 + 0.1 * case, 3.0, dt=float(np.float32(2.7 / steps)), max_steps=steps,
                                mode=strat, K=K, c2w=S.orbit_c2w(case, 12), roi=roi, seed=1234 + case,
                                model=A.HP_CAMERA_ORTHOGRAPHIC if case == 7 else A.HP_CAMERA_PINHOLE)
        bbox = ((0, 0, 0), (1, 1, 1)) if case % 2 == 0 else ((-0.1, 0.05, 0.0), (1.2, 0.9, 1.0))
        yield dict(case=case, desc=desc, sigma=sig, color=col, interp=interp, oob=oob, res=n, bmin=bbox[0],
                   bmax=bbox[1])
