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
# Grid gradients: the contract gate of SURVEY 8(d), |d| <= 1e-4 * max(|ref|, 1e-3 * max|ref|), against the float32
# reference.  A grid gradient is a float32 SUM of signed terms: the reference adds them in sample order, the GPU with
# float reds in arrival order, and two orders of one sum differ by rounding that scales with sum|terms|, not with the
# result.  Entries outside the contract gate are therefore not waved through by a wider floor but ADJUDICATED in the
# test (assert_grad_close) against the float64 sum of the reference's own float32 terms (oracle shadow,
# orc_render_shadowed): the entry passes only if it is inside the same contract gate of that exact sum, or within
# SUM_ORDER_ULPS float32 roundings of sum|terms| of it -- what any float32 summation order of those terms, the
# reference's included, can be off by.
GRAD_FLOOR_FRAC = 1e-3
SUM_ORDER_ALLOWANCE = 2.0 ** -20    # x sum|terms|: 16 float32 unit roundoffs
ADJUDICATED_MAX_FRAC = 0.02         # at most this share of the non-zero entries may need adjudication


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


def _grad_gate(got, ref, rtol, floor_frac):
    scale = np.maximum(np.abs(ref), floor_frac * (np.abs(ref).max() if ref.size else 0.0))
    err = np.abs(got - ref)
    return err, scale, err <= rtol * scale + 1e-30


def assert_grad_close(got, ref, what, key, rtol=GRAD_RTOL, box=None, res=None, scale_by=1.0):
    """Contract gate of one gradient grid (`key` = "sigma" or "color") against an oracle render made with
    shadow=True; entries outside it are adjudicated against the float64 shadow (see GRAD_FLOOR_FRAC above).

    got: flat array in the reference layout ([V] or [3V]).  ref: the oracle's result dict.  scale_by: factor applied
    to the oracle's values (accumulation / linearity tests).  When the shadow covers a box of the grid
    (ref["shadow_box"]) the comparison is made inside the box and everything outside must be exactly zero on both."""
    ch = 1 if key == "sigma" else 3
    got = np.asarray(got).reshape(-1)
    ref32 = np.asarray(ref[f"{key}_grad"]).reshape(-1)
    assert got.shape == ref32.shape, f"{what}: shape {got.shape} vs {ref32.shape}"
    have_shadow = f"{key}_sum" in ref
    if have_shadow and ref.get("shadow_box") is not None and res is not None:
        x0, y0, z0, bx, by, bz = ref["shadow_box"]
        nx, ny, nz = res
        if (bx, by, bz) != (nx, ny, nz):
            assert ref["shadow_misses"] == 0, f"{what}: {ref['shadow_misses']} oracle contributions outside the shadow box"
            g_in = got.reshape(nz, ny, nx, ch)[z0:z0 + bz, y0:y0 + by, x0:x0 + bx]
            r_in = ref32.reshape(nz, ny, nx, ch)[z0:z0 + bz, y0:y0 + by, x0:x0 + bx]
            # nothing outside the band's voxel box, on either side (counted, not copied: the grids can be GBs)
            assert np.count_nonzero(got) == np.count_nonzero(g_in), f"{what}: GPU gradient outside the voxel box of the band"
            assert np.count_nonzero(ref32) == np.count_nonzero(r_in), f"{what}: oracle gradient outside the voxel box of the band"
            got, ref32 = np.ascontiguousarray(g_in).reshape(-1), np.ascontiguousarray(r_in).reshape(-1)
    got = got.astype(np.float64)
    ref32 = ref32.astype(np.float64) * scale_by
    err, scale, ok = _grad_gate(got, ref32, rtol, GRAD_FLOOR_FRAC)
    if ok.all():
        return 0
    if not have_shadow:
        i = np.argmax(err / (scale + 1e-300))
        raise AssertionError(f"{what}: {(~ok).sum()} of {ref32.size} outside rtol={rtol} (floor {GRAD_FLOOR_FRAC}); worst rel "
                             f"{(err / (scale + 1e-300))[i]:.3e} (got {got[i]!r}, ref {ref32[i]!r}); no float64 shadow to adjudicate")
    ref64 = np.asarray(ref[f"{key}_sum"], np.float64) * scale_by
    mag = np.asarray(ref[f"{key}_abs"], np.float64) * abs(scale_by)
    assert ref64.shape == got.shape, f"{what}: shadow shape {ref64.shape} vs {got.shape}"
    err64, scale64, ok64 = _grad_gate(got, ref64, rtol, GRAD_FLOOR_FRAC)
    ok_order = err64 <= SUM_ORDER_ALLOWANCE * mag
    bad = ~ok & ~(ok64 | ok_order)
    n_adj = int((~ok).sum())
    nonzero = max(1, int((ref32 != 0).sum()))
    if bad.any():
        i = np.flatnonzero(bad)[np.argmax((err64 / (scale64 + 1e-300))[bad])]
        raise AssertionError(
            f"{what}: {bad.sum()} of {ref32.size} entries fail the contract gate AND the float64 adjudication; worst: got {got[i]!r}, "
            f"float32 ref {ref32[i]!r}, float64 sum {ref64[i]!r}, sum|terms| {mag[i]!r}, err/(1e-4 scale) {(err64[i] / (rtol * scale64[i])):.2f}, "
            f"err/sum|terms| {(err64[i] / max(mag[i], 1e-300)):.2e}")
    assert n_adj <= ADJUDICATED_MAX_FRAC * nonzero + 8, \
        f"{what}: {n_adj} of {nonzero} non-zero entries needed the float64 adjudication -- too many for rounding noise"
    return n_adj


def assert_grads(got_sigma, got_color, ref, what, res=None, scale_by=1.0):
    a = assert_grad_close(got_sigma, ref, what + " sigma_grad", "sigma", res=res, scale_by=scale_by)
    b = assert_grad_close(got_color, ref, what + " color_grad", "color", res=res, scale_by=scale_by)
    return a + b


def assert_camera_close(cam, ref16, mag16, what, rtol=GRAD_RTOL):
    """Camera gradient d/d c2w[12] | d/d {fx,fy,cx,cy} against the oracle's analytic adjoint (float64): the contract
    gate per block (floor 1e-3 of the block's largest entry), entries outside it adjudicated against the sum of the
    magnitudes of their terms (orc_camera_grad_mag) like the grid gradients."""
    cam = np.asarray(cam, np.float64)
    for name, sl in (("c2w", slice(0, 12)), ("intrinsics", slice(12, 16))):
        got, ref, mag = cam[sl], np.asarray(ref16, np.float64)[sl], np.asarray(mag16, np.float64)[sl]
        err = np.abs(got - ref)
        ok = err <= rtol * np.maximum(np.abs(ref), GRAD_FLOOR_FRAC * np.abs(ref).max()) + 1e-30
        ok |= err <= SUM_ORDER_ALLOWANCE * mag
        assert ok.all(), f"{what} {name}: got {got}, ref {ref}, err/|ref| {err / (np.abs(ref) + 1e-300)}, err/mag {err / (mag + 1e-300)}"


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
