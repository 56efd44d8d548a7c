"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference compiled into
oracle/_ref/libdvren_ref.so by `make -C oracle ref`):

    python tests/golden/make_golden.py

Every case drives the reference through its own C ABI with HOST tensors
(hp_ray -> hp_samp -> hp_int -> hp_img -> hp_diff, reference
src/render/renderer.cpp:259-415) and through dvren::Renderer Forward/Backward
(grid gradients via DenseGridField::AccumulateSampleGradients).  Inputs and
outputs are stored bit-exactly in one .npz per case; tests/test_oracle_pin.py
replays them against oracle/liboracle.so, and the GPU tests replay them against
the product library.  /root/reference is not needed at test time.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, os.path.join(REPO, "diff-volume-renderer_b200", "python"))
import hp_abi as A  # noqa: E402
import hp_host as H  # noqa: E402
import oracle as O  # noqa: E402
import synth as S  # noqa: E402


def cases():
    rng = np.random.default_rng(20261018)
    # 1. the reference's CLI example (examples/simple_volume.json): rays=16, samples=160 (README.md:94)
    density = np.array([0.1, 0.2, 0.3, 0.4, 0.4, 0.3, 0.2, 0.1], np.float32).reshape(2, 2, 2)
    color = np.array([1.0, 0.5, 0.5, 0.5, 1.0, 0.5, 0.5, 0.5, 1.0, 1.0, 1.0, 0.5, 0.5, 1.0, 1.0, 1.0, 0.5, 1.0, 0.8,
                      0.8, 0.8, 1.0, 1.0, 1.0], np.float32).reshape(2, 2, 2, 3)
    yield dict(name="cli_example", desc=A.make_plan_desc(4, 4, 0.0, 1.0, dt=0.1, max_steps=16), sigma=density,
               color=color, interp=A.HP_INTERP_LINEAR, oob=A.HP_OOB_ZERO, bmin=(0, 0, 0), bmax=(1, 1, 1))
    # 2. stratified marching through a dense volume: early termination, pinhole orbit camera
    sig = (rng.random((6, 5, 7), dtype=np.float32) * 35).astype(np.float32)
    col = rng.random((6, 5, 7, 3), dtype=np.float32)
    K = [1.2 * 12, 0, 6, 0, 1.2 * 12, 5, 0, 0, 1]
    yield dict(name="stratified_dense", sigma=sig, color=col, interp=A.HP_INTERP_LINEAR, oob=A.HP_OOB_ZERO,
               desc=A.make_plan_desc(12, 10, 0.9, 4.0, dt=1.5 / 40, max_steps=40, mode=A.HP_SAMPLING_STRATIFIED, K=K,
                                     c2w=S.orbit_c2w(1, 7), seed=42), bmin=(0, 0, 0), bmax=(1, 1, 1))
    # 3. nearest + clamp, ROI, non-unit scatter bbox
    sig = (rng.random((4, 9, 3), dtype=np.float32) * 3).astype(np.float32)
    col = rng.random((4, 9, 3, 3), dtype=np.float32)
    K = [14.0, 0, 8, 0, 15.0, 7, 0, 0, 1]
    yield dict(name="nearest_clamp_roi", sigma=sig, color=col, interp=A.HP_INTERP_NEAREST, oob=A.HP_OOB_CLAMP,
               desc=A.make_plan_desc(16, 14, 0.5, 3.0, dt=0.07, max_steps=50, K=K, c2w=S.orbit_c2w(3, 11),
                                     roi=(2, 3, 11, 9), seed=7), bmin=(-0.1, 0.05, 0.0), bmax=(1.2, 0.9, 1.0))
    # 4. linear + clamp, orthographic camera, ray end clipping (t_far inside the last step)
    sig = (rng.random((5, 5, 5), dtype=np.float32) * 8).astype(np.float32)
    col = rng.random((5, 5, 5, 3), dtype=np.float32)
    c2w = S.orbit_c2w(0, 1)
    c2w[:, 3] = (0.4, 0.6, -0.2)
    yield dict(name="ortho_clamp_clip", sigma=sig, color=col, interp=A.HP_INTERP_LINEAR, oob=A.HP_OOB_CLAMP,
               desc=A.make_plan_desc(6, 5, 0.1, 1.13, dt=0.1, max_steps=64, mode=A.HP_SAMPLING_STRATIFIED, c2w=c2w,
                                     model=A.HP_CAMERA_ORTHOGRAPHIC, seed=99), bmin=(0, 0, 0), bmax=(1, 1, 1))
    # 5. default-filled plan (zero K / c2w / dt / max_steps), thin hashed volume
    sig, col = S.hashed_volume(8, "thin")
    yield dict(name="plan_defaults", sigma=sig, color=col, interp=A.HP_INTERP_LINEAR, oob=A.HP_OOB_ZERO,
               desc=A.make_plan_desc(9, 7, 0.05, 2.5), bmin=(0, 0, 0), bmax=(1, 1, 1))


def main():
    assert O.ref_available(), "build oracle/_ref first: make -C oracle ref"
    ref = H.HpHostPipeline(O.ref_lib())
    for c in cases():
        desc = c["desc"]
        plan, rdesc = ref.plan(desc)
        n = rdesc.roi.width * rdesc.roi.height
        rays = ref.ray(plan, n)
        fs = ref.sigma_field(c["sigma"], c["interp"], c["oob"])
        fc = ref.color_field(c["color"], c["interp"], c["oob"])
        samp = ref.samp(plan, fs, fc, rays, rdesc.max_samples)
        intl = ref.integrate(plan, samp)
        img = ref.img(plan, rdesc, intl, rays)
        dl = S.hashed_image_grad(n)
        grads = ref.diff(plan, dl, samp, intl)
        nz, ny, nx = c["sigma"].shape
        sg, cg = O.ref_scatter((nx, ny, nz), c["bmin"], c["bmax"], c["interp"], c["oob"], samp["positions"],
                               grads["sigma"], grads["color"])
        rr = O.ref_render(desc, c["sigma"], c["color"], dl, c["interp"], c["oob"], c["bmin"], c["bmax"])
        assert rr["status"] == 0
        assert np.array_equal(rr["sigma_grad"], sg) and np.array_equal(rr["image"], img["image"])
        out = dict(
            desc_in=np.frombuffer(bytes(desc), np.uint8), desc_resolved=np.frombuffer(bytes(rdesc), np.uint8),
            sigma=c["sigma"], color=c["color"], interp=np.uint32(c["interp"]), oob=np.uint32(c["oob"]),
            bmin=np.asarray(c["bmin"], np.float32), bmax=np.asarray(c["bmax"], np.float32), dL_dI=dl,
            sample_count=np.uint64(samp["count"]), sigma_grad=sg, color_grad=cg)
        out.update({f"ray_{k}": v for k, v in rays.items()})
        out.update({f"samp_{k}": v for k, v in samp.items() if k != "count"})
        out.update({f"intl_{k}": v for k, v in intl.items()})
        out.update({f"img_{k}": v for k, v in img.items()})
        out.update({f"diff_{k}": v for k, v in grads.items()})
        path = os.path.join(HERE, c["name"] + ".npz")
        np.savez_compressed(path, **out)
        print(f"{c['name']}: rays={n} samples={samp['count']} -> {os.path.getsize(path)} bytes")
    ref.close()


if __name__ == "__main__":
    main()
