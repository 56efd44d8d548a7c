"""Golden frames for the command-line renderer (tests/golden/cli/*.json -> *.ppm, *.txt).

Run in the authoring container: needs oracle/_ref/bin/dvren_render_ref, the UNMODIFIED reference tool
(apps/dvren_render/main.cpp) compiled by `make -C oracle ref`.  Each JSON scene below is rendered by the
reference on its CPU path; the PPM bytes and the "rays=.. samples=.." counts are committed and
tests/test_cli.py requires the B200 tool to reproduce them byte for byte.

    python tests/golden/make_cli_golden.py
"""
import json
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "cli")
REF_TOOL = os.path.join(REPO, "oracle", "_ref", "bin", "dvren_render_ref")


def scenes():
    # 1. the values of the reference's example scene (examples/simple_volume.json; README.md:94: rays=16 samples=160)
    yield "example_2x2x2", {
        "render": {"width": 4, "height": 4, "t_near": 0.0, "t_far": 1.0, "dt": 0.1, "max_steps": 16,
                   "sampling_mode": "fixed", "seed": 0,
                   "options": {"use_fused_path": True, "enable_graph": False, "capture_stats": True}},
        "volume": {"size": [2, 2, 2], "density": [0.1, 0.2, 0.3, 0.4, 0.4, 0.3, 0.2, 0.1],
                   "color": [1.0, 0.5, 0.5, 0.5, 1.0, 0.5, 0.5, 0.5, 1.0, 1.0, 1.0, 0.5,
                             0.5, 1.0, 1.0, 1.0, 0.5, 1.0, 0.8, 0.8, 0.8, 1.0, 1.0, 1.0],
                   "bbox_min": [0.0, 0.0, 0.0], "bbox_max": [1.0, 1.0, 1.0]},
        "output": {"path": "simple.ppm"}}
    rng = np.random.default_rng(41)
    # 2. stratified marching, explicit pinhole camera looking into the cube, ROI, staged (non-fused) path
    nx, ny, nz = 6, 5, 4
    W, H = 24, 16
    yield "stratified_camera_roi", {
        "render": {"width": W, "height": H, "t_near": 0.9, "t_far": 4.0, "dt": 0.05, "max_steps": 40,
                   "sampling_mode": "stratified", "seed": 1234,
                   "roi": {"x": 2, "y": 1, "width": 19, "height": 13},
                   "camera": {"model": "pinhole", "K": [1.2 * W, 0, W / 2, 0, 1.2 * W, H / 2, 0, 0, 1],
                              "c2w": [1, 0, 0, 0.5, 0, 1, 0, 0.5, 0, 0, 1, -1.0]},
                   "options": {"use_fused_path": False}},
        "volume": {"size": [nx, ny, nz], "density": [round(float(v), 4) for v in rng.random(nx * ny * nz) * 6.0],
                   "color": [round(float(v), 4) for v in rng.random(nx * ny * nz * 3)]}}
    # 3. grey volume (no "color"), nearest + clamp, dense enough for early termination
    nx, ny, nz = 3, 4, 5
    W, H = 17, 11
    yield "grey_nearest_clamp", {
        "render": {"width": W, "height": H, "t_far": 3.0, "dt": 0.04, "max_steps": 64, "t_near": 0.5,
                   "camera": {"K": [20.0, 0, W / 2, 0, 21.0, H / 2, 0, 0, 1],
                              "c2w": [0.8, 0, 0.6, 0.1, 0, 1, 0, 0.45, -0.6, 0, 0.8, -0.7]}},
        "volume": {"size": [nx, ny, nz], "density": [round(float(v), 4) for v in rng.random(nx * ny * nz) * 1.5],
                   "interp": "nearest", "oob": "clamp"}}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, scene in scenes():
        cfg = os.path.join(OUT, name + ".json")
        with open(cfg, "w") as f:
            json.dump(scene, f, separators=(",", ":"))
        ppm = os.path.join(OUT, name + ".ppm")
        res = subprocess.run([REF_TOOL, cfg, ppm], check=True, capture_output=True, text=True)
        m = re.search(r"rays=(\d+) samples=(\d+)", res.stdout)
        with open(os.path.join(OUT, name + ".txt"), "w") as f:
            f.write(f"rays={m.group(1)} samples={m.group(2)}\n")
        print(name, m.group(0), os.path.getsize(ppm), "bytes")


if __name__ == "__main__":
    main()
