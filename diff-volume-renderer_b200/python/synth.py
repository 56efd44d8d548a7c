"""Synthetic volumes, cameras and plans shared by the tests and bench.py.

Everything is generated from one integer hash so that every harness (numpy
here, CUDA in csrc/synth.cu) produces identical bytes:
u(i, s) = top 24 bits of mix64((s ^ i) + GOLDEN) / 2^24.
"""
from __future__ import annotations

import math

import numpy as np

import hp_abi as A

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)


def mix64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * M1
        x = (x ^ (x >> np.uint64(27))) * M2
    return x ^ (x >> np.uint64(31))


def hash_unit(index: np.ndarray, seed: int) -> np.ndarray:
    """u(i, seed) in [0,1) as float32, exact multiples of 2^-24."""
    with np.errstate(over="ignore"):
        z = mix64((np.asarray(index, dtype=np.uint64) ^ np.uint64(seed)) + GOLDEN)
    return ((z >> np.uint64(40)).astype(np.float32)) * np.float32(1.0 / 16777216.0)


def hashed_volume(n, kind: str = "thin", seed: int = 1234):
    """(sigma[nz,ny,nx], color[nz,ny,nx,3]) hashed grids; kind thin (2u) or dense (40u)."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    idx = np.arange(nx * ny * nz, dtype=np.uint64)
    scale = {"thin": 2.0, "dense": 40.0}[kind]
    sigma = (hash_unit(idx, seed) * np.float32(scale)).reshape(nz, ny, nx)
    color = np.stack([hash_unit(idx, seed + 1 + c) for c in range(3)], axis=-1).reshape(nz, ny, nx, 3)
    return np.ascontiguousarray(sigma), np.ascontiguousarray(color)


def smooth_volume(n, blobs: int = 4, seed: int = 7, peak: float = 6.0):
    """Sum of Gaussian blobs (smooth, so finite differences of the camera are meaningful)."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.linspace(0, 1, nz), np.linspace(0, 1, ny), np.linspace(0, 1, nx), indexing="ij")
    sigma = np.zeros((nz, ny, nx), np.float64)
    color = np.zeros((nz, ny, nx, 3), np.float64)
    for _ in range(blobs):
        c = rng.uniform(0.25, 0.75, 3)
        w = rng.uniform(0.12, 0.25)
        g = np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (2 * w * w))
        sigma += peak * rng.uniform(0.5, 1.0) * g
        color += g[..., None] * rng.uniform(0.2, 1.0, 3)
    color = color / max(color.max(), 1e-6)
    # density must vanish on the cube faces: with the OOB-zero policy a non-zero face value makes
    # the rendered loss discontinuous in the camera, and finite differences meaningless
    window = (np.sin(np.pi * x) * np.sin(np.pi * y) * np.sin(np.pi * z)) ** 2
    sigma = sigma * window
    return sigma.astype(np.float32), np.ascontiguousarray(color.astype(np.float32))


def orbit_c2w(view: int = 0, views: int = 1, radius: float = 1.5, centre=(0.5, 0.5, 0.5)) -> np.ndarray:
    """3x4 camera-to-world looking at the cube centre; view 0 sits at (0.5,0.5,-1) with R = I."""
    ang = 2.0 * math.pi * view / max(views, 1)
    c, s = math.cos(ang), math.sin(ang)
    # rotate the canonical camera (origin centre-(0,0,radius), looking +z) about y
    R = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=np.float64)
    origin = np.asarray(centre) + R @ np.array([0.0, 0.0, -radius])
    m = np.zeros((3, 4), np.float32)
    m[:, :3] = R.astype(np.float32)
    m[:, 3] = origin.astype(np.float32)
    return m


def bench_plan(width: int, height: int, steps: int, stratified: bool, view: int = 0, views: int = 1,
               roi=None, seed: int = 42, max_samples: int = 0) -> A.hp_plan_desc:
    """The synthetic camera/plan of SURVEY 8(d): exactly `steps` samples on every ray."""
    K = [1.2 * width, 0, width / 2.0, 0, 1.2 * width, height / 2.0, 0, 0, 1]
    return A.make_plan_desc(width, height, 0.9, 4.0, dt=1.5 / steps, max_steps=steps,
                            mode=A.HP_SAMPLING_STRATIFIED if stratified else A.HP_SAMPLING_FIXED,
                            K=K, c2w=orbit_c2w(view, views), roi=roi, seed=seed, max_samples=max_samples)


def hashed_image_grad(n_rays: int, seed: int = 777) -> np.ndarray:
    """dL/dI (n_rays,3) = u(ray*3+c, seed) - 0.5."""
    idx = np.arange(n_rays * 3, dtype=np.uint64)
    return (hash_unit(idx, seed) - np.float32(0.5)).reshape(n_rays, 3)
