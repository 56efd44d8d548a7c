"""GPU parity of the hp.h entry points (the drop-in C ABI) against the pinned oracle and the
reference's golden vectors, through HOST tensors (staged to the GPU by the library) and DEVICE
tensors (used in place).  Indices, counts, positions and field lookups are bit-exact; integrals and
gradients are held to the north_star tolerances (alpha itself is bit-exact, see
test_alpha_matches_oracle_bitwise; aux[:,3] uses CUDA's logf, radiance sums are contracted in the
lean kernels)."""
import ctypes as C

import numpy as np
import pytest

import dvren_b200 as D
import hp_abi as A
import hp_host as H
import oracle as O
import synth as S
import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pipe():
    p = H.HpHostPipeline(D.load())
    yield p
    p.close()


def staged_vs_oracle(pipe, desc, sigma, color, interp, oob, res, bmin, bmax, what):
    plan, rdesc = pipe.plan(desc)
    st, odesc = O.plan_resolve(desc)
    assert st == 0 and bytes(rdesc) == bytes(odesc)
    n = rdesc.roi.width * rdesc.roi.height
    rays, orays = pipe.ray(plan, n), O.rays(odesc)
    for k in rays:
        U.assert_bits(rays[k], orays[k], f"{what} ray.{k}")
    fs, fc = pipe.sigma_field(sigma, interp, oob), pipe.color_field(color, interp, oob)
    gs, gc = U.oracle_grids(sigma, color, interp, oob)
    samp = pipe.samp(plan, fs, fc, rays, rdesc.max_samples)
    st, osamp = O.sample(odesc, gs, gc, orays, odesc.max_samples)
    assert st == 0 and samp["count"] == osamp["count"]
    for k in ("positions", "dt", "ray_offset", "sigma", "color"):
        U.assert_bits(samp[k], osamp[k], f"{what} samp.{k}")
    intl, ointl = pipe.integrate(plan, samp), O.integrate(odesc, osamp)
    for k in ("radiance", "transmittance", "opacity", "depth"):
        U.assert_close(intl[k], ointl[k], U.IMAGE_RTOL, f"{what} intl.{k}")
    U.assert_close(intl["aux"], ointl["aux"], U.IMAGE_RTOL, f"{what} aux")
    assert np.array_equal(intl["aux"] == 0, ointl["aux"] == 0), "early-stop rows must stay zero"
    img = pipe.img(plan, rdesc, intl, rays)
    st, oimg = O.image(odesc, orays, ointl)
    U.assert_bits(img["hitmask"], oimg["hitmask"], f"{what} hitmask")
    for k in ("image", "trans", "opacity", "depth"):
        U.assert_close(img[k], oimg[k], U.IMAGE_RTOL, f"{what} img.{k}")
    dl = S.hashed_image_grad(n)
    grads, ograds = pipe.diff(plan, dl, samp, intl), O.diff(dl, osamp, ointl)
    U.assert_close(grads["sigma"], ograds["sigma"], U.GRAD_RTOL, f"{what} diff.sigma")
    U.assert_close(grads["color"], ograds["color"], U.GRAD_RTOL, f"{what} diff.color")
    assert not grads["camera"].any()          # the ABI's (3,4) camera slot stays zero (diff_cpu.cpp:73-74)
    # fused == staged, exactly (reference hp_runner.cpp:1737-1760)
    fsamp, fintl = pipe.fused(plan, fs, fc, rays, rdesc.max_samples)
    for k in ("positions", "dt", "ray_offset", "sigma", "color"):
        U.assert_bits(fsamp[k], samp[k], f"{what} fused samp.{k}")
    for k in ("radiance", "transmittance", "opacity", "depth", "aux"):
        U.assert_bits(fintl[k], intl[k], f"{what} fused intl.{k}")
    return dict(samp=samp, intl=intl, img=img, grads=grads)


def test_alpha_matches_oracle_bitwise(pipe):
    """alpha_of (csrc/dv_device.cuh) against the reference's libm form (int_cpu.cpp:98-109) on 2 M
    optical depths: one single-sample ray each through hp_int, so aux = {alpha, alpha, 1, 0} and
    transmittance = max(1 - alpha, 0), all compared bit for bit."""
    rng = np.random.default_rng(5)
    n = 1 << 21
    od = np.exp(rng.uniform(np.log(2e-5), np.log(40.0), n)).astype(np.float32)
    od[:8] = [0.0, 9.9e-5, 1e-4, 17.4999, 17.5, 17.5001, 30.0, 1e-3]
    desc = A.make_plan_desc(n, 1, 0.0, 10.0, dt=1.0, max_steps=1, max_samples=n)
    plan, rdesc = pipe.plan(desc)
    st, odesc = O.plan_resolve(desc)
    samp = {"dt": np.ones(n, np.float32), "sigma": od, "color": np.full((n, 3), 0.5, np.float32),
            "positions": np.zeros((n, 3), np.float32), "ray_offset": np.arange(n + 1, dtype=np.uint32), "count": n}
    got, ref = pipe.integrate(plan, samp), O.integrate(odesc, samp)
    U.assert_bits(got["aux"][:, :3], ref["aux"][:, :3], "alpha / weight / T_prev")
    U.assert_bits(got["transmittance"], ref["transmittance"], "transmittance")
    U.assert_bits(got["opacity"], ref["opacity"], "opacity")


@pytest.mark.parametrize("case", list(U.random_cases(12, seed=0)), ids=lambda c: f"case{c['case']}")
def test_staged_abi_matches_oracle_random(pipe, case):
    staged_vs_oracle(pipe, case["desc"], case["sigma"], case["color"], case["interp"], case["oob"], case["res"],
                     case["bmin"], case["bmax"], f"case{case['case']}")


@pytest.mark.parametrize("path", U.golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_staged_abi_matches_reference_golden(pipe, path):
    g = U.load_golden(path)
    nz, ny, nx = g["sigma"].shape
    out = staged_vs_oracle(pipe, g["desc_in"], g["sigma"], g["color"], g["interp"], g["oob"], (nx, ny, nz), g["bmin"],
                           g["bmax"], "golden")
    U.assert_bits(out["samp"]["positions"], g["samp_positions"], "golden positions")
    U.assert_bits(out["samp"]["ray_offset"], g["samp_ray_offset"], "golden offsets")
    U.assert_bits(out["samp"]["sigma"], g["samp_sigma"], "golden sigma")
    U.assert_close(out["img"]["image"], g["img_image"], U.IMAGE_RTOL, "golden image")
    U.assert_close(out["grads"]["sigma"], g["diff_sigma"], U.GRAD_RTOL, "golden diff.sigma")


def test_sigma_only_and_color_only_fields(pipe):
    """hp_samp accepts a null sigma or colour field, not both (reference samp_cpu.cpp:161-163,255-289)."""
    c = next(iter(U.random_cases(2, seed=4)))
    plan, rdesc = pipe.plan(c["desc"])
    st, odesc = O.plan_resolve(c["desc"])
    rays, orays = pipe.ray(plan, rdesc.roi.width * rdesc.roi.height), O.rays(odesc)
    fs = pipe.sigma_field(c["sigma"], c["interp"], c["oob"])
    fc = pipe.color_field(c["color"], c["interp"], c["oob"])
    gs, gc = U.oracle_grids(c["sigma"], c["color"], c["interp"], c["oob"])
    a = pipe.samp(plan, fs, None, rays, rdesc.max_samples)
    _, oa = O.sample(odesc, gs, None, orays, odesc.max_samples)
    U.assert_bits(a["sigma"], oa["sigma"], "sigma only"); assert not a["color"].any()
    b = pipe.samp(plan, None, fc, rays, rdesc.max_samples)
    _, ob = O.sample(odesc, None, gc, orays, odesc.max_samples)
    U.assert_bits(b["color"], ob["color"], "color only"); assert not b["sigma"].any()
    with pytest.raises(H.HpError) as e:
        pipe.samp(plan, None, None, rays, rdesc.max_samples)
    assert e.value.status == A.HP_STATUS_INVALID_ARGUMENT


def test_mismatched_field_resolutions(pipe):
    """sigma and colour grids of different shape / policy take the unpacked lookup path."""
    rng = np.random.default_rng(11)
    sig = (rng.random((5, 7, 9), dtype=np.float32) * 4).astype(np.float32)
    col = rng.random((11, 3, 6, 3), dtype=np.float32)
    desc = S.bench_plan(20, 16, 48, stratified=True, view=3, views=7)
    plan, rdesc = pipe.plan(desc)
    st, odesc = O.plan_resolve(desc)
    rays, orays = pipe.ray(plan, 320), O.rays(odesc)
    fs = pipe.sigma_field(sig, A.HP_INTERP_LINEAR, A.HP_OOB_CLAMP)
    fc = pipe.color_field(col, A.HP_INTERP_NEAREST, A.HP_OOB_ZERO)
    gs = O.make_grid(sig, 1, A.HP_INTERP_LINEAR, A.HP_OOB_CLAMP)
    gc = O.make_grid(col, 3, A.HP_INTERP_NEAREST, A.HP_OOB_ZERO)
    a = pipe.samp(plan, fs, fc, rays, rdesc.max_samples)
    _, oa = O.sample(odesc, gs, gc, orays, odesc.max_samples)
    U.assert_bits(a["sigma"], oa["sigma"], "sigma"); U.assert_bits(a["color"], oa["color"], "color")


def test_capacity_overflow_and_ragged_rays(pipe):
    # capacity: 19 samples per ray against 16 per ray of capacity -> INVALID_ARGUMENT (samp_cpu.cpp:245-247)
    desc = A.make_plan_desc(8, 8, 0.1, 2.0, dt=0.1, max_steps=32, max_samples=8 * 8 * 16)
    plan, rdesc = pipe.plan(desc)
    sig, col = S.hashed_volume(4)
    fs, fc = pipe.sigma_field(sig), pipe.color_field(col)
    rays = pipe.ray(plan, 64)
    with pytest.raises(H.HpError) as e:
        pipe.samp(plan, fs, fc, rays, rdesc.max_samples)
    assert e.value.status == A.HP_STATUS_INVALID_ARGUMENT
    # ragged: override rays with per-ray t ranges, including empty and inverted ones
    desc = A.make_plan_desc(4, 3, 0.0, 2.0, dt=0.13, max_steps=40, mode=A.HP_SAMPLING_STRATIFIED, seed=5)
    plan, rdesc = pipe.plan(desc)
    st, odesc = O.plan_resolve(desc)
    base = O.rays(odesc)
    rng = np.random.default_rng(2)
    ov = {k: v.copy() for k, v in base.items()}
    ov["t_near"] = rng.uniform(0.0, 1.0, 12).astype(np.float32)
    ov["t_far"] = (ov["t_near"] + rng.uniform(-0.2, 1.5, 12)).astype(np.float32)
    ov["t_far"][3] = ov["t_near"][3]            # empty
    ov["origins"] += rng.uniform(-0.2, 0.2, (12, 3)).astype(np.float32)
    ov["pixel_ids"] = np.array([0, 1, 2, 3, 3, 3, 6, 7, 8, 9, 0, 11], np.uint32)   # repeated pixels
    rays = pipe.ray(plan, 12, override=ov)
    for k in ov:
        U.assert_bits(rays[k], ov[k], "override " + k)
    gs, gc = U.oracle_grids(sig, col, 1, 0)
    a = pipe.samp(plan, fs, fc, rays, rdesc.max_samples)
    _, oa = O.sample(odesc, gs, gc, ov, odesc.max_samples)
    assert a["count"] == oa["count"]
    for k in ("positions", "dt", "ray_offset", "sigma", "color"):
        U.assert_bits(a[k], oa[k], "ragged " + k)
    intl, ointl = pipe.integrate(plan, a), O.integrate(odesc, oa)
    # repeated pixel ids: first hit writes, later hits add / multiply / min in ray order (img_cpu.cpp:162-177)
    img = pipe.img(plan, rdesc, intl, rays)
    _, oimg = O.image(odesc, ov, ointl)
    U.assert_bits(img["hitmask"], oimg["hitmask"], "dup hitmask")
    for k in ("image", "trans", "opacity", "depth"):
        U.assert_close(img[k], oimg[k], U.IMAGE_RTOL, "dup " + k)
    # empty input: zero samples is legal
    empty = {"positions": np.zeros((0, 3), np.float32), "dt": np.zeros(0, np.float32), "sigma": np.zeros(0, np.float32),
             "color": np.zeros((0, 3), np.float32), "ray_offset": np.zeros(13, np.uint32), "count": 0}
    e = pipe.integrate(plan, empty)
    assert (e["transmittance"] == 1).all() and (e["depth"] == np.float32(2.0)).all() and not e["radiance"].any()


def test_bad_offsets_and_pixels_are_invalid_argument(pipe):
    desc = A.make_plan_desc(2, 2, 0.0, 1.0, dt=0.25, max_steps=4)
    plan, rdesc = pipe.plan(desc)
    samp = {"positions": np.zeros((8, 3), np.float32), "dt": np.full(8, 0.25, np.float32),
            "sigma": np.ones(8, np.float32), "color": np.ones((8, 3), np.float32),
            "ray_offset": np.array([0, 4, 2, 6, 8], np.uint32), "count": 8}
    with pytest.raises(H.HpError) as e:
        pipe.integrate(plan, samp)
    assert e.value.status == A.HP_STATUS_INVALID_ARGUMENT
    samp["ray_offset"] = np.array([0, 2, 4, 6, 8], np.uint32)
    intl = pipe.integrate(plan, samp)
    rays = pipe.ray(plan, 4)
    rays["pixel_ids"][2] = 99
    with pytest.raises(H.HpError) as e:
        pipe.img(plan, rdesc, intl, rays)
    assert e.value.status == A.HP_STATUS_INVALID_ARGUMENT


def test_workspace_allocation_contract(pipe):
    """Outputs whose .data is NULL come out of the caller's workspace in the reference's order with
    4-byte alignment; too small a workspace is OUT_OF_MEMORY (reference workspace.hpp:6-33)."""
    lib = pipe.lib
    desc = A.make_plan_desc(4, 4, 0.0, 1.0, dt=0.1, max_steps=16)
    plan, rdesc = pipe.plan(desc)
    n = 16
    ws = np.zeros(n * 36 + 8, np.uint8)
    rays = A.hp_rays_t()
    for k in ("origins", "directions", "t_near", "t_far", "pixel_ids"):
        setattr(rays, k, A.empty_tensor(A.HP_MEMSPACE_HOST))
    assert lib.hp_ray(plan, None, C.byref(rays), ws.ctypes.data, n * 36 - 4) == A.HP_STATUS_OUT_OF_MEMORY
    for k in ("origins", "directions", "t_near", "t_far", "pixel_ids"):
        setattr(rays, k, A.empty_tensor(A.HP_MEMSPACE_HOST))
    assert lib.hp_ray(plan, None, C.byref(rays), ws.ctypes.data, ws.nbytes) == 0
    base = ws.ctypes.data
    assert rays.origins.data == base and rays.directions.data == base + n * 12
    assert rays.t_near.data == base + n * 24 and rays.t_far.data == base + n * 28
    assert rays.pixel_ids.data == base + n * 32
    assert A.tensor_shape(rays.origins) == (n, 3) and rays.pixel_ids.dtype == A.HP_DTYPE_U32
    # fused: samples first (capacity sized, ray_offset last), integrator outputs carved after them
    sig, col = S.hashed_volume(4)
    fs, fc = pipe.sigma_field(sig), pipe.color_field(col)
    cap = rdesc.max_samples
    need = cap * 32 + (n + 1) * 4 + n * 24 + cap * 16
    ws2 = np.zeros(need, np.uint8)
    samp, intl = A.hp_samp_t(), A.hp_intl_t()
    st = lib.hp_samp_int_fused(plan, fs, fc, C.byref(rays), C.byref(samp), C.byref(intl), ws2.ctypes.data, need)
    assert st == 0
    b2 = ws2.ctypes.data
    m = int(samp.dt.shape[0])
    assert m == 160 and int(samp.ray_offset.shape[0]) == n + 1
    assert samp.positions.data == b2 and samp.dt.data == b2 + cap * 12 and samp.sigma.data == b2 + cap * 16
    assert samp.color.data == b2 + cap * 20 and samp.ray_offset.data == b2 + cap * 32
    assert intl.radiance.data == b2 + cap * 32 + (n + 1) * 4
    assert A.tensor_shape(intl.aux) == (m, 4)
    assert lib.hp_samp_int_fused(plan, fs, fc, C.byref(rays), C.byref(samp), C.byref(intl), None, 0) \
        == A.HP_STATUS_INVALID_ARGUMENT


def test_device_memspace_pipeline_and_graph(pipe):
    """The same entry points on DEVICE tensors (the reference's only GPU call pattern,
    tests/render/test_smoke_forward.cpp:137-234), plus hp_graph_* as a real captured graph."""
    import torch
    lib = pipe.lib
    c = list(U.random_cases(6, seed=9))[4]
    plan, rdesc = pipe.plan(c["desc"])
    st, odesc = O.plan_resolve(c["desc"])
    n = rdesc.roi.width * rdesc.roi.height
    cap = rdesc.max_samples
    fs = pipe.sigma_field(c["sigma"], c["interp"], c["oob"])
    fc = pipe.color_field(c["color"], c["interp"], c["oob"])
    dev = torch.device("cuda")
    t = {"origins": torch.zeros(n, 3, device=dev), "directions": torch.zeros(n, 3, device=dev),
         "t_near": torch.zeros(n, device=dev), "t_far": torch.zeros(n, device=dev),
         "pixel_ids": torch.zeros(n, dtype=torch.int32, device=dev)}
    rays = A.hp_rays_t()
    for k, v in t.items():
        x = A.hp_tensor(); x.data = v.data_ptr(); x.memspace = A.HP_MEMSPACE_DEVICE
        setattr(rays, k, x)
    assert lib.hp_ray(plan, None, C.byref(rays), None, 0) == 0
    orays = O.rays(odesc)
    U.assert_bits(t["origins"].cpu().numpy(), orays["origins"], "dev origins")
    U.assert_bits(t["directions"].cpu().numpy(), orays["directions"], "dev directions")
    U.assert_bits(t["pixel_ids"].cpu().numpy().view(np.uint32), orays["pixel_ids"], "dev pixel ids")
    ws_bytes = cap * 32 + (n + 1) * 4 + n * 24 + cap * 16
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
    samp, intl = A.hp_samp_t(), A.hp_intl_t()
    assert lib.hp_samp_int_fused(plan, fs, fc, C.byref(rays), C.byref(samp), C.byref(intl), ws.data_ptr(), ws_bytes) == 0
    m = int(samp.dt.shape[0])
    gs, gc = U.oracle_grids(c["sigma"], c["color"], c["interp"], c["oob"])
    _, osamp = O.sample(odesc, gs, gc, orays, odesc.max_samples)
    assert m == osamp["count"] and samp.sigma.memspace == A.HP_MEMSPACE_DEVICE
    raw = ws.cpu().numpy()
    off = samp.sigma.data - ws.data_ptr()
    U.assert_bits(raw[off:off + m * 4].view(np.float32), osamp["sigma"], "dev sigma")
    off = samp.ray_offset.data - ws.data_ptr()
    U.assert_bits(raw[off:off + (n + 1) * 4].view(np.uint32), osamp["ray_offset"], "dev offsets")
    ointl = O.integrate(odesc, osamp)
    off = intl.radiance.data - ws.data_ptr()
    U.assert_close(raw[off:off + n * 12].view(np.float32).reshape(n, 3), ointl["radiance"], U.IMAGE_RTOL, "dev radiance")
    # hp_diff on DEVICE: outputs the caller did not provide are cudaMalloc'ed (diff_cuda.cu:116-167)
    dl = torch.from_numpy(S.hashed_image_grad(n)).to(dev)
    g = A.hp_tensor(); g.data = dl.data_ptr(); g.memspace = A.HP_MEMSPACE_DEVICE; g.dtype = A.HP_DTYPE_F32
    g.rank = 2; g.shape[0], g.shape[1] = n, 3; g.stride[0], g.stride[1] = 3, 1
    grads = A.hp_grads_t()
    gsig = torch.zeros(m, device=dev); gcol = torch.zeros(m, 3, device=dev); gcam = torch.ones(12, device=dev)
    for k, v in (("sigma", gsig), ("color", gcol), ("camera", gcam)):
        x = A.hp_tensor(); x.data = v.data_ptr(); x.memspace = A.HP_MEMSPACE_DEVICE
        setattr(grads, k, x)
    assert lib.hp_diff(plan, C.byref(g), C.byref(samp), C.byref(intl), C.byref(grads), None, 0) == 0
    ograds = O.diff(dl.cpu().numpy(), osamp, ointl)
    U.assert_close(gsig.cpu().numpy(), ograds["sigma"], U.GRAD_RTOL, "dev diff.sigma")
    U.assert_close(gcol.cpu().numpy(), ograds["color"], U.GRAD_RTOL, "dev diff.color")
    assert not gcam.cpu().numpy().any()
    # graph: create -> capture(with dL/dI) -> execute twice -> same numbers as the direct calls
    handle = C.c_void_p()
    assert lib.hp_graph_create(plan, fs, fc, 0, 0, 0, 0, C.byref(handle)) == 0
    assert lib.hp_graph_execute(handle, None, None, None, None, None) == A.HP_STATUS_INVALID_ARGUMENT
    assert lib.hp_graph_capture(handle, plan, fs, fc, C.byref(g)) == 0
    o_r, o_s, o_i, o_m, o_g = A.hp_rays_t(), A.hp_samp_t(), A.hp_intl_t(), A.hp_img_t(), A.hp_grads_t()
    for _ in range(2):
        assert lib.hp_graph_execute(handle, C.byref(o_r), C.byref(o_s), C.byref(o_i), C.byref(o_m), C.byref(o_g)) == 0
    assert int(o_s.dt.shape[0]) == m and o_r.origins.data and o_m.image.data

    _, oimg = O.image(odesc, orays, ointl)

    def fetch(tensor, dtype=np.float32):
        arr = np.zeros(A.tensor_shape(tensor), dtype)
        assert lib.hpx_copy_to_host(pipe.ctx, arr.ctypes.data, tensor.data, arr.nbytes) == 0
        return arr

    U.assert_close(fetch(o_m.image), oimg["image"], U.IMAGE_RTOL, "graph image")
    U.assert_bits(fetch(o_m.hitmask, np.uint32), oimg["hitmask"], "graph hitmask")
    U.assert_close(fetch(o_g.sigma), ograds["sigma"], U.GRAD_RTOL, "graph diff.sigma")
    U.assert_bits(fetch(o_s.ray_offset, np.uint32), osamp["ray_offset"], "graph offsets")
    lib.hp_graph_release(handle)
