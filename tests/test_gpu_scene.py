"""Scene-level parity of the CUDA path against the oracle: LBVH topology (acceptance criterion 1), closest / any-hit
queries, light and texture sampling, per-pixel radiance at fixed seed (criterion 2), ray counts, film modes."""
import numpy as np
import pytest

from conftest import scene_path
from helpers import bits_equal, pixel_agreement, random_rays
from lumen_b200 import host, integrator
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

# (scene, width, height, max_depth, frames): BASELINE.json configs 1, 2, the all-BSDF scene and the
# directional-light + sky + texture variant, at sizes the oracle finishes in seconds
CASES = [
    ("cornell", 512, 512, 6, 16),
    ("caustics", 320, 180, 12, 8),
    ("materials", 256, 256, 10, 8),
    ("cornell_dir", 256, 256, 6, 4),
]


@pytest.fixture(scope="module")
def loaded(device):
    cache = {}

    def get(name, w, h):
        key = (name, w, h)
        if key not in cache:
            sc = host.Scene(scene_path(name), w, h)
            cache[key] = (sc, po.OracleScene(sc))
        sc, orc = cache[key]
        device.upload_scene(sc.desc)
        device.build_accel()
        return sc, orc

    return get


@pytest.mark.parametrize("name", ["cornell", "caustics", "materials"])
def test_lbvh_topology_bit_exact(device, loaded, name):
    sc, orc = loaded(name, 64, 64)
    g, c = device.lbvh(), orc.lbvh()
    assert g["left"].size == sc.info.n_triangles - 1
    for k in ("morton", "keys", "leaf_prim", "left", "right", "parent"):
        assert (g[k] == c[k]).all(), k
    assert g["aabb"].tobytes() == c["aabb"].tobytes()


@pytest.mark.parametrize("name", ["cornell", "caustics", "materials"])
def test_closest_and_any_hit_queries(device, loaded, name):
    sc, orc = loaded(name, 64, 64)
    rng = np.random.default_rng(21)
    box = orc.lbvh()["aabb"][:6]
    rays = random_rays(rng, box[:3], box[3:], 300000)
    gh, (ch, _) = device.trace_closest(rays), orc.trace_closest(rays)
    for k in ("prim", "t", "b1", "b2"):
        assert bits_equal(gh[k], ch[k]).all(), k
    assert (gh["prim"] != 0xFFFFFFFF).mean() > 0.2
    rays[:, 3] = 0.0
    rays[:, 7] = rng.uniform(0.05, 6.0, rays.shape[0])
    assert (device.trace_any(rays) == orc.trace_any(rays)[0]).all()


def test_axis_aligned_and_degenerate_rays(device, loaded):
    """Rays parallel to the cornell box walls, rays starting on surfaces, zero-length and NaN directions."""
    sc, orc = loaded("cornell", 64, 64)
    rng = np.random.default_rng(22)
    n = 6000
    rays = random_rays(rng, [-3, -1, -3], [3, 5, 3], n)
    rays[:2000, 4:7] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 2000)] * rng.choice([-1.0, 1.0], (2000, 1)).astype(np.float32)
    hits, _ = orc.trace_closest(rays)
    ok = hits["prim"] != 0xFFFFFFFF
    # restart from the hit points (t = 0 self hits must be rejected by tmin)
    rays[2000:4000] = rays[:2000]
    rays[2000:4000, :3] = rays[:2000, :3] + rays[:2000, 4:7] * np.where(ok[:2000], hits["t"][:2000], 0)[:, None]
    rays[4000, 4:7] = 0.0
    rays[4001, 4:7] = np.nan
    gh, (ch, _) = device.trace_closest(rays), orc.trace_closest(rays)
    assert (gh["prim"] == ch["prim"]).all() and bits_equal(gh["t"], ch["t"]).all()
    assert gh["prim"][4000] == 0xFFFFFFFF and gh["prim"][4001] == 0xFFFFFFFF


def test_non_finite_rays_hit_nothing(device, loaded, monkeypatch):
    """Hit definition (oracle/lbvh_cpu.h ray_finite): a NaN / inf component anywhere in origin or direction is a miss, for the
    wide walk, the binary walk and any-hit queries alike -- a partially NaN ray must not depend on the nodes a walk visits."""
    sc, orc = loaded("cornell", 64, 64)
    rng = np.random.default_rng(23)
    n = 4096
    rays = random_rays(rng, [-3, -1, -3], [3, 5, 3], n)
    bad = np.array([np.nan, np.inf, -np.inf], np.float32)
    cols = np.array([0, 1, 2, 4, 5, 6])
    for i in range(0, n, 2):  # every other ray gets one non-finite component; its neighbours stay valid
        rays[i, cols[rng.integers(0, 6)]] = bad[rng.integers(0, 3)]
    gh, (ch, _) = device.trace_closest(rays), orc.trace_closest(rays)
    assert (gh["prim"][0::2] == 0xFFFFFFFF).all() and (ch["prim"][0::2] == 0xFFFFFFFF).all()
    assert (gh["prim"] == ch["prim"]).all() and bits_equal(gh["t"], ch["t"]).all()
    assert (gh["prim"][1::2] != 0xFFFFFFFF).any()
    assert (device.trace_any(rays) == orc.trace_any(rays)[0]).all()
    assert not device.trace_any(rays)[0::2].any()


def test_light_and_texture_sampling(device, loaded):
    rng = np.random.default_rng(23)
    for name in ("cornell", "caustics", "materials", "cornell_dir"):
        sc, orc = loaded(name, 64, 64)
        r4 = rng.uniform(0, 1, (20000, 4)).astype(np.float32)
        p3 = rng.uniform(-3, 3, (20000, 3)).astype(np.float32)
        assert bits_equal(device.kat_sample_light(sc.info.n_lights, r4, p3), orc.sample_light(sc.info.n_lights, r4, p3)).all(), name
        if sc.info.n_textures:
            uv = rng.uniform(-3, 4, (20000, 2)).astype(np.float32)
            uv[:4] = [[0, 0], [1, 1], [0.5, 0.5], [-0.25, 7.75]]
            assert bits_equal(device.kat_texture(0, uv), orc.texture(0, uv)).all()


@pytest.mark.parametrize("name,w,h,depth,frames", CASES)
def test_per_pixel_radiance_parity(device, loaded, name, w, h, depth, frames):
    """Criterion 2: |gpu - cpu| <= 1e-4 * |cpu| per channel on >= 99.9 % of pixels, fp32 film, fixed seed; the
    deterministic-arithmetic design actually delivers bit equality, which is asserted too."""
    sc, orc = loaded(name, w, h)
    pc, ubo = sc.make_pc(depth, True), sc.make_ubo()
    device.init(w, h, 3)  # 3 frames in flight: batches of 3, 3, ... exercise the batch tail
    device.render(pc, ubo, 0, frames)
    gpu, gs = device.download(), device.stats()
    cpu, cs = orc.render(pc, ubo, 0, frames)
    assert (gs.rays_closest, gs.rays_shadow, gs.rays_probe) == (cs.rays_closest, cs.rays_shadow, cs.rays_probe)
    assert gs.nan_samples == cs.nan_pixels
    assert pixel_agreement(gpu, cpu) >= 0.999
    assert bits_equal(gpu, cpu).mean() >= 0.999
    assert (gpu[..., 3] == 1.0).all()
    assert gs.kernel_launches > 0 and gs.frames == frames


def test_progressive_equals_batched(device, loaded):
    """Path::render is called once per frame by Lumen; 1-frame calls must equal one batched call (frame_num drives RNG)."""
    sc, orc = loaded("cornell", 128, 128)
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    device.init(128, 128, 4)
    device.render(pc, ubo, 0, 6)
    batched = device.download()
    device.init(128, 128, 1)
    for f in range(6):
        device.render(pc, ubo, f, 1)
    assert bits_equal(device.download(), batched).all()


def test_direct_lighting_toggle_and_depth_one(device, loaded):
    """Path::gui exposes path length 0..12 and a direct-lighting toggle (Path.cpp:72-77)."""
    sc, orc = loaded("cornell", 96, 96)
    ubo = sc.make_ubo()
    for depth, direct in ((1, True), (2, False), (3, False), (0, True)):
        pc = sc.make_pc(depth if depth > 0 else 1, direct)
        pc.max_depth = depth
        device.init(96, 96, 2)
        device.render(pc, ubo, 0, 2)
        cpu, cs = orc.render(pc, ubo, 0, 2)
        gs = device.stats()
        assert bits_equal(device.download(), cpu).all(), (depth, direct)
        assert gs.rays == cs.rays


def test_sum_film_and_resolve_match_mean(device, loaded):
    """Sharded accumulation (sum + per-pixel sample count, resolved after the reduce) equals the running mean to fp32
    rounding (SURVEY.md 8e parity note: O(1e-7) relative)."""
    sc, orc = loaded("materials", 128, 128)
    pc, ubo = sc.make_pc(8, True), sc.make_ubo()
    device.init(128, 128, 4)
    device.render(pc, ubo, 0, 8)
    mean = device.download()
    shards = []
    for rank in range(2):  # frames {0,2,4,6} and {1,3,5,7}
        device.clear_film()
        device.render(pc, ubo, rank, 4, 2, integrator.FILM_SUM)
        shards.append(device.download())
    device.upload_film(shards[0] + shards[1])
    device.resolve()
    merged = device.download()
    count = shards[0][..., 3] + shards[1][..., 3]
    assert (count <= 8).all() and (count == 8).mean() > 0.99
    full = count == 8  # a NaN sample is skipped by both rules but leaves different weights behind (quirk Q9)
    assert np.allclose(merged[..., :3][full], mean[..., :3][full], rtol=2e-5, atol=1e-6)


def test_pixel_shard_rows_are_bit_identical_and_sum_to_the_image(device, loaded):
    """lmb_set_pixel_shard (SURVEY.md 8e secondary axis, BASELINE config 4): a shard renders exactly its interleaved rows,
    bit-identical to the same rows of the unsharded render (seeds are those of the full image), leaves the other rows
    untouched, traces exactly its share of the rays, and 3 pixel shards x 2 sample shards add up to the whole image."""
    sc, orc = loaded("materials", 96, 50)  # 50 rows over 3 shards: 17 + 17 + 16
    pc, ubo = sc.make_pc(8, True), sc.make_ubo()
    try:
        device.set_pixel_shard(0, 1)
        device.init(96, 50, 2)
        device.reset_stats()
        device.render(pc, ubo, 0, 4)
        full, full_rays = device.download(), device.stats().rays
        device.clear_film()
        device.render(pc, ubo, 0, 4, 1, integrator.FILM_SUM)
        full_sum = device.download()
        total, rays = np.zeros_like(full), 0
        for p in range(3):
            device.set_pixel_shard(p, 3)
            device.init(96, 50, 3)  # a batch size that does not divide the frame count
            device.reset_stats()
            device.render(pc, ubo, 0, 4)
            part = device.download()
            rays += device.stats().rays
            own = np.zeros(50, bool)
            own[p::3] = True
            assert part[own].tobytes() == full[own].tobytes()
            assert not part[~own].any()
            for s in range(2):  # sample shards inside the pixel shard: frames {0, 2} and {1, 3}
                device.clear_film()
                device.render(pc, ubo, s, 2, 2, integrator.FILM_SUM)
                shard = device.download()
                assert not shard[~own].any()
                total += shard
        assert rays == full_rays
        assert (total[..., 3] == full_sum[..., 3]).all()
        assert np.allclose(total[..., :3], full_sum[..., :3], rtol=2e-6, atol=1e-7)
        with pytest.raises(RuntimeError):
            device.set_pixel_shard(3, 3)
        device.set_pixel_shard(60, 64)  # owns no row of a 50-row image
        with pytest.raises(RuntimeError):
            device.init(96, 50, 1)
    finally:
        device.set_pixel_shard(0, 1)


def test_render_without_accel_fails_cleanly(device):
    sc = host.Scene(scene_path("cornell"), 32, 32)
    dev2 = integrator.Device(0)
    try:
        dev2.upload_scene(sc.desc)
        dev2.init(32, 32, 1)
        with pytest.raises(RuntimeError, match="lmb_build_accel"):
            dev2.render(sc.make_pc(6, True), sc.make_ubo(), 0, 1)
    finally:
        dev2.close()


def test_integrator_lifecycle_mirror(device):
    """PathB200 behind Lumen's init/render/update/destroy lifecycle (Path.cpp:4-70)."""
    sc = host.Scene(scene_path("cornell"), 64, 64)
    integ = integrator.PathB200(sc, device=0, frames_in_flight=2)
    integ.path_length = 6
    integ.init()
    assert integ.frame_num == 0
    for _ in range(3):
        integ.render()
        integ.update()
    assert integ.frame_num == 3
    out = integ.output()
    cpu, _ = po.OracleScene(sc).render(sc.make_pc(6, True), sc.make_ubo(), 0, 3)
    integ.destroy()
    assert bits_equal(out, cpu).all()


def test_upload_rejects_out_of_range_indices(device):
    """The C ABI validates every index the kernels will follow (material, vertex, texture, light mesh) and reports it instead of
    reading out of bounds on the device; the context stays usable."""
    import ctypes as C
    from lumen_b200._ctypes_types import SceneDesc, PrimMeshInfo
    sc = host.Scene(scene_path("cornell"), 32, 32)
    dev2 = integrator.Device(0)
    try:
        bad = SceneDesc.from_buffer_copy(sc.desc)
        infos = (PrimMeshInfo * sc.desc.n_prim_meshes).from_address(sc.desc.prim_infos)
        copy = (PrimMeshInfo * sc.desc.n_prim_meshes)(*infos)
        copy[1].material_index = sc.desc.n_materials + 3
        bad.prim_infos = C.addressof(copy)
        with pytest.raises(RuntimeError, match="material_index out of range"):
            dev2.upload_scene(bad)
        bad = SceneDesc.from_buffer_copy(sc.desc)
        bad.n_vertices = sc.desc.n_vertices - 5
        with pytest.raises(RuntimeError, match="vertex index out of range"):
            dev2.upload_scene(bad)
        bad = SceneDesc.from_buffer_copy(sc.desc)
        bad.n_indices = sc.desc.n_indices - 3
        with pytest.raises(RuntimeError, match="index range exceeds"):
            dev2.upload_scene(bad)
        dev2.upload_scene(sc.desc)  # the context is still good
        dev2.build_accel()
        dev2.init(32, 32, 1)
        dev2.render(sc.make_pc(6, True), sc.make_ubo(), 0, 1)
        assert dev2.stats().rays > 0
    finally:
        dev2.close()


def test_render_rejects_push_constants_that_do_not_fit_the_scene(device, loaded):
    """lmb_render / lmb_render_bdpt check the PCPath fields the kernels index or divide by (num_lights, dir_light_idx,
    light_triangle_count, max_depth) against the uploaded scene and report instead of reading lights[] out of bounds; an area light
    whose world_matrix is not its mesh's is refused at upload. The context stays usable."""
    import ctypes as C
    from lumen_b200._ctypes_types import Light, PCBdpt, SceneDesc
    sc, orc = loaded("cornell_dir", 48, 32)
    device.init(48, 32, 1)
    ubo = sc.make_ubo()

    def broken(**kw):
        pc = sc.make_pc(6, True)
        for k, v in kw.items():
            setattr(pc, k, v)
        return pc

    n = sc.info.n_lights
    for kw, msg in ((dict(num_lights=n + 1), "num_lights"), (dict(num_lights=-1), "num_lights"), (dict(dir_light_idx=n), "dir_light_idx"),
                    (dict(light_triangle_count=0), "light_triangle_count"), (dict(light_triangle_count=-3), "light_triangle_count"),
                    (dict(max_depth=-1), "max_depth"), (dict(max_depth=100000), "max_depth")):
        with pytest.raises(RuntimeError, match=msg):
            device.render(broken(**kw), ubo, 0, 1)
    bp = PCBdpt.from_path_pc(sc.make_pc(6, True))
    bp.num_lights = n + 7
    with pytest.raises(RuntimeError, match="num_lights"):
        device.render_bdpt(bp, ubo, 0, 1)
    bp = PCBdpt.from_path_pc(sc.make_pc(6, True))
    bp.max_depth = 65
    with pytest.raises(RuntimeError, match="max_depth"):
        device.render_bdpt(bp, ubo, 0, 1)
    device.render(sc.make_pc(6, True), ubo, 0, 1)  # still good
    got = device.download()
    want, _ = orc.render(sc.make_pc(6, True), ubo, 0, 1)
    assert bits_equal(got, want).all()
    # dir_light_idx naming a light that is not directional
    sc2, _ = loaded("cornell", 48, 32)
    pc = sc2.make_pc(6, True)
    pc.dir_light_idx = 0
    with pytest.raises(RuntimeError, match="directional"):
        device.render(pc, sc2.make_ubo(), 0, 1)
    # area light with a world matrix of its own
    dev2 = integrator.Device(0)
    try:
        bad = SceneDesc.from_buffer_copy(sc2.desc)
        lights = (Light * sc2.desc.n_lights).from_address(sc2.desc.lights)
        copy = (Light * sc2.desc.n_lights)(*lights)
        copy[0].world_matrix[12] += 1.0
        bad.lights = C.addressof(copy)
        with pytest.raises(RuntimeError, match="world_matrix differs"):
            dev2.upload_scene(bad)
    finally:
        dev2.close()


def test_film_add_from_reduces_sum_films(device, loaded):
    """lmb_film_add_from: dst.film += src.film, device to device -- the multi-GPU reduce of the C++ host (PathB200Multi). Two contexts
    render the even and the odd frames in sum mode; added and resolved they give the mean of all frames."""
    sc, orc = loaded("cornell", 64, 48)
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    a, b = integrator.Device(0), integrator.Device(0)
    try:
        for d, first in ((a, 0), (b, 1)):
            d.upload_scene(sc.desc)
            d.build_accel()
            d.init(64, 48, 2)
            d.render(pc, ubo, first, 3, 2, integrator.FILM_SUM)  # frames first, first + 2, first + 4
        sa, sb = a.download(), b.download()
        a.film_add_from(b)
        assert a.download().tobytes() == (sa + sb).tobytes()
        a.resolve()
        got = a.download()
        device.init(64, 48, 3)
        device.render(pc, ubo, 0, 6)
        want = device.download()
        assert np.allclose(got[..., :3], want[..., :3], rtol=3e-6, atol=1e-7) and (got[..., 3] == 1.0).all()
        with pytest.raises(RuntimeError, match="dst == src"):
            a.film_add_from(a)
        b.init(32, 32, 1)
        with pytest.raises(RuntimeError, match="image sizes differ"):
            a.film_add_from(b)
    finally:
        a.close()
        b.close()


def test_sorted_ray_batches_return_the_same_hits(device, loaded):
    """lmb_trace_closest_device_ex(sort_rays = 1): the batch is ordered by (origin cell, direction octant) on the device before it is
    walked (config 5's incoherent rays). Hits land at the rays' own indices and are those of the unsorted launch and of the oracle,
    bit for bit -- including rays outside the scene box, non-finite rays and a count that is not a multiple of 32."""
    import torch
    sc, orc = loaded("materials", 64, 64)
    rng = np.random.default_rng(77)
    lo, hi = np.float32([-6, -1, -6]), np.float32([6, 6, 6])
    rays = random_rays(rng, lo, hi, 200_003)
    rays[:50, :3] *= 40.0                    # far outside the box: clamped cells
    rays[50:60, 0] = np.nan                  # hit nothing by definition
    rays[60:70, 4] = np.inf
    d_rays = torch.from_numpy(rays).cuda()
    a = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
    b = torch.full((rays.shape[0], 4), -7.0, dtype=torch.float32, device="cuda")
    device.trace_closest_device(d_rays.data_ptr(), rays.shape[0], a.data_ptr(), 1)
    device.trace_closest_device(d_rays.data_ptr(), rays.shape[0], b.data_ptr(), 2, sort_rays=True)
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    want, _ = orc.trace_closest(rays)
    got = b.cpu().numpy()
    assert (got[:, 3].view(np.uint32) == want["prim"]).all() and bits_equal(got[:, 0][want["prim"] != 0xFFFFFFFF], want["t"][want["prim"] != 0xFFFFFFFF]).all()
    assert (want["prim"][50:70] == 0xFFFFFFFF).all()
