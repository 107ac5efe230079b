"""BASELINE.json configs 1-3 at their FULL resolution against the oracle (tests/test_gpu_scene.py runs reduced sizes): ray counts
equal, film bit-equal, on as many frames as the oracle renders in seconds on the box's host threads. Config 2 at this size is
where the culling rule of the hit definition was found not to be closed under the triangle test's rounding (3 rays of 26 M
differed between the two walks before tri_clamp_t)."""
import os
import sys

import pytest

from conftest import ROOT, scene_path
from helpers import bits_equal, pixel_agreement
from lumen_b200 import host
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _classroom():
    sys.path.insert(0, os.path.join(ROOT, "scenes"))
    import gen_classroom_standin as gen
    return gen.generate(os.path.join(ROOT, "scenes", "_generated", "classroom_standin"))[0]


CASES = [
    ("config 1: cornell_box path 512x512 16 spp depth 6", lambda: scene_path("cornell"), 512, 512, 6, 16),
    ("config 2: caustics 1280x720 depth 12 (4 of 64 spp)", lambda: scene_path("caustics"), 1280, 720, 12, 4),
    ("config 3: classroom stand-in 1920x1080 depth 8 (1 of 1024 spp)", _classroom, 1920, 1080, 8, 1),
]
# (first_frame, n_frames) windows late in each configuration's frame range: the seed (x, y, frame_num, 0) and the running-mean
# weight 1 / (frame_num + 1) at the far end, on top of a film that already holds something
LATE = [
    ("config 1: frames 14-15 of 16", lambda: scene_path("cornell"), 512, 512, 6, 14, 2),
    ("config 2: frame 63 of 64", lambda: scene_path("caustics"), 1280, 720, 12, 63, 1),
    ("config 3: frame 511 of 1024", _classroom, 1920, 1080, 8, 511, 1),
    ("config 3: frame 1023 of 1024", _classroom, 1920, 1080, 8, 1023, 1),
]


@pytest.mark.parametrize("label,path_fn,w,h,depth,frames", CASES, ids=[c[0].split(":")[0] for c in CASES])
def test_full_resolution_parity(device, label, path_fn, w, h, depth, frames):
    sc = host.Scene(path_fn(), w, h)
    orc = po.OracleScene(sc)
    pc, ubo = sc.make_pc(depth, True), sc.make_ubo()
    device.set_pixel_shard(0, 1)
    device.upload_scene(sc.desc)
    device.build_accel()
    device.init(w, h, 0)
    device.render(pc, ubo, 0, frames)
    gpu, gs = device.download(), device.stats()
    cpu, cs = orc.render(pc, ubo, 0, frames)
    assert (gs.rays_closest, gs.rays_shadow, gs.rays_probe) == (cs.rays_closest, cs.rays_shadow, cs.rays_probe), label
    assert gs.nan_samples == cs.nan_pixels
    assert pixel_agreement(gpu, cpu) >= 0.999
    assert bits_equal(gpu, cpu).mean() >= 0.9999


@pytest.mark.parametrize("label,path_fn,w,h,depth,first,frames", LATE, ids=[c[0] for c in LATE])
def test_full_resolution_parity_late_frames(device, label, path_fn, w, h, depth, first, frames):
    import numpy as np
    sc = host.Scene(path_fn(), w, h)
    orc = po.OracleScene(sc)
    pc, ubo = sc.make_pc(depth, True), sc.make_ubo()
    device.set_pixel_shard(0, 1)
    device.upload_scene(sc.desc)
    device.build_accel()
    device.init(w, h, 0)
    start = np.random.default_rng(first).uniform(0, 2, (h, w, 4)).astype(np.float32)  # stands for the mean of the frames before
    start[..., 3] = 1.0
    device.upload_film(start)
    device.reset_stats()
    device.render(pc, ubo, first, frames)
    gpu, gs = device.download(), device.stats()
    cpu, cs = orc.render(pc, ubo, first, frames, rgba=start.copy())
    assert (gs.rays_closest, gs.rays_shadow, gs.rays_probe) == (cs.rays_closest, cs.rays_shadow, cs.rays_probe), label
    assert pixel_agreement(gpu, cpu) >= 0.999
    assert bits_equal(gpu, cpu).mean() >= 0.9999
    assert not bits_equal(gpu, start).all()


def test_config4_pixel_x_sample_shards_at_4k(device):
    """BASELINE config 4's shape at its full resolution: 3840x2160, 2 pixel shards (interleaved rows) x 2 sample shards (frame index
    mod 2), four LMB_FILM_SUM films added and resolved. The oracle renders the rows of pixel shard 1 only (orc_set_row_shard): there
    the sum of the shards must be the mean of frames 0..3 (to fp32 association; every sample bit-equal), with equal ray counts."""
    import numpy as np
    from lumen_b200 import integrator
    w, h, depth, n_frames = 3840, 2160, 8, 4
    sc = host.Scene(_classroom(), w, h)
    orc = po.OracleScene(sc)
    pc, ubo = sc.make_pc(depth, True), sc.make_ubo()
    device.upload_scene(sc.desc)
    device.build_accel()
    total = np.zeros((h, w, 4), dtype=np.float32)
    rays_shard1 = [0, 0, 0]
    try:
        for p in range(2):
            device.set_pixel_shard(p, 2)
            device.init(w, h, 1)
            for s_ in range(2):
                device.clear_film()
                device.reset_stats()
                device.render(pc, ubo, s_, n_frames // 2, 2, integrator.FILM_SUM)  # frames s, s + 2
                film = device.download()
                assert (film[(1 - p)::2] == 0).all()  # the other shard's rows are never touched
                total += film
                if p == 1:
                    st = device.stats()
                    rays_shard1 = [a + b for a, b in zip(rays_shard1, (st.rays_closest, st.rays_shadow, st.rays_probe))]
    finally:
        device.set_pixel_shard(0, 1)
    assert (total[..., 3] == n_frames).all()
    resolved = total[..., :3] / total[..., 3:4]
    po.set_row_shard(1, 2)
    try:
        cpu, cs = orc.render(pc, ubo, 0, n_frames)
    finally:
        po.set_row_shard(0, 1)
    assert rays_shard1 == [cs.rays_closest, cs.rays_shadow, cs.rays_probe]
    assert np.allclose(resolved[1::2], cpu[1::2, :, :3], rtol=3e-6, atol=1e-7)
    assert (cpu[0::2] == 0).all()


@pytest.mark.parametrize("label,path_fn,w,h,depth,frames", [CASES[0][:5] + (2,), CASES[1][:5] + (1,), CASES[2][:5] + (1,)], ids=["config 1", "config 2", "config 3"])
def test_full_resolution_parity_bdpt(device, label, path_fn, w, h, depth, frames):
    """The same three configurations through lmb_render_bdpt against the oracle's BDPT: the pixels' own
    strategies bit-equal, the light-tracer image and the film within 1e-4 on >= 99.9 % of pixels, equal ray counts."""
    from lumen_b200._ctypes_types import PCBdpt
    sc = host.Scene(path_fn(), w, h)
    orc = po.OracleScene(sc)
    pc, ubo = PCBdpt.from_path_pc(sc.make_pc(depth, True)), sc.make_ubo()
    device.set_pixel_shard(0, 1)
    device.upload_scene(sc.desc)
    device.build_accel()
    device.init(w, h, 1)
    for frame in range(frames):
        device.reset_stats()
        col, splat = device.kat_bdpt_frame_raw(pc, ubo, frame)
        gs = device.stats()
        ocol, osplat, cs = orc.render_bdpt_frame_raw(pc, ubo, frame)
        assert (gs.rays_closest, gs.rays_shadow) == (cs.rays_closest, cs.rays_shadow), label
        assert bits_equal(col, ocol).all(axis=-1).mean() >= 0.9999, label
        assert pixel_agreement(splat, osplat) >= 0.999, label
    gpu = device.download()
    cpu, _ = orc.render_bdpt(pc, ubo, 0, frames)
    assert pixel_agreement(gpu, cpu) >= 0.999


def test_config5_full_size_lbvh_and_hits(device):
    """Config 5 at its full size (10 x 10 x 10 tori, 10 M triangles): the canonical LBVH of the GPU build is memcmp-equal to the
    CPU build (acceptance criterion 1), the 8-wide tree holds every triangle once, and 2^18 incoherent closest / any-hit rays
    return bit-identical records."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "scenes"))
    import gen_torus_grid as gen
    sc = gen.make_scene(10, 100, 50, 256, 256)
    assert sc.info.n_triangles == 10_000_000
    orc = po.OracleScene(sc)
    device.set_pixel_shard(0, 1)
    device.upload_scene(sc.desc)
    device.build_accel()
    g, c = device.lbvh(), orc.lbvh()
    for k in ("keys", "morton", "leaf_prim", "left", "right", "parent"):
        assert g[k].tobytes() == c[k].tobytes(), k
    assert g["aabb"].tobytes() == c["aabb"].tobytes()
    chk = device.wide_bvh_check()
    assert chk["errors"] == 0 and chk["dup_or_missing"] == 0 and chk["leaf_tris"] == 10_000_000 and chk["reachable"] == chk["nodes"]
    rays = gen.random_rays(1 << 18, 10)
    gh, (ch, _) = device.trace_closest(rays), orc.trace_closest(rays)
    assert (gh["prim"] == ch["prim"]).all() and bits_equal(gh["t"], ch["t"]).all() and bits_equal(gh["b1"], ch["b1"]).all()
    assert (gh["prim"] != 0xFFFFFFFF).mean() > 0.5
    assert (device.trace_any(rays) == orc.trace_any(rays)[0]).all()
