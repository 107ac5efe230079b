"""GPU parity of the remaining BASELINE.json configurations and of the traversal structure behind them:
the 8-wide BVH (structure, agreement with the binary LBVH walk), the classroom stand-in (config 3 path: Mitsuba XML,
sun + atmosphere sky, principled / conductor / diffuse), the synthetic torus grid (config 5) and the converged-image
RMSE criterion (BASELINE.json condition 3, both RMSE routines of SURVEY.md 5.5)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, scene_path
from helpers import bits_equal, pixel_agreement, random_rays
from lumen_b200 import host, integrator
from oracle import pyoracle as po

sys.path.insert(0, os.path.join(ROOT, "scenes"))
import gen_classroom_standin  # noqa: E402
import gen_torus_grid  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def classroom(tmp_path_factory):
    path, _ = gen_classroom_standin.generate(str(tmp_path_factory.mktemp("classroom_standin")))
    return path


@pytest.mark.parametrize("name", ["cornell", "caustics", "materials"])
def test_wide_bvh_structure(device, name):
    """Every triangle sits in exactly one leaf, every quantised child box contains its subtree, all nodes reachable."""
    sc = host.Scene(scene_path(name), 64, 64)
    device.upload_scene(sc.desc)
    device.build_accel()
    chk = device.wide_bvh_check()
    assert chk["errors"] == 0 and chk["dup_or_missing"] == 0, chk
    assert chk["nodes"] == chk["reachable"] == device.stats().wide_nodes
    assert chk["leaf_tris"] == sc.info.n_triangles
    assert chk["depth"] == device.stats().wide_levels


def test_wide_walk_equals_binary_walk(device, monkeypatch):
    """LMB_TRAVERSAL=bvh2 walks the canonical binary LBVH; both walks must return identical hits (t, barycentrics, prim)."""
    sc = host.Scene(scene_path("materials"), 64, 64)
    device.upload_scene(sc.desc)
    device.build_accel()
    box = device.lbvh()["aabb"][:6]
    rays = random_rays(np.random.default_rng(5), box[:3], box[3:], 200000)
    wide = device.trace_closest(rays)
    wide_any = device.trace_any(rays)
    monkeypatch.setenv("LMB_TRAVERSAL", "bvh2")
    dev2 = integrator.Device(0)
    try:
        dev2.upload_scene(sc.desc)
        dev2.build_accel()
        binary = dev2.trace_closest(rays)
        for k in ("prim", "t", "b1", "b2"):
            assert bits_equal(wide[k], binary[k]).all(), k
        assert (wide_any == dev2.trace_any(rays)).all()
    finally:
        dev2.close()


def test_pinned_and_unpinned_walkers_agree(monkeypatch):
    """k_trace has two instantiations (stack column address pinned in a register or not), picked by the BVH footprint against the
    L2; LMB_TRACE_PIN forces one. Same hits, same film, same traversal work."""
    from oracle import pyoracle as po
    sc = host.Scene(scene_path("cornell"), 96, 96)
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    rays = random_rays(np.random.default_rng(9), [-3, -1, -3], [3, 5, 3], 50000)
    out = []
    for pin in ("0", "1"):
        monkeypatch.setenv("LMB_TRACE_PIN", pin)
        dev = integrator.Device(0)
        try:
            dev.upload_scene(sc.desc)
            dev.build_accel()
            dev.init(96, 96, 3)
            dev.render(pc, ubo, 0, 5)
            st = dev.stats()
            out.append((dev.download(), dev.trace_closest(rays), dev.trace_any(rays), (st.rays_closest, st.rays_shadow, st.rays_probe, st.nodes_visited)))
        finally:
            dev.close()
    assert out[0][0].tobytes() == out[1][0].tobytes()
    assert out[0][1].tobytes() == out[1][1].tobytes() and (out[0][2] == out[1][2]).all()
    assert out[0][3] == out[1][3]
    cpu, _ = po.OracleScene(sc).render(pc, ubo, 0, 5)
    assert bits_equal(out[0][0], cpu).mean() >= 0.999


def test_tiny_and_degenerate_scenes(device):
    """1, 2 and 5 triangle scenes (root-only wide trees) and a scene of coplanar, axis-aligned triangles (flat boxes)."""
    from helpers import make_material
    rng = np.random.default_rng(9)
    mats = [make_material(albedo=(0.5, 0.5, 0.5), bsdf_type=1, bsdf_props=1 | 8)]
    for n_tri in (1, 2, 5, 64):
        v = np.zeros((n_tri * 3, 8), np.float32)
        if n_tri == 64:  # a flat 8 x 4 grid of quads in the plane y = 0.25
            k = 0
            for i in range(8):
                for j in range(4):
                    for tri in (((0, 0), (1, 0), (1, 1)), ((0, 0), (1, 1), (0, 1))):
                        for (a, b) in tri:
                            v[k, :3] = (i + a, 0.25, j + b)
                            k += 1
        else:
            v[:, :3] = rng.uniform(-1, 1, (n_tri * 3, 3))
        v[:, 3:6] = (0, 1, 0)
        sc = host.Scene.from_arrays(v, [n_tri], [0], mats, width=32, height=32)
        orc = po.OracleScene(sc)
        device.upload_scene(sc.desc)
        device.build_accel()
        chk = device.wide_bvh_check()
        assert chk["errors"] == 0 and chk["dup_or_missing"] == 0 and chk["leaf_tris"] == n_tri, (n_tri, chk)
        lo, hi = v[:, :3].min(0) - 0.5, v[:, :3].max(0) + 0.5
        rays = random_rays(rng, lo, hi, 50000)
        rays[:10000, 4:7] = (0.0, -1.0, 0.0)  # straight down: parallel to two slabs of every flat box
        gh, (ch, _) = device.trace_closest(rays), orc.trace_closest(rays)
        assert (gh["prim"] == ch["prim"]).all() and bits_equal(gh["t"], ch["t"]).all(), n_tri
        assert (device.trace_any(rays) == orc.trace_any(rays)[0]).all()


def test_classroom_standin_parity(device, classroom):
    """Config 3's code path at a size the oracle finishes in seconds: Mitsuba XML loader, sun + atmosphere sky on misses,
    principled / conductor / diffuse materials, depth 8."""
    w, h, depth, frames = 240, 135, 8, 2
    sc = host.Scene(classroom, w, h)
    orc = po.OracleScene(sc)
    device.upload_scene(sc.desc)
    device.build_accel()
    g, c = device.lbvh(), orc.lbvh()
    for k in ("keys", "left", "right", "parent"):
        assert (g[k] == c[k]).all(), k
    assert g["aabb"].tobytes() == c["aabb"].tobytes()
    chk = device.wide_bvh_check()
    assert chk["errors"] == 0 and chk["dup_or_missing"] == 0 and chk["leaf_tris"] == sc.info.n_triangles
    pc, ubo = sc.make_pc(depth, True), sc.make_ubo()
    assert pc.dir_light_idx != 0xFFFFFFFF  # sun + sky
    device.init(w, h, 2)
    device.render(pc, ubo, 0, frames)
    gpu, gs = device.download(), device.stats()
    cpu, cs = orc.render(pc, ubo, 0, frames)
    assert (gs.rays_closest, gs.rays_shadow, gs.rays_probe) == (cs.rays_closest, cs.rays_shadow, cs.rays_probe)
    assert pixel_agreement(gpu, cpu) >= 0.999
    assert bits_equal(gpu, cpu).mean() >= 0.999


@pytest.mark.parametrize("depth", [1, 2, 70])
def test_sky_march_ranges_at_odd_depths(device, classroom, depth):
    """k_miss hands out the escaped rays from cursors over ranges of the miss records -- one range per batch by default, one per bounce
    (63 with a range of their own, deeper ones share the last) with LMB_MISS_SIDE_BLOCKS=n, which profiles/ runs exercise: depth 1 (the
    only bounce is the last), 2, and 70 (beyond the ranges; Russian roulette keeps it cheap) against the oracle, over two batches."""
    w, h = 96, 54
    sc = host.Scene(classroom, w, h)
    orc = po.OracleScene(sc)
    device.upload_scene(sc.desc)
    device.build_accel()
    pc, ubo = sc.make_pc(depth, True), sc.make_ubo()
    device.init(w, h, 2)
    device.render(pc, ubo, 0, 3)  # two batches: the ranges start over
    gpu, gs = device.download(), device.stats()
    cpu, cs = orc.render(pc, ubo, 0, 3)
    assert (gs.rays_closest, gs.rays_shadow, gs.rays_probe) == (cs.rays_closest, cs.rays_shadow, cs.rays_probe)
    assert bits_equal(gpu, cpu).mean() >= 0.999


def test_torus_grid_config5_small(device):
    """Config 5 at 4 x 4 x 4 tori (82 k triangles): bit-exact LBVH, valid wide tree, closest / any-hit parity on the
    incoherent-ray generator of the sweep, and primary-ray render parity."""
    K = 4
    sc = gen_torus_grid.make_scene(K=K, nu=32, nv=20, width=128, height=128)
    orc = po.OracleScene(sc)
    device.upload_scene(sc.desc)
    device.build_accel()
    g, c = device.lbvh(), orc.lbvh()
    for k in ("morton", "keys", "leaf_prim", "left", "right", "parent"):
        assert (g[k] == c[k]).all(), k
    assert g["aabb"].tobytes() == c["aabb"].tobytes()
    chk = device.wide_bvh_check()
    assert chk["errors"] == 0 and chk["dup_or_missing"] == 0 and chk["leaf_tris"] == sc.info.n_triangles == K ** 3 * 2 * 32 * 20
    rays = gen_torus_grid.random_rays(1 << 18, K)
    gh, (ch, _) = device.trace_closest(rays), orc.trace_closest(rays)
    for k in ("prim", "t", "b1", "b2"):
        assert bits_equal(gh[k], ch[k]).all(), k
    assert 0.2 < (gh["prim"] != 0xFFFFFFFF).mean() < 0.99
    rays[:, 3], rays[:, 7] = 0.0, 1.5
    assert (device.trace_any(rays) == orc.trace_any(rays)[0]).all()
    pc, ubo = sc.make_pc(4, True), sc.make_ubo()
    device.init(128, 128, 2)
    device.render(pc, ubo, 0, 2)
    cpu, cs = orc.render(pc, ubo, 0, 2)
    assert device.stats().rays == cs.rays and cs.rays_shadow > 0.3 * 128 * 128 * 2  # the camera does see the grid
    assert bits_equal(device.download(), cpu).mean() >= 0.999


def test_converged_rmse_within_sampling_noise(device):
    """BASELINE.json condition 3: the RMSE of the GPU's converged image against the oracle's converged image (independent
    frames) is the oracle's own seed-to-seed noise -- with Lumen's literal RMSE routine (rmse/*.comp, quirks included) and
    with a true RMSE."""
    w = h = 256
    n = 512  # 512 spp at 256 x 256: the image is converged to ~1 % (noise falls as 1 / sqrt(n)); ~30 s of oracle time on the box
    sc = host.Scene(scene_path("cornell"), w, h)
    orc = po.OracleScene(sc)
    device.upload_scene(sc.desc)
    device.build_accel()
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    device.init(w, h, 64)
    device.render(pc, ubo, 0, n)
    gpu_a = device.download()
    cpu_a, _ = orc.render(pc, ubo, 0, n)
    # an independent estimate: frames [n, 2n) accumulated on their own (sum mode starts a fresh mean)
    device.clear_film()
    device.render(pc, ubo, n, n, 1, integrator.FILM_SUM)
    device.resolve()
    gpu_b = device.download()
    ref_noise_true = po.rmse_true(cpu_a, gpu_b)  # oracle [0,n) vs independent frames: the sampling noise level
    got_true = po.rmse_true(gpu_a, gpu_b)
    assert ref_noise_true > 0
    assert abs(got_true - ref_noise_true) <= 1e-3 * ref_noise_true
    assert po.rmse_true(gpu_a, cpu_a) <= 1e-3 * ref_noise_true  # same frames: far below the noise
    lit_ref, lit_got = po.rmse_literal(cpu_a, gpu_b), po.rmse_literal(gpu_a, gpu_b)
    assert abs(lit_got - lit_ref) <= 1e-3 * max(abs(lit_ref), 1e-12)
