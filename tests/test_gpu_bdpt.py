"""GPU parity of the BDPT integrator (bdpt.rgen + bdpt_commons.glsl) against the CPU oracle's restatement
(oracle/bdpt.h), through the C ABI (lmb_render_bdpt). The pixel's own strategies (t >= 2) are compared bit for bit; the
light-tracer image is a sum of float atomics on the GPU and a sum in source-pixel order in the oracle (quirk B2), so it and the
film are held to BASELINE.json's 1e-4 relative on >= 99.9 % of pixels."""
import numpy as np
import pytest

from conftest import scene_path
from helpers import bits_equal, pixel_agreement
from lumen_b200 import host, integrator
from lumen_b200._ctypes_types import PCBdpt
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

# scene, size, max_depth: area light + diffuse / mirror / dielectric; spot light + glass; all six BSDFs + 3 area lights;
# directional light (infinite: the g_term / pdf branches of the light walk) + constant sky
CASES = [("cornell", 96, 6), ("caustics", 96, 8), ("materials", 96, 7), ("cornell_dir", 64, 5)]


def _setup(device, name, size, depth, time=0):
    sc = host.Scene(scene_path(name), size, size)
    device.upload_scene(sc.desc)
    device.build_accel()
    device.init(size, size, 1)
    pc = PCBdpt.from_path_pc(sc.make_pc(depth, True), time)
    return sc, pc, sc.make_ubo()


@pytest.mark.parametrize("name,size,depth", CASES)
def test_bdpt_frame_matches_oracle(device, name, size, depth):
    sc, pc, ubo = _setup(device, name, size, depth)
    osc = po.OracleScene(sc)
    for frame in (0, 3):
        device.reset_stats()
        col, splat = device.kat_bdpt_frame_raw(pc, ubo, frame)
        st = device.stats()
        ocol, osplat, ost = osc.render_bdpt_frame_raw(pc, ubo, frame)
        assert (st.rays_closest, st.rays_shadow) == (ost.rays_closest, ost.rays_shadow)
        same = bits_equal(col, ocol).all(axis=-1)
        assert same.mean() >= 0.999, f"{name} frame {frame}: own strategies bit-equal on {same.mean():.5f} of pixels"
        assert pixel_agreement(col, ocol) >= 0.999
        assert pixel_agreement(splat, osplat) >= 0.999, f"{name} frame {frame}: light-tracer image"
        assert (splat.sum(-1) > 0).any() or name == "cornell_dir"
    osc.close()


@pytest.mark.parametrize("name,size,depth", [("materials", 96, 7), ("caustics", 80, 9), ("cornell_dir", 64, 5)])
def test_bdpt_three_pipelines_agree(device, monkeypatch, name, size, depth):
    """Default: staged pipeline with pair-parallel connections (k_bdpt_pair: one thread per (s, t) pair and pixel, MIS weights
    without the in-place vertex patches). LMB_BDPT=pixel: staged, connections looped per pixel with the GLSL's patch-and-restore.
    LMB_BDPT=mega: one kernel, rays traced in the thread over the binary LBVH. Same bits, same ray counts."""
    sc, pc, ubo = _setup(device, name, size, depth)
    device.reset_stats()
    col, splat = device.kat_bdpt_frame_raw(pc, ubo, 1)
    st = device.stats()
    for mode in ("pixel", "mega"):
        monkeypatch.setenv("LMB_BDPT", mode)
        device.reset_stats()
        mcol, msplat = device.kat_bdpt_frame_raw(pc, ubo, 1)
        mst = device.stats()
        assert bits_equal(col, mcol).all(), mode
        assert pixel_agreement(splat, msplat, rel=1e-5) >= 0.999, mode
        assert (st.rays_closest, st.rays_shadow) == (mst.rays_closest, mst.rays_shadow), mode


def test_bdpt_film_matches_oracle(device):
    """lmb_render_bdpt over several frames = the oracle's running-mean film; frames rendered one call at a time = one batched call."""
    sc, pc, ubo = _setup(device, "cornell", 96, 6)
    osc = po.OracleScene(sc)
    ofilm, ost = osc.render_bdpt(pc, ubo, 0, 6)
    device.clear_film()
    device.reset_stats()
    device.render_bdpt(pc, ubo, 0, 6)
    film = device.download()
    st = device.stats()
    assert pixel_agreement(film, ofilm) >= 0.999
    assert st.nan_samples == ost.nan_pixels
    assert (st.rays_closest, st.rays_shadow, st.rays_probe) == (ost.rays_closest, ost.rays_shadow, 0)
    device.clear_film()
    for f in range(6):
        device.render_bdpt(pc, ubo, f, 1)
    assert pixel_agreement(device.download(), film, rel=1e-5) >= 0.999
    osc.close()


@pytest.mark.parametrize("batch", ["1", "4"])
def test_bdpt_frames_in_flight_do_not_change_the_film(device, monkeypatch, batch):
    """Frames in flight (LMB_BDPT_BATCH; default up to 8): 6 frames as 4 + 2 (a short last batch changes the stride of every
    struct-of-arrays buffer) and one at a time give the film of the default run; the own strategies are bit-equal per frame, the
    light-tracer splats add in atomic order, hence the 1e-5 of the neighbouring test."""
    sc, pc, ubo = _setup(device, "caustics", 64, 7)
    device.clear_film()
    device.reset_stats()
    device.render_bdpt(pc, ubo, 3, 6, 2, integrator.FILM_SUM)  # frames 3, 5, .. 13 into a sum film (a strided running mean is refused)
    want, st = device.download(), device.stats()
    monkeypatch.setenv("LMB_BDPT_BATCH", batch)
    device.clear_film()
    device.reset_stats()
    device.render_bdpt(pc, ubo, 3, 6, 2, integrator.FILM_SUM)
    got, st2 = device.download(), device.stats()
    assert (got[..., 3] == want[..., 3]).all()
    assert pixel_agreement(got, want, rel=1e-5) >= 0.999
    assert (st2.rays_closest, st2.rays_shadow) == (st.rays_closest, st.rays_shadow)


def test_bdpt_time_enters_the_seed(device):
    """bdpt.rgen:36-37: seed.z = frame_num ^ pc.time. (frame 5, time 0) and (frame 4, time 1) are the same sample set."""
    sc, pc, ubo = _setup(device, "cornell", 64, 4)
    a, _ = device.kat_bdpt_frame_raw(pc, ubo, 5)
    pc1 = PCBdpt.from_path_pc(sc.make_pc(4, True), 1)
    b, _ = device.kat_bdpt_frame_raw(pc1, ubo, 4)
    c, _ = device.kat_bdpt_frame_raw(pc1, ubo, 5)
    assert bits_equal(a, b).all()
    assert not bits_equal(a, c).all()


def test_bdpt_lifecycle_class():
    """BDPTB200 mirrors BDPT::init / render / update / destroy (BDPT.cpp) and converges towards the Path integrator's image
    away from the emitter (oracle diagnostics: DESIGN.md section 8)."""
    sc = host.Scene(scene_path("cornell"), 64, 64)
    integ = integrator.BDPTB200(sc)
    integ.path_length = 4
    integ.init()
    for _ in range(4):
        integ.render(4)
        integ.update()
    assert integ.frame_num == 16
    img = integ.output()
    integ.destroy()
    assert np.isfinite(img).all() and img[..., 3].min() == 1.0 and img[..., :3].mean() > 0.05


def test_bdpt_rejects_pixel_shards():
    dev = integrator.Device(0)
    try:
        sc = host.Scene(scene_path("cornell"), 32, 32)
        dev.upload_scene(sc.desc)
        dev.build_accel()
        dev.set_pixel_shard(1, 2)
        dev.init(32, 32, 1)
        pc = PCBdpt.from_path_pc(sc.make_pc(4, True))
        with pytest.raises(RuntimeError, match="pixel shards"):
            dev.render_bdpt(pc, sc.make_ubo(), 0, 1)
        dev.set_pixel_shard(0, 1)
        dev.init(32, 32, 1)
        dev.render_bdpt(pc, sc.make_ubo(), 0, 1)  # the context stays usable
    finally:
        dev.close()


def test_cli_bdpt_equals_oracle_at_half_precision(tmp_path):
    """lumen_headless --integrator bdpt: BDPTB200 (lumen_b200/host/bdpt_b200.h) through init / create_accel / render / update /
    save_exr, with --time fixed so that the render is reproducible; the EXR equals the oracle's film at half precision."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "lumen_b200", "host", "lumen_headless")
    out = tmp_path / "bdpt.exr"
    r = subprocess.run([exe, scene_path("cornell"), "--integrator", "bdpt", "--time", "9", "--width", "80", "--height", "64", "--spp", "4", "--depth", "5",
                        "--batch", "2", "--out", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "BDPT depth 5" in r.stdout and "wrote" in r.stdout
    sc = host.Scene(scene_path("cornell"), 80, 64)
    pc = PCBdpt.from_path_pc(sc.make_pc(5, True), 9)
    cpu, _ = po.OracleScene(sc).render_bdpt(pc, sc.make_ubo(), 0, 4)
    want = cpu[..., :3].astype(np.float16).astype(np.float32)
    got = host.load_exr(str(out))[..., :3]
    close = np.abs(got - want) <= 2.0 ** -9 * np.maximum(np.abs(want), 1e-4)
    assert close.mean() > 0.999


def test_bdpt_sample_index_shards_sum_to_the_running_mean(device):
    """SURVEY.md 8e for the BDPT row: two contexts render the even / odd frames into LMB_FILM_SUM films (valid-sample count in
    alpha), lmb_film_add_from adds them, lmb_resolve divides -- the same image as one context's running mean, up to fp32 summation
    order (and the light tracer's float atomics)."""
    sc, pc, ubo = _setup(device, "cornell", 80, 5)
    device.clear_film()
    device.render_bdpt(pc, ubo, 0, 8)
    whole = device.download()
    other = integrator.Device(0)
    try:
        other.upload_scene(sc.desc)
        other.build_accel()
        other.init(80, 80, 1)
        device.clear_film()
        device.render_bdpt(pc, ubo, 0, 4, 2, integrator.FILM_SUM)
        other.render_bdpt(pc, ubo, 1, 4, 2, integrator.FILM_SUM)
        device.film_add_from(other)
        device.resolve()
        sharded = device.download()
    finally:
        other.close()
    assert pixel_agreement(sharded, whole, rel=1e-4) >= 0.999
    with pytest.raises(RuntimeError, match="frame_stride 1"):
        device.render_bdpt(pc, ubo, 0, 2, 2, integrator.FILM_RUNNING_MEAN)


def test_bdpt_scene_without_lights_matches_oracle(device):
    """No light buffer (num_lights = 0, light_triangle_count = 0): no light sub-path, no connection ray; material 0 made emissive
    after the scene was built, so that the s = 0 strategies (and the escaped-vertex rule, oracle/bdpt.h B3) return something."""
    import ctypes as C
    from helpers import MATERIALS, make_material
    from lumen_b200._ctypes_types import Material
    rng = np.random.default_rng(7)
    n_tri = 64
    v = np.zeros((n_tri * 3, 8), dtype=np.float32)
    v[:, :3] = rng.uniform(-1, 1, (n_tri * 3, 3))
    v[:, 3:6] = (0, 1, 0)
    sc = host.Scene.from_arrays(v, [n_tri], [0], [make_material(**MATERIALS["diffuse"])], width=48, height=48)
    assert sc.info.n_lights == 0
    mat0 = C.cast(sc.desc.materials, C.POINTER(Material))[0]  # both sides borrow / copy the scene's arrays after this patch
    mat0.emissive_factor[0], mat0.emissive_factor[1], mat0.emissive_factor[2] = 0.5, 0.25, 0.125
    device.upload_scene(sc.desc)
    device.build_accel()
    device.init(48, 48, 1)
    pc, ubo = PCBdpt.from_path_pc(sc.make_pc(4, True)), sc.make_ubo()
    osc = po.OracleScene(sc)
    device.reset_stats()
    col, splat = device.kat_bdpt_frame_raw(pc, ubo, 0)
    st = device.stats()
    ocol, osplat, ost = osc.render_bdpt_frame_raw(pc, ubo, 0)
    assert (st.rays_closest, st.rays_shadow) == (ost.rays_closest, ost.rays_shadow) and st.rays_shadow == 0
    assert bits_equal(col, ocol).all() and (splat == 0).all() and (osplat == 0).all()
    assert (ocol > 0).any()
    osc.close()
