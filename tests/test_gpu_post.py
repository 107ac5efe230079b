"""Device-side steps either side of the integrator, through the C ABI: EXR half planes,
Lumen's RMSE routine, checkpoint / resume, and the device-built per-triangle tables."""
import numpy as np
import pytest

from conftest import scene_path
from helpers import bits_equal
from lumen_b200 import host, integrator
from oracle import pyoracle as po
from test_post_cpu import _special_floats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cornell(device):
    sc = host.Scene(scene_path("cornell"), 200, 120)
    device.upload_scene(sc.desc)
    device.build_accel()
    return sc


def _render(device, sc, w, h, first, n, depth=6):
    pc, ubo = sc.make_pc(depth, True), sc.make_ubo()
    device.render(pc, ubo, first, n, 1, integrator.FILM_RUNNING_MEAN)


@pytest.mark.parametrize("w,h", [(200, 120), (37, 29), (1, 1)])
def test_half_planes_equal_tinyexr_conversion(device, cornell, w, h, tmp_path):
    """lmb_download_half_bgr == the oracle's (tinyexr-pinned) conversion of the film, bit for bit, on every branch of the
    conversion (film filled with special values) and on odd sizes (unaligned plane starts); the EXR written from the planes
    is byte-identical to the one written from the float film."""
    device.init(w, h, 1)
    rng = np.random.default_rng(w * 1000 + h)
    vals = _special_floats(rng, 3000)
    film = np.zeros((h, w, 4), np.float32)
    flat = rng.choice(vals, size=h * w * 3).astype(np.float32)
    film[..., :3] = flat.reshape(h, w, 3)
    device.upload_film(film)
    planes = device.download_half_bgr()
    want = np.stack([po.float_to_half(film[..., c]) for c in (2, 1, 0)])
    assert planes.shape == (3, h, w)
    assert (planes == want).all()
    a, b = str(tmp_path / "f.exr"), str(tmp_path / "h.exr")
    host.save_exr(device.download(), a)
    host.save_exr_half_bgr(planes, b)
    assert open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.parametrize("w,h", [(200, 120), (37, 29), (1920, 1080), (1030, 1)])
def test_rmse_routine_matches_oracle(device, cornell, w, h):
    """lmb_rmse: Lumen's literal RMSE routine bit-equal to the oracle's restatement (tail workgroups, two reduce levels at
    1080p: 2025 -> 2 -> 1 groups), true RMSE within 1e-9 relative (fp64 sum of 6 M terms, summation order differs)."""
    device.init(w, h, 1)
    rng = np.random.default_rng(h)
    film = np.exp(rng.uniform(-6, 3, (h, w, 4))).astype(np.float32)
    gt = (film * (1 + 0.1 * rng.normal(size=film.shape))).astype(np.float32)
    device.upload_film(film)
    with pytest.raises(RuntimeError):
        device.rmse()  # no reference image yet (has_gt)
    device.set_reference_image(gt)
    lit, tru = device.rmse()
    want_lit, want_true = po.rmse_literal(gt, film), po.rmse_true(gt, film)
    assert np.float32(lit).view(np.uint32) == np.float32(want_lit).view(np.uint32)
    assert abs(tru - want_true) <= 1e-9 * want_true
    device.set_reference_image(film)
    assert device.rmse() == (0.0, 0.0)


def test_rmse_of_a_render_against_oracle_image(device, cornell):
    """The progressive-service use: RMSE of the GPU film against a ground-truth image (here the oracle's render of other
    frames), literal routine bit-equal to the oracle's arithmetic on the same two images."""
    w, h = 200, 120
    sc = cornell
    device.init(w, h, 4)
    _render(device, sc, w, h, 0, 8)
    film = device.download()
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    gt, _ = po.OracleScene(sc).render(pc, ubo, 8, 8)
    device.set_reference_image(gt)
    lit, tru = device.rmse()
    assert np.float32(lit).view(np.uint32) == np.float32(po.rmse_literal(gt, film)).view(np.uint32)
    assert tru > 0 and abs(tru - po.rmse_true(gt, film)) <= 1e-9 * tru


def test_checkpoint_resume_is_bit_identical(device, cornell, tmp_path):
    """4 frames -> checkpoint file -> fresh film -> resume -> 4 more frames == 8 frames uninterrupted."""
    w, h = 200, 120
    sc = cornell
    device.init(w, h, 2)
    _render(device, sc, w, h, 0, 8)
    straight = device.download().copy()
    device.init(w, h, 2)
    _render(device, sc, w, h, 0, 4)
    p = str(tmp_path / "c.ckpt")
    host.save_checkpoint(p, device.download(), frames=4, path_length=6)
    device.init(w, h, 3)  # new film, different batching
    film, frames, depth = host.load_checkpoint(p)
    assert (frames, depth) == (4, 6)
    device.upload_film(film)
    _render(device, sc, w, h, frames, 4)
    assert device.download().tobytes() == straight.tobytes()


def test_async_download_equals_sync(device, cornell):
    import ctypes as C
    w, h = 200, 120
    device.init(w, h, 2)
    _render(device, cornell, w, h, 0, 2)
    want = device.download()
    got = np.zeros_like(want)
    device.download_async(got.ctypes.data)
    device.sync()
    assert got.tobytes() == want.tobytes()


def test_device_built_triangle_tables_multi_mesh(device):
    """The per-triangle tables (mesh, local id, vertex record, shade queue) are built on the device from PrimMeshInfo[] +
    indices[]: closest-hit records and a render over a many-mesh scene (material_test: 18 shapes, all BSDF types) equal the
    oracle's, which derives the same tables on the host."""
    sc = host.Scene(scene_path("materials"), 96, 96)
    orc = po.OracleScene(sc)
    device.upload_scene(sc.desc)
    device.build_accel()
    device.init(96, 96, 2)
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    device.reset_stats()
    device.render(pc, ubo, 0, 2, 1, integrator.FILM_RUNNING_MEAN)
    gpu, st = device.download(), device.stats()
    cpu, cst = orc.render(pc, ubo, 0, 2)
    assert st.rays == cst.rays
    assert bits_equal(gpu[..., :3], cpu[..., :3]).all(axis=-1).mean() >= 0.999
