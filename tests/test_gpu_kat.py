"""Function-level parity: every device function of the CUDA path against the CPU oracle, bit for bit, through the C-ABI
test hooks (include/lumen_b200_testhooks.h). Mirrors SURVEY.md section 4 item (1)."""
import numpy as np
import pytest

from helpers import MATERIALS, bits_equal, make_material, unit_vectors
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def test_pcg4d_and_rand_stream(device):
    rng = np.random.default_rng(11)
    v = rng.integers(0, 2**32, size=(4096, 4), dtype=np.uint32)
    v[:8] = [[0, 0, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1919, 1079, 1023, 0], [2**32 - 1] * 4, [511, 511, 15, 0]]
    assert (device.kat_pcg4d(v) == po.pcg4d(v)).all()
    assert bits_equal(device.kat_rand(v, 24), po.rand(v, 24)).all()


def test_detmath_bit_exact(device):
    rng = np.random.default_rng(12)
    x = np.concatenate([rng.uniform(-10, 10, 100000), rng.uniform(-150, 100, 100000), [0.0, -0.0, 1.0, 88.7, -88.0, 1e-30]]).astype(np.float32)
    y = rng.uniform(0, 9, x.size).astype(np.float32)
    g, c = device.kat_detmath(x, y), po.detmath(x, y)
    for k in ("sin", "cos", "exp", "pow"):
        assert bits_equal(g[k], c[k]).all(), k


def test_offset_ray(device):
    rng = np.random.default_rng(13)
    p = rng.uniform(-20, 20, (50000, 3)).astype(np.float32)
    p[::5] *= 1e-3  # exercise the |p| < 1/32 branch of utils.glsl:81-83
    p[::11, 1] = 0.0
    n = unit_vectors(rng, 50000)
    ga, gb = device.kat_offset_ray(p, n)
    ca, cb = po.offset_ray(p, n)
    assert bits_equal(ga, ca).all() and bits_equal(gb, cb).all()


@pytest.mark.parametrize("name", sorted(MATERIALS))
def test_bsdf_sample_and_eval(device, name):
    rng = np.random.default_rng(1000 + sorted(MATERIALS).index(name))  # fixed per material: a failure reproduces
    m = make_material(**MATERIALS[name])
    n = 60000
    ns, wo, wi = unit_vectors(rng, n), unit_vectors(rng, n), unit_vectors(rng, n)
    # grazing and exactly-aligned configurations
    wo[:100] = ns[:100]
    wi[100:200] = -wo[100:200]
    rr = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    rr[:50] = 0.0
    side = rng.integers(0, 2, n).astype(np.uint8)
    assert bits_equal(device.kat_sample_bsdf(m, ns, wo, rr, side), po.sample_bsdf(m, ns, wo, rr, side)).all()
    assert bits_equal(device.kat_eval_bsdf(m, ns, wo, wi, side), po.eval_bsdf(m, ns, wo, wi, side)).all()


def test_quirk_q1_eval_dielectric_transmission_is_zero(device):
    """Q1 (dielectric.glsl:153,171): the transmission branch returns the shadowed outer f, frozen to 0, with pdf > 0."""
    m = make_material(**MATERIALS["dielectric_rough"])
    ns = np.float32([[0, 0, 1]])
    wo = np.float32([[0.3, 0.1, 0.9486833]])
    wi = np.float32([[-0.2, -0.05, -0.9785193]])
    for out in (device.kat_eval_bsdf(m, ns, wo, wi, [1]), po.eval_bsdf(m, ns, wo, wi, [1])):
        assert (out[0, :3] == 0).all() and out[0, 3] > 0


def test_quirk_q8_unknown_bsdf_type_kills_path(device):
    """Q8: "disney" is not a recognised type -> bsdf_type 0 -> sample_bsdf returns pdf 0 (path.rgen:85-87 breaks)."""
    m = make_material(**MATERIALS["unknown_type"])
    rng = np.random.default_rng(5)
    ns, wo = unit_vectors(rng, 64), unit_vectors(rng, 64)
    out = device.kat_sample_bsdf(m, ns, wo, rng.uniform(0, 1, (64, 3)), np.ones(64, np.uint8))
    assert (out[:, 6] == 0).all() and (out[:, :3] == 0).all()


def test_atmosphere(device):
    rng = np.random.default_rng(14)
    o = rng.uniform(-8, 8, (4000, 3)).astype(np.float32)
    d = unit_vectors(rng, 4000)
    d[:10] = [0, 1, 0]
    d[10:20] = [0, -1, 0]
    g = device.kat_atmosphere(o, d, (0.48, 0.62, 0.62), (98.0, 82.0, 30.0))
    c = po.atmosphere(o, d, (0.48, 0.62, 0.62), (98.0, 82.0, 30.0))
    assert bits_equal(g, c).all()
    assert np.isfinite(c[:10]).all() and (c[:10] > 0).all()
