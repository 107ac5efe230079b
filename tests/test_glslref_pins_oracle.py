"""The oracle against the reference ITSELF (CPU, no GPU): oracle/_ref/libglslref.so is the reference's own shader source
(src/shaders/integrators/path/path.rgen and everything it includes, ray.rchit, ray.rmiss, ray_shadow.rmiss), translated
mechanically from the unmodified GLSL by oracle/glslref/glsl2cpp.py and compiled; oracle/liboracle.so is the hand-written
restatement the CUDA path is tested against. Function by function and image by image the two must agree BIT FOR BIT.

What the comparison cannot see (it is outside the shader source and enters both sides through the same callbacks):
ray/triangle intersection, instance transforms, bilinear texture filtering -- the Vulkan driver's and the hardware's
part, which this build DEFINES (oracle/lbvh_cpu.h). Definitions the translation has to take because
GLSL leaves them open: transcendental functions = include/lmb_detmath.h, no FMA contraction, uninitialised variables
read as zero (-ftrivial-auto-var-init=zero; quirks Q1 and the thin-dielectric `f`, see test_uninitialised_reads_*).
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, scene_path
from helpers import MATERIALS, bits_equal, make_material, unit_vectors
from lumen_b200 import host
from oracle import pyglslref as pr
from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libglslref.so not built (make -C oracle/glslref; needs /root/reference)")

N = 120000  # random inputs per function and material


class Pair:
    def __init__(self, path, w, h):
        self.scene = host.Scene(path, w, h)
        self.orc = po.OracleScene(self.scene)
        self.ref = pr.RefScene(self.scene, self.orc)


@pytest.fixture(scope="module")
def cornell():
    return Pair(scene_path("cornell"), 160, 160)


@pytest.fixture(scope="module")
def classroom(tmp_path_factory):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scenes"))
    import gen_classroom_standin
    path, _ = gen_classroom_standin.generate(str(tmp_path_factory.mktemp("classroom_standin")))
    return Pair(path, 192, 108)


def test_rng(cornell):
    """utils.glsl:121-154: pcg4d, uint_to_float, rand / rand2 / rand3 / rand4 (left-to-right draws)."""
    rng = np.random.default_rng(101)
    v = rng.integers(0, 2**32, size=(N, 4), dtype=np.uint32)
    v[:8] = [[0, 0, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1919, 1079, 1023, 0], [2**32 - 1] * 4, [511, 511, 15, 0]]
    assert (cornell.ref.pcg4d(v) == po.pcg4d(v)).all()
    # the translated side draws through rand4 / rand3 / rand2 / rand in the groupings the shaders use; the oracle one by one
    assert bits_equal(cornell.ref.rand(v[:20000], 37), po.rand(v[:20000], 37)).all()


def test_offset_ray(cornell):
    """utils.glsl:73-94 incl. the |p| < 1/32 branch and negative coordinates."""
    rng = np.random.default_rng(102)
    p = rng.uniform(-20, 20, (N, 3)).astype(np.float32)
    p[::5] *= 1e-3
    p[::11, 1] = 0.0
    p[::13, 2] = -0.0
    n = unit_vectors(rng, N)
    ra, rb = cornell.ref.offset_ray(p, n)
    oa, ob = po.offset_ray(p, n)
    assert bits_equal(ra, oa).all() and bits_equal(rb, ob).all()


def bsdf_inputs(seed):
    rng = np.random.default_rng(seed)
    ns, wo, wi = unit_vectors(rng, N), unit_vectors(rng, N), unit_vectors(rng, N)
    wo[:100] = ns[:100]                    # normal incidence
    wi[100:200] = -wo[100:200]             # straight-through transmission
    wi[200:300] = wo[200:300]              # retro-reflection
    t = unit_vectors(rng, 100)
    wo[300:400] = np.cross(ns[300:400], t)  # grazing: wo perpendicular to n_s
    wo[300:400] /= np.linalg.norm(wo[300:400], axis=1, keepdims=True)
    rr = rng.uniform(0, 1, (N, 3)).astype(np.float32)
    rr[:50] = 0.0
    rr[50:60] = np.float32(1.0) - np.float32(2.0**-24)
    side = rng.integers(0, 2, N).astype(np.uint8)
    return ns.astype(np.float32), wo.astype(np.float32), wi.astype(np.float32), rr, side


@pytest.mark.parametrize("k,name", list(enumerate(sorted(MATERIALS))))
def test_bsdf_sample_eval_pdf(cornell, k, name):
    """bsdf_commons.glsl:26-183 dispatch + bsdf/{diffuse,mirror,glass,dielectric,conductor,principled}.glsl,
    sampling_commons.glsl, microfacet_commons.glsl: sample_bsdf (f, wi, pdf, cos), eval_bsdf (f, pdf), bsdf_pdf."""
    m = make_material(**MATERIALS[name])
    ns, wo, wi, rr, side = bsdf_inputs(200 + k)
    assert bits_equal(cornell.ref.sample_bsdf(m, ns, wo, rr, side), po.sample_bsdf(m, ns, wo, rr, side)).all()
    assert bits_equal(cornell.ref.eval_bsdf(m, ns, wo, wi, side), po.eval_bsdf(m, ns, wo, wi, side)).all()
    assert bits_equal(cornell.ref.bsdf_pdf(m, ns, wo, wi, side), po.bsdf_pdf(m, ns, wo, wi, side)).all()


def test_bsdf_random_materials(cornell):
    """Random parameter vectors for every bsdf_type (not only the hand-picked ones of helpers.MATERIALS)."""
    rng = np.random.default_rng(303)
    ns, wo, wi, rr, side = [a[:4000] for a in bsdf_inputs(304)]
    for trial in range(48):
        bt = [1, 2, 4, 8, 16, 32][trial % 6]
        u = lambda lo=0.0, hi=1.0: float(rng.uniform(lo, hi))  # noqa: E731
        m = make_material(albedo=(u(), u(), u()), ior=u(1.0, 2.2), k=(u(0, 4), u(0, 4), u(0, 4)), roughness=u() ** 2, bsdf_type=bt,
                          bsdf_props=int(rng.integers(0, 32)), metallic=u(), spec_trans=u(), specular_tint=u(), sheen_tint=u(), clearcoat=u(),
                          clearcoat_gloss=u(), sheen=u(), subsurface=u(), flatness=u(), anisotropy=u(), diffuse_trans=u(),
                          thin=int(rng.integers(0, 2)))
        assert bits_equal(cornell.ref.sample_bsdf(m, ns, wo, rr, side), po.sample_bsdf(m, ns, wo, rr, side)).all(), (trial, bt)
        assert bits_equal(cornell.ref.eval_bsdf(m, ns, wo, wi, side), po.eval_bsdf(m, ns, wo, wi, side)).all(), (trial, bt)
        assert bits_equal(cornell.ref.bsdf_pdf(m, ns, wo, wi, side), po.bsdf_pdf(m, ns, wo, wi, side)).all(), (trial, bt)


def test_uninitialised_reads_are_the_only_definitions_taken(cornell):
    """Q1 (dielectric.glsl:153,171): eval_dielectric's transmission branch returns the outer, never-written `vec3 f`; and
    sample_dielectric's thin branch multiplies a never-written `f` (dielectric.glsl:70,93-104). The literal GLSL leaves both
    undefined; the oracle and the translation read them as zero. Here: f == 0 exactly there, with the pdf still positive."""
    m = make_material(**MATERIALS["dielectric_rough"])
    ns = np.float32([[0, 0, 1]])
    wo = np.float32([[0.3, 0.1, 0.9486833]])
    wi = np.float32([[-0.2, -0.05, -0.9785193]])
    for out in (cornell.ref.eval_bsdf(m, ns, wo, wi, [1]), po.eval_bsdf(m, ns, wo, wi, [1])):
        assert (out[0, :3] == 0).all() and out[0, 3] > 0


def test_atmosphere(cornell):
    """commons.glsl:156-168 shade_atmosphere + atmosphere/atmosphere.glsl:47-204 (64 x 8 step march)."""
    rng = np.random.default_rng(404)
    o = rng.uniform(-8, 8, (6000, 3)).astype(np.float32)
    d = unit_vectors(rng, 6000)
    d[:10] = [0, 1, 0]
    d[10:20] = [0, -1, 0]
    d[20:30] = [1, 0, 0]
    for ld, L in (((0.48, 0.62, 0.62), (98.0, 82.0, 30.0)), ((0.0, 1.0, 0.0), (100.0, 100.0, 100.0)), ((-0.7, 0.1, 0.7), (50.0, 40.0, 30.0))):
        # the translated shade_atmosphere derives light_dir = -normalize(light.to - light.pos) itself (commons.glsl:161); the oracle's
        # probe takes the direction, so hand it glm::normalize's result: v * (1 / sqrt(dot(v, v))) in fp32
        v = np.float32(ld)
        dot = np.float32(np.float32(v[0] * v[0]) + np.float32(v[1] * v[1])) + np.float32(v[2] * v[2])
        ldn = v * (np.float32(1.0) / np.sqrt(np.float32(dot)))
        assert bits_equal(cornell.ref.atmosphere(o, d, ld, L), po.atmosphere(o, d, ldn, L)).all()


@pytest.mark.parametrize("name", ["cornell", "materials", "caustics", "cornell_dir"])
def test_light_sampling(name):
    """commons.glsl:112-149 sample_triangle, :196-222 sample_area_light, :224-300 sample_light_Li (area / spot / directional)
    and :335-406 sample_light_Le (BDPT's emission sampling), on the lights of the reference's own scenes."""
    p = Pair(scene_path(name), 32, 32)
    i = p.scene.info
    rng = np.random.default_rng(505)
    r4 = rng.uniform(0, 1, (N // 4, 4)).astype(np.float32)
    r4[:16] = 0.0
    pts = rng.uniform(-3, 3, (N // 4, 3)).astype(np.float32)
    a, b = p.ref.sample_light(i.n_lights, r4, pts), p.orc.sample_light(i.n_lights, r4, pts)
    assert bits_equal(a, b).all()
    r6 = rng.uniform(0, 1, (N // 4, 6)).astype(np.float32)
    a, b = p.ref.light_Le(i.n_lights, i.total_light_triangle_cnt, r6), p.orc.light_Le(i.n_lights, i.total_light_triangle_cnt, r6)
    assert bits_equal(a, b).all()


def test_load_material_with_textures():
    """bsdf_commons.glsl:16-22: albedo *= texture(...).xyz for textured materials (the texel fetch itself is the callback);
    cornell_box_dir.json is the reference scene with a texture (wood1.jpg)."""
    classroom = Pair(scene_path("cornell_dir"), 32, 32)
    sc = classroom.scene
    rng = np.random.default_rng(606)
    n = 20000
    idx = rng.integers(0, sc.info.n_materials, n).astype(np.uint32)
    uv = rng.uniform(-2, 3, (n, 2)).astype(np.float32)
    out = classroom.ref.load_material(idx, uv)
    mats = np.frombuffer(out, dtype=np.uint8).reshape(n, 104)
    src = np.frombuffer((C.c_uint8 * (104 * sc.info.n_materials)).from_address(sc.desc.materials), dtype=np.uint8).reshape(-1, 104)
    tex_id = src[:, 48:52].copy().view(np.int32)[:, 0]
    assert (tex_id > -1).any(), "cornell_box_dir has a textured material"
    for mi in np.unique(idx):
        sel = idx == mi
        got = mats[sel]
        assert (got[:, 12:] == src[mi, 12:]).all()  # everything but the albedo is the stored record
        alb = got[:, :12].copy().view(np.float32)
        base = src[mi, :12].copy().view(np.float32)
        if tex_id[mi] > -1:
            want = base[None, :] * classroom.orc.texture(int(tex_id[mi]), uv[sel])
        else:
            want = np.broadcast_to(base, alb.shape)
        assert bits_equal(alb, np.ascontiguousarray(want, dtype=np.float32)).all()


RENDERS = [
    # scene, W, H, max_depth, frames: the reference's own scene files (BASELINE configs 1-2 + the two extra correctness scenes)
    ("cornell", 160, 160, 6, 6),
    ("caustics", 192, 108, 12, 6),
    ("materials", 160, 160, 10, 6),
    ("cornell_dir", 128, 128, 6, 4),
]


@pytest.mark.parametrize("name,w,h,depth,frames", RENDERS)
def test_path_rgen_film_is_bit_equal(name, w, h, depth, frames):
    """The whole pipeline: path.rgen main() per pixel and frame through traceRayEXT -> ray.rchit / ray.rmiss /
    ray_shadow.rmiss, running-mean film (path.rgen:102-112), against orc_render. Ray counts per type must match too."""
    p = Pair(scene_path(name), w, h)
    pc, ubo = p.scene.make_pc(depth, True), p.scene.make_ubo()
    ref, rays = p.ref.render(pc, ubo, 0, frames)
    orc, st = p.orc.render(pc, ubo, 0, frames)
    assert [int(r) for r in rays] == [st.rays_closest, st.rays_shadow, st.rays_probe]
    assert bits_equal(ref, orc).all()
    assert np.isfinite(orc[..., :3]).all() and orc[..., :3].max() > 0


def test_path_rgen_direct_lighting_off_and_late_frames(cornell):
    """pc.direct_lighting = 0 (path.rgen:51,58,77) and frame numbers far from 0 (seed, running-mean weight 1/(frame+1))."""
    pc, ubo = cornell.scene.make_pc(6, False), cornell.scene.make_ubo()
    ref, _ = cornell.ref.render(pc, ubo, 0, 2)
    orc, _ = cornell.orc.render(pc, ubo, 0, 2)
    assert bits_equal(ref, orc).all()
    pc = cornell.scene.make_pc(6, True)
    start = np.random.default_rng(7).uniform(0, 1, (160, 160, 4)).astype(np.float32)
    a, b = start.copy(), start.copy()
    cornell.ref.render(pc, ubo, 1022, 2, rgba=a)
    cornell.orc.render(pc, ubo, 1022, 2, rgba=b)
    assert bits_equal(a, b).all() and not bits_equal(a, start).all()


def test_path_rgen_classroom_standin(classroom):
    """Config 3's code path: Mitsuba loader output, sun + sky (atmosphere on every escaped ray), textures, principled + conductor."""
    pc, ubo = classroom.scene.make_pc(8, True), classroom.scene.make_ubo()
    ref, rays = classroom.ref.render(pc, ubo, 0, 2)
    orc, st = classroom.orc.render(pc, ubo, 0, 2)
    assert [int(r) for r in rays] == [st.rays_closest, st.rays_shadow, st.rays_probe]
    assert bits_equal(ref, orc).all()


BDPT_RENDERS = [("cornell", 96, 72, 6), ("caustics", 96, 54, 8), ("materials", 80, 80, 6), ("cornell_dir", 72, 72, 5)]


@pytest.mark.parametrize("name,w,h,depth", BDPT_RENDERS)
def test_bdpt_rgen_matches_the_bdpt_oracle(name, w, h, depth):
    """bdpt.rgen + integrators/bdpt_commons.glsl (light / eye walks, calc_mis_weight, bdpt_connect_cam,
    bdpt_connect) translated from the unmodified GLSL, one dispatch per frame, against oracle/bdpt.h. The GLSL's cross-pixel splat
    (`tmp_col.d[idx] += splat`, non-atomic, read and cleared by other invocations of the same dispatch) makes the reference's own
    per-pixel value depend on scheduling, so each translated invocation runs on a private colour storage: it stores (own strategies +
    the splats it sends to itself) and the splats it sends elsewhere are harvested per target pixel. Then
      * every traceRayEXT the reference issues, the oracle issues (closest and any-hit counts equal);
      * on pixels no splat lands on, the stored value IS the oracle's own-strategy radiance, bit for bit;
      * everywhere, stored + harvested = the oracle's col + splat to fp32 summation order (1e-4 relative)."""
    from lumen_b200._ctypes_types import PCBdpt
    p = Pair(scene_path(name), w, h)
    pc, ubo = PCBdpt.from_path_pc(p.scene.make_pc(depth, True), 1234567), p.scene.make_ubo()
    for frame in (0, 5):
        img, spl, rays = p.ref.render_bdpt_frame(pc, ubo, frame)
        col, osp, st = p.orc.render_bdpt_frame_raw(pc, ubo, frame)
        assert [int(rays[0]), int(rays[1])] == [st.rays_closest, st.rays_shadow] and rays[2] == 0
        quiet = (osp == 0).all(axis=2) & (spl == 0).all(axis=2)
        assert quiet.mean() > 0.3
        assert bits_equal(img[..., :3], col).all(axis=2)[quiet].all()
        got, want = img[..., :3] + spl, col + osp
        ok = (np.abs(got - want) <= 1e-4 * np.maximum(np.abs(want), 1e-6)) | (np.isnan(got) & np.isnan(want))
        assert ok.all()
        assert np.nanmax(osp) > 0 and np.nanmax(col) > 0  # both kinds of strategies are exercised
