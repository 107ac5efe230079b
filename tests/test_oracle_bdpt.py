"""CPU tests of the BDPT restatement (oracle/bdpt.h: bdpt.rgen + bdpt_commons.glsl) and of the stand-alone bsdf_pdf functions
(bsdf_commons.glsl:26-66). The reference ships no test for either; what is checked here is internal consistency
(two independent routes to the same density), determinism, the frozen quirks, agreement with the Path oracle where both are
unbiased, and the committed golden vectors."""
import os

import numpy as np
import pytest

from conftest import scene_path
from helpers import MATERIALS, bits_equal, make_material, unit_vectors
from lumen_b200 import host
from lumen_b200._ctypes_types import PCBdpt
from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", list(MATERIALS))
def test_bsdf_pdf_equals_the_pdf_eval_bsdf_reports(name):
    """*_pdf (diffuse.glsl:85, dielectric.glsl:191, conductor.glsl:74, principled.glsl:338) and the pdf_w that eval_* writes
    (diffuse.glsl:58, dielectric.glsl:110, conductor.glsl:45, principled.glsl:289) are separate GLSL functions meant to return
    the same density; the two restatements agree to 1e-4 relative for every material, side and direction pair."""
    rng = np.random.default_rng(11)
    n = 3000
    mat = make_material(**MATERIALS[name])
    n_s, wo, wi = unit_vectors(rng, n), unit_vectors(rng, n), unit_vectors(rng, n)
    for side in (1, 0):
        sd = np.full(n, side, np.uint8)
        p = po.bsdf_pdf(mat, n_s, wo, wi, sd)
        e = po.eval_bsdf(mat, n_s, wo, wi, sd)[:, 3]
        big = np.maximum(np.abs(p), np.abs(e))
        assert ((np.abs(p - e) <= 1e-4 * big) | (np.isnan(p) & np.isnan(e))).all()


@pytest.mark.parametrize("name", ["diffuse", "dielectric_rough", "conductor_rough", "dielectric_reflect_only"])
def test_bsdf_pdf_equals_the_sampling_pdf(name):
    """For the single-lobe materials the density sample_bsdf reports for its own direction is bsdf_pdf of that direction (the
    few percent that differ are directions where sample_* and *_pdf reject on different sides of a grazing test)."""
    rng = np.random.default_rng(12)
    n = 3000
    mat = make_material(**MATERIALS[name])
    n_s, wo = unit_vectors(rng, n), unit_vectors(rng, n)
    wo[np.sum(n_s * wo, axis=1) < 0] *= -1
    s = po.sample_bsdf(mat, n_s, wo, rng.uniform(0, 1, (n, 3)).astype(np.float32), np.ones(n, np.uint8))
    wi, pdf = s[:, 3:6], s[:, 6]
    p = po.bsdf_pdf(mat, n_s, wo, wi, np.ones(n, np.uint8))
    ok = pdf > 0
    assert ok.mean() > 0.9
    assert (np.abs(p - pdf)[ok] <= 1e-3 * pdf[ok]).mean() > 0.97


def _bdpt(name, size, depth, time=0):
    sc = host.Scene(scene_path(name), size, size)
    return sc, po.OracleScene(sc), PCBdpt.from_path_pc(sc.make_pc(depth, True), time), sc.make_ubo()


def test_bdpt_is_deterministic_in_the_thread_count():
    """Splats are applied in source-pixel order after the parallel loop (quirk B2), so the image does not depend on scheduling."""
    sc, orc, pc, ubo = _bdpt("cornell", 48, 5)
    a, sa, st_a = orc.render_bdpt_frame_raw(pc, ubo, 1, threads=1)
    b, sb, st_b = orc.render_bdpt_frame_raw(pc, ubo, 1, threads=0)
    assert bits_equal(a, b).all() and bits_equal(sa, sb).all()
    assert (st_a.rays_closest, st_a.rays_shadow) == (st_b.rays_closest, st_b.rays_shadow)
    assert (sa.sum(-1) > 0).mean() > 0.2  # the light tracer reaches a good part of the image


def test_bdpt_seed_is_frame_xor_time():
    """bdpt.rgen:36-37 (quirk B1: `time` is an input)."""
    sc, orc, pc, ubo = _bdpt("cornell", 32, 4)
    a, _, _ = orc.render_bdpt_frame_raw(pc, ubo, 5)
    pc1 = PCBdpt.from_path_pc(sc.make_pc(4, True), 1)
    b, _, _ = orc.render_bdpt_frame_raw(pc1, ubo, 4)
    c, _, _ = orc.render_bdpt_frame_raw(pc1, ubo, 5)
    assert bits_equal(a, b).all() and not bits_equal(a, c).all()


def test_bdpt_escaped_vertex_takes_material_zero():
    """Quirk B3: the camera walk's escaped vertex keeps material_idx 0 of the zeroed buffer, so the s = 0 strategy adds material 0's
    emissive_factor for a primary ray that leaves the scene (the Path integrator adds the sky colour there)."""
    import ctypes as C
    from lumen_b200._ctypes_types import Material
    sc, orc, pc, ubo = _bdpt("cornell", 32, 4)
    inv_view = np.array(list(ubo.inv_view), dtype=np.float64).reshape(4, 4).T  # column-major -> numpy
    inv_proj = np.array(list(ubo.inv_projection), dtype=np.float64).reshape(4, 4).T
    xs, ys = np.meshgrid(np.arange(32) + 0.5, np.arange(32) + 0.5)
    d = np.stack([xs / 32 * 2 - 1, ys / 32 * 2 - 1, np.ones_like(xs), np.ones_like(xs)], axis=-1)  # bdpt.rgen:40-46: pixel centres
    target = (d @ inv_proj.T)[..., :3]
    target /= np.linalg.norm(target, axis=-1, keepdims=True)
    dirs = (np.concatenate([target, np.zeros_like(xs)[..., None]], axis=-1) @ inv_view.T)[..., :3]
    origin = inv_view[:3, 3]
    rays = np.concatenate([np.broadcast_to(origin, dirs.shape), np.full(xs.shape + (1,), 1e-3), dirs, np.full(xs.shape + (1,), 1e6)], axis=-1)
    hits, _ = orc.trace_closest(rays.reshape(-1, 8))
    escaped = (hits["prim"] == 0xFFFFFFFF).reshape(32, 32)
    assert escaped.sum() > 20
    mat0 = C.cast(sc.desc.materials, C.POINTER(Material))[0]  # the oracle borrows the scene's arrays: patch in place
    assert tuple(mat0.emissive_factor) == (0.0, 0.0, 0.0)
    mat0.emissive_factor[0], mat0.emissive_factor[1], mat0.emissive_factor[2] = 0.25, 0.5, 0.75
    col, _, _ = orc.render_bdpt_frame_raw(pc, ubo, 0)
    assert (col[escaped] == np.array([0.25, 0.5, 0.75], dtype=np.float32)).all()


def test_bdpt_single_strategies_converge_to_each_other():
    """The camera walk and next-event estimation are each unbiased on their own: with weight 1, the s = 0 strategies (emission
    found by BSDF sampling) and the s = 1 strategies (light sampling) must converge to the same image wherever the emitter is
    not directly visible. This checks the walk, the throughputs and the light pdfs of the restatement without trusting the
    reference's MIS weights."""
    sc, orc, pc, ubo = _bdpt("cornell", 64, 3)
    try:
        po.bdpt_set_only_s(0)
        a, _ = orc.render_bdpt(pc, ubo, 0, 128)
        po.bdpt_set_only_s(1)
        b, _ = orc.render_bdpt(pc, ubo, 0, 128)
    finally:
        po.bdpt_set_only_s(-1)
    lower = slice(24, 64)  # rows below the light
    ma, mb = a[lower, :, :3].mean(), b[lower, :, :3].mean()
    assert mb > 0.05 and abs(ma / mb - 1) < 0.05, (ma, mb)


def test_bdpt_is_close_to_path_on_direct_light():
    """At max_depth 2 both integrators estimate emission + direct light on what the camera sees. Away from the emitter (whose
    pixels carry the Path integrator's stale-payload term, pt_commons.glsl:33) the images agree to a few percent -- the
    reference's BDPT weights do not sum to exactly one, so this is a sanity bound, not an equality."""
    sc, orc, pc, ubo = _bdpt("cornell", 64, 2)
    b, _ = orc.render_bdpt(pc, ubo, 0, 96)
    p, _ = orc.render(sc.make_pc(2, True), ubo, 0, 96)
    lower = slice(24, 64)
    mb, mp = b[lower, :, :3].mean(), p[lower, :, :3].mean()
    assert abs(mb / mp - 1) < 0.08, (mb, mp)


def test_bdpt_film_is_the_running_mean_of_col_plus_splat():
    sc, orc, pc, ubo = _bdpt("cornell", 32, 4)
    film, st = orc.render_bdpt(pc, ubo, 0, 3)
    acc = None
    for f in range(3):
        col, splat, _ = orc.render_bdpt_frame_raw(pc, ubo, f)
        c = col + splat
        acc = c if f == 0 else acc * np.float32(1 - np.float32(1.0) / np.float32(f + 1)) + c * (np.float32(1.0) / np.float32(f + 1))
    assert st.nan_pixels == 0
    assert bits_equal(film[..., :3], acc.astype(np.float32)).all() and (film[..., 3] == 1).all()


def test_bdpt_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "oracle_bdpt_golden.npz"))
    for name in ("cornell", "materials", "caustics", "cornell_dir"):
        size, depth, time, frame = (int(v) for v in g[f"{name}_cfg"])
        sc, orc, pc, ubo = _bdpt(name, size, depth, time)
        col, splat, st = orc.render_bdpt_frame_raw(pc, ubo, frame)
        assert bits_equal(col, g[f"{name}_col"]).all(), name
        assert bits_equal(splat, g[f"{name}_splat"]).all(), name
        assert (st.rays_closest, st.rays_shadow) == tuple(int(v) for v in g[f"{name}_rays"])
        orc.close()


def test_bdpt_on_a_scene_without_lights():
    """num_lights = 0, light_triangle_count = 0 (LumenScene.cpp:134 creates no light buffer): sample_light_Le reads the all-zero
    Light, pdf_dir = 0, so there is no light sub-path and no connection ray; only the s = 0 strategies run. Must not crash, must
    be deterministic, and traces exactly the camera walk."""
    rng = np.random.default_rng(7)
    n_tri = 64
    v = np.zeros((n_tri * 3, 8), dtype=np.float32)
    v[:, :3] = rng.uniform(-1, 1, (n_tri * 3, 3))
    v[:, 3:6] = (0, 1, 0)
    mats = [make_material(**MATERIALS["diffuse"])]  # not emissive: an emissive material would make the mesh an area light (LumenScene.cpp:83-95)
    sc = host.Scene.from_arrays(v, [n_tri], [0], mats, width=32, height=32)
    assert sc.info.n_lights == 0
    orc = po.OracleScene(sc)
    pc, ubo = PCBdpt.from_path_pc(sc.make_pc(4, True)), sc.make_ubo()
    a, sa, st = orc.render_bdpt_frame_raw(pc, ubo, 0)
    b, sb, _ = orc.render_bdpt_frame_raw(pc, ubo, 0, threads=1)
    assert bits_equal(a, b).all() and bits_equal(sa, sb).all()
    assert st.rays_shadow == 0 and st.rays_closest >= 32 * 32
    assert (sa == 0).all() and (a == 0).all()  # nothing emits: every s = 0 strategy multiplies a zero emissive_factor


def test_calc_mis_weight_restores_every_vertex_it_patches():
    """bdpt_commons.glsl:311-340 patches up to eight vertex fields in place and :441-466 puts them back. The CUDA path evaluates
    the (s, t) pairs of a pixel in parallel (k_bdpt_pair) and substitutes the patched values in registers, which is only the
    same computation if the GLSL's restore is complete -- checked here byte for byte after every call, on all four scenes. The
    same switch checks the kernels' second assumption: the rand4 of the s = 1 strategy of camera vertex t is the (t - 2)-th
    draw after the walks."""
    po.bdpt_check_restore(True)
    try:
        before = po.bdpt_restore_violations()
        for name, depth in (("cornell", 6), ("caustics", 8), ("materials", 7), ("cornell_dir", 5)):
            sc, orc, pc, ubo = _bdpt(name, 40, depth)
            orc.render_bdpt_frame_raw(pc, ubo, 0)
            orc.close()
        assert po.bdpt_restore_violations() == before
    finally:
        po.bdpt_check_restore(False)


def _lights(sc):
    import ctypes as C
    from lumen_b200._ctypes_types import Light
    return C.cast(sc.desc.lights, C.POINTER(Light))


def test_light_emission_sampling_spot_against_numpy():
    """sample_light_Le, LIGHT_SPOT (commons.glsl:362-385): uniform cone of 30 degrees around the light's axis, rotated out of the
    local frame by the inverse of to_local_quat (utils.glsl:157-173). Independent fp64 restatement."""
    sc, orc, pc, ubo = _bdpt("caustics", 16, 4)
    L = _lights(sc)[0]
    assert L.light_flags & 7 == 1
    rng = np.random.default_rng(21)
    r = rng.uniform(0, 1, (4000, 6)).astype(np.float32)
    out = orc.light_Le(pc.num_lights, pc.light_triangle_count, r)
    pos, to = np.array(list(L.pos), np.float64), np.array(list(L.to), np.float64)
    axis = (to - pos) / np.linalg.norm(to - pos)
    cw = np.cos(np.float32(30 * np.float32(3.14159265359) / 180), dtype=np.float64)
    u, v = r[:, 4].astype(np.float64), r[:, 5].astype(np.float64)
    ct = (1 - u) + u * cw
    st = np.sqrt(1 - ct * ct)
    phi = v * 6.28318530718
    local = np.stack([np.cos(phi) * st, np.sin(phi) * st, ct], axis=1)
    q = np.array([axis[1], -axis[0], 0.0, 1.0 + axis[2]])
    q /= np.linalg.norm(q)
    qi = np.array([-q[0], -q[1], -q[2], q[3]])  # invert_quat
    qa = qi[:3]
    want = 2 * (local @ qa)[:, None] * qa + (qi[3] ** 2 - qa @ qa) * local + 2 * qi[3] * np.cross(qa, local)
    assert np.abs(out[:, 6:9] - want).max() < 5e-6  # fp32 detmath sin / cos against fp64
    assert np.abs(out[:, 12] - ct).max() < 5e-6  # cos_from_light = the sampled cos(theta): the rotation maps z onto the axis
    assert (out[:, 3:6] == np.array(list(L.pos), np.float32)).all()
    assert np.allclose(out[:, 14], 1 / (6.28318530718 * (1 - cw)), rtol=1e-6) and (out[:, 13] == 1.0 / pc.light_triangle_count).all()
    assert bits_equal(out[:, 9:12], out[:, 6:9]).all()  # n = wi


def test_light_emission_sampling_directional_and_area_properties():
    """LIGHT_DIRECTIONAL (commons.glsl:386-400): a point of the disk of radius world_radius that faces the scene, direction -dir,
    pdf_pos = 1 / (pi r^2) / total. LIGHT_AREA (:347-361): cosine-weighted direction about the sampled normal, pdf_dir = cos / pi,
    pdf_pos = 1 / (triangle area) / total."""
    sc, orc, pc, ubo = _bdpt("cornell_dir", 16, 4)
    L = _lights(sc)[0]
    assert L.light_flags & 7 == 3
    rng = np.random.default_rng(22)
    out = orc.light_Le(pc.num_lights, pc.light_triangle_count, rng.uniform(0, 1, (4000, 6)).astype(np.float32))
    d = np.array(list(L.pos), np.float64) - np.array(list(L.to), np.float64)
    d /= np.linalg.norm(d)  # -normalize(to - pos)
    assert np.abs(out[:, 6:9] + d).max() < 1e-6  # wi = -dir
    rel = out[:, 3:6].astype(np.float64) - (np.array(list(L.world_center), np.float64) + d * L.world_radius)
    assert np.abs(rel @ d).max() < 1e-4 * L.world_radius and np.linalg.norm(rel, axis=1).max() <= L.world_radius * (1 + 1e-5)
    assert np.allclose(out[:, 13], 1 / (np.pi * L.world_radius ** 2) / pc.light_triangle_count, rtol=1e-5)
    assert (out[:, 14] == 1).all() and (out[:, 12] == 1).all()

    sc, orc, pc, ubo = _bdpt("cornell", 16, 4)
    assert _lights(sc)[0].light_flags & 7 == 2
    out = orc.light_Le(pc.num_lights, pc.light_triangle_count, rng.uniform(0, 1, (4000, 6)).astype(np.float32))
    wi, n = out[:, 6:9].astype(np.float64), out[:, 9:12].astype(np.float64)
    cos = np.sum(wi * n, axis=1)
    assert (cos >= -1e-6).all() and np.abs(np.linalg.norm(wi, axis=1) - 1).max() < 1e-5
    assert np.allclose(out[:, 14], cos / np.pi, atol=1e-6) and np.allclose(out[:, 12], np.maximum(cos, 0), atol=1e-6)
    areas = 1 / (out[:, 13].astype(np.float64) * pc.light_triangle_count)
    assert (areas > 0).all() and areas.max() <= pc.total_light_area and len(np.unique(np.round(areas, 4))) <= _lights(sc)[0].num_triangles
    # a cosine-weighted hemisphere has E[cos] = 2/3
    assert abs(cos.mean() - 2 / 3) < 0.02


def test_mis_weights_add_up_to_one_in_practice():
    """calc_mis_weight (bdpt_commons.glsl:288-470) is the most intricate function of the restatement. If its weights form a
    partition of unity over the strategies of a path, the full estimator converges to what each single strategy converges to
    with weight 1; a slip in an index or a pdf would bias it. Cornell box, depth 5, rows below the light: within 4 %."""
    sc, orc, pc, ubo = _bdpt("cornell", 64, 5)
    lower = slice(24, 64)
    means = {}
    try:
        for only_s in (-1, 0, 1):
            po.bdpt_set_only_s(only_s)
            img, _ = orc.render_bdpt(pc, ubo, 0, 128)
            means[only_s] = float(img[lower, :, :3].mean())
    finally:
        po.bdpt_set_only_s(-1)
    assert abs(means[-1] / means[1] - 1) < 0.04 and abs(means[-1] / means[0] - 1) < 0.04, means
