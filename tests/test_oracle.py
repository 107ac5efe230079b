"""CPU checks of the oracle itself: structure of the canonical LBVH, traversal against brute force, RNG against an
independent numpy restatement, determinism, the committed golden vectors, and the reference's RMSE arithmetic."""
import os

import numpy as np
import pytest

from conftest import scene_path
from helpers import bits_equal, random_rays
from lumen_b200 import host
from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def cornell():
    sc = host.Scene(scene_path("cornell"), 64, 64)
    return sc, po.OracleScene(sc)


def pcg4d_numpy(v):
    """Independent restatement of utils.glsl:121-134 with numpy uint32 wraparound."""
    v = v.astype(np.uint32).copy()
    with np.errstate(over="ignore"):
        v = v * np.uint32(1664525) + np.uint32(1013904223)
        x, y, z, w = v[:, 0], v[:, 1], v[:, 2], v[:, 3]
        x += y * w; y += z * x; z += x * y; w += y * z
        x ^= x >> np.uint32(16); y ^= y >> np.uint32(16); z ^= z >> np.uint32(16); w ^= w >> np.uint32(16)
        x += y * w; y += z * x; z += x * y; w += y * z
    return np.stack([x, y, z, w], axis=1)


def test_pcg4d_matches_independent_numpy():
    v = np.random.default_rng(0).integers(0, 2**32, size=(2000, 4), dtype=np.uint32)
    assert (po.pcg4d(v) == pcg4d_numpy(v)).all()


def test_rand_is_counter_based_and_in_unit_interval():
    """Appendix A: draw k of a pixel-frame is uint_to_float(pcg4d(x, y, frame, k).x), state.w incremented first."""
    seeds = np.array([[3, 5, 7, 0], [1919, 1079, 1023, 0]], dtype=np.uint32)
    r = po.rand(seeds, 8)
    assert ((r >= 0) & (r < 1)).all()
    for i, s in enumerate(seeds):
        for k in range(8):
            h = pcg4d_numpy(np.array([[s[0], s[1], s[2], k + 1]], dtype=np.uint32))[0, 0]
            expect = (np.uint32(0x3F800000) | (h >> np.uint32(9))).view(np.float32) - np.float32(1.0)
            assert r[i, k] == expect


def test_scene_loader_counts(cornell):
    sc, _ = cornell
    i = sc.info
    # cornell_box = 8 shapes / 3,912 tris / 7 materials / 1 area light (12 tris)
    assert (i.n_prim_meshes, i.n_triangles, i.n_materials, i.n_lights, i.total_light_triangle_cnt) == (8, 3912, 7, 1, 12)
    assert i.integrator == b"path" and i.path_length == 10 and i.dir_light_idx == 0xFFFFFFFF
    assert i.bsdf_types == 1 | 2 | 8
    ca = host.Scene(scene_path("caustics"), 64, 36)
    assert (ca.info.n_triangles, ca.info.n_lights, ca.info.integrator) == (18893, 1, b"vcm")  # F5: file says vcm
    mt = host.Scene(scene_path("materials"), 32, 32)
    assert (mt.info.n_prim_meshes, mt.info.n_triangles, mt.info.n_lights) == (18, 23112, 3)
    assert mt.info.bsdf_types == 1 | 2 | 8 | 16 | 32
    cd = host.Scene(scene_path("cornell_dir"), 32, 32)
    assert cd.info.n_textures == 1 and cd.info.dir_light_idx == 0 and cd.info.n_lights == 1


def test_lbvh_structure(cornell):
    _, orc = cornell
    b = orc.lbvh()
    n = orc.n_tris
    assert (np.diff(b["keys"].astype(np.uint64)) > 0).all()  # strictly sorted (unique keys)
    assert sorted(b["leaf_prim"].tolist()) == list(range(n))
    assert ((b["keys"] >> np.uint64(32)).astype(np.uint32) == b["morton"][b["leaf_prim"]]).all()
    # every node except the root has exactly one parent; children point back
    children = np.concatenate([b["left"], b["right"]])
    assert sorted(children.tolist()) == list(range(1, 2 * n - 1))
    assert b["parent"][0] == 0xFFFFFFFF
    assert (b["parent"][b["left"]] == np.arange(n - 1)).all() and (b["parent"][b["right"]] == np.arange(n - 1)).all()
    # parent boxes are the exact union of the child boxes
    bb = b["aabb"].reshape(-1, 6)
    assert (bb[: n - 1, :3] == np.minimum(bb[b["left"], :3], bb[b["right"], :3])).all()
    assert (bb[: n - 1, 3:] == np.maximum(bb[b["left"], 3:], bb[b["right"], 3:])).all()


def brute_force_closest(tri, rays):
    """Moeller-Trumbore in float64 over all triangles (independent of the oracle's watertight test and of the BVH)."""
    o, d = rays[:, :3].astype(np.float64), rays[:, 4:7].astype(np.float64)
    v0, v1, v2 = (tri[:, k].astype(np.float64) for k in range(3))
    e1, e2 = v1 - v0, v2 - v0
    best_t = np.full(rays.shape[0], np.inf)
    best_p = np.full(rays.shape[0], -1, dtype=np.int64)
    for i in range(rays.shape[0]):
        p = np.cross(d[i], e2)
        det = (e1 * p).sum(1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            s = o[i] - v0
            u = (s * p).sum(1) * inv
            q = np.cross(s, e1)
            v = (q * d[i]).sum(1) * inv
            t = (e2 * q).sum(1) * inv
        ok = (np.abs(det) > 1e-14) & (u >= -1e-9) & (v >= -1e-9) & (u + v <= 1 + 1e-9) & (t > rays[i, 3]) & (t < rays[i, 7])
        if ok.any():
            t = np.where(ok, t, np.inf)
            best_p[i] = int(np.argmin(t))
            best_t[i] = t[best_p[i]]
    return best_t, best_p


def test_traversal_matches_brute_force(cornell):
    sc, orc = cornell
    d = sc.desc
    import ctypes as C
    verts = np.ctypeslib.as_array(C.cast(d.vertices, C.POINTER(C.c_float)), shape=(d.n_vertices, 8))
    tri = verts[:, :3].reshape(-1, 3, 3)  # de-indexed, identity transforms, prim meshes in order
    rays = random_rays(np.random.default_rng(3), [-3, -1, -3], [3, 5, 3], 600)
    hits, _ = orc.trace_closest(rays)
    bt, bp = brute_force_closest(tri, rays)
    hit_o = hits["prim"] != 0xFFFFFFFF
    assert (hit_o == (bp >= 0)).mean() > 0.995  # edge grazers may differ between fp32 watertight and fp64 MT
    both = hit_o & (bp >= 0)
    assert np.allclose(hits["t"][both], bt[both], rtol=2e-4, atol=1e-5)
    occ, _ = orc.trace_any(rays)
    assert (occ.astype(bool) == hit_o).all()


def test_render_is_deterministic_across_thread_counts(cornell):
    sc, orc = cornell
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    a, sa = orc.render(pc, ubo, 0, 2, threads=1)
    b, sb = orc.render(pc, ubo, 0, 2, threads=4)
    assert a.tobytes() == b.tobytes() and sa.rays == sb.rays
    # progressive == batched (running mean keyed by frame_num)
    c, _ = orc.render(pc, ubo, 0, 1)
    c, _ = orc.render(pc, ubo, 1, 1, rgba=c)
    assert c.tobytes() == a.tobytes()


def test_ray_budget_per_path(cornell):
    """at depth 6 with an area light a path traces at most 6 + 5*2 = 16 rays."""
    sc, orc = cornell
    pc, ubo = sc.make_pc(6, True), sc.make_ubo()
    _, st = orc.render(pc, ubo, 0, 1)
    n_paths = 64 * 64
    assert n_paths <= st.rays_closest <= 6 * n_paths
    assert st.rays_shadow <= 5 * n_paths and st.rays_probe <= st.rays_shadow


def test_golden_vectors():
    """tests/golden/*.npz were minted from this oracle by tests/golden/make_golden.py (the reference ships no vectors,
    SURVEY.md F2); they freeze the oracle against accidental change."""
    g = np.load(os.path.join(GOLDEN, "oracle_golden.npz"))
    assert (po.pcg4d(g["pcg_in"]) == g["pcg_out"]).all()
    assert bits_equal(po.rand(g["rand_seed"], 12), g["rand_out"]).all()
    sc = host.Scene(scene_path("cornell"), 48, 48)
    orc = po.OracleScene(sc)
    img, st = orc.render(sc.make_pc(6, True), sc.make_ubo(), 0, 2)
    assert bits_equal(img, g["cornell_48_d6_f2"]).all()
    assert st.rays == int(g["cornell_48_d6_f2_rays"])
    b = orc.lbvh()
    assert (b["left"] == g["cornell_left"]).all() and (b["right"] == g["cornell_right"]).all()
    sc2 = host.Scene(scene_path("materials"), 40, 40)
    img2, _ = po.OracleScene(sc2).render(sc2.make_pc(10, True), sc2.make_ubo(), 0, 2)
    assert bits_equal(img2, g["materials_40_d10_f2"]).all()


def test_rmse_literal_and_true():
    rng = np.random.default_rng(4)
    a = rng.uniform(0, 1, (5000, 4)).astype(np.float32)
    b = a.copy()
    assert po.rmse_true(a, b) == 0 and po.rmse_literal(a, b) == 0
    b[:, :3] += 0.1
    assert abs(po.rmse_true(a, b) - 0.1) < 1e-6
    # literal routine (calc_rmse.comp:37 subgroupMin): per 1024-pixel workgroup the MIN of the 32 subgroup sums survives,
    # and output_rmse.comp:23 divides sqrt(sum) by 3N
    d2 = ((a[:, :3] - b[:, :3]) ** 2).sum(1)
    wg_vals = []
    for wg in range(0, 5000, 1024):
        sums = [d2[s:min(s + 32, 5000)].sum(dtype=np.float32) if s < 5000 else np.float32(0) for s in range(wg, wg + 1024, 32)]
        wg_vals.append(min(sums))
    expect = np.sqrt(np.float32(sum(wg_vals))) / (5000 * 3)
    assert abs(po.rmse_literal(a, b) - expect) < 1e-6 * expect + 1e-12


def test_exr_roundtrip_is_half_precision(tmp_path):
    """ImageUtils::save_exr stores HALF (F8): the 1e-4 parity check must never be made on a written EXR."""
    rng = np.random.default_rng(6)
    img = rng.uniform(0, 4, (20, 30, 4)).astype(np.float32)
    img[..., 3] = 1
    p = str(tmp_path / "out.exr")
    host.save_exr(img, p)
    back = host.load_exr(p)
    assert back.shape == (20, 30, 4)
    rel = np.abs(back[..., :3] - img[..., :3]) / img[..., :3]
    assert rel.max() < 1e-3 and rel.max() > 1e-5


def _edge_rays(tri, rng, n):
    """Rays aimed at points ON triangle edges and vertices (where neighbouring triangles tie or nearly tie in t)."""
    k = rng.integers(0, tri.shape[0], n)
    a, b = tri[k, 0], tri[k, rng.integers(1, 3, n)]
    s = rng.uniform(0, 1, (n, 1)).astype(np.float32)
    s[: n // 4] = 0.0  # exactly a vertex
    target = a + s * (b - a)
    lo, hi = tri.reshape(-1, 3).min(0), tri.reshape(-1, 3).max(0)
    org = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = target - org
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    return np.concatenate([org, np.full((n, 1), 1e-3, np.float32), d.astype(np.float32), np.full((n, 1), 1e4, np.float32)], axis=1).astype(np.float32)


@pytest.mark.parametrize("name", ["caustics", "cornell"])
def test_hits_are_independent_of_the_tree(name):
    """The hit definition (watertight test + tri_clamp_t + closest-t / lowest-id rule) evaluated over ALL triangles with no tree
    must equal the LBVH walk bit for bit -- also for rays through edges, vertices and wall corners, where a culling rule that
    is not closed under the triangle test's rounding makes the answer depend on visiting order (found on caustics 1280x720:
    the walk held t = 19.496059 from the floor and culled the box of a wall triangle whose watertight t was 19.496048)."""
    sc = host.Scene(scene_path(name), 64, 64)
    orc = po.OracleScene(sc)
    d = sc.desc
    import ctypes as C
    verts = np.ctypeslib.as_array(C.cast(d.vertices, C.POINTER(C.c_float)), shape=(d.n_vertices, 8))
    tri = verts[:, :3].reshape(-1, 3, 3)
    rng = np.random.default_rng(31)
    lo, hi = tri.reshape(-1, 3).min(0), tri.reshape(-1, 3).max(0)
    rays = np.concatenate([_edge_rays(tri, rng, 6000), random_rays(rng, lo, hi, 2000)])
    if name == "caustics":
        rays[0] = [-9.94975662, 4.86726284, 0.856632948, 1e-3, 0.837418616, -0.503551424, 0.212522879, 1e4]
    walk, _ = orc.trace_closest(rays)
    brute = orc.trace_closest_brute(rays)
    assert (walk["prim"] == brute["prim"]).all()
    assert (walk["t"].view(np.uint32) == brute["t"].view(np.uint32)).all()
    assert (walk["b1"].view(np.uint32) == brute["b1"].view(np.uint32)).all()
