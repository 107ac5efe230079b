"""Diagnostic parity sweep CUDA vs oracle (prints, does not assert). Run on the GPU box: python tests/diag_gpu_vs_oracle.py"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
from lumen_b200 import host, integrator
from lumen_b200._ctypes_types import Material
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rng = np.random.default_rng(1)
dev = integrator.Device(0)

def cmp(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)) if a.dtype == np.float32 else (a == b)
    print(f"{name:28s} equal {same.mean()*100:8.4f}%  n={a.size}", "" if same.all() else f" first mismatch idx {np.argwhere(~same)[0]} gpu={a[tuple(np.argwhere(~same)[0])]} cpu={b[tuple(np.argwhere(~same)[0])]}")
    return same.all()

# --- KATs
v = rng.integers(0, 2**32, size=(1000, 4), dtype=np.uint32)
cmp("pcg4d", dev.kat_pcg4d(v), po.pcg4d(v))
cmp("rand", dev.kat_rand(v, 16), po.rand(v, 16))
x = np.concatenate([rng.uniform(-10, 10, 50000), rng.uniform(-100, 100, 50000)]).astype(np.float32)
y = rng.uniform(0, 9, x.size).astype(np.float32)
g, c = dev.kat_detmath(x, y), po.detmath(x, y)
for k in g: cmp("detmath." + k, g[k], c[k])
p = rng.uniform(-20, 20, (20000, 3)).astype(np.float32); p[::7] *= 0.001
n = rng.normal(size=(20000, 3)).astype(np.float32); n /= np.linalg.norm(n, axis=1, keepdims=True)
ga, gb = dev.kat_offset_ray(p, n); ca, cb = po.offset_ray(p, n)
cmp("offset_ray", ga, ca); cmp("offset_ray2", gb, cb)

def mat(**kw):
    m = Material(); m.texture_id = -1
    for k, val in kw.items():
        if isinstance(val, (tuple, list)):
            for i, vv in enumerate(val): getattr(m, k)[i] = vv
        else: setattr(m, k, val)
    return m
mats = {
 "diffuse": mat(albedo=(0.7, 0.5, 0.3), bsdf_type=1, bsdf_props=1 | 8),
 "mirror": mat(albedo=(1, 1, 1), bsdf_type=2, bsdf_props=2 | 8),
 "glass": mat(albedo=(1, 1, 1), ior=1.5, bsdf_type=4, bsdf_props=2 | 16),
 "dielectric_smooth": mat(albedo=(1, 1, 1), ior=1.52, bsdf_type=8, bsdf_props=2 | 8 | 16),
 "dielectric_rough": mat(albedo=(0.9, 0.8, 1), ior=1.52, roughness=0.3, bsdf_type=8, bsdf_props=4 | 8 | 16),
 "dielectric_thin": mat(albedo=(0.9, 0.8, 1), ior=1.3, roughness=0.4, thin=1, bsdf_type=8, bsdf_props=4 | 8 | 16),
 "conductor_rough": mat(albedo=(1, 1, 1), k=(3.0, 2.5, 2.0), roughness=0.3, bsdf_type=16, bsdf_props=4 | 8),
 "conductor_smooth": mat(albedo=(0.2, 0.9, 1.1), k=(3.0, 2.5, 2.0), roughness=0.0, bsdf_type=16, bsdf_props=2 | 8),
 "principled": mat(albedo=(0.8, 0.3, 0.2), ior=1.45, roughness=0.4, metallic=0.3, spec_trans=0.3, specular_tint=0.2, clearcoat=0.7, clearcoat_gloss=0.6, flatness=0.2, anisotropy=0.4, sheen_tint=0.5, subsurface=0.1, bsdf_type=32, bsdf_props=4 | 8 | 16),
 "principled_plastic": mat(albedo=(0.5, 0.5, 0.5), ior=1.5, roughness=0.3, metallic=1.0, subsurface=0.1, spec_trans=0.5, thin=1, bsdf_type=32, bsdf_props=1 | 8 | 16 | 4),
}
N = 20000
ns = rng.normal(size=(N, 3)).astype(np.float32); ns /= np.linalg.norm(ns, axis=1, keepdims=True)
wo = rng.normal(size=(N, 3)).astype(np.float32); wo /= np.linalg.norm(wo, axis=1, keepdims=True)
wi = rng.normal(size=(N, 3)).astype(np.float32); wi /= np.linalg.norm(wi, axis=1, keepdims=True)
rr = rng.uniform(0, 1, (N, 3)).astype(np.float32)
side = rng.integers(0, 2, N).astype(np.uint8)
for name, m in mats.items():
    cmp("sample." + name, dev.kat_sample_bsdf(m, ns, wo, rr, side), po.sample_bsdf(m, ns, wo, rr, side))
    cmp("eval." + name, dev.kat_eval_bsdf(m, ns, wo, wi, side), po.eval_bsdf(m, ns, wo, wi, side))
o3 = rng.uniform(-5, 5, (2000, 3)).astype(np.float32); o3[:, 1] = np.abs(o3[:, 1])
d3 = rng.normal(size=(2000, 3)).astype(np.float32); d3 /= np.linalg.norm(d3, axis=1, keepdims=True)
cmp("atmosphere", dev.kat_atmosphere(o3, d3, (0.5, 0.7, 0.2), (98, 82, 30)), po.atmosphere(o3, d3, (0.5, 0.7, 0.2), (98, 82, 30)))

def scene_checks(path, W, H, depth, frames, force_lights=None):
    print("=== scene", path, W, H, "depth", depth, "frames", frames)
    sc = host.Scene(os.path.join(ROOT, path), W, H)
    orc = po.OracleScene(sc)
    dev.upload_scene(sc.desc); dev.build_accel()
    gl, cl = dev.lbvh(), orc.lbvh()
    for k in cl: cmp("lbvh." + k, gl[k], cl[k])
    st = dev.stats(); print("build ms", st.ms_build_accel, st.ms_build_morton, st.ms_build_sort, st.ms_build_tree, st.ms_build_refit)
    # rays
    nr = 200000
    lo = cl["aabb"][:3]; hi = cl["aabb"][3:6]
    org = rng.uniform(lo, hi, (nr, 3)).astype(np.float32)
    d = rng.normal(size=(nr, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, np.full((nr, 1), 1e-3, np.float32), d, np.full((nr, 1), 1e4, np.float32)], axis=1)
    gh = dev.trace_closest(rays); ch, _ = orc.trace_closest(rays)
    for k in ("t", "b1", "b2", "prim"): cmp("closest." + k, gh[k], ch[k])
    rays[:, 7] = rng.uniform(0.1, 5, nr)
    cmp("any", dev.trace_any(rays), orc.trace_any(rays)[0])
    if sc.info.n_lights:
        r4 = rng.uniform(0, 1, (5000, 4)).astype(np.float32); p3 = rng.uniform(lo, hi, (5000, 3)).astype(np.float32)
        cmp("sample_light", dev.kat_sample_light(sc.info.n_lights, r4, p3), orc.sample_light(sc.info.n_lights, r4, p3))
    if sc.info.n_textures:
        uv = rng.uniform(-2, 3, (5000, 2)).astype(np.float32)
        cmp("texture", dev.kat_texture(0, uv), orc.texture(0, uv))
    pc = sc.make_pc(depth, True); ubo = sc.make_ubo()
    dev.init(W, H, 4)
    dev.set_profile_stages(True)
    t = time.time(); dev.render(pc, ubo, 0, frames); tg = time.time() - t
    gimg = dev.download(); gs = dev.stats()
    t = time.time(); cimg, cs = orc.render(pc, ubo, 0, frames); tc = time.time() - t
    print(f"gpu rays {gs.rays} ({gs.rays_closest},{gs.rays_shadow},{gs.rays_probe}) cpu rays {cs.rays} ({cs.rays_closest},{cs.rays_shadow},{cs.rays_probe}) nan gpu {gs.nan_samples} cpu {cs.nan_pixels}")
    print(f"gpu nodes/ray {gs.nodes_visited/max(gs.rays,1):.2f} tris/ray {gs.tris_tested/max(gs.rays,1):.2f} | cpu {cs.nodes_visited/max(cs.rays,1):.2f} {cs.tris_tested/max(cs.rays,1):.2f}")
    print(f"gpu {gs.ms_render:.1f} ms ({gs.rays/gs.ms_render/1e3:.1f} Mrays/s; trace {gs.ms_extend:.1f} shade {gs.ms_shade:.1f} connect {gs.ms_connect:.1f} raygen+sky+film {gs.ms_film:.1f}) wall {tg:.2f}s | cpu {tc:.2f}s ({cs.rays/cs.seconds/1e6:.2f} Mrays/s, {cs.threads} thr)")
    a, b = gimg[..., :3], cimg[..., :3]
    exact = (a.view(np.uint32) == b.view(np.uint32)).all(axis=2)
    rel = np.abs(a - b) <= 1e-4 * np.maximum(np.abs(b), 1e-6)
    ok = rel.all(axis=2)
    print(f"pixels bit-exact {exact.mean()*100:.4f}%  within 1e-4 rel {ok.mean()*100:.4f}%  max abs diff {np.abs(a-b).max():.3g}")
    if not ok.all():
        ys, xs = np.nonzero(~ok); print(" first bad pixels", list(zip(xs[:5], ys[:5])), a[ys[0], xs[0]], b[ys[0], xs[0]])
    return sc, orc

scene_checks("scenes/cornell_box/cornell_box_path.json", 256, 256, 6, 4)
scene_checks("scenes/material_test/materials.json", 256, 256, 10, 4)
scene_checks("scenes/caustics.json", 320, 180, 12, 4)
scene_checks("scenes/cornell_box/cornell_box_dir.json", 256, 256, 6, 4)
print("DONE")
