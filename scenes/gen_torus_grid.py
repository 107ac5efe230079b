"""BASELINE.json config 5: synthetic K x K x K grid of instanced tori, flattened to world space (SURVEY.md 8d "Config 5").

Constants: major radius 0.35, minor radius 0.12, grid pitch 1.0, tessellation nu x nv quads (2 * nu * nv triangles per
torus). Every instance is rotated about a jittered axis and shifted by up to 0.15 pitch; the jitter of instance (i, j, k)
comes from pcg4d(i, j, k, 0) -- Lumen's own hash (src/shaders/utils.glsl:121-134) -- so the scene is a pure function of
(K, nu, nv). The default 10 x 10 x 10 x (100 x 50) gives 10.0 M triangles.

    scene = make_scene(K=10, nu=100, nv=50)            # lumen_b200.host.Scene via lmh_scene_from_arrays
"""
import numpy as np

MAJOR, MINOR, PITCH, JITTER = 0.35, 0.12, 1.0, 0.15


def pcg4d(v):
    v = np.array(v, dtype=np.uint32).reshape(-1, 4).copy()
    with np.errstate(over="ignore"):
        v = v * np.uint32(1664525) + np.uint32(1013904223)
        v[:, 0] += v[:, 1] * v[:, 3]; v[:, 1] += v[:, 2] * v[:, 0]; v[:, 2] += v[:, 0] * v[:, 1]; v[:, 3] += v[:, 1] * v[:, 2]
        v ^= v >> np.uint32(16)
        v[:, 0] += v[:, 1] * v[:, 3]; v[:, 1] += v[:, 2] * v[:, 0]; v[:, 2] += v[:, 0] * v[:, 1]; v[:, 3] += v[:, 1] * v[:, 2]
    return v


def unit_torus(nu, nv):
    """De-indexed triangles of one torus: (2*nu*nv*3, 8) float32 rows = pos, normal, uv."""
    th = np.linspace(0, 2 * np.pi, nu + 1)
    ph = np.linspace(0, 2 * np.pi, nv + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    ring = np.stack([np.cos(T), np.zeros_like(T), np.sin(T)], -1)
    nrm = np.cos(P)[..., None] * ring + np.sin(P)[..., None] * np.array([0, 1.0, 0])
    pos = MAJOR * ring + MINOR * nrm
    uv = np.stack([T / (2 * np.pi), P / (2 * np.pi)], -1)
    vert = np.concatenate([pos, nrm, uv], -1).reshape(-1, 8)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * (nv + 1) + j).ravel()
    b, c, d = a + (nv + 1), a + (nv + 1) + 1, a + 1
    f = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
    return vert[f.ravel()].astype(np.float32)


def instance_transforms(K):
    ijk = np.stack(np.meshgrid(np.arange(K), np.arange(K), np.arange(K), indexing="ij"), -1).reshape(-1, 3)
    h = pcg4d(np.concatenate([ijk, np.zeros((ijk.shape[0], 1), np.int64)], 1))
    u = (h >> np.uint32(8)).astype(np.float64) / float(1 << 24)  # four uniforms in [0, 1) per instance
    axis = np.stack([2 * u[:, 0] - 1, 2 * u[:, 1] - 1, 2 * u[:, 2] - 1], -1) + 1e-3
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    ang = 2 * np.pi * u[:, 3]
    c, s = np.cos(ang)[:, None, None], np.sin(ang)[:, None, None]
    x, y, z = axis[:, 0], axis[:, 1], axis[:, 2]
    Kx = np.zeros((ijk.shape[0], 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0], Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -z, y, z, -x, -y, x
    R = np.eye(3)[None] + s * Kx + (1 - c) * (Kx @ Kx)
    shift = (ijk - (K - 1) / 2.0) * PITCH + (u[:, :3] - 0.5) * 2 * JITTER * PITCH
    return R, shift


def make_vertices(K=10, nu=100, nv=50):
    base = unit_torus(nu, nv)
    R, shift = instance_transforms(K)
    out = np.empty((R.shape[0], base.shape[0], 8), np.float32)
    for n in range(R.shape[0]):  # one instance at a time keeps the peak memory at the output size
        out[n, :, 0:3] = base[:, 0:3] @ R[n].T.astype(np.float32) + shift[n].astype(np.float32)
        out[n, :, 3:6] = base[:, 3:6] @ R[n].T.astype(np.float32)
        out[n, :, 6:8] = base[:, 6:8]
    return out.reshape(-1, 8), R.shape[0], base.shape[0] // 3


def make_scene(K=10, nu=100, nv=50, width=1024, height=1024):
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from helpers import make_material
    from lumen_b200 import host
    from lumen_b200._ctypes_types import Light
    verts, n_inst, tris_per = make_vertices(K, nu, nv)
    mats = [make_material(albedo=(0.7, 0.7, 0.7), bsdf_type=1, bsdf_props=1 | 8)]
    sun = Light()
    sun.pos[:] = (3.0, 10.0, 5.0)
    sun.to[:] = (0.0, 0.0, 0.0)
    sun.L[:] = (3.0, 3.0, 3.0)
    sun.light_flags = 3 | (1 << 5)  # directional, delta (LumenScene.cpp:500-510)
    ext = K * PITCH  # Lumen's camera quirk (Camera.h:67-85, see cornell_box_path.json): "dir" +z looks down -z
    return host.Scene.from_arrays(verts, np.full(n_inst, tris_per, np.uint32), np.zeros(n_inst, np.uint32), mats, [sun], fov=45.0,
                                  cam_pos=(0.0, 0.0, 1.6 * ext), cam_dir=(0.0, 0.0, 1.0), path_length=4, sky_col=(0.5, 0.6, 0.8), width=width,
                                  height=height)


def random_rays(n, K, seed=1):
    """Uniformly random origins and directions inside the grid's bounding box (incoherent-ray sweep), seeded by pcg4d."""
    idx = np.arange(n, dtype=np.uint32)
    a = pcg4d(np.stack([idx, np.full(n, seed, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)], 1))
    b = pcg4d(np.stack([idx, np.full(n, seed, np.uint32), np.ones(n, np.uint32), np.zeros(n, np.uint32)], 1))
    ua = (a >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)
    ub = (b >> np.uint32(8)).astype(np.float32) / np.float32(1 << 24)
    half = 0.5 * K * PITCH
    org = (ua[:, :3] * 2 - 1) * half
    z = ub[:, 0] * 2 - 1
    phi = 2 * np.pi * ub[:, 1]
    r = np.sqrt(np.maximum(0, 1 - z * z))
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z], 1)
    rays = np.concatenate([org, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], 1)
    return np.ascontiguousarray(rays, dtype=np.float32)
