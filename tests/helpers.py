import ctypes as C
import os

import numpy as np

from lumen_b200._ctypes_types import Material

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_material(**kw):
    m = Material()
    m.texture_id = -1
    for k, val in kw.items():
        if isinstance(val, (tuple, list)):
            for i, vv in enumerate(val):
                getattr(m, k)[i] = vv
        else:
            setattr(m, k, val)
    return m


# One material per BSDF code path of src/shaders/bsdf/*.glsl
MATERIALS = {
    "diffuse": dict(albedo=(0.7, 0.5, 0.3), bsdf_type=1, bsdf_props=1 | 8),
    "mirror": dict(albedo=(1, 1, 1), bsdf_type=2, bsdf_props=2 | 8),
    "glass": dict(albedo=(1, 1, 1), ior=1.5, bsdf_type=4, bsdf_props=2 | 16),
    "dielectric_smooth": dict(albedo=(1, 1, 1), ior=1.52, bsdf_type=8, bsdf_props=2 | 8 | 16),
    "dielectric_rough": dict(albedo=(0.9, 0.8, 1), ior=1.52, roughness=0.3, bsdf_type=8, bsdf_props=4 | 8 | 16),
    "dielectric_thin": dict(albedo=(0.9, 0.8, 1), ior=1.3, roughness=0.4, thin=1, bsdf_type=8, bsdf_props=4 | 8 | 16),
    "dielectric_reflect_only": dict(albedo=(1, 1, 1), ior=1.4, roughness=0.2, bsdf_type=8, bsdf_props=4 | 8),
    "conductor_rough": dict(albedo=(1, 1, 1), k=(3.0, 2.5, 2.0), roughness=0.3, bsdf_type=16, bsdf_props=4 | 8),
    "conductor_smooth": dict(albedo=(0.2, 0.9, 1.1), k=(3.0, 2.5, 2.0), roughness=0.0, bsdf_type=16, bsdf_props=2 | 8),
    "principled": dict(albedo=(0.8, 0.3, 0.2), ior=1.45, roughness=0.4, metallic=0.3, spec_trans=0.3, specular_tint=0.2, clearcoat=0.7,
                       clearcoat_gloss=0.6, flatness=0.2, anisotropy=0.4, sheen_tint=0.5, subsurface=0.1, bsdf_type=32, bsdf_props=4 | 8 | 16),
    "principled_default": dict(albedo=(1, 1, 1), ior=1.0, roughness=0.5, sheen_tint=0.5, clearcoat_gloss=1.0, bsdf_type=32, bsdf_props=4 | 8),
    "principled_plastic": dict(albedo=(0.5, 0.5, 0.5), ior=1.5, roughness=0.3, metallic=1.0, subsurface=0.1, spec_trans=0.5, thin=1, bsdf_type=32,
                               bsdf_props=1 | 8 | 16 | 4),
    "principled_mirrorlike": dict(albedo=(0.9, 0.9, 0.9), ior=1.5, roughness=0.02, metallic=1.0, bsdf_type=32, bsdf_props=2 | 8),
    "unknown_type": dict(albedo=(1, 1, 1), bsdf_type=0, bsdf_props=0),
}


def unit_vectors(rng, n):
    v = rng.normal(size=(n, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32)


def bits_equal(a, b):
    """Bit-exact float comparison that treats NaN == NaN (any payload) as equal."""
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype == np.float32:
        return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    return a == b


def random_rays(rng, lo, hi, n, tmin=1e-3, tmax=1e4):
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    pad = 0.1 * (hi - lo)
    org = rng.uniform(lo - pad, hi + pad, (n, 3)).astype(np.float32)
    d = unit_vectors(rng, n)
    return np.concatenate([org, np.full((n, 1), tmin, np.float32), d, np.full((n, 1), tmax, np.float32)], axis=1).astype(np.float32)


def pixel_agreement(gpu, cpu, rel=1e-4):
    """Fraction of pixels whose RGB agree within rel (BASELINE.json: 1e-4 relative on >= 99.9 % of pixels)."""
    a, b = gpu[..., :3], cpu[..., :3]
    ok = (np.abs(a - b) <= rel * np.maximum(np.abs(b), 1e-6)) | (np.isnan(a) & np.isnan(b))
    return float(ok.all(axis=-1).mean())
