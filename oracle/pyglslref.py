"""ORACLE -- TEST INFRASTRUCTURE ONLY. ctypes face of oracle/_ref/libglslref.so (oracle/glslref/glslref.h): the reference's
own Path shaders translated mechanically from the unmodified GLSL (oracle/glslref/glsl2cpp.py) and compiled on the CPU.

It pins oracle/liboracle.so to the reference; only tests/ import it. Ray/triangle intersection and texture filtering are not
part of the shader source (Vulkan driver / hardware): they are callbacks, pointed here at liboracle.so's orc_trace1 /
orc_kat_texture, so both sides see the same hits and texels and every difference comes from the shading code.
"""
import ctypes as C
import os

import numpy as np

from . import pyoracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libglslref.so"))


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_ref", "libglslref.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make -C oracle/glslref` (needs the reference checkout)")
        L = C.CDLL(path)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int32
        L.ref_scene_create.argtypes = [vp, vp, vp, vp, C.POINTER(vp)]
        L.ref_scene_destroy.argtypes = [vp]
        L.ref_render_path.argtypes = [vp, vp, vp, u32, u32, vp, vp, i32]
        L.ref_render_bdpt_frame.argtypes = [vp, vp, vp, u32, vp, vp, vp, i32]
        L.ref_kat_pcg4d.argtypes = [vp, vp, u32, vp]
        L.ref_kat_rand.argtypes = [vp, vp, u32, u32, vp]
        L.ref_kat_offset_ray.argtypes = [vp, vp, vp, u32, vp, vp]
        L.ref_kat_sample_bsdf.argtypes = [vp, vp, vp, vp, vp, vp, u32, vp]
        L.ref_kat_eval_bsdf.argtypes = [vp, vp, vp, vp, vp, vp, u32, vp]
        L.ref_kat_bsdf_pdf.argtypes = [vp, vp, vp, vp, vp, vp, u32, vp]
        L.ref_kat_atmosphere.argtypes = [vp, vp, vp, vp, vp, u32, vp]
        L.ref_kat_sample_light.argtypes = [vp, i32, vp, vp, u32, vp]
        L.ref_kat_light_Le.argtypes = [vp, i32, i32, vp, u32, vp]
        L.ref_kat_load_material.argtypes = [vp, vp, vp, u32, vp]
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefScene:
    """The translated reference shaders bound to a host Scene; intersection + texels come from `oracle_scene`."""

    def __init__(self, scene, oracle_scene):
        self.scene, self.oracle_scene = scene, oracle_scene
        O = pyoracle.lib()
        self._h = C.c_void_p()
        trace1 = C.cast(O.orc_trace1, C.c_void_p)
        texture = C.cast(O.orc_kat_texture, C.c_void_p)
        if lib().ref_scene_create(C.addressof(scene.desc), oracle_scene._h, trace1, texture, C.byref(self._h)) != 0:
            raise RuntimeError("ref_scene_create failed")

    def render(self, pc, ubo, first_frame, n_frames, rgba=None, threads=0):
        """(film, [closest, shadow, probe] ray counts) after frames [first_frame, first_frame + n_frames) of path.rgen."""
        W, H = pc.size_x, pc.size_y
        if rgba is None:
            rgba = np.zeros((H, W, 4), dtype=np.float32)
        assert rgba.dtype == np.float32 and rgba.flags.c_contiguous
        rays = np.zeros(3, dtype=np.uint64)
        lib().ref_render_path(self._h, C.addressof(pc), C.addressof(ubo), first_frame, n_frames, rgba.ctypes.data, rays.ctypes.data, threads)
        return rgba, rays

    def render_bdpt_frame(self, pc, ubo, frame, threads=0):
        """One dispatch of bdpt.rgen (pc is a PCBdpt): (image (H, W, 4) = own strategies + self-splats as main() stores them,
        splat (H, W, 3) = the splats the pixels send to OTHER pixels, [closest, shadow, -] ray counts)."""
        W, H = pc.size_x, pc.size_y
        image = np.zeros((H, W, 4), dtype=np.float32)
        splat = np.zeros((H, W, 3), dtype=np.float32)
        rays = np.zeros(3, dtype=np.uint64)
        lib().ref_render_bdpt_frame(self._h, C.addressof(pc), C.addressof(ubo), int(frame), image.ctypes.data, splat.ctypes.data, rays.ctypes.data, threads)
        return image, splat, rays

    def pcg4d(self, v4):
        v = np.ascontiguousarray(v4, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros_like(v)
        lib().ref_kat_pcg4d(self._h, v.ctypes.data, v.shape[0], out.ctypes.data)
        return out

    def rand(self, seed4, draws):
        s = np.ascontiguousarray(seed4, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros((s.shape[0], draws), dtype=np.float32)
        lib().ref_kat_rand(self._h, s.ctypes.data, s.shape[0], draws, out.ctypes.data)
        return out

    def offset_ray(self, p, n):
        p, n = _f32(p).reshape(-1, 3), _f32(n).reshape(-1, 3)
        a, b = np.zeros_like(p), np.zeros_like(p)
        lib().ref_kat_offset_ray(self._h, p.ctypes.data, n.ctypes.data, p.shape[0], a.ctypes.data, b.ctypes.data)
        return a, b

    def sample_bsdf(self, mat, n_s, wo, rands, side):
        n_s, wo, rands = _f32(n_s).reshape(-1, 3), _f32(wo).reshape(-1, 3), _f32(rands).reshape(-1, 3)
        side = np.ascontiguousarray(side, dtype=np.uint8)
        out = np.zeros((n_s.shape[0], 8), dtype=np.float32)
        lib().ref_kat_sample_bsdf(self._h, C.addressof(mat), n_s.ctypes.data, wo.ctypes.data, rands.ctypes.data, side.ctypes.data,
                                  n_s.shape[0], out.ctypes.data)
        return out

    def eval_bsdf(self, mat, n_s, wo, wi, side):
        n_s, wo, wi = _f32(n_s).reshape(-1, 3), _f32(wo).reshape(-1, 3), _f32(wi).reshape(-1, 3)
        side = np.ascontiguousarray(side, dtype=np.uint8)
        out = np.zeros((n_s.shape[0], 4), dtype=np.float32)
        lib().ref_kat_eval_bsdf(self._h, C.addressof(mat), n_s.ctypes.data, wo.ctypes.data, wi.ctypes.data, side.ctypes.data,
                                n_s.shape[0], out.ctypes.data)
        return out

    def bsdf_pdf(self, mat, n_s, wo, wi, side):
        n_s, wo, wi = _f32(n_s).reshape(-1, 3), _f32(wo).reshape(-1, 3), _f32(wi).reshape(-1, 3)
        side = np.ascontiguousarray(side, dtype=np.uint8)
        out = np.zeros(n_s.shape[0], dtype=np.float32)
        lib().ref_kat_bsdf_pdf(self._h, C.addressof(mat), n_s.ctypes.data, wo.ctypes.data, wi.ctypes.data, side.ctypes.data,
                               n_s.shape[0], out.ctypes.data)
        return out

    def atmosphere(self, origin, direction, light_dir, light_L):
        o, d = _f32(origin).reshape(-1, 3), _f32(direction).reshape(-1, 3)
        ld, lL = _f32(light_dir).reshape(3), _f32(light_L).reshape(3)
        out = np.zeros_like(o)
        lib().ref_kat_atmosphere(self._h, o.ctypes.data, d.ctypes.data, ld.ctypes.data, lL.ctypes.data, o.shape[0], out.ctypes.data)
        return out

    def sample_light(self, num_lights, rands4, p3):
        r, p = _f32(rands4).reshape(-1, 4), _f32(p3).reshape(-1, 3)
        out = np.zeros((r.shape[0], 16), dtype=np.float32)
        lib().ref_kat_sample_light(self._h, num_lights, r.ctypes.data, p.ctypes.data, r.shape[0], out.ctypes.data)
        return out

    def light_Le(self, num_lights, total_light, rands6):
        r = _f32(rands6).reshape(-1, 6)
        out = np.zeros((r.shape[0], 16), dtype=np.float32)
        lib().ref_kat_light_Le(self._h, num_lights, total_light, r.ctypes.data, r.shape[0], out.ctypes.data)
        return out

    def load_material(self, material_idx, uv):
        from lumen_b200._ctypes_types import Material
        idx = np.ascontiguousarray(material_idx, dtype=np.uint32)
        uv = _f32(uv).reshape(-1, 2)
        out = (Material * idx.size)()
        lib().ref_kat_load_material(self._h, idx.ctypes.data, uv.ctypes.data, idx.size, C.addressof(out))
        return out

    def close(self):
        if self._h:
            lib().ref_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
