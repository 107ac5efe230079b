"""Python face of the C++ host shell (liblumen_host.so, include/lumen_host.h).

Mirrors the part of Lumen's `RayTracer` / `LumenScene` that feeds the Path integrator: `Scene(path, w, h)` is
`LumenScene::load_scene` (src/RayTracer/LumenScene.cpp:52-229) with the window size made explicit; `make_pc` is the
push-constant fill of `Path::render` (src/RayTracer/Path.cpp:27-38); `make_ubo` is
`Integrator::update_uniform_buffers` (src/RayTracer/Integrator.cpp:60-72); `save_exr` is `ImageUtils::save_exr`.
"""
import ctypes as C
import os

import numpy as np

from ._ctypes_types import Light, Material, PCPath, SceneDesc, SceneInfo, SceneUBO, Vertex

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "host", "liblumen_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs the vendored parsers)")
        L = C.CDLL(path)
        L.lmh_last_error.restype = C.c_char_p
        L.lmh_scene_load.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
        L.lmh_scene_from_arrays.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                            C.c_void_p, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_uint32,
                                            C.c_uint32, C.POINTER(C.c_void_p)]
        L.lmh_scene_destroy.argtypes = [C.c_void_p]
        L.lmh_scene_get_desc.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        L.lmh_scene_get_info.argtypes = [C.c_void_p, C.POINTER(SceneInfo)]
        L.lmh_scene_make_pc.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(PCPath)]
        L.lmh_scene_make_ubo.argtypes = [C.c_void_p, C.POINTER(SceneUBO)]
        L.lmh_save_exr.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p]
        L.lmh_save_exr_half_bgr.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p]
        L.lmh_save_checkpoint.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.lmh_load_checkpoint.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)] + [C.POINTER(C.c_uint32)] * 4
        L.lmh_load_exr.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.lmh_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class Scene:
    """Loaded scene: owns the host arrays; `desc` is the lmb_scene_desc to hand to lmb_upload_scene."""

    def __init__(self, path=None, width=1920, height=1080, _handle=None):
        self._h = C.c_void_p()
        self.width, self.height = int(width), int(height)
        if _handle is not None:
            self._h = _handle
        else:
            if lib().lmh_scene_load(os.fsencode(path), self.width, self.height, C.byref(self._h)) != 0:
                raise RuntimeError("lmh_scene_load: " + lib().lmh_last_error().decode())
        self.desc = SceneDesc()
        lib().lmh_scene_get_desc(self._h, C.byref(self.desc))
        self.info = SceneInfo()
        lib().lmh_scene_get_info(self._h, C.byref(self.info))

    @classmethod
    def from_arrays(cls, vertices, mesh_tri_counts, mesh_materials, materials, analytic_lights=(), fov=45.0, cam_pos=(0, 0, 5),
                    cam_dir=(0, 0, -1), path_length=6, sky_col=(0, 0, 0), width=512, height=512, mesh_world=None):
        """vertices: structured (n,8) float32 array = pos, normal, uv per de-indexed vertex."""
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 8)
        tc = np.ascontiguousarray(mesh_tri_counts, dtype=np.uint32)
        mm = np.ascontiguousarray(mesh_materials, dtype=np.uint32)
        mats = (Material * len(materials))(*materials)
        lights = (Light * max(1, len(analytic_lights)))(*analytic_lights)
        mw = None if mesh_world is None else np.ascontiguousarray(mesh_world, dtype=np.float32)
        pos = (C.c_float * 3)(*cam_pos)
        d = (C.c_float * 3)(*cam_dir)
        sky = (C.c_float * 3)(*sky_col)
        h = C.c_void_p()
        rc = lib().lmh_scene_from_arrays(v.ctypes.data, v.shape[0], tc.ctypes.data, mm.ctypes.data, None if mw is None else mw.ctypes.data,
                                         len(tc), C.addressof(mats), len(materials), C.addressof(lights), len(analytic_lights), float(fov),
                                         C.addressof(pos), C.addressof(d), int(path_length), C.addressof(sky), int(width), int(height), C.byref(h))
        if rc != 0:
            raise RuntimeError("lmh_scene_from_arrays: " + lib().lmh_last_error().decode())
        return cls(width=width, height=height, _handle=h)

    def make_pc(self, max_depth=0, direct_lighting=True):
        pc = PCPath()
        lib().lmh_scene_make_pc(self._h, int(max_depth), 1 if direct_lighting else 0, C.byref(pc))
        return pc

    def make_ubo(self):
        ubo = SceneUBO()
        lib().lmh_scene_make_ubo(self._h, C.byref(ubo))
        return ubo

    def close(self):
        if self._h:
            lib().lmh_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def save_exr(rgba, path):
    a = np.ascontiguousarray(rgba, dtype=np.float32)
    h, w = a.shape[0], a.shape[1]
    if lib().lmh_save_exr(a.ctypes.data, w, h, os.fsencode(path)) != 0:
        raise RuntimeError("lmh_save_exr: " + lib().lmh_last_error().decode())


def save_exr_half_bgr(planes, path):
    """planes: (3, H, W) uint16 = B, G, R half planes from Device.download_half_bgr()."""
    a = np.ascontiguousarray(planes, dtype=np.uint16)
    h, w = a.shape[1], a.shape[2]
    if lib().lmh_save_exr_half_bgr(a.ctypes.data, w, h, os.fsencode(path)) != 0:
        raise RuntimeError("lmh_save_exr_half_bgr: " + lib().lmh_last_error().decode())


def save_checkpoint(path, rgba, frames, path_length):
    a = np.ascontiguousarray(rgba, dtype=np.float32)
    if lib().lmh_save_checkpoint(os.fsencode(path), a.ctypes.data, a.shape[1], a.shape[0], int(frames), int(path_length)) != 0:
        raise RuntimeError("lmh_save_checkpoint: " + lib().lmh_last_error().decode())


def load_checkpoint(path):
    """-> (rgba (H, W, 4) float32, frames, path_length)"""
    p = C.c_void_p()
    w, h, fr, pl = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    if lib().lmh_load_checkpoint(os.fsencode(path), C.byref(p), C.byref(w), C.byref(h), C.byref(fr), C.byref(pl)) != 0:
        raise RuntimeError("lmh_load_checkpoint: " + lib().lmh_last_error().decode())
    n = w.value * h.value * 4
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).copy().reshape(h.value, w.value, 4)
    lib().lmh_free(p)
    return out, fr.value, pl.value


def load_exr(path):
    p, w, h = C.c_void_p(), C.c_int32(), C.c_int32()
    if lib().lmh_load_exr(os.fsencode(path), C.byref(p), C.byref(w), C.byref(h)) != 0:
        raise RuntimeError("lmh_load_exr: " + lib().lmh_last_error().decode())
    n = w.value * h.value * 4
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).copy().reshape(h.value, w.value, 4)
    lib().lmh_free(p)
    return out
