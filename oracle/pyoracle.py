"""ORACLE -- TEST INFRASTRUCTURE ONLY. ctypes face of oracle/liboracle.so (oracle/oracle.h).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs. Never imported by
the lumen_b200 package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Stats(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_probe", C.c_uint64), ("nodes_visited", C.c_uint64),
                ("tris_tested", C.c_uint64), ("nan_pixels", C.c_uint64), ("seconds", C.c_double), ("threads", C.c_int32)]

    @property
    def rays(self):
        return self.rays_closest + self.rays_shadow + self.rays_probe


HIT_DTYPE = np.dtype([("t", np.float32), ("b1", np.float32), ("b2", np.float32), ("prim", np.uint32)])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make -C oracle`")
        L = C.CDLL(path)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int32
        L.orc_scene_create.argtypes = [vp, C.POINTER(vp)]
        L.orc_scene_destroy.argtypes = [vp]
        for name, rt in [("left", u32), ("right", u32), ("parent", u32), ("leaf_prim", u32), ("morton", u32), ("keys", C.c_uint64),
                         ("aabb", C.c_float)]:
            fn = getattr(L, "orc_lbvh_" + name)
            fn.argtypes = [vp]
            fn.restype = C.POINTER(rt)
        L.orc_lbvh_num_tris.argtypes = [vp]
        L.orc_lbvh_num_tris.restype = u32
        L.orc_render.argtypes = [vp, vp, vp, u32, u32, vp, vp, i32]
        L.orc_set_row_shard.argtypes = [u32, u32]
        L.orc_render_frame_raw.argtypes = [vp, vp, vp, u32, vp, vp, i32]
        L.orc_render_bdpt.argtypes = [vp, vp, vp, u32, u32, vp, vp, i32]
        L.orc_render_bdpt_frame_raw.argtypes = [vp, vp, vp, u32, vp, vp, vp, i32]
        L.orc_bdpt_set_only_s.argtypes = [i32]
        L.orc_bdpt_set_check_restore.argtypes = [i32]
        L.orc_bdpt_restore_violations.restype = C.c_longlong
        L.orc_trace_closest.argtypes = [vp, vp, u32, vp, vp, i32]
        L.orc_trace_any.argtypes = [vp, vp, u32, vp, vp, i32]
        L.orc_trace_closest_brute.argtypes = [vp, vp, u32, vp, i32]
        L.orc_kat_pcg4d.argtypes = [vp, u32, vp]
        L.orc_kat_rand.argtypes = [vp, u32, u32, vp]
        L.orc_kat_detmath.argtypes = [vp, vp, u32, vp, vp, vp, vp]
        L.orc_kat_offset_ray.argtypes = [vp, vp, u32, vp, vp]
        L.orc_kat_sample_bsdf.argtypes = [vp, vp, vp, vp, vp, u32, vp]
        L.orc_kat_eval_bsdf.argtypes = [vp, vp, vp, vp, vp, u32, vp]
        L.orc_kat_bsdf_pdf.argtypes = [vp, vp, vp, vp, vp, u32, vp]
        L.orc_kat_atmosphere.argtypes = [vp, vp, vp, vp, u32, vp]
        L.orc_kat_sample_light.argtypes = [vp, i32, vp, vp, u32, vp]
        L.orc_kat_light_Le.argtypes = [vp, i32, i32, vp, u32, vp]
        L.orc_kat_texture.argtypes = [vp, u32, vp, u32, vp]
        L.orc_rmse_literal.argtypes = [vp, vp, u32]
        L.orc_rmse_literal.restype = C.c_float
        L.orc_rmse_true.argtypes = [vp, vp, u32]
        L.orc_rmse_true.restype = C.c_double
        L.orc_float_to_half.argtypes = [vp, u32, vp]
        L.orc_max_threads.restype = i32
        _LIB = L
    return _LIB


def set_row_shard(row_first, row_stride):
    """OracleScene.render then covers only the image rows row_first, row_first + row_stride, ... ((0, 1) = every row)."""
    lib().orc_set_row_shard(int(row_first), int(row_stride))


def bdpt_set_only_s(s):
    """Diagnostic switch of oracle/bdpt.h: only the strategies with s light vertices, weight 1 (-1 restores the MIS weights)."""
    lib().orc_bdpt_set_only_s(int(s))


def bdpt_check_restore(on):
    """Diagnostic switch of oracle/bdpt.h: verify that calc_mis_weight restores every vertex it patches."""
    lib().orc_bdpt_set_check_restore(1 if on else 0)


def bdpt_restore_violations():
    return int(lib().orc_bdpt_restore_violations())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleScene:
    """CPU LBVH + renderer over a host Scene (keeps the Scene alive: the oracle borrows its arrays)."""

    def __init__(self, scene):
        self.scene = scene
        self._h = C.c_void_p()
        if lib().orc_scene_create(C.addressof(scene.desc), C.byref(self._h)) != 0:
            raise RuntimeError("orc_scene_create failed")
        self.n_tris = lib().orc_lbvh_num_tris(self._h)

    def lbvh(self):
        n, L, h = self.n_tris, lib(), self._h

        def arr(fn, count, dt):
            if count == 0:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(fn(h), shape=(count,)).copy()

        ni = max(n - 1, 0)
        nn = 2 * n - 1 if n > 0 else 0
        return dict(left=arr(L.orc_lbvh_left, ni, np.uint32), right=arr(L.orc_lbvh_right, ni, np.uint32),
                    parent=arr(L.orc_lbvh_parent, nn, np.uint32), leaf_prim=arr(L.orc_lbvh_leaf_prim, n, np.uint32),
                    morton=arr(L.orc_lbvh_morton, n, np.uint32), keys=arr(L.orc_lbvh_keys, n, np.uint64),
                    aabb=arr(L.orc_lbvh_aabb, 6 * nn, np.float32))

    def render(self, pc, ubo, first_frame, n_frames, rgba=None, threads=0):
        W, H = pc.size_x, pc.size_y
        if rgba is None:
            rgba = np.zeros((H, W, 4), dtype=np.float32)
        assert rgba.dtype == np.float32 and rgba.flags.c_contiguous
        st = Stats()
        lib().orc_render(self._h, C.addressof(pc), C.addressof(ubo), first_frame, n_frames, rgba.ctypes.data, C.addressof(st), threads)
        return rgba, st

    def render_frame_raw(self, pc, ubo, frame, threads=0):
        W, H = pc.size_x, pc.size_y
        rgb = np.zeros((H, W, 3), dtype=np.float32)
        st = Stats()
        lib().orc_render_frame_raw(self._h, C.addressof(pc), C.addressof(ubo), frame, rgb.ctypes.data, C.addressof(st), threads)
        return rgb, st

    def render_bdpt(self, pc, ubo, first_frame, n_frames, rgba=None, threads=0):
        """bdpt.rgen restated (oracle/bdpt.h); pc is a PCBdpt."""
        W, H = pc.size_x, pc.size_y
        if rgba is None:
            rgba = np.zeros((H, W, 4), dtype=np.float32)
        assert rgba.dtype == np.float32 and rgba.flags.c_contiguous
        st = Stats()
        rc = lib().orc_render_bdpt(self._h, C.addressof(pc), C.addressof(ubo), first_frame, n_frames, rgba.ctypes.data, C.addressof(st), threads)
        if rc != 0:
            raise RuntimeError(f"orc_render_bdpt failed ({rc})")
        return rgba, st

    def render_bdpt_frame_raw(self, pc, ubo, frame, threads=0):
        """(col, splat, stats) of one frame: the pixel's own strategies and the light-tracer image, before the film."""
        W, H = pc.size_x, pc.size_y
        col = np.zeros((H, W, 3), dtype=np.float32)
        splat = np.zeros((H, W, 3), dtype=np.float32)
        st = Stats()
        rc = lib().orc_render_bdpt_frame_raw(self._h, C.addressof(pc), C.addressof(ubo), frame, col.ctypes.data, splat.ctypes.data,
                                             C.addressof(st), threads)
        if rc != 0:
            raise RuntimeError(f"orc_render_bdpt_frame_raw failed ({rc})")
        return col, splat, st

    def trace_closest(self, rays, threads=0):
        rays = _f32(rays).reshape(-1, 8)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        st = Stats()
        lib().orc_trace_closest(self._h, rays.ctypes.data, rays.shape[0], hits.ctypes.data, C.addressof(st), threads)
        return hits, st

    def trace_closest_brute(self, rays, threads=0):
        """The hit definition over all triangles with no tree (O(rays x triangles))."""
        rays = _f32(rays).reshape(-1, 8)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        lib().orc_trace_closest_brute(self._h, rays.ctypes.data, rays.shape[0], hits.ctypes.data, threads)
        return hits

    def trace_any(self, rays, threads=0):
        rays = _f32(rays).reshape(-1, 8)
        occ = np.zeros(rays.shape[0], dtype=np.uint8)
        st = Stats()
        lib().orc_trace_any(self._h, rays.ctypes.data, rays.shape[0], occ.ctypes.data, C.addressof(st), threads)
        return occ, st

    def sample_light(self, num_lights, rands4, p3):
        r, p = _f32(rands4).reshape(-1, 4), _f32(p3).reshape(-1, 3)
        out = np.zeros((r.shape[0], 16), dtype=np.float32)
        lib().orc_kat_sample_light(self._h, num_lights, r.ctypes.data, p.ctypes.data, r.shape[0], out.ctypes.data)
        return out

    def light_Le(self, num_lights, total_light, rands6):
        """sample_light_Le (commons.glsl:335-406): (n, 16) = L, pos, wi, n, cos_from_light, pdf_pos_a, pdf_dir_w, flags."""
        r = _f32(rands6).reshape(-1, 6)
        out = np.zeros((r.shape[0], 16), dtype=np.float32)
        lib().orc_kat_light_Le(self._h, num_lights, total_light, r.ctypes.data, r.shape[0], out.ctypes.data)
        return out

    def texture(self, tex, uv):
        uv = _f32(uv).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), dtype=np.float32)
        lib().orc_kat_texture(self._h, tex, uv.ctypes.data, uv.shape[0], out.ctypes.data)
        return out

    def close(self):
        if self._h:
            lib().orc_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pcg4d(v4):
    v = np.ascontiguousarray(v4, dtype=np.uint32).reshape(-1, 4)
    out = np.zeros_like(v)
    lib().orc_kat_pcg4d(v.ctypes.data, v.shape[0], out.ctypes.data)
    return out


def rand(seed4, draws):
    s = np.ascontiguousarray(seed4, dtype=np.uint32).reshape(-1, 4)
    out = np.zeros((s.shape[0], draws), dtype=np.float32)
    lib().orc_kat_rand(s.ctypes.data, s.shape[0], draws, out.ctypes.data)
    return out


def detmath(x, y):
    x, y = _f32(x), _f32(y)
    outs = [np.zeros_like(x) for _ in range(4)]
    lib().orc_kat_detmath(x.ctypes.data, y.ctypes.data, x.size, *[o.ctypes.data for o in outs])
    return dict(sin=outs[0], cos=outs[1], exp=outs[2], pow=outs[3])


def offset_ray(p, n):
    p, n = _f32(p).reshape(-1, 3), _f32(n).reshape(-1, 3)
    a, b = np.zeros_like(p), np.zeros_like(p)
    lib().orc_kat_offset_ray(p.ctypes.data, n.ctypes.data, p.shape[0], a.ctypes.data, b.ctypes.data)
    return a, b


def sample_bsdf(mat, n_s, wo, rands, side):
    n_s, wo, rands = _f32(n_s).reshape(-1, 3), _f32(wo).reshape(-1, 3), _f32(rands).reshape(-1, 3)
    side = np.ascontiguousarray(side, dtype=np.uint8)
    out = np.zeros((n_s.shape[0], 8), dtype=np.float32)
    lib().orc_kat_sample_bsdf(C.addressof(mat), n_s.ctypes.data, wo.ctypes.data, rands.ctypes.data, side.ctypes.data, n_s.shape[0], out.ctypes.data)
    return out


def eval_bsdf(mat, n_s, wo, wi, side):
    n_s, wo, wi = _f32(n_s).reshape(-1, 3), _f32(wo).reshape(-1, 3), _f32(wi).reshape(-1, 3)
    side = np.ascontiguousarray(side, dtype=np.uint8)
    out = np.zeros((n_s.shape[0], 4), dtype=np.float32)
    lib().orc_kat_eval_bsdf(C.addressof(mat), n_s.ctypes.data, wo.ctypes.data, wi.ctypes.data, side.ctypes.data, n_s.shape[0], out.ctypes.data)
    return out


def bsdf_pdf(mat, n_s, wo, wi, side):
    """bsdf_pdf of bsdf_commons.glsl:26-66 (the stand-alone pdf functions): (n,) float32."""
    n_s, wo, wi = _f32(n_s).reshape(-1, 3), _f32(wo).reshape(-1, 3), _f32(wi).reshape(-1, 3)
    side = np.ascontiguousarray(side, dtype=np.uint8)
    out = np.zeros(n_s.shape[0], dtype=np.float32)
    lib().orc_kat_bsdf_pdf(C.addressof(mat), n_s.ctypes.data, wo.ctypes.data, wi.ctypes.data, side.ctypes.data, n_s.shape[0], out.ctypes.data)
    return out


def atmosphere(origin, direction, light_dir, light_L):
    o, d = _f32(origin).reshape(-1, 3), _f32(direction).reshape(-1, 3)
    ld, lL = _f32(light_dir).reshape(3), _f32(light_L).reshape(3)
    out = np.zeros_like(o)
    lib().orc_kat_atmosphere(o.ctypes.data, d.ctypes.data, ld.ctypes.data, lL.ctypes.data, o.shape[0], out.ctypes.data)
    return out


def rmse_literal(a, b):
    a, b = _f32(a).reshape(-1, 4), _f32(b).reshape(-1, 4)
    return float(lib().orc_rmse_literal(a.ctypes.data, b.ctypes.data, a.shape[0]))


def float_to_half(x):
    """tinyexr's float -> half conversion (what ImageUtils::save_exr stores); returns uint16 of x's shape."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(a.shape, dtype=np.uint16)
    lib().orc_float_to_half(a.ctypes.data, a.size, out.ctypes.data)
    return out


def rmse_true(a, b):
    a, b = _f32(a).reshape(-1, 4), _f32(b).reshape(-1, 4)
    return float(lib().orc_rmse_true(a.ctypes.data, b.ctypes.data, a.shape[0]))


def max_threads():
    return int(lib().orc_max_threads())
