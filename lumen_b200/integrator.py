"""ctypes face of liblumen_b200.so (include/lumen_b200.h) shaped like Lumen's integrator lifecycle.

Reference interface mirrored: `class Integrator` (src/RayTracer/Integrator.h:14-33) and `class Path`
(src/RayTracer/Path.h:4-23, Path.cpp:4-70): init() / render() / update() / destroy(), the public `frame_num`, and
`output_tex` (here: `output()` downloads the RGBA32F film). There is NO CPU fallback: if the CUDA library or a CUDA
device is missing, construction raises.
"""
import ctypes as C
import os

import numpy as np

from ._ctypes_types import Material, PCPath, SceneDesc, SceneUBO

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FILM_RUNNING_MEAN, FILM_SUM = 0, 1
HIT_DTYPE = np.dtype([("t", np.float32), ("b1", np.float32), ("b2", np.float32), ("prim", np.uint32)])


class Stats(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_probe", C.c_uint64), ("nodes_visited", C.c_uint64),
                ("tris_tested", C.c_uint64), ("nan_samples", C.c_uint64), ("frames", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ms_render", C.c_float), ("ms_extend", C.c_float), ("ms_shade", C.c_float), ("ms_connect", C.c_float), ("ms_film", C.c_float),
                ("ms_build_accel", C.c_float), ("ms_build_morton", C.c_float), ("ms_build_sort", C.c_float), ("ms_build_tree", C.c_float),
                ("ms_build_refit", C.c_float), ("ms_build_wide", C.c_float), ("wide_nodes", C.c_uint32), ("wide_levels", C.c_uint32),
                ("ploc_iterations", C.c_uint32), ("ms_build_ploc", C.c_float), ("tree_cost_ratio", C.c_float),
                ("trace_warp_iters", C.c_uint64), ("trace_node_trips", C.c_uint64), ("trace_tri_rounds", C.c_uint64), ("trace_refills", C.c_uint64)]

    @property
    def rays(self):
        return self.rays_closest + self.rays_shadow + self.rays_probe


def lib():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        # LMB_LIB: another build of the same library (A/B runs of kernel variants, tools/gpu_variants.sh)
        path = os.environ.get("LMB_LIB") or os.path.join(_HERE, "csrc", "liblumen_b200.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make -C lumen_b200/csrc` (nvcc, sm_100a). lumen_b200 has no CPU fallback.")
        L = C.CDLL(path)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int32
        L.lmb_last_error.argtypes = [vp]
        L.lmb_last_error.restype = C.c_char_p
        L.lmb_create.argtypes = [C.POINTER(vp), i32]
        L.lmb_destroy.argtypes = [vp]
        L.lmb_upload_scene.argtypes = [vp, vp]
        L.lmb_build_accel.argtypes = [vp]
        L.lmb_init.argtypes = [vp, u32, u32, u32]
        L.lmb_render.argtypes = [vp, vp, vp, u32, u32, u32, i32]
        L.lmb_render_bdpt.argtypes = [vp, vp, vp, u32, u32, u32, i32]
        L.lmb_clear_film.argtypes = [vp]
        L.lmb_set_pixel_shard.argtypes = [vp, C.c_uint32, C.c_uint32]
        L.lmb_resolve.argtypes = [vp]
        L.lmb_download.argtypes = [vp, vp]
        L.lmb_upload_film.argtypes = [vp, vp]
        L.lmb_film_add_from.argtypes = [vp, vp]
        L.lmb_comm_get_unique_id.argtypes = [vp]
        L.lmb_comm_init.argtypes = [vp, vp, i32, i32]
        L.lmb_comm_init_all.argtypes = [vp, i32]
        L.lmb_comm_destroy.argtypes = [vp]
        L.lmb_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.lmb_film_allreduce.argtypes = [vp, vp, i32]
        L.lmb_film_reduce.argtypes = [vp, i32, vp, i32]
        L.lmb_download_async.argtypes = [vp, vp]
        L.lmb_sync.argtypes = [vp]
        L.lmb_download_half_bgr.argtypes = [vp, vp]
        L.lmb_set_reference_image.argtypes = [vp, vp]
        L.lmb_rmse.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_double)]
        L.lmb_film_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
        L.lmb_stream.argtypes = [vp, C.POINTER(vp)]
        L.lmb_set_profile_stages.argtypes = [vp, i32]
        L.lmb_get_stats.argtypes = [vp, vp]
        L.lmb_reset_stats.argtypes = [vp]
        L.lmb_trace_closest.argtypes = [vp, vp, u32, vp]
        L.lmb_trace_any.argtypes = [vp, vp, u32, vp]
        L.lmb_trace_closest_device.argtypes = [vp, vp, u32, vp, u32, C.POINTER(C.c_float)]
        L.lmb_trace_closest_device_ex.argtypes = [vp, vp, u32, vp, u32, i32, C.POINTER(C.c_float)]
        L.lmb_accel_num_tris.argtypes = [vp, C.POINTER(u32)]
        L.lmb_accel_download.argtypes = [vp] * 8
        L.lmb_kat_pcg4d.argtypes = [vp, vp, u32, vp]
        L.lmb_kat_rand.argtypes = [vp, vp, u32, u32, vp]
        L.lmb_kat_detmath.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp]
        L.lmb_kat_offset_ray.argtypes = [vp, vp, vp, u32, vp, vp]
        L.lmb_kat_sample_bsdf.argtypes = [vp, vp, vp, vp, vp, vp, u32, vp]
        L.lmb_kat_eval_bsdf.argtypes = [vp, vp, vp, vp, vp, vp, u32, vp]
        L.lmb_kat_atmosphere.argtypes = [vp, vp, vp, vp, vp, u32, vp]
        L.lmb_kat_sample_light.argtypes = [vp, i32, vp, vp, u32, vp]
        L.lmb_kat_texture.argtypes = [vp, u32, vp, u32, vp]
        L.lmb_kat_wide_bvh_check.argtypes = [vp, vp]
        L.lmb_kat_bdpt_frame_raw.argtypes = [vp, vp, vp, u32, vp, vp]
        _LIB = L
    return _LIB


EXPORTS = ["lmb_create", "lmb_destroy", "lmb_last_error", "lmb_upload_scene", "lmb_build_accel", "lmb_init", "lmb_render", "lmb_render_bdpt", "lmb_set_pixel_shard", "lmb_clear_film",
           "lmb_resolve", "lmb_download", "lmb_upload_film", "lmb_film_add_from", "lmb_film_device_ptr", "lmb_stream", "lmb_set_profile_stages", "lmb_get_stats",
           "lmb_reset_stats", "lmb_trace_closest", "lmb_trace_any", "lmb_trace_closest_device", "lmb_accel_num_tris", "lmb_accel_download",
           "lmb_download_async", "lmb_sync", "lmb_download_half_bgr", "lmb_set_reference_image", "lmb_rmse",
           "lmb_trace_closest_device_ex", "lmb_comm_get_unique_id", "lmb_comm_init", "lmb_comm_init_all", "lmb_comm_destroy", "lmb_comm_info", "lmb_film_allreduce", "lmb_film_reduce"]
TESTHOOK_EXPORTS = ["lmb_kat_pcg4d", "lmb_kat_rand", "lmb_kat_detmath", "lmb_kat_offset_ray", "lmb_kat_sample_bsdf", "lmb_kat_eval_bsdf",
                    "lmb_kat_atmosphere", "lmb_kat_sample_light", "lmb_kat_texture", "lmb_kat_wide_bvh_check", "lmb_kat_bdpt_frame_raw"]


COMM_ID_BYTES = 128


def comm_unique_id():
    """ncclGetUniqueId through the C ABI: rank 0 calls it and hands the bytes to the other ranks."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    rc = lib().lmb_comm_get_unique_id(C.addressof(buf))
    if rc != 0:
        raise RuntimeError(f"lmb_comm_get_unique_id failed ({rc}): " + lib().lmb_last_error(None).decode())
    return bytes(buf)


def comm_init_all(devices):
    """One communicator over the Devices of THIS process (one per GPU): rank i = devices[i]."""
    arr = (C.c_void_p * len(devices))(*[d._h for d in devices])
    rc = lib().lmb_comm_init_all(C.addressof(arr), len(devices))
    if rc != 0:
        msgs = [lib().lmb_last_error(d._h).decode() for d in devices]  # the message sits in the context that was refused
        raise RuntimeError(f"lmb_comm_init_all failed ({rc}): " + "; ".join(m for m in msgs if m))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Device:
    """An lmb_ctx: one GPU, one stream."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().lmb_create(C.byref(self._h), int(device))
        if rc != 0:
            raise RuntimeError(f"lmb_create failed ({rc}): " + lib().lmb_last_error(None).decode())
        self.device = int(device)

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): " + lib().lmb_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().lmb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene / accel
    def upload_scene(self, desc):
        self._ck(lib().lmb_upload_scene(self._h, C.addressof(desc)), "lmb_upload_scene")

    def build_accel(self):
        self._ck(lib().lmb_build_accel(self._h), "lmb_build_accel")

    def lbvh(self):
        n = C.c_uint32()
        self._ck(lib().lmb_accel_num_tris(self._h, C.byref(n)), "lmb_accel_num_tris")
        n = n.value
        ni, nn = max(n - 1, 0), (2 * n - 1 if n else 0)
        out = dict(left=np.zeros(ni, np.uint32), right=np.zeros(ni, np.uint32), parent=np.zeros(nn, np.uint32), leaf_prim=np.zeros(n, np.uint32),
                   morton=np.zeros(n, np.uint32), keys=np.zeros(n, np.uint64), aabb=np.zeros(6 * nn, np.float32))
        self._ck(lib().lmb_accel_download(self._h, *[out[k].ctypes.data for k in ("left", "right", "parent", "leaf_prim", "morton", "keys", "aabb")]),
                 "lmb_accel_download")
        return out

    # ---- film / render
    def init(self, width, height, frames_in_flight=0):
        self.width, self.height = int(width), int(height)
        self._ck(lib().lmb_init(self._h, self.width, self.height, int(frames_in_flight)), "lmb_init")

    def set_pixel_shard(self, row_first, row_stride):
        """This context renders only image rows row_first + k * row_stride. Call before init()."""
        self._ck(lib().lmb_set_pixel_shard(self._h, int(row_first), int(row_stride)), "lmb_set_pixel_shard")

    def render(self, pc, ubo, first_frame, n_frames, frame_stride=1, film_mode=FILM_RUNNING_MEAN):
        self._ck(lib().lmb_render(self._h, C.addressof(pc), C.addressof(ubo), int(first_frame), int(n_frames), int(frame_stride), int(film_mode)),
                 "lmb_render")

    def render_bdpt(self, pc, ubo, first_frame, n_frames, frame_stride=1, film_mode=FILM_RUNNING_MEAN):
        """BDPT::render for frames first_frame + k * frame_stride, k < n_frames; pc is a PCBdpt."""
        self._ck(lib().lmb_render_bdpt(self._h, C.addressof(pc), C.addressof(ubo), int(first_frame), int(n_frames), int(frame_stride), int(film_mode)),
                 "lmb_render_bdpt")

    def kat_bdpt_frame_raw(self, pc, ubo, frame):
        """One BDPT frame (film updated) + the two images it is made of: (col (H, W, 3), splat (H, W, 3))."""
        col = np.zeros((pc.size_y, pc.size_x, 4), dtype=np.float32)
        splat = np.zeros((pc.size_y, pc.size_x, 3), dtype=np.float32)
        self._ck(lib().lmb_kat_bdpt_frame_raw(self._h, C.addressof(pc), C.addressof(ubo), int(frame), col.ctypes.data, splat.ctypes.data),
                 "lmb_kat_bdpt_frame_raw")
        return col[..., :3].copy(), splat

    def clear_film(self):
        self._ck(lib().lmb_clear_film(self._h), "lmb_clear_film")

    def resolve(self):
        self._ck(lib().lmb_resolve(self._h), "lmb_resolve")

    def download(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(lib().lmb_download(self._h, out.ctypes.data), "lmb_download")
        return out

    def download_into(self, host_ptr):
        self._ck(lib().lmb_download(self._h, host_ptr), "lmb_download")

    def download_async(self, host_ptr):
        self._ck(lib().lmb_download_async(self._h, host_ptr), "lmb_download_async")

    def sync(self):
        self._ck(lib().lmb_sync(self._h), "lmb_sync")

    def download_half_bgr(self):
        """(3, H, W) uint16: the B, G, R half planes of the EXR writer, converted on the device."""
        out = np.empty((3, self.height, self.width), dtype=np.uint16)
        self._ck(lib().lmb_download_half_bgr(self._h, out.ctypes.data), "lmb_download_half_bgr")
        return out

    def set_reference_image(self, rgba):
        a = _f32(rgba)
        assert a.size == self.width * self.height * 4
        self._ck(lib().lmb_set_reference_image(self._h, a.ctypes.data), "lmb_set_reference_image")

    def rmse(self):
        """(literal, true): Lumen's RMSE routine and a true RMSE of the film against the reference image."""
        lit, tru = C.c_float(), C.c_double()
        self._ck(lib().lmb_rmse(self._h, C.byref(lit), C.byref(tru)), "lmb_rmse")
        return lit.value, tru.value

    def upload_film(self, rgba):
        a = _f32(rgba)
        assert a.size == self.width * self.height * 4
        self._ck(lib().lmb_upload_film(self._h, a.ctypes.data), "lmb_upload_film")

    def film_add_from(self, other):
        """self.film += other.film (sum films of a multi-GPU render; device-to-device copy)."""
        self._ck(lib().lmb_film_add_from(self._h, other._h), "lmb_film_add_from")

    # ---- the exchange step of a multi-GPU render (include/lumen_b200.h lmb_comm_*): NCCL sum of the LMB_FILM_SUM films + resolve
    def comm_init(self, unique_id, rank, n_ranks):
        """Collective: every rank passes the 128 bytes rank 0 got from comm_unique_id()."""
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._ck(lib().lmb_comm_init(self._h, C.addressof(buf), int(rank), int(n_ranks)), "lmb_comm_init")

    def comm_destroy(self):
        self._ck(lib().lmb_comm_destroy(self._h), "lmb_comm_destroy")

    def comm_info(self):
        """(rank, n_ranks, nccl_version); n_ranks = 0 without a communicator."""
        r, n, v = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(lib().lmb_comm_info(self._h, C.byref(r), C.byref(n), C.byref(v)), "lmb_comm_info")
        return r.value, n.value, v.value

    def film_allreduce(self, out_ptr=None, clear_film=False):
        """film -> NCCL sum over the ranks -> rgb / alpha. out_ptr None: in place on the render stream; else the snapshot is reduced
        on the comm stream and copied to out_ptr (host or device address) while rendering goes on; sync() waits."""
        self._ck(lib().lmb_film_allreduce(self._h, out_ptr, 1 if clear_film else 0), "lmb_film_allreduce")

    def film_reduce(self, root, out_ptr=None, clear_film=False):
        """film_allreduce with ONE receiver (ncclReduce): rank `root` gets the resolved image, the others only contribute (their
        out_ptr is ignored; clear_film alone selects the snapshot form there)."""
        self._ck(lib().lmb_film_reduce(self._h, int(root), out_ptr, 1 if clear_film else 0), "lmb_film_reduce")

    def film_device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(lib().lmb_film_device_ptr(self._h, C.byref(p), C.byref(n)), "lmb_film_device_ptr")
        return p.value, n.value

    def stream(self):
        p = C.c_void_p()
        self._ck(lib().lmb_stream(self._h, C.byref(p)), "lmb_stream")
        return p.value or 0

    def set_profile_stages(self, on):
        self._ck(lib().lmb_set_profile_stages(self._h, 1 if on else 0), "lmb_set_profile_stages")

    def stats(self):
        st = Stats()
        self._ck(lib().lmb_get_stats(self._h, C.addressof(st)), "lmb_get_stats")
        return st

    def reset_stats(self):
        self._ck(lib().lmb_reset_stats(self._h), "lmb_reset_stats")

    # ---- ray queries
    def trace_closest(self, rays):
        rays = _f32(rays).reshape(-1, 8)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        self._ck(lib().lmb_trace_closest(self._h, rays.ctypes.data, rays.shape[0], hits.ctypes.data), "lmb_trace_closest")
        return hits

    def trace_any(self, rays):
        rays = _f32(rays).reshape(-1, 8)
        occ = np.zeros(rays.shape[0], dtype=np.uint8)
        self._ck(lib().lmb_trace_any(self._h, rays.ctypes.data, rays.shape[0], occ.ctypes.data), "lmb_trace_any")
        return occ

    def trace_closest_device(self, d_rays_ptr, n, d_hits_ptr, repeat=1, sort_rays=False):
        """ms of `repeat` launches over device arrays; sort_rays: order the rays by (origin cell, direction bin) inside every launch"""
        ms = C.c_float()
        self._ck(lib().lmb_trace_closest_device_ex(self._h, d_rays_ptr, int(n), d_hits_ptr, int(repeat), 1 if sort_rays else 0, C.byref(ms)),
                 "lmb_trace_closest_device_ex")
        return ms.value

    # ---- known-answer probes (include/lumen_b200_testhooks.h)
    def kat_pcg4d(self, v4):
        v = np.ascontiguousarray(v4, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros_like(v)
        self._ck(lib().lmb_kat_pcg4d(self._h, v.ctypes.data, v.shape[0], out.ctypes.data), "lmb_kat_pcg4d")
        return out

    def kat_rand(self, seed4, draws):
        s = np.ascontiguousarray(seed4, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros((s.shape[0], draws), dtype=np.float32)
        self._ck(lib().lmb_kat_rand(self._h, s.ctypes.data, s.shape[0], draws, out.ctypes.data), "lmb_kat_rand")
        return out

    def kat_detmath(self, x, y):
        x, y = _f32(x), _f32(y)
        outs = [np.zeros_like(x) for _ in range(4)]
        self._ck(lib().lmb_kat_detmath(self._h, x.ctypes.data, y.ctypes.data, x.size, *[o.ctypes.data for o in outs]), "lmb_kat_detmath")
        return dict(sin=outs[0], cos=outs[1], exp=outs[2], pow=outs[3])

    def kat_offset_ray(self, p, n):
        p, n = _f32(p).reshape(-1, 3), _f32(n).reshape(-1, 3)
        a, b = np.zeros_like(p), np.zeros_like(p)
        self._ck(lib().lmb_kat_offset_ray(self._h, p.ctypes.data, n.ctypes.data, p.shape[0], a.ctypes.data, b.ctypes.data), "lmb_kat_offset_ray")
        return a, b

    def kat_sample_bsdf(self, mat, n_s, wo, rands, side):
        n_s, wo, rands = _f32(n_s).reshape(-1, 3), _f32(wo).reshape(-1, 3), _f32(rands).reshape(-1, 3)
        side = np.ascontiguousarray(side, dtype=np.uint8)
        out = np.zeros((n_s.shape[0], 8), dtype=np.float32)
        self._ck(lib().lmb_kat_sample_bsdf(self._h, C.addressof(mat), n_s.ctypes.data, wo.ctypes.data, rands.ctypes.data, side.ctypes.data,
                                           n_s.shape[0], out.ctypes.data), "lmb_kat_sample_bsdf")
        return out

    def kat_eval_bsdf(self, mat, n_s, wo, wi, side):
        n_s, wo, wi = _f32(n_s).reshape(-1, 3), _f32(wo).reshape(-1, 3), _f32(wi).reshape(-1, 3)
        side = np.ascontiguousarray(side, dtype=np.uint8)
        out = np.zeros((n_s.shape[0], 4), dtype=np.float32)
        self._ck(lib().lmb_kat_eval_bsdf(self._h, C.addressof(mat), n_s.ctypes.data, wo.ctypes.data, wi.ctypes.data, side.ctypes.data,
                                         n_s.shape[0], out.ctypes.data), "lmb_kat_eval_bsdf")
        return out

    def kat_atmosphere(self, origin, direction, light_dir, light_L):
        o, d = _f32(origin).reshape(-1, 3), _f32(direction).reshape(-1, 3)
        ld, lL = _f32(light_dir).reshape(3), _f32(light_L).reshape(3)
        out = np.zeros_like(o)
        self._ck(lib().lmb_kat_atmosphere(self._h, o.ctypes.data, d.ctypes.data, ld.ctypes.data, lL.ctypes.data, o.shape[0], out.ctypes.data),
                 "lmb_kat_atmosphere")
        return out

    def kat_sample_light(self, num_lights, rands4, p3):
        r, p = _f32(rands4).reshape(-1, 4), _f32(p3).reshape(-1, 3)
        out = np.zeros((r.shape[0], 16), dtype=np.float32)
        self._ck(lib().lmb_kat_sample_light(self._h, int(num_lights), r.ctypes.data, p.ctypes.data, r.shape[0], out.ctypes.data), "lmb_kat_sample_light")
        return out

    def kat_texture(self, tex, uv):
        uv = _f32(uv).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), dtype=np.float32)
        self._ck(lib().lmb_kat_texture(self._h, int(tex), uv.ctypes.data, uv.shape[0], out.ctypes.data), "lmb_kat_texture")
        return out

    def wide_bvh_check(self):
        """Host-side structural validation of the 8-wide traversal BVH (see lumen_b200_testhooks.h)."""
        out = np.zeros(8, dtype=np.uint64)
        self._ck(lib().lmb_kat_wide_bvh_check(self._h, out.ctypes.data), "lmb_kat_wide_bvh_check")
        keys = ("nodes", "reachable", "depth", "errors", "dup_or_missing", "internal_children", "leaf_children", "leaf_tris")
        return dict(zip(keys, (int(v) for v in out)))


class BDPTB200:
    """Drop-in for Lumen's `BDPT` integrator behind the same lifecycle (BDPT.cpp:4-108).

    init()    -> Integrator::init + BDPT::init + create_accel (the vertex buffers of BDPT.cpp:7-24 are allocated by the first render)
    render(n) -> BDPT::render for frames [frame_num, frame_num + n): PCBDPT filled as BDPT.cpp:56-64; `time` is this object's
                 (BDPT.cpp:57 draws rand() % UINT_MAX: set .time before render() to do the same; the default 0 keeps renders reproducible)
    update()  -> BDPT::update: advances frame_num
    destroy() -> BDPT::destroy
    """

    def __init__(self, scene, device=0):
        self.scene = scene
        self.dev = Device(device)
        self.frame_num = 0
        self.path_length = scene.info.path_length
        self.time = 0
        self._pending = 0

    def init(self):
        self.dev.upload_scene(self.scene.desc)
        self.dev.build_accel()
        self.dev.init(self.scene.width, self.scene.height, 1)
        self.ubo = self.scene.make_ubo()
        self.frame_num = 0

    def render(self, n_frames=1):
        from ._ctypes_types import PCBdpt
        pc = PCBdpt.from_path_pc(self.scene.make_pc(self.path_length, True), self.time)
        pc.frame_num = self.frame_num
        self.dev.render_bdpt(pc, self.ubo, self.frame_num, n_frames)
        self._pending = n_frames

    def update(self):
        self.frame_num += self._pending
        self._pending = 0
        return False

    def output(self):
        return self.dev.download()

    def destroy(self):
        self.dev.close()


class PathB200:
    """Drop-in for Lumen's `Path` integrator behind the same lifecycle (Path.cpp:4-70).

    init()    -> Integrator::init + Path::init + create_accel: upload scene, build the LBVH, allocate the film
    render(n) -> Path::render for frames [frame_num, frame_num + n): fills PCPath from the scene exactly as Path.cpp:27-38
    update()  -> Path::update: advances frame_num (returns False: the headless camera never moves)
    destroy() -> Path::destroy
    """

    def __init__(self, scene, device=0, frames_in_flight=0):
        self.scene = scene
        self.dev = Device(device)
        self.frames_in_flight = frames_in_flight
        self.frame_num = 0
        self.path_length = scene.info.path_length
        self.direct_lighting = True
        self._pending = 0

    def init(self):
        self.dev.upload_scene(self.scene.desc)
        self.dev.build_accel()
        self.dev.init(self.scene.width, self.scene.height, self.frames_in_flight)
        self.ubo = self.scene.make_ubo()
        self.frame_num = 0

    def render(self, n_frames=1):
        pc = self.scene.make_pc(self.path_length, self.direct_lighting)
        pc.frame_num = self.frame_num
        self.dev.render(pc, self.ubo, self.frame_num, n_frames, 1, FILM_RUNNING_MEAN)
        self._pending = n_frames

    def update(self):
        self.frame_num += self._pending
        self._pending = 0
        return False

    def output(self):
        return self.dev.download()

    def destroy(self):
        self.dev.close()
