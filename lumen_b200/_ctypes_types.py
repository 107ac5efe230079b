"""ctypes mirrors of include/lmb_types.h (reference layouts: src/shaders/commons.h:180-340, path_commons.h:3-14)."""
import ctypes as C

f32, u32, i32 = C.c_float, C.c_uint32, C.c_int32


class Vertex(C.Structure):
    _fields_ = [("pos", f32 * 3), ("normal", f32 * 3), ("uv0", f32 * 2)]


class Light(C.Structure):
    _fields_ = [("world_matrix", f32 * 16), ("pos", f32 * 3), ("prim_mesh_idx", u32), ("to", f32 * 3), ("num_triangles", u32),
                ("L", f32 * 3), ("light_flags", u32), ("world_center", f32 * 3), ("world_radius", f32)]


class Material(C.Structure):
    _fields_ = [("albedo", f32 * 3), ("ior", f32), ("emissive_factor", f32 * 3), ("bsdf_type", u32), ("bsdf_props", u32),
                ("k", f32 * 3), ("texture_id", i32), ("roughness", f32), ("diffuse_trans", f32), ("spec_trans", f32),
                ("metallic", f32), ("specular_tint", f32), ("sheen_tint", f32), ("clearcoat", f32), ("clearcoat_gloss", f32),
                ("sheen", f32), ("subsurface", f32), ("flatness", f32), ("anisotropy", f32), ("thin", u32)]


class PrimMeshInfo(C.Structure):
    _fields_ = [("index_offset", u32), ("vertex_offset", u32), ("material_index", u32), ("pad", u32), ("min_pos", f32 * 4),
                ("max_pos", f32 * 4)]


class PCPath(C.Structure):
    _fields_ = [("sky_col", f32 * 3), ("frame_num", u32), ("size_x", u32), ("size_y", u32), ("num_lights", i32), ("time", u32),
                ("max_depth", i32), ("total_light_area", f32), ("light_triangle_count", i32), ("dir_light_idx", u32),
                ("direct_lighting", u32)]


class PCBdpt(C.Structure):
    """PCBDPT, integrators/bdpt/bdpt_commons.h:5-16 (48 B): PCPath without direct_lighting."""
    _fields_ = [("sky_col", f32 * 3), ("frame_num", u32), ("size_x", u32), ("size_y", u32), ("num_lights", i32), ("time", u32),
                ("max_depth", i32), ("total_light_area", f32), ("light_triangle_count", i32), ("dir_light_idx", u32)]

    @classmethod
    def from_path_pc(cls, pc, time=0):
        """BDPT.cpp:55-66 fills the same fields Path.cpp:27-38 does (time is the caller's: oracle/bdpt.h quirk B1)."""
        out = cls()
        C.memmove(C.addressof(out), C.addressof(pc), C.sizeof(cls))
        out.time = time
        return out


class SceneUBO(C.Structure):
    _fields_ = [("projection", f32 * 16), ("view", f32 * 16), ("model", f32 * 16), ("inv_view", f32 * 16),
                ("inv_projection", f32 * 16), ("light_pos", f32 * 4), ("view_pos", f32 * 4), ("prev_view", f32 * 16),
                ("prev_projection", f32 * 16), ("clicked_pos", i32 * 2), ("debug_click", i32)]


class Texture(C.Structure):
    _fields_ = [("rgba8", C.c_void_p), ("width", u32), ("height", u32)]


class SceneDesc(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("n_vertices", u32), ("indices", C.c_void_p), ("n_indices", u32),
                ("materials", C.c_void_p), ("n_materials", u32), ("prim_infos", C.c_void_p), ("n_prim_meshes", u32),
                ("prim_idx_counts", C.c_void_p), ("world_matrices", C.c_void_p), ("inv_world_matrices", C.c_void_p),
                ("lights", C.c_void_p), ("n_lights", u32), ("textures", C.c_void_p), ("n_textures", u32)]


class SceneInfo(C.Structure):
    _fields_ = [("path_length", i32), ("sky_col", f32 * 3), ("integrator", C.c_char * 32), ("n_triangles", u32),
                ("n_prim_meshes", u32), ("n_materials", u32), ("n_lights", u32), ("n_textures", u32),
                ("total_light_triangle_cnt", u32), ("total_light_area", f32), ("dir_light_idx", u32), ("bsdf_types", u32),
                ("world_radius", f32)]


SIZES = {Vertex: 32, Light: 128, Material: 104, PrimMeshInfo: 48, PCPath: 52, SceneUBO: 492}
for _t, _n in SIZES.items():
    assert C.sizeof(_t) == _n, (_t, C.sizeof(_t), _n)

BSDF_DIFFUSE, BSDF_MIRROR, BSDF_GLASS, BSDF_DIELECTRIC, BSDF_CONDUCTOR, BSDF_PRINCIPLED = 1, 2, 4, 8, 16, 32
FLAG_DIFFUSE, FLAG_SPECULAR, FLAG_GLOSSY, FLAG_REFLECTION, FLAG_TRANSMISSION = 1, 2, 4, 8, 16
LIGHT_SPOT, LIGHT_AREA, LIGHT_DIRECTIONAL = 1, 2, 3
LIGHT_FINITE_BIT, LIGHT_DELTA_BIT = 1 << 4, 1 << 5
