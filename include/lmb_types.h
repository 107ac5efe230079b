/*
 * lmb_types.h -- plain-C byte layouts shared by the host shell, the CUDA library and the CPU oracle.
 *
 * These are the structures Lumen uploads for its Path integrator, restated as tightly packed C structs
 * (GLSL `scalar` block layout == C layout when every member is 4 bytes wide):
 *   Vertex / Light / Material / PrimMeshInfo / SceneUBO   reference: src/shaders/commons.h:180-340
 *   PCPath (push constants)                              reference: src/shaders/integrators/path/path_commons.h:3-14
 * Sizes are pinned by static asserts below and by tests/test_layout.py.
 */
#ifndef LMB_TYPES_H
#define LMB_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* BSDF type bits (commons.h:13-18) */
#define LMB_BSDF_DIFFUSE 1u
#define LMB_BSDF_MIRROR 2u
#define LMB_BSDF_GLASS 4u
#define LMB_BSDF_DIELECTRIC 8u
#define LMB_BSDF_CONDUCTOR 16u
#define LMB_BSDF_PRINCIPLED 32u

/* BSDF property bits (commons.h:22-26) */
#define LMB_FLAG_DIFFUSE 1u
#define LMB_FLAG_SPECULAR 2u
#define LMB_FLAG_GLOSSY 4u
#define LMB_FLAG_REFLECTION 8u
#define LMB_FLAG_TRANSMISSION 16u

/* Light types: low 3 bits of light_flags; bit 4 = finite, bit 5 = delta (commons.h:38-40, commons.glsl:42-46) */
#define LMB_LIGHT_SPOT 1u
#define LMB_LIGHT_AREA 2u
#define LMB_LIGHT_DIRECTIONAL 3u
#define LMB_LIGHT_FINITE_BIT (1u << 4)
#define LMB_LIGHT_DELTA_BIT (1u << 5)

typedef struct lmb_vertex { /* 32 B, commons.h:194-198 */
	float pos[3];
	float normal[3];
	float uv0[2];
} lmb_vertex;

typedef struct lmb_light { /* 128 B, commons.h:200-210 */
	float world_matrix[16]; /* column-major */
	float pos[3];
	uint32_t prim_mesh_idx;
	float to[3];
	uint32_t num_triangles;
	float L[3];
	uint32_t light_flags;
	float world_center[3];
	float world_radius;
} lmb_light;

typedef struct lmb_material { /* 104 B, commons.h:212-234 */
	float albedo[3];
	float ior;
	float emissive_factor[3];
	uint32_t bsdf_type;
	uint32_t bsdf_props;
	float k[3];
	int32_t texture_id;
	float roughness;
	float diffuse_trans;
	float spec_trans;
	float metallic;
	float specular_tint;
	float sheen_tint;
	float clearcoat;
	float clearcoat_gloss;
	float sheen;
	float subsurface;
	float flatness;
	float anisotropy;
	uint32_t thin;
} lmb_material;

typedef struct lmb_prim_mesh_info { /* 48 B, commons.h:333-340 */
	uint32_t index_offset;
	uint32_t vertex_offset;
	uint32_t material_index;
	uint32_t pad;
	float min_pos[4];
	float max_pos[4];
} lmb_prim_mesh_info;

typedef struct lmb_pc_path { /* 52 B, path_commons.h:3-14 */
	float sky_col[3];
	uint32_t frame_num;
	uint32_t size_x;
	uint32_t size_y;
	int32_t num_lights;
	uint32_t time;
	int32_t max_depth;
	float total_light_area;
	int32_t light_triangle_count;
	uint32_t dir_light_idx;
	uint32_t direct_lighting;
} lmb_pc_path;

typedef struct lmb_pc_bdpt { /* 48 B, integrators/bdpt/bdpt_commons.h:5-16 (PCPath without direct_lighting) */
	float sky_col[3];
	uint32_t frame_num;
	uint32_t size_x;
	uint32_t size_y;
	int32_t num_lights;
	uint32_t time; /* BDPT.cpp:57 rand() % UINT_MAX per frame; enters the RNG seed as frame_num ^ time (bdpt.rgen:36-37) */
	int32_t max_depth;
	float total_light_area;
	int32_t light_triangle_count;
	uint32_t dir_light_idx;
} lmb_pc_bdpt;

typedef struct lmb_scene_ubo { /* 492 B, commons.h:180-192; matrices column-major */
	float projection[16];
	float view[16];
	float model[16];
	float inv_view[16];
	float inv_projection[16];
	float light_pos[4];
	float view_pos[4];
	float prev_view[16];
	float prev_projection[16];
	int32_t clicked_pos[2];
	int32_t debug_click;
} lmb_scene_ubo;

/* One RGBA8 texture, sRGB-encoded colour (LumenScene.cpp:204-213: VK_FORMAT_R8G8B8A8_SRGB, stbi 4 channels). */
typedef struct lmb_texture {
	const uint8_t* rgba8;
	uint32_t width;
	uint32_t height;
} lmb_texture;

/*
 * Everything LumenScene::load_scene uploads for the Path integrator (LumenScene.cpp:134-216) plus the two per-mesh
 * items Lumen hands to the acceleration-structure build (Integrator.cpp:137-160): the instance transform and the
 * index count. Pointers are HOST pointers owned by the caller; lmb_upload_scene copies them.
 */
typedef struct lmb_scene_desc {
	const lmb_vertex* vertices; /* "Compact Vertices Buffer" */
	uint32_t n_vertices;
	const uint32_t* indices; /* "Index Buffer"; mesh-local, add vertex_offset */
	uint32_t n_indices;
	const lmb_material* materials;
	uint32_t n_materials;
	const lmb_prim_mesh_info* prim_infos; /* "Prim Lookup Buffer" */
	uint32_t n_prim_meshes;
	const uint32_t* prim_idx_counts;   /* LumenPrimMesh::idx_count per mesh */
	const float* world_matrices;       /* n_prim_meshes x 16, column-major (LumenPrimMesh::world_matrix) */
	const float* inv_world_matrices;   /* n_prim_meshes x 16, glm::inverse(world_matrix): stands in for gl_WorldToObjectEXT */
	const lmb_light* lights;           /* "Mesh Lights Buffer" (gpu_lights) */
	uint32_t n_lights;
	const lmb_texture* textures;
	uint32_t n_textures;
} lmb_scene_desc;

#ifdef __cplusplus
}
static_assert(sizeof(lmb_vertex) == 32, "Vertex layout");
static_assert(sizeof(lmb_light) == 128, "Light layout");
static_assert(sizeof(lmb_material) == 104, "Material layout");
static_assert(sizeof(lmb_prim_mesh_info) == 48, "PrimMeshInfo layout");
static_assert(sizeof(lmb_pc_path) == 52, "PCPath layout");
static_assert(sizeof(lmb_pc_bdpt) == 48, "PCBDPT layout");
static_assert(sizeof(lmb_scene_ubo) == 492, "SceneUBO layout");
#endif

#endif /* LMB_TYPES_H */
