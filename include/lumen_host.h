/*
 * lumen_host.h -- C entry points of the host shell (liblumen_host.so): scene loading, camera / push-constant setup and
 * EXR output, i.e. the parts of Lumen's RayTracer / LumenScene / ImageUtils that sit either side of the Path
 * integrator. Plain C so that Python (ctypes) tests and bench.py drive exactly the code the C++ headless renderer uses.
 */
#ifndef LUMEN_HOST_H
#define LUMEN_HOST_H
#include <stdint.h>
#include "lmb_types.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lmh_scene lmh_scene;

typedef struct lmh_scene_info {
	int32_t path_length;  /* SceneConfig::path_length */
	float sky_col[3];
	char integrator[32];  /* lower-cased integrator.type from the file (e.g. "vcm" for scenes/caustics.json) */
	uint32_t n_triangles, n_prim_meshes, n_materials, n_lights, n_textures;
	uint32_t total_light_triangle_cnt;
	float total_light_area;
	uint32_t dir_light_idx;
	uint32_t bsdf_types;  /* OR of Material::bsdf_type over the scene (ENABLE_* macros, LumenScene.cpp:217-228) */
	float world_radius;
} lmh_scene_info;

/* RayTracer::init -> LumenScene::load_scene (RayTracer.cpp:60, LumenScene.cpp:52). *.json or *.xml. */
int lmh_scene_load(const char* path, uint32_t width, uint32_t height, lmh_scene** out);
/* Procedural ingest: caller supplies de-indexed geometry (3 vertices per triangle) per mesh plus materials, analytic
 * lights and camera; runs the same finalize step as the file loaders. mesh_tri_counts[n_meshes], mesh_materials[n_meshes],
 * mesh_world[n_meshes*16] (column-major, may be NULL for identity). cam_pos/cam_dir as in the JSON schema. */
int lmh_scene_from_arrays(const lmb_vertex* vertices, uint32_t n_vertices, const uint32_t* mesh_tri_counts, const uint32_t* mesh_materials,
						  const float* mesh_world, uint32_t n_meshes, const lmb_material* materials, uint32_t n_materials,
						  const lmb_light* analytic_lights, uint32_t n_analytic, float fov, const float* cam_pos, const float* cam_dir,
						  int32_t path_length, const float* sky_col, uint32_t width, uint32_t height, lmh_scene** out);
void lmh_scene_destroy(lmh_scene* s);
void lmh_scene_get_desc(const lmh_scene* s, lmb_scene_desc* out);
void lmh_scene_get_info(const lmh_scene* s, lmh_scene_info* out);
/* Path::render's PCPath fill (Path.cpp:27-38); max_depth <= 0 takes the scene's path_length. */
void lmh_scene_make_pc(const lmh_scene* s, int32_t max_depth, int32_t direct_lighting, lmb_pc_path* out);
/* Integrator::update_uniform_buffers (Integrator.cpp:60-72) */
void lmh_scene_make_ubo(lmh_scene* s, lmb_scene_ubo* out);
int lmh_save_exr(const float* rgba, int32_t width, int32_t height, const char* path);
/* The same EXR from B, G, R planes already converted to HALF (lmb_download_half_bgr); file bytes equal lmh_save_exr's. */
int lmh_save_exr_half_bgr(const uint16_t* planes, int32_t width, int32_t height, const char* path);
/* Checkpoint of a progressive render (running-mean film + frames accumulated + path length); load allocates *rgba_out
 * (free with lmh_free). Resuming = lmb_upload_film + continuing at frame `frames`: bit-identical to an uninterrupted run. */
int lmh_save_checkpoint(const char* path, const float* rgba, uint32_t width, uint32_t height, uint32_t frames, uint32_t path_length);
int lmh_load_checkpoint(const char* path, float** rgba_out, uint32_t* width, uint32_t* height, uint32_t* frames, uint32_t* path_length);
int lmh_load_exr(const char* path, float** rgba_out, int32_t* width, int32_t* height); /* free with lmh_free */
void lmh_free(void* p);
const char* lmh_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
