/* ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or executed from lumen_b200/).
 *
 * C entry points of oracle/_ref/libglslref.so: the reference's OWN shader source for the Path integrator
 * (src/shaders/integrators/path/path.rgen + everything it includes, ray.rchit, ray.rmiss, ray_shadow.rmiss), translated
 * mechanically to C++ by glsl2cpp.py from the unmodified files under /root/reference and compiled here. It exists to PIN
 * oracle/liboracle.so (the hand-written restatement that the CUDA path is tested against) to the reference: tests/
 * test_glslref_*.py compare the two function by function and image by image. What the shaders do not contain --
 * ray/triangle intersection, instance transforms, texture filtering (Vulkan driver / hardware) -- enters through two
 * callbacks that the tests point at liboracle.so (orc_trace1, orc_kat_texture).
 */
#ifndef GLSLREF_H
#define GLSLREF_H
#include <stdint.h>
#include "lmb_types.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ref_scene ref_scene;
/* same signatures as orc_trace1 / orc_kat_texture (oracle/oracle.h) */
typedef int (*ref_trace1_fn)(const void* user, const float* ray8, int any_hit, void* hit16, uint32_t* mesh, uint32_t* local);
typedef void (*ref_texture_fn)(const void* user, uint32_t tex, const float* uv2, uint32_t n, float* out3);

/* Binds the scene arrays exactly as Path::init / Path::render bind them (Path.cpp:6-11, 49-57): SceneDesc addresses,
 * lights at binding 3, textures at binding 4. `sd` pointers must outlive the handle. `user` is passed to the callbacks. */
int ref_scene_create(const lmb_scene_desc* sd, const void* user, ref_trace1_fn trace1, ref_texture_fn texture, ref_scene** out);
void ref_scene_destroy(ref_scene* s);

/* One vkCmdTraceRaysKHR(W, H, 1) of path.rgen per frame in [first_frame, first_frame + n_frames) with pc.frame_num = frame
 * (Path.cpp:27-59). rgba = the RGBA32F storage image (read when first_frame > 0). rays[3] += closest (cull mask 0xFF),
 * shadow (terminate-on-first-hit), probe (closest, cull mask 0x1) traceRayEXT calls. */
int ref_render_path(ref_scene* s, const lmb_pc_path* pc, const lmb_scene_ubo* ubo, uint32_t first_frame, uint32_t n_frames, float* rgba,
					uint64_t* rays3, int n_threads);

/* One dispatch of bdpt.rgen (src/RayTracer/BDPT.cpp:55-95) for frame `frame`, seed (x, y, frame ^ pc->time, 0). Every invocation
 * runs on a private, zeroed colour storage (the GLSL's non-atomic cross-pixel `tmp_col.d[idx] += splat` makes the reference's own
 * result depend on scheduling): image_rgba[W*H*4] = what main() stores = the pixel's own (s, t >= 2) strategies + the splats it
 * sends to itself; splat_rgb[W*H*3] = the splats it sends to other pixels, summed per target pixel. rays3 as ref_render_path. */
int ref_render_bdpt_frame(ref_scene* s, const lmb_pc_bdpt* pc, const lmb_scene_ubo* ubo, uint32_t frame, float* image_rgba, float* splat_rgb,
						  uint64_t* rays3, int n_threads);

/* Function probes with the signatures of oracle.h's orc_kat_* (the scene handle provides the bindings the stage needs
 * to exist; BSDF / RNG / sky probes do not read it). */
void ref_kat_pcg4d(ref_scene* s, const uint32_t* in4, uint32_t n, uint32_t* out4);
void ref_kat_rand(ref_scene* s, const uint32_t* seed4, uint32_t n, uint32_t draws, float* out);
void ref_kat_offset_ray(ref_scene* s, const float* p3, const float* n3, uint32_t n, float* out3, float* out3_b);
void ref_kat_sample_bsdf(ref_scene* s, const lmb_material* mat, const float* n_s3, const float* wo3, const float* rands3, const uint8_t* side,
						 uint32_t n, float* out8);
void ref_kat_eval_bsdf(ref_scene* s, const lmb_material* mat, const float* n_s3, const float* wo3, const float* wi3, const uint8_t* side,
					   uint32_t n, float* out4);
void ref_kat_bsdf_pdf(ref_scene* s, const lmb_material* mat, const float* n_s3, const float* wo3, const float* wi3, const uint8_t* side,
					  uint32_t n, float* out);
void ref_kat_atmosphere(ref_scene* s, const float* origin3, const float* dir3, const float* light_dir3, const float* light_L3, uint32_t n,
						float* out3);
void ref_kat_sample_light(ref_scene* s, int32_t num_lights, const float* rands4, const float* p3, uint32_t n, float* out16);
void ref_kat_light_Le(ref_scene* s, int32_t num_lights, int32_t total_light, const float* rands6, uint32_t n, float* out16);
/* load_material (bsdf_commons.glsl:16-22): out = n Material records (104 B each) */
void ref_kat_load_material(ref_scene* s, const uint32_t* material_idx, const float* uv2, uint32_t n, lmb_material* out);

#ifdef __cplusplus
}
#endif
#endif
