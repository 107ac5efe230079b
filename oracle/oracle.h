/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * C entry points of the CPU oracle (liboracle.so): a C++/glm restatement of Lumen's Path integrator
 * (src/shaders/integrators/path/path.rgen + includes) over a canonical CPU LBVH. Loaded through ctypes by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by lumen_b200/.
 *
 * Parity status: PINNED TO THE REFERENCE ITSELF RUN HERE. The reference ships no tests, golden images or known-answer
 * vectors for this path and its executable needs a Vulkan RT GPU, but its shader SOURCE compiles on the CPU
 * after a mechanical GLSL -> C++ pass (oracle/glslref/glsl2cpp.py over the unmodified files, built into
 * oracle/_ref/libglslref.so). tests/test_glslref_pins_oracle.py holds this oracle to that library bit for bit: RNG, ray
 * offsets, every BSDF's sample / eval / pdf, light sampling, the sky model, load_material, and whole path.rgen and
 * bdpt.rgen dispatches (films and ray counts) on the reference's scenes. Outside the shader source, hence DEFINED by this
 * build on both sides: ray/triangle intersection and instance transforms (Vulkan driver), bilinear texture filtering
 * (hardware), transcendental functions (include/lmb_detmath.h), no FMA contraction, uninitialised variables read as zero.
 * tests/golden/ additionally freezes vectors minted from this oracle so that it cannot drift unnoticed on a box without
 * the reference checkout.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#include "lmb_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene; /* scene view + CPU LBVH */

typedef struct orc_stats {
	uint64_t rays_closest; /* continuation (camera + bounce) rays, path.rgen:48 */
	uint64_t rays_shadow;  /* any-hit visibility rays, pt_commons.glsl:21 */
	uint64_t rays_probe;   /* MIS BSDF-probe closest-hit rays, pt_commons.glsl:32 */
	uint64_t nodes_visited;
	uint64_t tris_tested;
	uint64_t nan_pixels;   /* samples dropped by the NaN guard, path.rgen:102-104 */
	double seconds;        /* wall time of the render loop */
	int32_t threads;
} orc_stats;

typedef struct orc_hit {
	float t, b1, b2;
	uint32_t prim; /* global triangle id, 0xFFFFFFFF = miss */
} orc_hit;

/* Builds the canonical LBVH; `sd` pointers must stay valid for the lifetime of the handle. */
int orc_scene_create(const lmb_scene_desc* sd, orc_scene** out);
void orc_scene_destroy(orc_scene* s);

/* LBVH arrays for bit-exact comparison with the GPU build. Sizes: n = n_tris; left/right n-1; parent 2n-1;
 * aabb 6*(2n-1); morton n (by global id); keys n (sorted); leaf_prim n. */
uint32_t orc_lbvh_num_tris(const orc_scene* s);
const uint32_t* orc_lbvh_left(const orc_scene* s);
const uint32_t* orc_lbvh_right(const orc_scene* s);
const uint32_t* orc_lbvh_parent(const orc_scene* s);
const uint32_t* orc_lbvh_leaf_prim(const orc_scene* s);
const uint32_t* orc_lbvh_morton(const orc_scene* s);
const uint64_t* orc_lbvh_keys(const orc_scene* s);
const float* orc_lbvh_aabb(const orc_scene* s);

/* Renders frames [first_frame, first_frame + n_frames) into `rgba` (W*H*4 floats) with the reference's running-mean
 * film update (path.rgen:102-112). `rgba` is read when first_frame > 0. pc->frame_num is ignored (set per frame). */
int orc_render(const orc_scene* s, const lmb_pc_path* pc, const lmb_scene_ubo* ubo, uint32_t first_frame, uint32_t n_frames,
			   float* rgba, orc_stats* stats, int n_threads);
/* Restricts orc_render to the image rows row_first, row_first + row_stride, ... (the pixel shard of lmb_set_pixel_shard; other rows of
 * `rgba` are left untouched). Process-wide; (0, 1) = every row. */
void orc_set_row_shard(uint32_t row_first, uint32_t row_stride);
/* Per-sample radiance of one frame without the film update: out[W*H*3], NaN samples are kept as NaN. */
int orc_render_frame_raw(const orc_scene* s, const lmb_pc_path* pc, const lmb_scene_ubo* ubo, uint32_t frame, float* rgb,
						 orc_stats* stats, int n_threads);

/* BDPT (bdpt.rgen + bdpt_commons.glsl, quirks B1-B6 in oracle/bdpt.h). Film update as orc_render.
 * RNG seed of a sample is (x, y, frame ^ pc->time, 0); pc->frame_num is ignored. */
int orc_render_bdpt(const orc_scene* s, const lmb_pc_bdpt* pc, const lmb_scene_ubo* ubo, uint32_t first_frame, uint32_t n_frames,
					float* rgba, orc_stats* stats, int n_threads);
/* One frame without the film: col_rgb[W*H*3] = the pixel's own (s, t >= 2) strategies, splat_rgb[W*H*3] = light-tracer image. */
int orc_render_bdpt_frame_raw(const orc_scene* s, const lmb_pc_bdpt* pc, const lmb_scene_ubo* ubo, uint32_t frame, float* col_rgb,
							  float* splat_rgb, orc_stats* stats, int n_threads);

/* Diagnostic: keep only the BDPT strategies with `s` light vertices, weight 1 (-1 = all strategies, the reference's MIS weights). */
void orc_bdpt_set_only_s(int s);
/* Diagnostic: check after every calc_mis_weight that its in-place vertex patches were restored completely (the pair-parallel CUDA
 * kernels rely on it); orc_bdpt_restore_violations() = calls so far that left a vertex changed. */
void orc_bdpt_set_check_restore(int on);
long long orc_bdpt_restore_violations(void);

/* Ray queries: rays = n x 8 floats (ox, oy, oz, tmin, dx, dy, dz, tmax). */
int orc_trace_closest(const orc_scene* s, const float* rays, uint32_t n, orc_hit* hits, orc_stats* stats, int n_threads);
/* one ray without OpenMP (callback of oracle/glslref's traceRayEXT); returns 1 on a hit and fills mesh / local with
 * gl_InstanceCustomIndexEXT / gl_PrimitiveID */
int orc_trace1(const orc_scene* s, const float* ray8, int any_hit, orc_hit* hit, uint32_t* mesh, uint32_t* local);
/* the hit definition evaluated over ALL triangles, no tree: what orc_trace_closest must equal bit for bit */
int orc_trace_closest_brute(const orc_scene* s, const float* rays, uint32_t n, orc_hit* hits, int n_threads);
int orc_trace_any(const orc_scene* s, const float* rays, uint32_t n, uint8_t* occluded, orc_stats* stats, int n_threads);

/* Known-answer probes for individual shader functions (all arrays are n-element, tightly packed). */
void orc_kat_pcg4d(const uint32_t* in4, uint32_t n, uint32_t* out4);
void orc_kat_rand(const uint32_t* seed4, uint32_t n, uint32_t draws, float* out);
void orc_kat_detmath(const float* x, const float* y, uint32_t n, float* out_sin, float* out_cos, float* out_exp, float* out_pow);
void orc_kat_offset_ray(const float* p3, const float* n3, uint32_t n, float* out3, float* out3_b);
/* sample: out 8 floats per item = f.xyz, wi.xyz, pdf, cos_theta.  eval: out 4 floats = f.xyz, pdf. */
void orc_kat_sample_bsdf(const lmb_material* mat, const float* n_s3, const float* wo3, const float* rands3, const uint8_t* side,
						 uint32_t n, float* out8);
/* bsdf_pdf of bsdf_commons.glsl:26-66 (the stand-alone pdf functions BDPT uses): out[n] */
void orc_kat_bsdf_pdf(const lmb_material* mat, const float* n_s3, const float* wo3, const float* wi3, const uint8_t* side, uint32_t n, float* out);
void orc_kat_eval_bsdf(const lmb_material* mat, const float* n_s3, const float* wo3, const float* wi3, const uint8_t* side,
					   uint32_t n, float* out4);
/* sky: out 3 floats per item */
void orc_kat_atmosphere(const float* origin3, const float* dir3, const float* light_dir3, const float* light_L3, uint32_t n,
						float* out3);
/* light sampling at shading points p: out 16 floats = Le.xyz, wi.xyz, wi_len, pdf_w, pdf_a, cos_from_light,
 * light_idx, flags, triangle_idx, instance_idx, bary.xy */
void orc_kat_sample_light(const orc_scene* s, int32_t num_lights, const float* rands4, const float* p3, uint32_t n, float* out16);
/* light EMISSION sampling of the BDPT light walk (sample_light_Le, commons.glsl:335-406): rands6 = rands_pos.xyzw, rands_dir.xy;
 * out 16 floats = L.xyz, pos.xyz, wi.xyz, n.xyz, cos_from_light, pdf_pos_a (already / total_light), pdf_dir_w, flags */
void orc_kat_light_Le(const orc_scene* s, int32_t num_lights, int32_t total_light, const float* rands6, uint32_t n, float* out16);
/* texture fetch: out 3 floats */
void orc_kat_texture(const orc_scene* s, uint32_t tex, const float* uv2, uint32_t n, float* out3);

/* Image difference metrics. literal = the reference's rmse/ *.comp arithmetic including its quirks
 * (calc_rmse.comp:37 subgroupMin, output_rmse.comp:23 sqrt(S)/(3N)); true = sqrt(mean((a-b)^2)) over RGB. */
float orc_rmse_literal(const float* rgba_a, const float* rgba_b, uint32_t n_pixels);
double orc_rmse_true(const float* rgba_a, const float* rgba_b, uint32_t n_pixels);

/* float -> half exactly as the EXR writer does it (tinyexr float_to_half_full, libs/tinyexr.h:898-934). */
void orc_float_to_half(const float* in, uint32_t n, uint16_t* out);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
