/*
 * lumen_b200_testhooks.h -- known-answer probes into the DEVICE functions of liblumen_b200.so, one kernel launch per
 * call, host arrays in / out. They exist so that tests/ can compare each shader-level function of the CUDA path with the
 * CPU oracle bit for bit (the reference has no tests, the pyramid is authored here). Array shapes
 * match the orc_kat_* probes of oracle/oracle.h.
 */
#ifndef LUMEN_B200_TESTHOOKS_H
#define LUMEN_B200_TESTHOOKS_H
#include "lumen_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int lmb_kat_pcg4d(lmb_ctx* ctx, const uint32_t* in4, uint32_t n, uint32_t* out4);                 /* utils.glsl:121-134 */
int lmb_kat_rand(lmb_ctx* ctx, const uint32_t* seed4, uint32_t n, uint32_t draws, float* out);      /* utils.glsl:145-148 */
int lmb_kat_detmath(lmb_ctx* ctx, const float* x, const float* y, uint32_t n, float* out_sin, float* out_cos, float* out_exp, float* out_pow);
int lmb_kat_offset_ray(lmb_ctx* ctx, const float* p3, const float* n3, uint32_t n, float* out3, float* out3_b); /* utils.glsl:73-94 */
int lmb_kat_sample_bsdf(lmb_ctx* ctx, const lmb_material* mat, const float* n_s3, const float* wo3, const float* rands3, const uint8_t* side,
						uint32_t n, float* out8);                                                     /* bsdf_commons.glsl:68-121 */
int lmb_kat_eval_bsdf(lmb_ctx* ctx, const lmb_material* mat, const float* n_s3, const float* wo3, const float* wi3, const uint8_t* side,
					  uint32_t n, float* out4);                                                       /* bsdf_commons.glsl:123-183 */
int lmb_kat_atmosphere(lmb_ctx* ctx, const float* origin3, const float* dir3, const float* light_dir3, const float* light_L3, uint32_t n,
					   float* out3);                                                                  /* atmosphere.glsl:148-204 */
int lmb_kat_sample_light(lmb_ctx* ctx, int32_t num_lights, const float* rands4, const float* p3, uint32_t n, float* out16); /* commons.glsl:224-300 */
int lmb_kat_texture(lmb_ctx* ctx, uint32_t tex, const float* uv2, uint32_t n, float* out3);          /* bsdf_commons.glsl:19 */
/* Structural check of the 8-wide traversal BVH on the host: out8 = nodes allocated, nodes reachable from the root, depth,
 * errors (child box not containing its subtree, bad meta / imask), duplicate + missing triangles, internal children,
 * leaf children, leaf triangles. A valid tree has out8[0] == out8[1], out8[3] == out8[4] == 0, out8[7] == triangle count. */
int lmb_kat_wide_bvh_check(lmb_ctx* ctx, uint64_t* out8);
/* One BDPT frame (lmb_render_bdpt, film update included) that also hands back the two images the film is made of:
 * col_rgba[W*H*4] = radiance of each pixel's own (s, t >= 2) strategies, splat_rgb[W*H*3] = the light-tracer image. */
int lmb_kat_bdpt_frame_raw(lmb_ctx* ctx, const lmb_pc_bdpt* pc, const lmb_scene_ubo* ubo, uint32_t frame, float* col_rgba, float* splat_rgb);
#ifdef __cplusplus
}
#endif
#endif
