// bdpt.cu -- Lumen's bidirectional path tracer on the Path integrator's scene, BVH and BSDF code.
//
// Replaces: BDPT::render (src/RayTracer/BDPT.cpp:55-95) = one vkCmdTraceRaysKHR of src/shaders/integrators/bdpt/bdpt.rgen:39-90
// with src/shaders/integrators/bdpt_commons.glsl:13-641 (random walks, calc_mis_weight, bdpt_connect_cam, bdpt_connect),
// sample_light_Le / light_pdf / light_pdf_a_to_w (src/shaders/commons.glsl:40-111, 335-406) and the stand-alone bsdf_pdf
// functions (bsdf_commons.glsl:26-66, diffuse.glsl:85-90, dielectric.glsl:191-240, conductor.glsl:74-90,
// principled.glsl:177-181, 226-251, 338-360).
//
// Frozen where the reference is undefined (the same definitions as the CPU oracle, oracle/bdpt.h B1-B6): the RNG seed is
// (x, y, frame ^ pc.time, 0) with `time` an input; vertex storage is zeroed before every frame exactly as BDPT.cpp:79-80 does;
// the light tracer's splats (t == 1) of frame f all land in frame f: they are added to a splat image with float atomics and
// k_bdpt_film, a later launch, adds that image to the pixel's own strategies before the film update (the reference reads and
// clears the splat buffer inside the same dispatch that is still writing it).
//
// Layout: the two sub-paths of a pixel live in HBM as a struct of arrays over pixels -- word w of vertex i of pixel p at
// verts[((i * 23 + w) * n_pix) + p] -- so every vertex field access of a warp is one coalesced 128-byte transaction;
// 2 x (max_depth + 1) x 92 B per pixel (2.7 GB at 1080p, depth 6).
//
// Three pipelines over the same device functions, bit-equal to each other (tests/test_gpu_bdpt.py) -- DESIGN.md section 8:
//   default      staged: the per-pixel GLSL is cut at every ray (k_bdpt_begin, k_bdpt_walk, k_bdpt_mid), rays go through the Path
//                integrator's persistent 8-wide walker as slots (NaN origin = no ray), connections run one thread per
//                ((s, t) pair, pixel) with register substitutes for calc_mis_weight's in-place patches (k_bdpt_pair, k_bdpt_gather)
//   LMB_BDPT=pixel   staged walks, connections looped per pixel with the GLSL's patch-and-restore (k_bdpt_connect)
//   LMB_BDPT=mega    one kernel per frame, rays walked in the thread over the binary LBVH (k_bdpt): the first correct state of
//                    this row, kept as the plainest statement of the algorithm (3.3-6 x slower: profiles/r01j_summary.md)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "context.h"
#include "scene_device.cuh"
#include "trace.cuh"

namespace lmb {

namespace {

// ------------------------------------------------------------------------------------------------ bsdf_pdf
// A/B variant (off): -DLMB_BDPT_V3_BYVAL=1 hands the direction vectors of the noinline pdf functions over by value, so that the
// callers need not park them in local memory to take their address.
#ifdef LMB_BDPT_V3_BYVAL
typedef V3 V3arg;
#else
typedef const V3& V3arg;
#endif
// diffuse.glsl:85-90
LMB_D float lambertian_diffuse_pdf(const V3& wo, const V3& wi) {
	if (gmin(wi.z, wo.z) <= 0.0f) return 0.0f;
	return wi.z * LMB_INV_PI;
}
// dielectric.glsl:191-240
LMB_DN float dielectric_pdf(const lmb_material& mat, V3arg wo, V3arg wi, bool forward_facing) {
	const float roughness = mat.thin == 1 ? modify_thin_roughness(mat.ior, mat.roughness) : mat.roughness;
	const float alpha = roughness * roughness;
	if (alpha == 0 || mat.ior == 1) return 0.0f;
	const bool has_reflection = has_prop(mat.bsdf_props, LMB_FLAG_REFLECTION);
	const bool has_transmission = has_prop(mat.bsdf_props, LMB_FLAG_TRANSMISSION);
	if (!has_reflection && !has_transmission) return 0.0f;
	const bool is_reflection = wi.z * wo.z > 0;
	float eta = 1.0f;
	if (!is_reflection) eta = forward_facing ? mat.ior : 1.0f / mat.ior;
	V3 h = normalize(wo + wi * eta);
	h *= gsign(h.z);
	if (wi.z == 0 || wo.z == 0 || dot(h, h) == 0) return 0.0f;
	if (dot(wi, h) * wi.z < 0 || dot(wo, h) * wo.z < 0) return 0.0f;
	const float F = fresnel_dielectric(dot(wo, h), mat.ior, forward_facing);
	const float pr = has_reflection ? F : 0.0f;
	const float pt = has_transmission ? (1.0f - F) : 0.0f;
	float D;
	float pdf_w = vndf_pdf_iso(alpha, wo, h, D);
	if (is_reflection) {
		const float jacobian = 1.0f / (4.0f * fabsf(dot(wo, h)));
		const float prob_reflection = pr / (pr + pt);
		pdf_w = pdf_w * jacobian * prob_reflection;
	} else {
		float jacobian_denom = dot(wi, h) + dot(wo, h) / eta;
		jacobian_denom = jacobian_denom * jacobian_denom;
		const float jacobian = fabsf(dot(wi, h)) / jacobian_denom;
		const float prob_refraction = pt / (pr + pt);
		pdf_w = pdf_w * jacobian * prob_refraction;
	}
	return pdf_w;
}
// conductor.glsl:74-90
LMB_DN float conductor_pdf(const lmb_material& mat, V3arg wo, V3arg wi) {
	const float alpha = mat.roughness * mat.roughness;
	if (effectively_delta(alpha)) return 0.0f;
	if (wo.z * wi.z < 0) return 0.0f;
	if (wo.z == 0 || wi.z == 0) return 0.0f;
	V3 h = normalize(wo + wi);
	h *= gsign(h.z);
	float D;
	return vndf_pdf_iso(alpha, wo, h, D) / (4.0f * dot(wo, h));
}
// principled.glsl:177-181
LMB_D float clearcoat_pdf(const lmb_material& mat, const V3& wo, const V3& wi) {
	const V3 h = normalize(wo + wi);
	const float D = d_ggx_iso(mix(0.1f, 0.001f, mat.clearcoat_gloss), h.z);
	return D / (4.0f * dot(wo, h));
}
// principled.glsl:226-251
LMB_DN float principled_brdf_pdf(const lmb_material& mat, V3arg wo, V3arg wi) {
	const V2 alpha = calc_anisotropy(mat.roughness, mat.anisotropy);
	if (effectively_delta(alpha)) return 0.0f;
	if (wo.z * wi.z < 0) return 0.0f;
	if (wo.z == 0 || wi.z == 0) return 0.0f;
	V3 h = normalize(wo + wi);
	h *= gsign(h.z);
	float D;
	return vndf_pdf_aniso(alpha, wo, h, D) / (4.0f * dot(wo, h));
}
// principled.glsl:338-360
LMB_DN float principled_pdf(const lmb_material& mat, V3arg wo, V3arg wi, bool forward_facing) {
	float pdf = 0.0f;
	const float F = fresnel_dielectric(wo.z, mat.ior, forward_facing);
	const LobeProbs p = sampling_probs(mat, F, forward_facing);
	if (p.spec > 0) pdf += p.spec * principled_brdf_pdf(mat, wo, wi);
	const bool upper = gmin(wi.z, wo.z) > 0;
	if (upper) {
		if (p.diff > 0) pdf += p.diff * lambertian_diffuse_pdf(wo, wi);
		if (p.clearcoat > 0) pdf += p.clearcoat * clearcoat_pdf(mat, wo, wi);
	}
	if (p.spec_trans > 0) pdf += p.spec_trans * dielectric_pdf(mat, wo, wi, forward_facing);
	return pdf;
}
// bsdf_commons.glsl:26-66
LMB_DN float bsdf_pdf(const lmb_material& mat, V3arg n_s, V3arg wo_world, V3arg wi_world, bool forward_facing) {
	V3 T, B;
	branchless_onb(n_s, T, B);
	const V3 wo = to_local(wo_world, T, B, n_s);
	const V3 wi = to_local(wi_world, T, B, n_s);
	switch (mat.bsdf_type) {
		case LMB_BSDF_DIFFUSE:
			return lambertian_diffuse_pdf(wo, wi);
		case LMB_BSDF_DIELECTRIC:
			return dielectric_pdf(mat, wo, wi, forward_facing);
		case LMB_BSDF_CONDUCTOR:
			return conductor_pdf(mat, wo, wi);
		case LMB_BSDF_PRINCIPLED:
			return principled_pdf(mat, wo, wi, forward_facing);
		default:
			break;
	}
	return 0.0f;
}

// ------------------------------------------------------------------------------------------------ lights
LMB_D bool is_light_finite(uint32_t f) { return ((f >> 4) & 1u) != 0; }  // commons.glsl:42
LMB_D bool is_light_delta(uint32_t f) { return ((f >> 5) & 1u) != 0; }   // commons.glsl:44
LMB_D float uniform_cone_pdf(float cos_max) { return 1.0f / (LMB_TWO_PI * (1 - cos_max)); }  // commons.glsl:40
LMB_D bool same_hemisphere(const V3& wi, const V3& wo, const V3& n) { return dot(wi, n) > 0 && dot(wo, n) > 0; }
LMB_D bool is_zero(const V3& f) { return f.x == 0 && f.y == 0 && f.z == 0; }

// commons.glsl:81-95
LMB_D float light_pdf(uint32_t light_flags, const V3& n_s, const V3& wi) {
	const float cos_width = lmb_cosf(30 * LMB_PI / 180);
	switch (light_flags & 0x7u) {
		case LMB_LIGHT_AREA:
			return gmax(dot(n_s, wi) / LMB_PI, 0.0f);
		case LMB_LIGHT_SPOT:
			return uniform_cone_pdf(cos_width);
		default:
			break;
	}
	return 0.0f;
}
// commons.glsl:64-79
LMB_D float light_pdf_a_to_w(uint32_t light_flags, float pdf_a, float wi_len_sqr, float cos_from_light) {
	switch (light_flags & 0x7u) {
		case LMB_LIGHT_AREA:
			return pdf_a * wi_len_sqr / cos_from_light;
		case LMB_LIGHT_SPOT:
			return wi_len_sqr / cos_from_light;
		case LMB_LIGHT_DIRECTIONAL:
			return 1.0f;
		default:
			break;
	}
	return 0.0f;
}
// utils.glsl:157-173
LMB_D V4 to_local_quat(const V3& v) {
	if (v.z < -0.99999f) return v4(1, 0, 0, 0);
	const V4 q = v4(v.y, -v.x, 0.0f, 1.0f + v.z);
	const float inv = 1.0f / sqrtf((q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w));
	return v4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
}
LMB_D V3 rot_quat(const V4& q, const V3& v) {
	const V3 q_axis = v3(q.x, q.y, q.z);
	return (2.0f * dot(q_axis, v)) * q_axis + (q.w * q.w - dot(q_axis, q_axis)) * v + (2.0f * q.w) * cross(q_axis, v);
}
// utils.glsl:175-182
LMB_D void make_coord_system(const V3& v1, V3& v2o, V3& v3o) {
	if (fabsf(v1.x) > fabsf(v1.y))
		v2o = normalize(v3(-v1.z, 0, v1.x));
	else
		v2o = normalize(v3(0, v1.z, -v1.y));
	v3o = cross(v1, v2o);
}
// commons.glsl:217-222
LMB_D V3 uniform_sample_cone(const V2& uv, float cos_max) {
	const float cos_theta = (1.0f - uv.x) + uv.x * cos_max;
	const float sin_theta = sqrtf(1 - cos_theta * cos_theta);
	const float phi = uv.y * LMB_TWO_PI;
	return v3(lmb_cosf(phi) * sin_theta, lmb_sinf(phi) * sin_theta, cos_theta);
}
// sampling_commons.glsl:147-161 (concentric-disk mode), explicit normal
LMB_D V3 sample_hemisphere_n(const V2& xi, const V3& n) {
	V3 T, B;
	branchless_onb(n, T, B);
	const V2 d = concentric_sample_disk(xi);
	const float z = sqrtf(gmax(0.f, 1.f - dot(d, d)));
	return to_world(v3(d.x, d.y, z), T, B, n);
}

struct TriangleRecord {
	V3 pos, n_s;
	float triangle_pdf;
};
// commons.glsl:112-149 (Q7: w = 1 on the edge vectors and on the normal); same expressions as the AREA branch of sample_light_Li
LMB_DN TriangleRecord sample_triangle(const DeviceScene& sc, const lmb_light& light, float r_tri, float r_u, float r_v, V2& uv, uint32_t& material_idx) {
	TriangleRecord r;
	const uint32_t pm = light.prim_mesh_idx;
	const lmb_prim_mesh_info& pinfo = sc.prim_infos[pm];
	material_idx = pinfo.material_index;
	const uint32_t triangle_idx = (uint32_t)(r_tri * (float)light.num_triangles);
	const M4 wm = load_m4(light.world_matrix);
	const M4 inv_tr = transpose(load_m4(sc.inv_world_matrices + 16 * pm));
	const uint32_t index_offset = pinfo.index_offset + 3 * triangle_idx;
	const uint32_t vo = pinfo.vertex_offset;
	const lmb_vertex a0 = sc.vertices[sc.indices[index_offset + 0] + vo];
	const lmb_vertex a1 = sc.vertices[sc.indices[index_offset + 1] + vo];
	const lmb_vertex a2 = sc.vertices[sc.indices[index_offset + 2] + vo];
	const V3 q0 = vtx_pos(a0), q1 = vtx_pos(a1), q2 = vtx_pos(a2);
	const V3 n0 = vtx_nrm(a0), n1 = vtx_nrm(a1), n2 = vtx_nrm(a2);
	const float sq = sqrtf(r_u);
	uv = v2(1 - sq, r_v * sq);
	const V3 bary = v3(1.0f - uv.x - uv.y, uv.x, uv.y);
	const V4 etmp0 = mul(wm, v4(q1 - q0, 1.0f));
	const V4 etmp1 = mul(wm, v4(q2 - q0, 1.0f));
	const V3 pos = q0 * bary.x + q1 * bary.y + q2 * bary.z;
	const V3 nrm = normalize(n0 * bary.x + n1 * bary.y + n2 * bary.z);
	const V4 world_pos = mul(wm, v4(pos, 1.0f));
	r.n_s = normalize(xyz(mul(inv_tr, v4(nrm, 1.0f))));
	r.triangle_pdf = 2.0f / length(cross(xyz(etmp0), xyz(etmp1)));
	r.pos = xyz(world_pos);
	return r;
}

// `out vec3 n, out vec3 pos` of sample_light_Li (commons.glsl:224-300), which scene_device.cuh's version (shared with the Path
// kernels) does not return: recomputed from the same inputs with the same expressions.
LMB_DN void light_sample_n_pos(const DeviceScene& sc, const V4& rands, const V3& p, int num_lights, const LightSample& ls, V3& n, V3& pos) {
	n = v3(0.0f), pos = v3(0.0f);
	const uint32_t light_idx = (uint32_t)(rands.x * (float)num_lights);
	const lmb_light& light = sc.lights[light_idx];
	switch (light.light_flags & 0x7u) {
		case LMB_LIGHT_AREA: {
			V2 uv;
			uint32_t material_idx;
			const TriangleRecord rec = sample_triangle(sc, light, rands.y, rands.z, rands.w, uv, material_idx);
			n = rec.n_s;
			pos = rec.pos;
		} break;
		case LMB_LIGHT_SPOT:
			n = -ls.wi;
			pos = v3(light.pos);
			break;
		case LMB_LIGHT_DIRECTIONAL: {
			const V3 dir = normalize(v3(light.pos) - v3(light.to));
			n = -ls.wi;
			pos = p + dir * (2 * light.world_radius);
		} break;
		default:
			break;
	}
}

struct LightEmission {
	V3 L, pos, wi, n;
	float cos_from_light, pdf_pos_a, pdf_dir_w;
	uint32_t flags;
};
// commons.glsl:335-406
LMB_DN LightEmission sample_light_Le(const DeviceScene& sc, const V4& rands_pos, const V2& rands_dir, int num_lights, int total_light) {
	LightEmission o;
	o.L = v3(0.0f), o.pos = v3(0.0f), o.wi = v3(0.0f), o.n = v3(0.0f);
	o.cos_from_light = 0, o.pdf_pos_a = 0, o.pdf_dir_w = 0;
	const uint32_t light_idx = (uint32_t)(rands_pos.x * (float)num_lights);
	const lmb_light& light = sc.lights[light_idx];
	o.flags = light.light_flags;
	switch (light.light_flags & 0x7u) {
		case LMB_LIGHT_AREA: {
			V2 bary;
			uint32_t material_idx;
			const TriangleRecord rec = sample_triangle(sc, light, rands_pos.y, rands_pos.z, rands_pos.w, bary, material_idx);
			const lmb_material light_mat = load_material(sc, material_idx, bary);
			o.pos = rec.pos;
			o.wi = sample_hemisphere_n(rands_dir, rec.n_s);
			o.L = v3(light_mat.emissive_factor);
			o.cos_from_light = gmax(dot(rec.n_s, o.wi), 0.0f);
			o.pdf_pos_a = rec.triangle_pdf;
			o.pdf_dir_w = dot(o.wi, rec.n_s) / LMB_PI;
			o.n = rec.n_s;
		} break;
		case LMB_LIGHT_SPOT: {
			const float cos_width = lmb_cosf(30 * LMB_PI / 180);
			const float cos_faloff = lmb_cosf(25 * LMB_PI / 180);
			const V3 light_dir = normalize(v3(light.to) - v3(light.pos));
			const V4 lq = to_local_quat(light_dir);
			o.wi = rot_quat(v4(-lq.x, -lq.y, -lq.z, lq.w), uniform_sample_cone(rands_dir, cos_width));
			o.pos = v3(light.pos);
			o.cos_from_light = dot(o.wi, light_dir);
			float faloff;
			if (o.cos_from_light < cos_width) {
				faloff = 0;
			} else if (o.cos_from_light >= cos_faloff) {
				faloff = 1;
			} else {
				const float d = (o.cos_from_light - cos_width) / (cos_faloff - cos_width);
				faloff = (d * d) * (d * d);
			}
			o.L = v3(light.L) * faloff;
			o.pdf_pos_a = 1.0f;
			o.pdf_dir_w = uniform_cone_pdf(cos_width);
			o.n = o.wi;
		} break;
		case LMB_LIGHT_DIRECTIONAL: {
			const V3 dir = -normalize(v3(light.to) - v3(light.pos));
			V3 v1, v2_;
			make_coord_system(dir, v1, v2_);
			const V2 uv = concentric_sample_disk(rands_dir);
			const V3 l_pos = v3(light.world_center) + light.world_radius * (uv.x * v1 + uv.y * v2_);
			o.pos = l_pos + dir * light.world_radius;
			o.wi = -dir;
			o.L = v3(light.L);
			o.pdf_pos_a = 1.0f / (LMB_PI * light.world_radius * light.world_radius);
			o.pdf_dir_w = 1;
			o.cos_from_light = 1;
			o.n = o.wi;
		} break;
		default:
			break;
	}
	o.pdf_pos_a /= (float)total_light;
	return o;
}

// payload.area of ray.rchit:76-83 (the Path kernels' build_hit leaves it out: path.rgen never reads it)
LMB_DN float hit_area(const DeviceScene& sc, uint32_t prim_global) {
	const uint4 rec = __ldg(&sc.tri_rec[prim_global]);
	const V3 q0 = vtx_pos(sc.vertices[rec.x]), q1 = vtx_pos(sc.vertices[rec.y]), q2 = vtx_pos(sc.vertices[rec.z]);
	const M4 o2w = load_m4(sc.world_matrices + 16 * rec.w);
	const V3 e0t = xyz(mul(o2w, v4(q2 - q0, 0.0f)));
	const V3 e1t = xyz(mul(o2w, v4(q1 - q0, 0.0f)));
	return 0.5f * length(cross(e0t, e1t));
}

// load_material (bsdf_commons.glsl:16-22) without the 104-byte copy when there is no texture to fold into the albedo: the callee reads
// the scene's own record (the BSDF functions take the material by reference). Same values either way; the copies of the connection
// kernels were most of their local-memory traffic (profiles/r01j_summary.md).
LMB_D const lmb_material& material_at(const DeviceScene& sc, uint32_t material_idx, const V2& uv, lmb_material& scratch) {
	const lmb_material& m = sc.materials[material_idx];
	if (m.texture_id > -1) {
		scratch = load_material(sc, material_idx, uv);
		return scratch;
	}
	return m;
}

// ------------------------------------------------------------------------------------------------ vertex storage
// PathVertex (bdpt_commons.h:18-33), 23 words, struct of arrays over pixels.
enum VertexWord { W_DIR = 0, W_NS = 3, W_POS = 6, W_UV = 9, W_THR = 11, W_LFLAGS = 14, W_LIDX = 15, W_MAT = 16, W_DELTA = 17, W_SIDE = 18, W_MODE = 19,
				  W_AREA = 20, W_PFWD = 21, W_PREV = 22, W_COUNT = 23 };

struct Verts {
	float* base;    // word 0 of vertex 0 of this pixel
	size_t stride;  // pixels
	LMB_D float& f(int i, int w) const { return base[((size_t)i * W_COUNT + w) * stride]; }
	LMB_D uint32_t u(int i, int w) const { return __float_as_uint(f(i, w)); }
	LMB_D void su(int i, int w, uint32_t v) const { f(i, w) = __uint_as_float(v); }
	LMB_D V3 v(int i, int w) const { return v3(f(i, w), f(i, w + 1), f(i, w + 2)); }
	LMB_D void sv(int i, int w, const V3& a) const { f(i, w) = a.x, f(i, w + 1) = a.y, f(i, w + 2) = a.z; }
	LMB_D V2 uv(int i) const { return v2(f(i, W_UV), f(i, W_UV + 1)); }
};

struct Sampled {  // the fields of `PathVertex sampled` that calc_mis_weight reads (oracle/bdpt.h B4: zero unless a strategy writes them)
	V3 pos, n_s;
	float pdf_fwd;
};

struct BdptParams {
	M4 inv_view, inv_proj, view, neg_proj;
	// n_pix = frames in flight x n_real: a "pixel" of the kernels below is a (frame of the batch, image pixel) pair, frame-major; frame fb of
	// the batch is frame_first + fb * frame_stride and seeds with that number ^ time (bdpt.rgen:36-37)
	uint32_t width, height, n_pix, n_real, frame_first, frame_stride, time;
	int32_t num_lights, max_depth, light_triangle_count;
	float* light_verts;
	float* camera_verts;
	float4* col;    // per pixel: radiance of the pixel's own strategies (t >= 2)
	float* splat;   // 3 floats per pixel: light-tracer image of this frame
	unsigned long long* stats;
	// staged pipeline only
	float* walk;      // WalkWord struct of arrays: the random walk in flight
	uint32_t* misc;   // MiscWord struct of arrays
	float4* rays;     // ray slots, 2 float4 each: n_pix for a walk step, n_conn_slots * n_pix for the connections
	float4* hits;     // closest hit of each pixel's walk ray
	uint8_t* occ;     // any-hit result of each connection slot
	float4* contrib;  // pair-parallel resolve: weighted radiance of each (slot, pixel), .w = splat target pixel (bits) or ~0
	uint32_t n_conn_slots;
	// walkers in flight as dense pixel lists (set per launch): k_bdpt_walk reads alive_in[0 .. *n_alive_in) and appends the pixels whose
	// walk goes on to alive_out; k_bdpt_begin / k_bdpt_mid start the lists
	const uint32_t* alive_in;
	const uint32_t* n_alive_in;
	uint32_t* alive_out;
	uint32_t* n_alive_out;
};

enum WalkWord { WW_POS = 0, WW_WI = 3, WW_THR = 6, WW_PDF = 9, WW_B = 10, WW_ALIVE = 11, WW_COUNT = 12 };
enum MiscWord { MW_RNG = 0, MW_LPDFPOS = 1, MW_NLIGHT = 2, MW_COUNT = 3 };

struct WalkSt {  // the loop-carried variables of bdpt_random_walk_*
	V3 ray_pos, wi, thr;
	float pdf_fwd;
	int b;
	bool alive;
};

struct Kctx {  // per-thread state of one pixel
	const BdptParams& P;
	const DeviceScene& sc;
	const BvhView& bvh;
	Rng seed;
	Verts lig, cam;
	float light_pdf_pos;
	float screen_size;
	uint32_t pix, slot;  // staged: this pixel and the connection slot of the (t, s) pair being evaluated
	uint32_t n_closest, n_shadow, n_nodes, n_tris;
};

constexpr float BDPT_T_MIN = 0.001f;  // bdpt_commons.glsl:24-25
constexpr float BDPT_T_MAX = 1e6f;

LMB_DN Hit trace_closest(Kctx& k, const V3& o, const V3& d, float tmin, float tmax) {
	k.n_closest++;
	return trace_ray<false>(k.bvh, o, d, tmin, tmax, k.n_nodes, k.n_tris);
}
LMB_DN bool occluded(Kctx& k, const V3& o, const V3& d, float tmax) {
	k.n_shadow++;
	return trace_ray<true>(k.bvh, o, d, 0.0f, tmax, k.n_nodes, k.n_tris).prim != 0xFFFFFFFFu;
}

// Visibility test of a connection. MODE 0 (megakernel): trace here. MODE 1 (staged, emit pass): write the ray into this pair's slot
// and report "not visible", so nothing downstream is evaluated yet. MODE 2 (staged, resolve pass): read the traced result.
template <int MODE>
LMB_D bool shadow_visible(Kctx& k, const V3& o, const V3& d, float tmax) {
	if (MODE == 0) return !occluded(k, o, d, tmax);
	const size_t i = (size_t)k.slot * k.P.n_pix + k.pix;
	if (MODE == 1) {
		k.P.rays[2 * i] = make_float4(o.x, o.y, o.z, 0.0f);
		k.P.rays[2 * i + 1] = make_float4(d.x, d.y, d.z, tmax);
		k.n_shadow++;
		return false;
	}
	return k.P.occ[i] == 0;
}

// One trip of the `while (true)` of bdpt_random_walk_light (bdpt_commons.glsl:13-111, EYE = false) / _eye (:113-216, EYE = true),
// given the closest hit `h` of the ray in `st`. st.alive = false where the GLSL breaks. V.x(i + 1, ..) is the GLSL's vtx(i, ..).
template <bool EYE>
LMB_DN void walk_step(Kctx& k, const Verts& V, int max_depth, WalkSt& st, const Hit& h) {
	int b = st.b;
	const int prev = b - 1;
	st.alive = false;
	if (h.prim == 0xFFFFFFFFu) {
		if (EYE) {
			V.sv(b + 1, W_THR, st.thr);
			V.f(b + 1, W_PFWD) = st.pdf_fwd;
			st.b = b + 1;
		}
		return;
	}
	const bool finite_light = is_light_finite(V.u(0, W_LFLAGS));
	const HitPayload payload = build_hit(k.sc, h.prim, h.b1, h.b2);
	V3 wo = V.v(prev + 1, W_POS) - payload.pos;
	const float wo_len = length(wo);
	wo /= wo_len;
	V3 n_s = payload.n_s;
	bool side = true;
	V3 n_g = payload.n_g;
	if (dot(payload.n_g, wo) < 0.0f) n_g = -n_g;
	if (dot(n_g, n_s) < 0) {
		n_s *= -1.0f;
		side = false;
	}
	V.f(b + 1, W_PFWD) = st.pdf_fwd * fabsf(dot(wo, n_s)) / (wo_len * wo_len);
	V.sv(b + 1, W_NS, n_s);
	V.f(b + 1, W_AREA) = hit_area(k.sc, h.prim);
	V.sv(b + 1, W_POS, payload.pos);
	V.f(b + 1, W_UV) = payload.uv.x, V.f(b + 1, W_UV + 1) = payload.uv.y;
	V.su(b + 1, W_MAT, payload.material_idx);
	V.sv(b + 1, W_THR, st.thr);
	V.su(b + 1, W_SIDE, side ? 1u : 0u);
	V.su(b + 1, W_MODE, EYE ? 1u : 0u);
	lmb_material mat_tex;
	const lmb_material& mat = material_at(k.sc, payload.material_idx, payload.uv, mat_tex);
	const bool mat_specular = (mat.bsdf_props & LMB_FLAG_SPECULAR) == LMB_FLAG_SPECULAR;
	const bool mat_transmissive = (mat.bsdf_props & LMB_FLAG_TRANSMISSION) == LMB_FLAG_TRANSMISSION;
	V.su(b + 1, W_DELTA, mat_specular ? 1u : 0u);
	st.b = ++b;
	if (b >= max_depth) return;
	const V3 r3 = rand3(k.seed);
	const BsdfSample bs = sample_bsdf(n_s, wo, mat, EYE ? 1u : 0u, side, r3);
	st.wi = bs.wi;
	st.pdf_fwd = bs.pdf;
	const bool same_hem = same_hemisphere(bs.wi, wo, n_s);
	if (is_zero(bs.f) || bs.pdf == 0 || (!same_hem && !mat_transmissive)) return;
	st.thr *= bs.f * fabsf(bs.cos_theta) / bs.pdf;
	float pdf_rev = bs.pdf;
	if (!mat_specular) pdf_rev = bsdf_pdf(mat, n_s, bs.wi, wo, side);
	const bool g_term = EYE ? true : (prev > -1 || finite_light);
	if (g_term) pdf_rev *= fabsf(dot(V.v(prev + 1, W_NS), wo)) / (wo_len * wo_len);
	V.f(prev + 1, W_PREV) = pdf_rev;
	st.ray_pos = offset_ray(payload.pos, n_g);
	st.alive = true;
}

// megakernel: the whole walk in this thread
template <bool EYE>
LMB_DN int random_walk(Kctx& k, const Verts& V, int max_depth, const V3& throughput, float pdf) {
	if (max_depth == 0) return 0;
	WalkSt st{V.v(0, W_POS), V.v(0, W_DIR), throughput, pdf, 0, true};
	while (st.alive) {
		const Hit h = trace_closest(k, st.ray_pos, st.wi, BDPT_T_MIN, BDPT_T_MAX);
		walk_step<EYE>(k, V, max_depth, st, h);
	}
	return st.b;
}

// bdpt_generate_light_subpath, bdpt_commons.glsl:218-249: emission sample and vertex 0. false = no light sub-path (pdf_dir <= 0)
LMB_DN bool light_begin(Kctx& k, V3& throughput, float& pdf_dir) {
	const V4 rands_pos = rand4(k.seed);
	const float d0 = rand1(k.seed);
	const float d1 = rand1(k.seed);
	const LightEmission le = sample_light_Le(k.sc, rands_pos, v2(d0, d1), k.P.num_lights, k.P.light_triangle_count);
	if (le.pdf_dir_w <= 0) return false;
	k.light_pdf_pos = le.pdf_pos_a;
	const Verts& L = k.lig;
	L.sv(0, W_POS, le.pos);
	L.su(0, W_LFLAGS, le.flags);
	L.su(0, W_DELTA, 0);
	L.sv(0, W_DIR, le.wi);
	L.f(0, W_PFWD) = le.pdf_pos_a;
	L.sv(0, W_NS, le.n);
	L.su(0, W_SIDE, 1);
	L.su(0, W_MODE, 0);
	throughput = le.L * le.cos_from_light / (le.pdf_dir_w * le.pdf_pos_a);
	L.sv(0, W_THR, le.L);
	pdf_dir = le.pdf_dir_w;
	return true;
}
// ... :252-259, after the walk. pdf_pos and wi are read back from where light_begin stored them.
LMB_D void light_end(Kctx& k) {
	const Verts& L = k.lig;
	const uint32_t flags = L.u(0, W_LFLAGS);
	if (!is_light_finite(flags)) L.f(1, W_PFWD) = k.light_pdf_pos * fabsf(dot(L.v(0, W_DIR), L.v(1, W_NS)));
	if (is_light_delta(flags)) L.f(0, W_PFWD) = 0.0f;
}

LMB_D M4 neg_m4(const M4& m) {
	M4 r;
#pragma unroll
	for (int c = 0; c < 4; c++) r.c[c] = v4(-m.c[c].x, -m.c[c].y, -m.c[c].z, -m.c[c].w);
	return r;
}

// bdpt.rgen:40-53 + bdpt_generate_camera_subpath, bdpt_commons.glsl:262-284: camera vertex 0; returns the pdf the eye walk starts with
LMB_DN float camera_begin(Kctx& k, uint32_t px, uint32_t py) {
	const BdptParams& P = k.P;
	const V2 size = v2((float)P.width, (float)P.height);
	const V2 pixel = v2((float)px, (float)py) + 0.5f;
	const V2 in_uv = pixel / size;
	const V2 d = in_uv * 2.0f - 1.0f;
	const V4 origin = mul(P.inv_view, v4(0, 0, 0, 1));
	V4 area_int = mul(P.inv_proj, v4(2.0f / (float)P.width, 2.0f / (float)P.height, 0, 1));
	area_int = v4(area_int.x / area_int.w, area_int.y / area_int.w, area_int.z / area_int.w, area_int.w / area_int.w);
	const float cam_area = fabsf(area_int.x * area_int.y);
	const Verts& C = k.cam;
	C.sv(0, W_POS, xyz(origin));
	const V4 target = mul(P.inv_proj, v4(d.x, d.y, 1, 1));
	const V3 dir = xyz(mul(P.inv_view, v4(normalize(xyz(target)), 0)));  // sample_camera, commons.glsl:30-33
	C.sv(0, W_DIR, dir);
	C.f(0, W_AREA) = cam_area;
	C.sv(0, W_THR, v3(1.0f));
	C.su(0, W_DELTA, 0);
	const V3 n_s = xyz(mul(neg_m4(P.inv_view), v4(0, 0, 1, 0)));
	C.sv(0, W_NS, n_s);
	C.su(0, W_SIDE, 1);
	k.lig.su(0, W_MODE, 1);  // sic, :272
	const float cos_theta = dot(dir, n_s);
	return 1 / (cam_area * k.screen_size * cos_theta * cos_theta * cos_theta);
}

LMB_D float remap0(float v) { return v != 0.0f ? v : 1.0f; }

// bdpt_commons.glsl:288-470. The GLSL patches vertices in place and restores them afterwards; so does this.
LMB_DN float calc_mis_weight(Kctx& k, int s, int t, const Sampled& sampled) {
	const Verts& cam = k.cam;
	const Verts& lig = k.lig;
	bool s_0_changed = false, t_0_changed = false;
	float s_0_pdf = 0;
	V3 s_0_pdf_pos = v3(0.0f), s_0_pdf_nrm = v3(0.0f);
	uint32_t idx_1 = 0xFFFFFFFFu, idx_2 = 0xFFFFFFFFu, idx_3 = 0xFFFFFFFFu, idx_4 = 0xFFFFFFFFu;
	float idx_1_val = 0, idx_2_val = 0, idx_3_val = 0, idx_4_val = 0;
	uint32_t delta_t_old = 0, delta_s_old = 0;
	if (s + t == 2) return 1.0f;
	if (s == 1) {
		s_0_pdf = lig.f(0, W_PFWD), s_0_pdf_pos = lig.v(0, W_POS), s_0_pdf_nrm = lig.v(0, W_NS);
		lig.f(0, W_PFWD) = sampled.pdf_fwd, lig.sv(0, W_POS, sampled.pos), lig.sv(0, W_NS, sampled.n_s);
		s_0_changed = true;
	}
	if (t == 1) {
		s_0_pdf = cam.f(0, W_PFWD), s_0_pdf_pos = cam.v(0, W_POS), s_0_pdf_nrm = cam.v(0, W_NS);
		cam.f(0, W_PFWD) = sampled.pdf_fwd, cam.sv(0, W_POS, sampled.pos), cam.sv(0, W_NS, sampled.n_s);
		t_0_changed = true;
	}
	if (t > 0) {
		delta_t_old = cam.u(t - 1, W_DELTA);
		cam.su(t - 1, W_DELTA, 0);
	}
	if (s > 0) {
		delta_s_old = lig.u(s - 1, W_DELTA);
		lig.su(s - 1, W_DELTA, 0);
	}
	if (t > 0) {
		idx_1_val = cam.f(t - 1, W_PREV);
		idx_1 = (uint32_t)t;
		if (s > 0) {
			V3 dir = cam.v(t - 1, W_POS) - lig.v(s - 1, W_POS);
			const float dir_len = length(dir);
			dir /= dir_len;
			float pdf_rev = 0;
			if (s >= 2) {
				lmb_material mat_tex;
				const lmb_material& mat = material_at(k.sc, lig.u(s - 1, W_MAT), lig.uv(s - 1), mat_tex);
				const V3 wo = normalize(lig.v(s - 2, W_POS) - lig.v(s - 1, W_POS));
				pdf_rev = bsdf_pdf(mat, lig.v(s - 1, W_NS), wo, dir, lig.u(s - 1, W_SIDE) == 1);
				pdf_rev *= fabsf(dot(dir, cam.v(t - 1, W_NS))) / (dir_len * dir_len);
			} else {
				if (!is_light_finite(lig.u(0, W_LFLAGS))) {
					pdf_rev = k.light_pdf_pos;
					pdf_rev *= fabsf(dot(dir, cam.v(t - 1, W_NS)));
				} else {
					pdf_rev = light_pdf(lig.u(0, W_LFLAGS), lig.v(0, W_NS), dir);
					pdf_rev *= fabsf(dot(dir, cam.v(t - 1, W_NS))) / (dir_len * dir_len);
				}
			}
			cam.f(t - 1, W_PREV) = pdf_rev;
		} else {
			cam.f(t - 1, W_PREV) = 1.0f / ((float)k.P.light_triangle_count * cam.f(t - 1, W_AREA));
		}
	}
	if (t > 1) {
		idx_2_val = cam.f(t - 2, W_PREV);
		idx_2 = (uint32_t)t;
		V3 dir = cam.v(t - 2, W_POS) - cam.v(t - 1, W_POS);
		const float dir_len = length(dir);
		dir /= dir_len;
		if (s > 0) {
			lmb_material mat_tex;
			const lmb_material& mat = material_at(k.sc, cam.u(t - 1, W_MAT), cam.uv(t - 1), mat_tex);
			const V3 wo = normalize(lig.v(s - 1, W_POS) - cam.v(t - 1, W_POS));
			float pr = bsdf_pdf(mat, cam.v(t - 1, W_NS), wo, dir, cam.u(t - 1, W_SIDE) == 1);
			if (pr != 0) pr *= fabsf(dot(dir, cam.v(t - 2, W_NS))) / (dir_len * dir_len);
			cam.f(t - 2, W_PREV) = pr;
		} else {
			const float cos_x = dot(cam.v(t - 1, W_NS), dir);
			const float cos_y = dot(cam.v(t - 2, W_NS), dir);
			cam.f(t - 2, W_PREV) = fabsf(cos_x * cos_y) / (LMB_PI * dir_len * dir_len);
		}
	}
	if (s > 0) {
		idx_3_val = lig.f(s - 1, W_PREV);
		idx_3 = (uint32_t)s;
		V3 dir = lig.v(s - 1, W_POS) - cam.v(t - 1, W_POS);
		const float dir_len = length(dir);
		dir /= dir_len;
		if (t == 1) {
			const float cos_theta = dot(cam.v(0, W_NS), dir);
			float pdf = 1.0f / (cam.f(0, W_AREA) * k.screen_size * cos_theta * cos_theta * cos_theta);
			pdf *= fabsf(dot(dir, lig.v(s - 1, W_NS))) / (dir_len * dir_len);
			lig.f(s - 1, W_PREV) = pdf;
		} else {
			const V3 wo = normalize(cam.v(t - 2, W_POS) - cam.v(t - 1, W_POS));
			lmb_material mat_tex;
			const lmb_material& mat = material_at(k.sc, cam.u(t - 1, W_MAT), cam.uv(t - 1), mat_tex);
			float pr = bsdf_pdf(mat, cam.v(t - 1, W_NS), wo, dir, cam.u(t - 1, W_SIDE) == 1);
			if ((s == 1 && is_light_finite(lig.u(0, W_LFLAGS))) || s > 1) pr *= fabsf(dot(dir, lig.v(s - 1, W_NS))) / (dir_len * dir_len);
			lig.f(s - 1, W_PREV) = pr;
		}
	}
	if (s > 1) {
		idx_4_val = lig.f(s - 2, W_PREV);
		idx_4 = (uint32_t)s;
		V3 dir = lig.v(s - 2, W_POS) - lig.v(s - 1, W_POS);
		const V3 wo = normalize(cam.v(t - 1, W_POS) - lig.v(s - 1, W_POS));
		const float dir_len = length(dir);
		dir /= dir_len;
		lmb_material mat_tex;
		const lmb_material& mat = material_at(k.sc, lig.u(s - 1, W_MAT), lig.uv(s - 1), mat_tex);
		float pr = bsdf_pdf(mat, lig.v(s - 1, W_NS), wo, dir, lig.u(s - 1, W_SIDE) == 1);
		if ((s == 2 && is_light_finite(lig.u(0, W_LFLAGS))) || s > 2) pr *= fabsf(dot(dir, lig.v(s - 2, W_NS))) / (dir_len * dir_len);
		lig.f(s - 2, W_PREV) = pr;
	}
	float sum_ri = 0.0f;
	float weight = 1.0f;
	for (int i = t - 1; i > 0; i--) {
		weight *= remap0(cam.f(i, W_PREV)) / remap0(cam.f(i, W_PFWD));
		if (cam.u(i, W_DELTA) == 0 && cam.u(i - 1, W_DELTA) == 0) sum_ri += weight;
	}
	weight = 1.0f;
	for (int i = s - 1; i >= 0; i--) {
		weight *= remap0(lig.f(i, W_PREV)) / remap0(lig.f(i, W_PFWD));
		const bool delta_prev = i > 0 ? lig.u(i - 1, W_DELTA) == 1 : is_light_delta(lig.u(0, W_LFLAGS));
		if (lig.u(i, W_DELTA) == 0 && !delta_prev) sum_ri += weight;
	}
	if (s_0_changed) lig.f(0, W_PFWD) = s_0_pdf, lig.sv(0, W_POS, s_0_pdf_pos), lig.sv(0, W_NS, s_0_pdf_nrm);
	if (t_0_changed) cam.f(0, W_PFWD) = s_0_pdf, cam.sv(0, W_POS, s_0_pdf_pos), cam.sv(0, W_NS, s_0_pdf_nrm);
	if (idx_1 != 0xFFFFFFFFu) {
		cam.f(idx_1 - 1, W_PREV) = idx_1_val;
		cam.su(idx_1 - 1, W_DELTA, delta_t_old);
	}
	if (idx_2 != 0xFFFFFFFFu) cam.f(idx_2 - 2, W_PREV) = idx_2_val;
	if (idx_3 != 0xFFFFFFFFu) {
		lig.f(idx_3 - 1, W_PREV) = idx_3_val;
		lig.su(idx_3 - 1, W_DELTA, delta_s_old);
	}
	if (idx_4 != 0xFFFFFFFFu) lig.f(idx_4 - 2, W_PREV) = idx_4_val;
	return 1 / (1 + sum_ri);
}

// calc_mis_weight without the in-place patches: the values the GLSL writes into the vertices (and restores) live in registers and
// are substituted where the GLSL reads them back, so threads working on different (s, t) pairs of one pixel can run side by side
// (k_bdpt_pair). Same arithmetic, same order. t == 1 patches cam[0].pos / n_s with copies of themselves (connect_cam), and
// cam[0].pdf_fwd is not read by the camera sum (i > 0), so only the s == 1 patch of light vertex 0 needs substituting.
LMB_DN float calc_mis_weight_ro(Kctx& k, int s, int t, const Sampled& sampled) {
	const Verts& cam = k.cam;
	const Verts& lig = k.lig;
	if (s + t == 2) return 1.0f;
	const bool s1 = s == 1;
	auto lpos = [&](int i) { return (s1 && i == 0) ? sampled.pos : lig.v(i, W_POS); };
	auto lns = [&](int i) { return (s1 && i == 0) ? sampled.n_s : lig.v(i, W_NS); };
	float new1 = 0, new2 = 0, new3 = 0, new4 = 0;  // cam[t-1].pdf_rev, cam[t-2].pdf_rev, lig[s-1].pdf_rev, lig[s-2].pdf_rev
	const uint32_t lflags = lig.u(0, W_LFLAGS);
	{
		if (s > 0) {
			V3 dir = cam.v(t - 1, W_POS) - lpos(s - 1);
			const float dir_len = length(dir);
			dir /= dir_len;
			float pdf_rev = 0;
			if (s >= 2) {
				lmb_material mat_tex;
				const lmb_material& mat = material_at(k.sc, lig.u(s - 1, W_MAT), lig.uv(s - 1), mat_tex);
				const V3 wo = normalize(lig.v(s - 2, W_POS) - lig.v(s - 1, W_POS));
				pdf_rev = bsdf_pdf(mat, lig.v(s - 1, W_NS), wo, dir, lig.u(s - 1, W_SIDE) == 1);
				pdf_rev *= fabsf(dot(dir, cam.v(t - 1, W_NS))) / (dir_len * dir_len);
			} else {
				if (!is_light_finite(lflags)) {
					pdf_rev = k.light_pdf_pos;
					pdf_rev *= fabsf(dot(dir, cam.v(t - 1, W_NS)));
				} else {
					pdf_rev = light_pdf(lflags, lns(0), dir);
					pdf_rev *= fabsf(dot(dir, cam.v(t - 1, W_NS))) / (dir_len * dir_len);
				}
			}
			new1 = pdf_rev;
		} else {
			new1 = 1.0f / ((float)k.P.light_triangle_count * cam.f(t - 1, W_AREA));
		}
	}
	if (t > 1) {
		V3 dir = cam.v(t - 2, W_POS) - cam.v(t - 1, W_POS);
		const float dir_len = length(dir);
		dir /= dir_len;
		if (s > 0) {
			lmb_material mat_tex;
			const lmb_material& mat = material_at(k.sc, cam.u(t - 1, W_MAT), cam.uv(t - 1), mat_tex);
			const V3 wo = normalize(lpos(s - 1) - cam.v(t - 1, W_POS));
			float pr = bsdf_pdf(mat, cam.v(t - 1, W_NS), wo, dir, cam.u(t - 1, W_SIDE) == 1);
			if (pr != 0) pr *= fabsf(dot(dir, cam.v(t - 2, W_NS))) / (dir_len * dir_len);
			new2 = pr;
		} else {
			const float cos_x = dot(cam.v(t - 1, W_NS), dir);
			const float cos_y = dot(cam.v(t - 2, W_NS), dir);
			new2 = fabsf(cos_x * cos_y) / (LMB_PI * dir_len * dir_len);
		}
	}
	if (s > 0) {
		V3 dir = lpos(s - 1) - cam.v(t - 1, W_POS);
		const float dir_len = length(dir);
		dir /= dir_len;
		if (t == 1) {
			const float cos_theta = dot(cam.v(0, W_NS), dir);
			float pdf = 1.0f / (cam.f(0, W_AREA) * k.screen_size * cos_theta * cos_theta * cos_theta);
			pdf *= fabsf(dot(dir, lns(s - 1))) / (dir_len * dir_len);
			new3 = pdf;
		} else {
			const V3 wo = normalize(cam.v(t - 2, W_POS) - cam.v(t - 1, W_POS));
			lmb_material mat_tex;
			const lmb_material& mat = material_at(k.sc, cam.u(t - 1, W_MAT), cam.uv(t - 1), mat_tex);
			float pr = bsdf_pdf(mat, cam.v(t - 1, W_NS), wo, dir, cam.u(t - 1, W_SIDE) == 1);
			if ((s == 1 && is_light_finite(lflags)) || s > 1) pr *= fabsf(dot(dir, lns(s - 1))) / (dir_len * dir_len);
			new3 = pr;
		}
	}
	if (s > 1) {
		V3 dir = lig.v(s - 2, W_POS) - lig.v(s - 1, W_POS);
		const V3 wo = normalize(cam.v(t - 1, W_POS) - lig.v(s - 1, W_POS));
		const float dir_len = length(dir);
		dir /= dir_len;
		lmb_material mat_tex;
		const lmb_material& mat = material_at(k.sc, lig.u(s - 1, W_MAT), lig.uv(s - 1), mat_tex);
		float pr = bsdf_pdf(mat, lig.v(s - 1, W_NS), wo, dir, lig.u(s - 1, W_SIDE) == 1);
		if ((s == 2 && is_light_finite(lflags)) || s > 2) pr *= fabsf(dot(dir, lig.v(s - 2, W_NS))) / (dir_len * dir_len);
		new4 = pr;
	}
	float sum_ri = 0.0f;
	float weight = 1.0f;
	for (int i = t - 1; i > 0; i--) {
		const float prev_i = i == t - 1 ? new1 : (i == t - 2 ? new2 : cam.f(i, W_PREV));
		weight *= remap0(prev_i) / remap0(cam.f(i, W_PFWD));
		const uint32_t delta_i = i == t - 1 ? 0u : cam.u(i, W_DELTA);
		if (delta_i == 0 && cam.u(i - 1, W_DELTA) == 0) sum_ri += weight;
	}
	weight = 1.0f;
	for (int i = s - 1; i >= 0; i--) {
		const float prev_i = i == s - 1 ? new3 : (i == s - 2 ? new4 : lig.f(i, W_PREV));
		const float pfwd_i = (s1 && i == 0) ? sampled.pdf_fwd : lig.f(i, W_PFWD);
		weight *= remap0(prev_i) / remap0(pfwd_i);
		const bool delta_prev = i > 0 ? lig.u(i - 1, W_DELTA) == 1 : is_light_delta(lflags);
		const uint32_t delta_i = i == s - 1 ? 0u : lig.u(i, W_DELTA);
		if (delta_i == 0 && !delta_prev) sum_ri += weight;
	}
	return 1 / (1 + sum_ri);
}

// `ivec2(...)` of a float with no int value is undefined in GLSL: the splat is dropped (oracle/bdpt.h B5)
LMB_D bool splat_coord(float v, int& out) {
	if (!(fabsf(v) < 1e9f)) return false;
	out = (int)v;
	return true;
}

// bdpt_commons.glsl:472-530
// connect_cam_eval / connect_eval return the UNWEIGHTED radiance and what calc_mis_weight needs of `sampled`; connect_cam / connect
// apply the MIS weight. The pair kernels call the halves themselves, with a block barrier in between (k_bdpt_pair).
template <int MODE>
LMB_DN V3 connect_cam_eval(Kctx& k, int s, int& cx, int& cy, Sampled& sampled) {
	const Verts& cam = k.cam;
	const Verts& lig = k.lig;
	sampled = Sampled{v3(0.0f), v3(0.0f), 0.0f};
	V3 L = v3(0.0f);
	cx = cy = -1;
	const V3 cam_pos = cam.v(0, W_POS), cam_n = cam.v(0, W_NS);
	const V3 lpos = lig.v(s - 1, W_POS), ln = lig.v(s - 1, W_NS);
	V3 dir = cam_pos - lpos;
	const float len = length(dir);
	dir /= len;
	const float cos_y = dot(dir, ln);
	const float cos_theta = dot(cam_n, -dir);
	if (cos_theta <= 0.0f) return v3(0.0f);
	const float cos_3_theta = cos_theta * cos_theta * cos_theta;
	const float cam_pdf_ratio = fabsf(cos_y) / (cam.f(0, W_AREA) * cos_3_theta * len * len);
	const V3 ray_origin = offset_ray2(lpos, ln);
	lmb_material mat_tex;
	const lmb_material& mat = material_at(k.sc, lig.u(s - 1, W_MAT), lig.uv(s - 1), mat_tex);
	const V3 wo = normalize(lig.v(s - 2, W_POS) - lpos);
	float unused_pdf;
	const V3 f = eval_bsdf(ln, wo, mat, lig.u(s - 1, W_SIDE) == 1, dir, unused_pdf);
	if (is_zero(f)) return L;
	if (cam_pdf_ratio > 0.0f) {
		if (shadow_visible<MODE>(k, ray_origin, dir, len - LMB_EPS)) {
			sampled.pos = cam_pos;
			sampled.n_s = cam_n;
			L = lig.v(s - 1, W_THR) * cam_pdf_ratio * f / k.screen_size;
		}
	}
	dir = -dir;
	V4 target = mul(k.P.view, v4(dir.x, dir.y, dir.z, 0));
	target = v4(target.x / target.z, target.y / target.z, target.z / target.z, target.w / target.z);
	target = mul(k.P.neg_proj, target);
	const V2 cf = 0.5f * (v2(target.x, target.y) + 1.0f) * v2((float)k.P.width, (float)k.P.height) - 0.5f;
	if (!splat_coord(cf.x, cx) || !splat_coord(cf.y, cy)) {
		cx = cy = -1;
		return v3(0.0f);
	}
	if (cx < 0 || (uint32_t)cx >= k.P.width || cy < 0 || (uint32_t)cy >= k.P.height || dot(dir, cam_n) < 0) return v3(0.0f);
	return L;
}
template <int MODE>
LMB_D V3 mis_weighted_cam(Kctx& k, int s, const V3& L, const Sampled& sampled) {
	float mis_weight = 1.0f;
	if (luminance(L) != 0.0f) mis_weight = MODE == 3 ? calc_mis_weight_ro(k, s, 1, sampled) : calc_mis_weight(k, s, 1, sampled);
	return mis_weight * L;
}
template <int MODE>
LMB_D V3 connect_cam(Kctx& k, int s, int& cx, int& cy) {
	Sampled sampled;
	const V3 L = connect_cam_eval<MODE>(k, s, cx, cy, sampled);
	return mis_weighted_cam<MODE>(k, s, L, sampled);
}

// bdpt_commons.glsl:532-641
template <int MODE>
LMB_DN V3 connect_eval(Kctx& k, int s, int t, Sampled& sampled) {
	const Verts& cam = k.cam;
	const Verts& lig = k.lig;
	V3 L = v3(0.0f);
	sampled = Sampled{v3(0.0f), v3(0.0f), 0.0f};
	if (s == 0) {
		const lmb_material& mat = k.sc.materials[cam.u(t - 1, W_MAT)];  // `materials.m[mat_idx]`: un-textured, and material 0 for an escaped vertex (B3)
		L = v3(mat.emissive_factor) * cam.v(t - 1, W_THR);
	} else if (s == 1) {
		const V4 r4 = rand4(k.seed);
		const V3 cpos = cam.v(t - 1, W_POS), cn = cam.v(t - 1, W_NS);
		const LightSample ls = sample_light_Li(k.sc, r4, cpos, k.P.num_lights);
		const float cos_x = fabsf(dot(ls.wi, cn));
		const V3 ray_origin = offset_ray2(cpos, cn);
		const V3 wo = normalize(cam.v(t - 2, W_POS) - cpos);
		lmb_material mat_tex;
		const lmb_material& mat = material_at(k.sc, cam.u(t - 1, W_MAT), cam.uv(t - 1), mat_tex);
		float unused_pdf;
		const V3 f = eval_bsdf(cn, wo, mat, cam.u(t - 1, W_SIDE) == 1, ls.wi, unused_pdf);
		if (!is_zero(f)) {
			if (shadow_visible<MODE>(k, ray_origin, ls.wi, ls.wi_len - LMB_EPS)) {
				const float pdf_light_w = light_pdf_a_to_w(ls.flags, ls.pdf_a, ls.wi_len * ls.wi_len, ls.cos_from_light) / (float)k.P.light_triangle_count;
				sampled.pdf_fwd = ls.pdf_a / (float)k.P.light_triangle_count;
				light_sample_n_pos(k.sc, r4, cpos, k.P.num_lights, ls, sampled.n_s, sampled.pos);
				L = cam.v(t - 1, W_THR) * f * fabsf(cos_x) * ls.Le / pdf_light_w;
			}
		}
	} else {
		const V3 n_s = lig.v(s - 1, W_NS);
		const V3 n_t = cam.v(t - 1, W_NS);
		const V3 cpos = cam.v(t - 1, W_POS), lpos = lig.v(s - 1, W_POS);
		V3 d = lpos - cpos;
		const float len = length(d);
		d /= len;
		const float G = dot(n_s, -d) * dot(n_t, d) / (len * len);
		if (G > 0) {
			lmb_material mat_1_tex;
			const lmb_material& mat_1 = material_at(k.sc, cam.u(t - 1, W_MAT), cam.uv(t - 1), mat_1_tex);
			lmb_material mat_2_tex;
			const lmb_material& mat_2 = material_at(k.sc, lig.u(s - 1, W_MAT), lig.uv(s - 1), mat_2_tex);
			const V3 wo_1 = normalize(cam.v(t - 2, W_POS) - cpos);
			const V3 wo_2 = normalize(lig.v(s - 2, W_POS) - lpos);
			float unused_pdf;
			const V3 brdf1 = eval_bsdf(n_t, wo_1, mat_1, cam.u(t - 1, W_SIDE) == 1, d, unused_pdf);
			const V3 brdf2 = eval_bsdf(n_s, wo_2, mat_2, lig.u(s - 1, W_SIDE) == 1, -d, unused_pdf);
			if (!is_zero(brdf1) && !is_zero(brdf2)) {
				const V3 ray_origin = offset_ray2(cpos, n_t);
				if (shadow_visible<MODE>(k, ray_origin, d, len - LMB_EPS)) L = lig.v(s - 1, W_THR) * G * brdf1 * brdf2 * cam.v(t - 1, W_THR);
			}
		}
	}
	return L;
}
template <int MODE>
LMB_D V3 mis_weighted(Kctx& k, int s, int t, V3 L, const Sampled& sampled) {
	if (luminance(L) != 0.0f) {
		const float mis_weight = MODE == 3 ? calc_mis_weight_ro(k, s, t, sampled) : calc_mis_weight(k, s, t, sampled);
		L *= mis_weight;
	}
	return L;
}
template <int MODE>
LMB_D V3 connect(Kctx& k, int s, int t) {
	Sampled sampled;
	const V3 L = connect_eval<MODE>(k, s, t, sampled);
	return mis_weighted<MODE>(k, s, t, L, sampled);
}

// bdpt.rgen:55-75: the (s, t) loop. Every pair the loop body reaches owns one connection slot (staged pipeline).
template <int MODE>
LMB_D V3 connect_all(Kctx& k, int num_light_paths, int num_cam_paths) {
	const BdptParams& P = k.P;
	V3 col = v3(0.0f);
	k.slot = 0;
	const float4 dead = make_float4(__int_as_float(0x7FC00000), 0.0f, 0.0f, 0.0f);  // NaN origin: the ray hits nothing (ray_finite)
	for (int t = 1; t <= num_cam_paths; t++) {
		for (int s = 0; s <= num_light_paths; s++) {
			const int depth = s + t - 2;
			if (depth > (P.max_depth - 1) || depth < 0 || (s == 1 && t == 1)) continue;
			if (MODE == 1) P.rays[2 * ((size_t)k.slot * P.n_pix + k.pix)] = dead;
			if (t == 1) {
				int cx, cy;
				const V3 splat_col = connect_cam<MODE>(k, s, cx, cy);
				if (MODE != 1 && luminance(splat_col) > 0) {
					float* o = P.splat + 3 * ((size_t)(k.pix / P.n_real) * P.n_real + (size_t)cy * P.width + cx);  // the splat image of this pixel's frame
					atomicAdd(o + 0, splat_col.x), atomicAdd(o + 1, splat_col.y), atomicAdd(o + 2, splat_col.z);
				}
			} else {
				col += connect<MODE>(k, s, t);
			}
			k.slot++;
		}
	}
	if (MODE == 1)
		for (uint32_t c = k.slot; c < P.n_conn_slots; c++) P.rays[2 * ((size_t)c * P.n_pix + k.pix)] = dead;
	return col;
}

LMB_D void flush_counts(unsigned long long* stats, uint32_t n_closest, uint32_t n_shadow, uint32_t n_nodes, uint32_t n_tris) {
	for (int o = 16; o > 0; o >>= 1) {  // one atomic per warp and counter
		n_closest += __shfl_down_sync(0xFFFFFFFFu, n_closest, o);
		n_shadow += __shfl_down_sync(0xFFFFFFFFu, n_shadow, o);
		n_nodes += __shfl_down_sync(0xFFFFFFFFu, n_nodes, o);
		n_tris += __shfl_down_sync(0xFFFFFFFFu, n_tris, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (n_closest) atomicAdd(&stats[ST_CLOSEST], (unsigned long long)n_closest);
		if (n_shadow) atomicAdd(&stats[ST_SHADOW], (unsigned long long)n_shadow);
		if (n_nodes) atomicAdd(&stats[ST_NODES], (unsigned long long)n_nodes);
		if (n_tris) atomicAdd(&stats[ST_TRIS], (unsigned long long)n_tris);
	}
}

LMB_D Kctx make_kctx(const BdptParams& P, const DeviceScene& sc, const BvhView& bvh, uint32_t pix, uint32_t rng_w) {
	const uint32_t fb = pix / P.n_real, rpix = pix - fb * P.n_real;
	return Kctx{P, sc, bvh, Rng{rpix % P.width, rpix / P.width, (P.frame_first + fb * P.frame_stride) ^ P.time, rng_w}, Verts{P.light_verts + pix, P.n_pix},
				Verts{P.camera_verts + pix, P.n_pix}, 0.0f, (float)(P.width * P.height), pix, 0u, 0u, 0u, 0u, 0u};
}

// ---------------------------------------------------------------------------------------------- megakernel (LMB_BDPT=mega)
// bdpt.rgen:39-75 for one pixel of one frame, rays traced in the thread
__global__ void __launch_bounds__(128) k_bdpt(const __grid_constant__ BdptParams P, const __grid_constant__ DeviceScene sc, const __grid_constant__ BvhView bvh) {
	const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t n_closest = 0, n_shadow = 0, n_nodes = 0, n_tris = 0;
	if (pix < P.n_pix) {
		Kctx k = make_kctx(P, sc, bvh, pix, 0u);
		int num_light_paths = 0;
		V3 thr;
		float pdf_dir;
		if (light_begin(k, thr, pdf_dir)) {
			num_light_paths = random_walk<false>(k, k.lig, P.max_depth, thr, pdf_dir) + 1;
			light_end(k);
		}
		const float pdf = camera_begin(k, (pix % P.n_real) % P.width, (pix % P.n_real) / P.width);
		const int num_cam_paths = random_walk<true>(k, k.cam, P.max_depth, v3(1.0f), pdf) + 1;
		const V3 col = connect_all<0>(k, num_light_paths, num_cam_paths);
		P.col[pix] = make_float4(col.x, col.y, col.z, 0.0f);
		n_closest = k.n_closest, n_shadow = k.n_shadow, n_nodes = k.n_nodes, n_tris = k.n_tris;
	}
	flush_counts(P.stats, n_closest, n_shadow, n_nodes, n_tris);
}

// ---------------------------------------------------------------------------------------------- staged pipeline (default)
// The same per-pixel code cut at every ray: k_bdpt_begin -> [trace, k_bdpt_walk<light>] x max_depth -> k_bdpt_mid ->
// [trace, k_bdpt_walk<eye>] x max_depth -> k_bdpt_connect<emit> -> trace (any-hit) -> k_bdpt_connect<resolve>. Rays go through the
// Path integrator's persistent 8-wide walker (k_trace_array) as per-pixel slots; a slot without a ray holds a NaN origin.
LMB_D void store_walk(const BdptParams& P, uint32_t pix, const WalkSt& st, uint32_t& n_closest) {
	float* w = P.walk + pix;
	const size_t n = P.n_pix;
	w[(WW_POS + 0) * n] = st.ray_pos.x, w[(WW_POS + 1) * n] = st.ray_pos.y, w[(WW_POS + 2) * n] = st.ray_pos.z;
	w[(WW_WI + 0) * n] = st.wi.x, w[(WW_WI + 1) * n] = st.wi.y, w[(WW_WI + 2) * n] = st.wi.z;
	w[(WW_THR + 0) * n] = st.thr.x, w[(WW_THR + 1) * n] = st.thr.y, w[(WW_THR + 2) * n] = st.thr.z;
	w[WW_PDF * n] = st.pdf_fwd;
	w[WW_B * n] = __int_as_float(st.b);
	w[WW_ALIVE * n] = __int_as_float(st.alive ? 1 : 0);
	if (st.alive) {
		P.rays[2 * (size_t)pix] = make_float4(st.ray_pos.x, st.ray_pos.y, st.ray_pos.z, BDPT_T_MIN);
		P.rays[2 * (size_t)pix + 1] = make_float4(st.wi.x, st.wi.y, st.wi.z, BDPT_T_MAX);
		n_closest++;
	} else {
		P.rays[2 * (size_t)pix] = make_float4(__int_as_float(0x7FC00000), 0.0f, 0.0f, 0.0f);
	}
}
// Every lane of the warp calls this: the pixels whose walk goes on are appended to the next list with one atomic per warp, in lane order.
// On the classroom stand-in 3 of 32 light walkers and 13 of 32 eye walkers are alive after the first bounces (profiles/r02c): walking and
// tracing per-pixel slots ran the ~2500-instruction step at that lane count.
LMB_D void append_alive(const BdptParams& P, bool alive, uint32_t pix) {
	const uint32_t m = __ballot_sync(0xFFFFFFFFu, alive);
	if (m == 0u) return;
	const int lane = threadIdx.x & 31;
	uint32_t at = 0;
	if (lane == 0) at = atomicAdd(P.n_alive_out, (uint32_t)__popc(m));
	at = __shfl_sync(0xFFFFFFFFu, at, 0);
	if (alive) P.alive_out[at + (uint32_t)__popc(m & ((1u << lane) - 1u))] = pix;
}
LMB_D WalkSt load_walk(const BdptParams& P, uint32_t pix) {
	const float* w = P.walk + pix;
	const size_t n = P.n_pix;
	WalkSt st;
	st.ray_pos = v3(w[(WW_POS + 0) * n], w[(WW_POS + 1) * n], w[(WW_POS + 2) * n]);
	st.wi = v3(w[(WW_WI + 0) * n], w[(WW_WI + 1) * n], w[(WW_WI + 2) * n]);
	st.thr = v3(w[(WW_THR + 0) * n], w[(WW_THR + 1) * n], w[(WW_THR + 2) * n]);
	st.pdf_fwd = w[WW_PDF * n];
	st.b = __float_as_int(w[WW_B * n]);
	st.alive = __float_as_int(w[WW_ALIVE * n]) != 0;
	return st;
}

__global__ void __launch_bounds__(128) k_bdpt_begin(const __grid_constant__ BdptParams P, const __grid_constant__ DeviceScene sc) {
	const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t n_closest = 0;
	bool alive = false;
	if (pix < P.n_pix) {
		const BvhView none{nullptr, nullptr, 0};
		Kctx k = make_kctx(P, sc, none, pix, 0u);
		WalkSt st{v3(0.0f), v3(0.0f), v3(0.0f), 0.0f, 0, false};
		float pdf_dir;
		const bool ok = light_begin(k, st.thr, pdf_dir);
		if (ok) st.ray_pos = k.lig.v(0, W_POS), st.wi = k.lig.v(0, W_DIR), st.pdf_fwd = pdf_dir, st.alive = true;  // walk max_depth = pc.max_depth >= 1
		store_walk(P, pix, st, n_closest);
		P.misc[MW_RNG * (size_t)P.n_pix + pix] = k.seed.w;
		P.misc[MW_LPDFPOS * (size_t)P.n_pix + pix] = __float_as_uint(k.light_pdf_pos);
		P.misc[MW_NLIGHT * (size_t)P.n_pix + pix] = ok ? 1u : 0u;
		alive = st.alive;
	}
	append_alive(P, alive, pix);
	flush_counts(P.stats, n_closest, 0, 0, 0);
}

#ifndef LMB_BDPT_WALK_THREADS
#define LMB_BDPT_WALK_THREADS 256  // 128 / 256 / 384 / 768: classroom stand-in 19.29 / 19.04 / 18.99 / 19.04 ms per frame, cornell 2.75 / 2.78 / 2.80 / 2.82
#endif
template <bool EYE>
__global__ void __launch_bounds__(LMB_BDPT_WALK_THREADS, 768 / LMB_BDPT_WALK_THREADS) k_bdpt_walk(const __grid_constant__ BdptParams P, const __grid_constant__ DeviceScene sc) {
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if ((j & ~31u) >= *P.n_alive_in) return;  // warp-uniform: the whole warp is past the end of the list
	uint32_t n_closest = 0, pix = 0;
	bool alive = false;
	if (j < *P.n_alive_in) {
		pix = P.alive_in[j];
		WalkSt st = load_walk(P, pix);
		const BvhView none{nullptr, nullptr, 0};
		Kctx k = make_kctx(P, sc, none, pix, P.misc[MW_RNG * (size_t)P.n_pix + pix]);
		const float4 h4 = P.hits[pix];
		const Hit h{h4.x, h4.y, h4.z, __float_as_uint(h4.w)};
		walk_step<EYE>(k, EYE ? k.cam : k.lig, P.max_depth, st, h);
		store_walk(P, pix, st, n_closest);
		P.misc[MW_RNG * (size_t)P.n_pix + pix] = k.seed.w;
		alive = st.alive;
	}
	append_alive(P, alive, pix);
	flush_counts(P.stats, n_closest, 0, 0, 0);
}

__global__ void __launch_bounds__(128) k_bdpt_mid(const __grid_constant__ BdptParams P, const __grid_constant__ DeviceScene sc) {
	const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t n_closest = 0;
	if (pix < P.n_pix) {
		const BvhView none{nullptr, nullptr, 0};
		Kctx k = make_kctx(P, sc, none, pix, 0u);
		k.light_pdf_pos = __uint_as_float(P.misc[MW_LPDFPOS * (size_t)P.n_pix + pix]);
		const WalkSt lst = load_walk(P, pix);
		uint32_t num_light_paths = 0;
		if (P.misc[MW_NLIGHT * (size_t)P.n_pix + pix]) {
			num_light_paths = (uint32_t)lst.b + 1;
			light_end(k);
		}
		P.misc[MW_NLIGHT * (size_t)P.n_pix + pix] = num_light_paths;
		const float pdf = camera_begin(k, (pix % P.n_real) % P.width, (pix % P.n_real) / P.width);
		const WalkSt st{k.cam.v(0, W_POS), k.cam.v(0, W_DIR), v3(1.0f), pdf, 0, true};
		store_walk(P, pix, st, n_closest);
	}
	append_alive(P, pix < P.n_pix, pix);
	flush_counts(P.stats, n_closest, 0, 0, 0);
}

template <int MODE>
__global__ void __launch_bounds__(128) k_bdpt_connect(const __grid_constant__ BdptParams P, const __grid_constant__ DeviceScene sc) {
	const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t n_shadow = 0;
	if (pix < P.n_pix) {
		const BvhView none{nullptr, nullptr, 0};
		Kctx k = make_kctx(P, sc, none, pix, P.misc[MW_RNG * (size_t)P.n_pix + pix]);
		k.light_pdf_pos = __uint_as_float(P.misc[MW_LPDFPOS * (size_t)P.n_pix + pix]);
		const int num_light_paths = (int)P.misc[MW_NLIGHT * (size_t)P.n_pix + pix];
		const int num_cam_paths = __float_as_int(P.walk[WW_B * (size_t)P.n_pix + pix]) + 1;
		const V3 col = connect_all<MODE>(k, num_light_paths, num_cam_paths);
		if (MODE == 2) P.col[pix] = make_float4(col.x, col.y, col.z, 0.0f);
		n_shadow = k.n_shadow;
	}
	flush_counts(P.stats, 0, n_shadow, 0, 0);
}

// Pair-parallel connections (default): one thread per (connection slot, pixel), slot-major, so a warp works on ONE (s, t) strategy
// for 32 neighbouring pixels -- the per-pixel loop of k_bdpt_connect runs ~30 pairs in sequence with every lane in another branch
// (profiles/r01j: 57 % of the frame). Slots are numbered over the pairs bdpt.rgen:57-62 admits for full-length sub-paths (pair_ts),
// a pixel whose sub-paths are shorter leaves the rest dead. The rand4 of the s == 1 strategy of camera vertex t is the (t - 2)-th
// draw after the walks: every earlier t has that strategy too. MODE 1 emits the shadow rays, MODE 3 weights what was visible;
// k_bdpt_gather adds a pixel's pairs in slot order = the order of the GLSL loop, so the float sum is the same.
#ifndef LMB_BDPT_PAIR_SYNC
#define LMB_BDPT_PAIR_SYNC 1  // barriers at the start of a trip and before the MIS half: see k_bdpt_pair
#endif
#ifndef LMB_BDPT_PAIR_THREADS
#define LMB_BDPT_PAIR_THREADS 384  // 2 blocks per SM at 80 registers; 128 / 256 / 384 / 768: classroom stand-in 20.1 / 19.5 / 19.4 / 19.7 ms per frame
#endif
#ifndef LMB_BDPT_PAIR_BLOCKS
#define LMB_BDPT_PAIR_BLOCKS 6  // 4 / 6 / 8 blocks per SM: classroom stand-in 52.9 / 48.8 / 47.0 ms, cornell 3.98 / 4.02 / 4.32 ms per frame
#endif
// Which (slot, pixel) entries have work in pass MODE, as a dense list. A warp of the slot-major grid holds 32 neighbouring pixels of ONE
// pair, but only the pixels whose sub-paths are long enough have that pair at all, and in the resolve pass only those whose shadow ray
// was emitted and found nothing: on the classroom stand-in k_bdpt_pair<3> ran at 4.6 of 32 lanes with its ~3000 instructions of MIS
// code out of the instruction cache (no_instruction 10.7 warps per issue, profiles/r02c). The list keeps the slot-major, pixel-ascending
// order inside a block, so a warp of the pair kernel still works on one pair for pixels that are close: vertex reads stay mostly
// coalesced and the (s, t) loops converge. Entries without work get their zero contribution here.
template <int MODE>
__global__ void __launch_bounds__(256) k_bdpt_worklist(const __grid_constant__ BdptParams P, const uint8_t* __restrict__ pair_ts, uint32_t* __restrict__ list, uint32_t* __restrict__ count) {
	__shared__ uint32_t s_warp[8], s_base;
	const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t c = blockIdx.y;
	const int t = pair_ts[2 * c], s = pair_ts[2 * c + 1];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	bool work = false;
	uint32_t i = 0;
	if (pix < P.n_pix) {
		i = c * P.n_pix + pix;
		const int num_light_paths = (int)P.misc[MW_NLIGHT * (size_t)P.n_pix + pix];
		const int num_cam_paths = __float_as_int(P.walk[WW_B * (size_t)P.n_pix + pix]) + 1;
		work = t <= num_cam_paths && s <= num_light_paths;
		// every strategy with a light vertex is zero unless its shadow ray was emitted and found nothing: skip the re-evaluation. The host
		// sets occ to 0xFF before the emit pass and the any-hit launch writes 0 / 1 for the rays that were emitted (only those are traced)
		if (MODE == 3 && work && s > 0) work = P.occ[i] == 0;
		if (MODE == 3 && !work) P.contrib[i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu));
	}
	const uint32_t b = __ballot_sync(0xFFFFFFFFu, work);
	if (lane == 0) s_warp[warp] = (uint32_t)__popc(b);
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t total = 0;
		for (int w = 0; w < 8; w++) {
			const uint32_t n = s_warp[w];
			s_warp[w] = total;
			total += n;
		}
		s_base = total ? atomicAdd(count, total) : 0u;
	}
	__syncthreads();
	if (work) list[s_base + s_warp[warp] + (uint32_t)__popc(b & ((1u << lane) - 1u))] = i;
}

// One (pair, pixel) of the work list per thread: MODE 1 emits its shadow ray (and notes the entry in emit_list: the any-hit launch walks
// that list instead of all n_conn_slots * n_pix slots, most of which hold no ray), MODE 3 weights what was visible.
template <int MODE>
__global__ void __launch_bounds__(LMB_BDPT_PAIR_THREADS, LMB_BDPT_PAIR_BLOCKS * 128 / LMB_BDPT_PAIR_THREADS) k_bdpt_pair(const __grid_constant__ BdptParams P, const __grid_constant__ DeviceScene sc, const uint8_t* __restrict__ pair_ts,
																		   const uint32_t* __restrict__ list, const uint32_t* __restrict__ count, uint32_t* __restrict__ emit_list,
																		   uint32_t* __restrict__ emit_count) {
	const uint32_t n = *count;
	const int lane = threadIdx.x & 31;
	uint32_t n_shadow = 0;
	for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {  // block-uniform trip count
#if LMB_BDPT_PAIR_SYNC
		__syncthreads();  // the warps of a block start every trip together: they walk the same ~170 KB of code, 32 KB of which the instruction cache holds
#endif
		const uint32_t j = base + threadIdx.x;
		const bool valid = j < n;
		uint32_t i = 0, c = 0, pix = 0;
		int t = 1, s = 0, cx = -1, cy = -1;
		bool emitted = false;
		if (valid) {
			i = list[j];
			c = i / P.n_pix, pix = i - c * P.n_pix;
			t = pair_ts[2 * c], s = pair_ts[2 * c + 1];
		}
		const BvhView none{nullptr, nullptr, 0};
		Kctx k = make_kctx(P, sc, none, pix, valid ? P.misc[MW_RNG * (size_t)P.n_pix + pix] + (s == 1 ? 4u * (uint32_t)(t - 2) : 0u) : 0u);
		Sampled sampled{v3(0.0f), v3(0.0f), 0.0f};
		V3 L = v3(0.0f);
		if (valid) {  // the connection itself: geometry, BSDFs, light sample; MODE 1 writes the shadow ray, MODE 3 reads its result
			k.light_pdf_pos = __uint_as_float(P.misc[MW_LPDFPOS * (size_t)P.n_pix + pix]);
			k.slot = c;
			L = t == 1 ? connect_cam_eval<MODE>(k, s, cx, cy, sampled) : connect_eval<MODE>(k, s, t, sampled);
			n_shadow += k.n_shadow;
			emitted = k.n_shadow != 0;  // shadow_visible<1> counts the ray it writes
		}
		if (MODE == 3) {
#if LMB_BDPT_PAIR_SYNC
			__syncthreads();  // the MIS weights are the other half of the code: the block enters it together
#endif
			if (valid) {
				float4 out = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu));
				if (t == 1) {
					const V3 sp = mis_weighted_cam<MODE>(k, s, L, sampled);
					if (luminance(sp) > 0) out = make_float4(sp.x, sp.y, sp.z, __uint_as_float((uint32_t)cy * P.width + (uint32_t)cx));
				} else {
					L = mis_weighted<MODE>(k, s, t, L, sampled);
					out = make_float4(L.x, L.y, L.z, __uint_as_float(0xFFFFFFFFu));
				}
				P.contrib[i] = out;
			}
		}
		if (MODE == 1) {
			const uint32_t m = __ballot_sync(0xFFFFFFFFu, emitted);
			if (m) {
				uint32_t at = 0;
				if (lane == 0) at = atomicAdd(emit_count, (uint32_t)__popc(m));
				at = __shfl_sync(0xFFFFFFFFu, at, 0);
				if (emitted) emit_list[at + (uint32_t)__popc(m & ((1u << lane) - 1u))] = i;
			}
		}
	}
	flush_counts(P.stats, 0, n_shadow, 0, 0);
}

__global__ void __launch_bounds__(256) k_bdpt_gather(const __grid_constant__ BdptParams P, const uint8_t* __restrict__ pair_ts) {
	const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
	if (pix >= P.n_pix) return;
	const int num_light_paths = (int)P.misc[MW_NLIGHT * (size_t)P.n_pix + pix];
	const int num_cam_paths = __float_as_int(P.walk[WW_B * (size_t)P.n_pix + pix]) + 1;
	V3 col = v3(0.0f);
	for (uint32_t c = 0; c < P.n_conn_slots; c++) {
		const int t = pair_ts[2 * c], s = pair_ts[2 * c + 1];
		if (t > num_cam_paths || s > num_light_paths) continue;
		const float4 v = P.contrib[(size_t)c * P.n_pix + pix];
		if (t == 1) {
			const uint32_t target = __float_as_uint(v.w);
			if (target != 0xFFFFFFFFu) {
				float* o = P.splat + 3 * ((size_t)(pix / P.n_real) * P.n_real + target);  // the splat image of this pixel's frame
				atomicAdd(o + 0, v.x), atomicAdd(o + 1, v.y), atomicAdd(o + 2, v.z);
			}
		} else {
			col += v3(v.x, v.y, v.z);
		}
	}
	P.col[pix] = make_float4(col.x, col.y, col.z, 0.0f);
}

// bdpt.rgen:76-89: own strategies + this frame's splats -> running-mean film (NaN samples leave the pixel untouched) or sum film;
// clears the splat image for the next frame.
__global__ void __launch_bounds__(256) k_bdpt_film(uint32_t n_real, uint32_t n_batch_frames, uint32_t frame_first, uint32_t frame_stride, int film_mode, const float4* __restrict__ colb,
													float* __restrict__ splat, float4* __restrict__ film, unsigned long long* stats) {
	uint32_t nan_count = 0;
	for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n_real; pix += gridDim.x * blockDim.x) {
		float4 acc = film[pix];
		for (uint32_t fb = 0; fb < n_batch_frames; fb++) {  // the frames of the batch in order: the running mean is the reference's
			const size_t vp = (size_t)fb * n_real + pix;
			const uint32_t frame = frame_first + fb * frame_stride;
			V3 col = xyz(V4{colb[vp].x, colb[vp].y, colb[vp].z, 0.0f});
			col += v3(splat[3 * vp + 0], splat[3 * vp + 1], splat[3 * vp + 2]);
			splat[3 * vp + 0] = 0.0f, splat[3 * vp + 1] = 0.0f, splat[3 * vp + 2] = 0.0f;
			const float lum = luminance(col);
			if (lum != lum) {
				nan_count++;
				continue;
			}
			if (film_mode == LMB_FILM_SUM) {  // un-normalised sum, valid-sample count in alpha (multi-GPU sample-index shards; lmb_resolve divides)
				acc = make_float4(acc.x + col.x, acc.y + col.y, acc.z + col.z, acc.w + 1.0f);
			} else if (frame > 0) {
				const float w = 1.0f / float(frame + 1);
				const V3 m = mix(v3(acc.x, acc.y, acc.z), col, w);
				acc = make_float4(m.x, m.y, m.z, 1.0f);
			} else {
				acc = make_float4(col.x, col.y, col.z, 1.0f);
			}
		}
		film[pix] = acc;
	}
	if (nan_count) atomicAdd(&stats[ST_NAN], (unsigned long long)nan_count);
}

}  // namespace

void bdpt_free(lmb_ctx* ctx) {
	BdptState& b = ctx->bdpt;
	cudaFree(b.light_verts), cudaFree(b.camera_verts), cudaFree(b.col), cudaFree(b.splat);
	cudaFree(b.walk), cudaFree(b.misc), cudaFree(b.rays), cudaFree(b.hits), cudaFree(b.occ), cudaFree(b.contrib), cudaFree(b.pair_ts), cudaFree(b.work_list), cudaFree(b.emit_list), cudaFree(b.work_count);
	cudaFree(b.alive_list[0]), cudaFree(b.alive_list[1]), cudaFree(b.alive_count);
	b = BdptState{};
}

// connection slots a pixel can use: the pairs bdpt.rgen:57-62 lets through when both sub-paths have their full length
static uint32_t count_conn_slots(int max_depth) {
	uint32_t n = 0;
	for (int t = 1; t <= max_depth + 1; t++)
		for (int s = 0; s <= max_depth + 1; s++) {
			const int depth = s + t - 2;
			if (depth > (max_depth - 1) || depth < 0 || (s == 1 && t == 1)) continue;
			n++;
		}
	return n;
}

int bdpt_render(lmb_ctx* ctx, const lmb_pc_bdpt& pc, const lmb_scene_ubo& ubo, uint32_t first_frame, uint32_t n_frames, uint32_t frame_stride, int film_mode,
				float* raw_col, float* raw_splat) {
	if (ctx->row_stride != 1) return set_error(ctx, LMB_ERR_INVALID, "lmb_render_bdpt: pixel shards are not supported (light-tracer splats cross rows)");
	if (pc.max_depth < 1 || pc.max_depth > 63) return set_error(ctx, LMB_ERR_INVALID, "lmb_render_bdpt: max_depth must be in [1, 63]");
	if (!ctx->wf.stats) return set_error(ctx, LMB_ERR_INVALID, "lmb_render_bdpt: call lmb_init first");
	// LMB_BDPT=mega: the one-kernel version (rays traced in the thread over the binary LBVH); default: the staged pipeline
	const char* mode_env = getenv("LMB_BDPT");
	const bool mega = mode_env && strcmp(mode_env, "mega") == 0;
	const bool per_pixel = mode_env && strcmp(mode_env, "pixel") == 0;  // staged, connections looped per pixel (k_bdpt_connect)
	BdptState& b = ctx->bdpt;
	cudaStream_t st = ctx->stream;
	const uint32_t n_real = ctx->width * ctx->height;
	const uint32_t n_verts = (uint32_t)pc.max_depth + 1;
	const uint32_t n_conn_slots = count_conn_slots(pc.max_depth);
	if ((uint64_t)n_real * n_conn_slots > 0x7FFFFFFFull) return set_error(ctx, LMB_ERR_INVALID, "lmb_render_bdpt: width * height * connection slots out of range");
	// Frames in flight: the kernels work on (frame of the batch, pixel) pairs, so that the ~17 closest / any-hit launches and the tails of
	// every kernel are paid once per BATCH. Sized by memory (~4 KB per pixel and frame at depth 8: two sub-paths, ray slots, contributions,
	// lists; 40 GB or half of what the device has free, whichever is less), at most 8, within the 31-bit slot index; LMB_BDPT_BATCH overrides; the single-frame test hook renders one at a time.
	const uint64_t bytes_per_pixel = 2ull * n_verts * W_COUNT * 4 + (uint64_t)n_conn_slots * (32 + 16 + 4 + 4 + 1) + (WW_COUNT + MW_COUNT) * 4 + 16 + 8 + 16 + 12;
	size_t mem_free = 0, mem_total = 0;
	cudaMemGetInfo(&mem_free, &mem_total);
	const uint64_t have = (uint64_t)mem_free + (b.n_pix ? (uint64_t)b.n_pix * bytes_per_pixel : 0);  // what is free now + what this state already holds
	const uint64_t budget = std::min<uint64_t>(40ull << 30, have / 2);
	uint32_t batch = (uint32_t)std::min<uint64_t>(8, std::max<uint64_t>(1, budget / (bytes_per_pixel * n_real)));
	batch = (uint32_t)std::min<uint64_t>(batch, 0x7FFFFFFFull / ((uint64_t)n_real * n_conn_slots));
	if (const char* e = getenv("LMB_BDPT_BATCH")) batch = std::max(1, atoi(e));
	if (raw_col || mega || per_pixel) batch = 1;
	uint32_t n_pix = n_real * batch;  // capacity of the buffers in (frame, pixel) pairs
	// all or nothing: a failed allocation (the batch takes tens of GB) must not leave a half-built state behind for the next call
	auto allocate = [&]() -> int {
		if (b.n_pix != n_pix || b.n_verts != n_verts) {
			bdpt_free(ctx);
			LMB_CUDA(ctx, cudaMalloc((void**)&b.light_verts, (size_t)n_pix * n_verts * W_COUNT * 4));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.camera_verts, (size_t)n_pix * n_verts * W_COUNT * 4));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.col, (size_t)n_pix * 16));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.splat, (size_t)n_pix * 12));
			b.n_pix = n_pix, b.n_verts = n_verts;
		}
		if (!mega && !b.rays) {
			LMB_CUDA(ctx, cudaMalloc((void**)&b.walk, (size_t)n_pix * WW_COUNT * 4));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.misc, (size_t)n_pix * MW_COUNT * 4));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.rays, (size_t)n_pix * n_conn_slots * 32));
			// a dead slot is marked in its first float4 only; the walker loads both halves of every slot, so the second must be defined
			LMB_CUDA(ctx, cudaMemsetAsync(b.rays, 0, (size_t)n_pix * n_conn_slots * 32, st));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.hits, (size_t)n_pix * 16));
			for (int i = 0; i < 2; i++) LMB_CUDA(ctx, cudaMalloc((void**)&b.alive_list[i], (size_t)n_pix * 4));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.alive_count, (2 * 64 + 4) * 4));  // max_depth <= 64 (check_bdpt_args)
			LMB_CUDA(ctx, cudaMalloc((void**)&b.occ, (size_t)n_pix * n_conn_slots));
			b.n_conn_slots = n_conn_slots;
		}
		if (!mega && !per_pixel && !b.contrib) {
			std::vector<uint8_t> ts;
			for (int t = 1; t <= pc.max_depth + 1; t++)
				for (int s = 0; s <= pc.max_depth + 1; s++) {
					const int depth = s + t - 2;
					if (depth > (pc.max_depth - 1) || depth < 0 || (s == 1 && t == 1)) continue;
					ts.push_back((uint8_t)t), ts.push_back((uint8_t)s);
				}
			LMB_CUDA(ctx, cudaMalloc((void**)&b.contrib, (size_t)n_pix * n_conn_slots * 16));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.pair_ts, ts.size()));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.work_list, (size_t)n_pix * n_conn_slots * 4));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.emit_list, (size_t)n_pix * n_conn_slots * 4));
			LMB_CUDA(ctx, cudaMalloc((void**)&b.work_count, 12));
			LMB_CUDA(ctx, cudaMemcpy(b.pair_ts, ts.data(), ts.size(), cudaMemcpyHostToDevice));
		}
		return 0;
	};
	for (;;) {
		const int rc_alloc = allocate();
		if (rc_alloc == 0) break;
		bdpt_free(ctx);
		cudaGetLastError();  // an out-of-memory error is not sticky: clear it
		if (batch == 1) return rc_alloc;
		batch = 1, n_pix = n_real;  // fall back to one frame in flight
	}
	LMB_CUDA(ctx, cudaMemsetAsync(b.splat, 0, (size_t)n_pix * 12, st));
	auto load = [](const float* p) {
		M4 m;
		for (int c = 0; c < 4; c++) m.c[c] = V4{p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]};
		return m;
	};
	BdptParams P;
	P.inv_view = load(ubo.inv_view), P.inv_proj = load(ubo.inv_projection), P.view = load(ubo.view);
	P.neg_proj = load(ubo.projection);
	for (int c = 0; c < 4; c++) P.neg_proj.c[c] = V4{-P.neg_proj.c[c].x, -P.neg_proj.c[c].y, -P.neg_proj.c[c].z, -P.neg_proj.c[c].w};
	P.width = ctx->width, P.height = ctx->height, P.n_real = n_real, P.frame_stride = frame_stride, P.time = pc.time;
	P.num_lights = pc.num_lights, P.max_depth = pc.max_depth, P.light_triangle_count = pc.light_triangle_count;
	P.light_verts = b.light_verts, P.camera_verts = b.camera_verts, P.col = b.col, P.splat = b.splat, P.stats = ctx->wf.stats;
	P.walk = b.walk, P.misc = b.misc, P.rays = b.rays, P.hits = b.hits, P.occ = b.occ, P.contrib = b.contrib, P.n_conn_slots = n_conn_slots;
	const BvhView bvh{ctx->bvh.nodes, ctx->bvh.tris, ctx->bvh.n};
	int rc;
	cudaEventRecord(ctx->ev[0], st);
	for (uint32_t done = 0; done < n_frames;) {
		const uint32_t nb = std::min(batch, n_frames - done);
		const uint32_t n_pix = nb * n_real;  // (frame, pixel) pairs of THIS batch: also the stride of every struct-of-arrays below
		const size_t vert_bytes = (size_t)n_pix * n_verts * W_COUNT * 4;
		const uint32_t grid = (n_pix + 127) / 128;
		P.n_pix = n_pix, P.frame_first = first_frame + done * frame_stride;
		// BDPT.cpp:79-80: both vertex buffers are zeroed before every frame
		LMB_CUDA(ctx, cudaMemsetAsync(b.light_verts, 0, vert_bytes, st));
		LMB_CUDA(ctx, cudaMemsetAsync(b.camera_verts, 0, vert_bytes, st));
		if (mega) {
			k_bdpt<<<grid, 128, 0, st>>>(P, ctx->scene, bvh);
			ctx->stats.kernel_launches += 1;
		} else {
			// walkers travel as dense pixel lists: list `cur` with its length at alive_count[step]; a step traces and walks that list and
			// appends the survivors to the other buffer
			const uint32_t n_steps = 2u * (uint32_t)pc.max_depth + 2u;
			LMB_CUDA(ctx, cudaMemsetAsync(b.alive_count, 0, n_steps * 4, st));
			uint32_t step = 0;
			int cur = 0;
			auto lists = [&](bool has_in) {
				P.alive_in = has_in ? b.alive_list[cur] : nullptr, P.n_alive_in = has_in ? b.alive_count + step : nullptr;
				if (has_in) cur ^= 1, step++;
				P.alive_out = b.alive_list[cur], P.n_alive_out = b.alive_count + step;
			};
			lists(false);
			k_bdpt_begin<<<grid, 128, 0, st>>>(P, ctx->scene);
			for (int d = 0; d < pc.max_depth; d++) {
				if ((rc = launch_trace_slot_list(ctx, b.rays, b.alive_list[cur], b.alive_count + step, n_pix, b.hits, nullptr, false))) return rc;
				lists(true);
				k_bdpt_walk<false><<<(n_pix + LMB_BDPT_WALK_THREADS - 1) / LMB_BDPT_WALK_THREADS, LMB_BDPT_WALK_THREADS, 0, st>>>(P, ctx->scene);
			}
			step++;  // the eye walk starts a list of its own
			lists(false);
			k_bdpt_mid<<<grid, 128, 0, st>>>(P, ctx->scene);
			for (int d = 0; d < pc.max_depth; d++) {
				if ((rc = launch_trace_slot_list(ctx, b.rays, b.alive_list[cur], b.alive_count + step, n_pix, b.hits, nullptr, false))) return rc;
				lists(true);
				k_bdpt_walk<true><<<(n_pix + LMB_BDPT_WALK_THREADS - 1) / LMB_BDPT_WALK_THREADS, LMB_BDPT_WALK_THREADS, 0, st>>>(P, ctx->scene);
			}
			if (per_pixel) {
				k_bdpt_connect<1><<<grid, 128, 0, st>>>(P, ctx->scene);
				if ((rc = launch_trace_slots(ctx, b.rays, n_pix * n_conn_slots, nullptr, b.occ, true))) return rc;
				k_bdpt_connect<2><<<grid, 128, 0, st>>>(P, ctx->scene);
				ctx->stats.kernel_launches += 5 + 4 * (uint64_t)pc.max_depth;
			} else {
				const dim3 wgrid((n_pix + 255) / 256, n_conn_slots);
				const int pair_grid = ctx->sm_count * LMB_BDPT_PAIR_BLOCKS * 2;
				LMB_CUDA(ctx, cudaMemsetAsync(b.work_count, 0, 12, st));
				LMB_CUDA(ctx, cudaMemsetAsync(b.occ, 0xFF, (size_t)n_pix * n_conn_slots, st));  // "no ray emitted" until the any-hit launch says otherwise
				k_bdpt_worklist<1><<<wgrid, 256, 0, st>>>(P, b.pair_ts, b.work_list, b.work_count);
				k_bdpt_pair<1><<<pair_grid * 128 / LMB_BDPT_PAIR_THREADS, LMB_BDPT_PAIR_THREADS, 0, st>>>(P, ctx->scene, b.pair_ts, b.work_list, b.work_count, b.emit_list, b.work_count + 2);
				if ((rc = launch_trace_slot_list(ctx, b.rays, b.emit_list, b.work_count + 2, n_pix * n_conn_slots, nullptr, b.occ, true))) return rc;
				k_bdpt_worklist<3><<<wgrid, 256, 0, st>>>(P, b.pair_ts, b.work_list, b.work_count + 1);
				k_bdpt_pair<3><<<pair_grid * 128 / LMB_BDPT_PAIR_THREADS, LMB_BDPT_PAIR_THREADS, 0, st>>>(P, ctx->scene, b.pair_ts, b.work_list, b.work_count + 1, nullptr, nullptr);
				k_bdpt_gather<<<(n_pix + 255) / 256, 256, 0, st>>>(P, b.pair_ts);
				ctx->stats.kernel_launches += 8 + 4 * (uint64_t)pc.max_depth;
			}
		}
		if (raw_col) {  // test hook: the two images before the film update
			LMB_CUDA(ctx, cudaMemcpyAsync(raw_col, b.col, (size_t)n_pix * 16, cudaMemcpyDeviceToHost, st));
			LMB_CUDA(ctx, cudaMemcpyAsync(raw_splat, b.splat, (size_t)n_pix * 12, cudaMemcpyDeviceToHost, st));
		}
		k_bdpt_film<<<ctx->sm_count * 8, 256, 0, st>>>(n_real, nb, P.frame_first, frame_stride, film_mode, b.col, b.splat, ctx->film, ctx->wf.stats);
		ctx->stats.kernel_launches += 1;
		done += nb;
	}
	cudaEventRecord(ctx->ev[5], st);
	LMB_CUDA(ctx, cudaStreamSynchronize(st));
	LMB_CUDA(ctx, cudaGetLastError());
	float ms = 0;
	cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[5]);
	ctx->stats.ms_render += ms;
	ctx->stats.frames += n_frames;
	return 0;
}

}  // namespace lmb
