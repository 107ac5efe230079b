// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl_compat.h).
//
// shading.h: CPU restatement (C++/glm) of the GLSL the reference's Path integrator executes per hit:
// RNG, ray offsets, BSDF sample/eval for all six material types, light sampling, texture fetch and the sky model.
// Every function names the reference lines it follows. Arithmetic is fp32, evaluated in the order the GLSL
// expression trees prescribe; transcendental functions come from include/lmb_detmath.h.
//
// Frozen quirks: Q1 eval_dielectric's transmission branch shadows `f` and returns the outer,
// uninitialised vec3 -> defined here as vec3(0); Q2 sample_clearcoat derives sin_t from the inout cos_theta (0 on
// entry); Q3 sample_principled returns the un-weighted lobe value; Q5 light pick pdf = 1/light_triangle_count;
// Q7 sample_triangle applies the world matrix to edge vectors and normals with w = 1.
#pragma once
#include "glsl_compat.h"

namespace orc {

// ------------------------------------------------------------------------------------------------ RNG
// utils.glsl:121-134
inline uvec4 pcg4d(uvec4 v) {
	v = v * 1664525u + 1013904223u;
	v.x += v.y * v.w;
	v.y += v.z * v.x;
	v.z += v.x * v.y;
	v.w += v.y * v.z;
	v = v ^ (v >> 16u);
	v.x += v.y * v.w;
	v.y += v.z * v.x;
	v.z += v.x * v.y;
	v.w += v.y * v.z;
	return v;
}
// utils.glsl:137
inline float uint_to_float(uint x) { return lmb_bits2f(0x3f800000u | (x >> 9)) - 1.0f; }
// utils.glsl:145-154. GLSL evaluates constructor arguments left to right; C++ does not, hence the temporaries.
inline float rand1(uvec4& s) {
	s.w++;
	return uint_to_float(pcg4d(s).x);
}
inline vec2 rand2(uvec4& s) {
	const float a = rand1(s);
	const float b = rand1(s);
	return vec2(a, b);
}
inline vec3 rand3(uvec4& s) {
	const vec2 ab = rand2(s);
	const float c = rand1(s);
	return vec3(ab, c);
}
inline vec4 rand4(uvec4& s) {
	const vec3 abc = rand3(s);
	const float d = rand1(s);
	return vec4(abc, d);
}

// ------------------------------------------------------------------------------------------------ small utils
// utils.glsl:73-85 (Ray Tracing Gems ch. 6)
inline vec3 offset_ray(const vec3& p, const vec3& n) {
	const float origin = 1.0f / 32.0f;
	const float float_scale = 1.0f / 65536.0f;
	const float int_scale = 256.0f;
	const ivec3 of_i((int)(int_scale * n.x), (int)(int_scale * n.y), (int)(int_scale * n.z));
	auto bump = [](float v, int o) { return lmb_bits2f((uint32_t)((int32_t)lmb_f2bits(v) + ((v < 0) ? -o : o))); };
	const vec3 p_i(bump(p.x, of_i.x), bump(p.y, of_i.y), bump(p.z, of_i.z));
	return vec3(std::fabs(p.x) < origin ? p.x + float_scale * n.x : p_i.x,
				std::fabs(p.y) < origin ? p.y + float_scale * n.y : p_i.y,
				std::fabs(p.z) < origin ? p.z + float_scale * n.z : p_i.z);
}
// utils.glsl:91-94
inline vec3 offset_ray2(const vec3& p, const vec3& n) {
	const float float_scale = 2.0f / 65536.0f;
	return p + float_scale * n;
}
// utils.glsl:100
inline float luminance(const vec3& rgb) { return glm::dot(rgb, vec3(0.2126f, 0.7152f, 0.0722f)); }
// utils.glsl:185-191
inline void branchless_onb(const vec3& n, vec3& b1, vec3& b2) {
	const float sign = n.z >= 0.0f ? 1.0f : -1.0f;
	const float a = -1.0f / (sign + n.z);
	const float b = n.x * n.y * a;
	b1 = vec3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
	b2 = vec3(b, sign + n.y * n.y * a, -n.y);
}
// utils.glsl:193-195
inline vec3 to_world(const vec3& v, const vec3& T, const vec3& B, const vec3& N) { return v.x * T + v.y * B + v.z * N; }
inline vec3 to_local(const vec3& v, const vec3& T, const vec3& B, const vec3& N) {
	return vec3(glm::dot(v, T), glm::dot(v, B), glm::dot(v, N));
}
// utils.glsl:215-229
inline vec2 concentric_sample_disk(const vec2& rands) {
	const vec2 offset = 2.0f * rands - 1.0f;
	if (offset.x == 0 && offset.y == 0) return vec2(0);
	float theta, r;
	if (std::fabs(offset.x) > std::fabs(offset.y)) {
		r = offset.x;
		theta = 0.25f * PI * offset.y / offset.x;
	} else {
		r = offset.y;
		theta = PI * (0.5f - 0.25f * offset.x / offset.y);
	}
	float s, c;
	lmb_sincosf(theta, &s, &c);
	return r * vec2(c, s);
}
inline bool has_prop(uint props, uint flag) { return (props & flag) != 0; }

// ------------------------------------------------------------------------------------------------ sampling_commons.glsl
// :16-20
inline bool effectively_delta(float alpha) { return alpha <= 0.0064f; }
inline bool effectively_delta(const vec2& alpha) { return glm::min(alpha.x, alpha.y) <= 0.0064f; }

// :22-46 refract(); always returns true in the reference (TIR falls back to reflect)
inline void refract_dir(const vec3& n_s, const vec3& wo, bool forward_facing, float eta, uint mode, vec3& wi, vec3& f,
						float& inv_eta) {
	const float cos_i = glm::dot(n_s, wo);
	inv_eta = forward_facing ? 1.0f / eta : eta;
	const float sin2_t = inv_eta * inv_eta * (1.0f - cos_i * cos_i);
	if (sin2_t >= 1.0f) {
		wi = glm::reflect(-wo, n_s);
	} else {
		const float cos_t = std::sqrt(1 - sin2_t);
		wi = -inv_eta * wo + (inv_eta * cos_i - cos_t) * n_s;
	}
	f = mode == 1 ? vec3(inv_eta * inv_eta) : vec3(1);
}

// :48-65
inline float fresnel_dielectric(float cos_i, float eta, bool forward_facing) {
	cos_i = glm::clamp(cos_i, -1.0f, 1.0f);
	if (!forward_facing) eta = 1.0f / eta;
	if (cos_i < 0) {
		eta = 1.0f / eta;
		cos_i = -cos_i;
	}
	const float sin2_i = 1 - cos_i * cos_i;
	const float sin2_t = sin2_i / (eta * eta);
	if (sin2_t >= 1) return 1.f;
	const float cos_t = std::sqrt(1 - sin2_t);
	const float r_parallel = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
	const float r_perp = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
	return 0.5f * (r_parallel * r_parallel + r_perp * r_perp);
}

// :68-82
inline float fresnel_conductor(float cos_i, float eta, float k) {
	const float cos_sqr = cos_i * cos_i;
	const float sin_sqr = glm::max(1.0f - cos_sqr, 0.0f);
	const float sin_4 = sin_sqr * sin_sqr;
	const float inner_term = eta * eta - k * k - sin_sqr;
	const float a_sq_p_b_sq = std::sqrt(glm::max(inner_term * inner_term + 4.0f * eta * eta * k * k, 0.0f));
	const float a = std::sqrt(glm::max((a_sq_p_b_sq + inner_term) * 0.5f, 0.0f));
	const float rs = ((a_sq_p_b_sq + cos_sqr) - (2.0f * a * cos_i)) / ((a_sq_p_b_sq + cos_sqr) + (2.0f * a * cos_i));
	const float rp = ((cos_sqr * a_sq_p_b_sq + sin_4) - (2.0f * a * cos_i * sin_sqr)) /
					 ((cos_sqr * a_sq_p_b_sq + sin_4) + (2.0f * a * cos_i * sin_sqr));
	return 0.5f * (rs + rs * rp);
}
// :141-144
inline vec3 fresnel_conductor(float cos_i, const vec3& eta, const vec3& k) {
	return vec3(fresnel_conductor(cos_i, eta.x, k.x), fresnel_conductor(cos_i, eta.y, k.y),
				fresnel_conductor(cos_i, eta.z, k.z));
}
// :84-92
inline float fresnel_schlick(float f0, float f90, float ns) {
	return f0 + (f90 - f0) * g_pow(glm::max(1.0f - ns, 0.0f), 5.0f);
}
inline vec3 fresnel_schlick(const vec3& f0, const vec3& f90, float ns) {
	return f0 + (f90 - f0) * g_pow(glm::max(1.0f - ns, 0.0f), 5.0f);
}
// :95-98
inline float eta_to_schlick_R0(float eta) {
	const float val = (eta - 1.0f) / (eta + 1.0f);
	return val * val;
}
// :102-111
inline float disney_fresnel(const vec3& wi, const vec3& wo, float roughness, float& f_wi, float& f_wo) {
	const vec3 h = glm::normalize(wi + wo);
	const float wo_dot_h = glm::dot(wo, h);
	const float fd90 = 0.5f + 2.0f * wo_dot_h * wo_dot_h * roughness;
	const float fd0 = 1.f;
	f_wi = fresnel_schlick(fd0, fd90, wi.z);
	f_wo = fresnel_schlick(fd0, fd90, wo.z);
	return f_wi * f_wo;
}
// :120-123
inline vec3 calc_tint(const vec3& albedo) {
	const float lum = luminance(albedo);
	return lum > 0 ? albedo / lum : vec3(1);
}
// :126-139
inline vec3 disney_fresnel(const lmb_material& mat, const vec3& wo, const vec3& h, const vec3& /*wi*/, float eta) {
	const float wo_dot_h = glm::dot(wo, h);
	const vec3 albedo = v3(mat.albedo);
	vec3 R0 = eta_to_schlick_R0(eta) * glm::mix(vec3(1.0f), calc_tint(albedo), mat.specular_tint);
	R0 = glm::mix(R0, albedo, mat.metallic);
	const float fr_dielectric = fresnel_dielectric(wo_dot_h, eta, true);
	const vec3 fr_metallic = fresnel_schlick(R0, vec3(1), wo_dot_h);
	return glm::mix(vec3(fr_dielectric), fr_metallic, mat.metallic);
}
// :164-174 (SAMPLING_MODE == concentric disk)
inline vec3 sample_hemisphere(const vec2& xi) {
	const vec2 d = concentric_sample_disk(xi);
	const float z = std::sqrt(glm::max(0.f, 1.f - glm::dot(d, d)));
	return vec3(d, z);
}

// ------------------------------------------------------------------------------------------------ microfacet_commons.glsl
// :5-12
inline float smith_lambda_iso(float alpha_sqr, float cos_theta) {
	if (cos_theta == 0) return 0;
	const float cos_sqr = cos_theta * cos_theta;
	const float tan_sqr = glm::max(1.0f - cos_sqr, 0.0f) / cos_sqr;
	return 0.5f * (std::sqrt(1.0f + alpha_sqr * tan_sqr) - 1);
}
// :13-24
inline float smith_lambda_aniso(const vec3& w, const vec2& alpha) {
	const float cos_sqr = w.z * w.z;
	const float sin_sqr = glm::max(1.0f - cos_sqr, 0.0f);
	const float tan_sqr = sin_sqr / cos_sqr;
	if (g_isinf(tan_sqr)) return 0.0f;
	const vec2 cos_phi_sqr = sin_sqr == 0.0f ? vec2(1.0f, 0.0f) : glm::clamp(vec2(w.x * w.x, w.y * w.y), 0.0f, 1.0f) / sin_sqr;
	const float alpha_sqr = glm::dot(cos_phi_sqr, alpha * alpha);
	return 0.5f * (std::sqrt(1.0f + alpha_sqr * tan_sqr) - 1);
}
// :26
inline float g1_ggx_aniso(const vec3& w, const vec2& alpha) { return 1.0f / (1.0f + smith_lambda_aniso(w, alpha)); }
// :28-31
inline float g_ggx_corr_iso(float alpha, const vec3& wo, const vec3& wi) {
	const float alpha_sqr = alpha * alpha;
	return 1.0f / (1.0f + smith_lambda_iso(alpha_sqr, wo.z) + smith_lambda_iso(alpha_sqr, wi.z));
}
// :33-35
inline float g_ggx_corr_aniso(const vec2& alpha, const vec3& wo, const vec3& wi) {
	return 1.0f / (1.0f + smith_lambda_aniso(wo, alpha) + smith_lambda_aniso(wi, alpha));
}
// :37-53
inline float d_ggx_aniso(const vec2& alpha, const vec3& h) {
	const float cos_sqr = h.z * h.z;
	const float sin_sqr = glm::max(1.0f - cos_sqr, 0.0f);
	const float tan_sqr = sin_sqr / cos_sqr;
	if (g_isinf(tan_sqr)) return 0.0f;
	const float cos_4 = cos_sqr * cos_sqr;
	if (cos_4 < 1e-16f) return 0.0f;
	const vec2 phi_sqr = sin_sqr == 0.0f ? vec2(1.0f, 0.0f) : glm::clamp(vec2(h.x * h.x, h.y * h.y), 0.0f, 1.0f) / sin_sqr;
	const vec2 alpha_sqr = phi_sqr / (alpha * alpha);
	const float e = tan_sqr * (alpha_sqr.x + alpha_sqr.y);
	return 1.0f / (PI * alpha.x * alpha.y * cos_4 * (1.0f + e) * (1.0f + e));
}
// :56-59
inline float d_ggx_iso(float alpha_sqr, float cos_theta) {
	const float d = ((cos_theta * alpha_sqr - cos_theta) * cos_theta + 1);
	return alpha_sqr / (d * d * PI);
}
// :62-66
inline float g1_ggx_iso(float alpha_sqr, float cos_theta) {
	const float cos_sqr = cos_theta * cos_theta;
	const float tan_sqr = glm::max(1.0f - cos_sqr, 0.0f) / cos_sqr;
	return 2.0f / (1.0f + std::sqrt(1.0f + alpha_sqr * tan_sqr));
}
// :69-78
inline float vndf_pdf_iso(float alpha, const vec3& wo, const vec3& h, float& D) {
	D = 0.0f;
	if (wo.z <= 0) return 0.0f;
	const float alpha_sqr = alpha * alpha;
	const float G1 = g1_ggx_iso(alpha_sqr, wo.z);
	D = d_ggx_iso(alpha_sqr, h.z);
	return G1 * D * glm::max(0.0f, glm::dot(wo, h)) / std::fabs(wo.z);
}
// :86-94
inline float vndf_pdf_aniso(const vec2& alpha, const vec3& wo, const vec3& h, float& D) {
	D = 0.0f;
	if (wo.z <= 0) return 0.0f;
	const float G1 = g1_ggx_aniso(wo, alpha);
	D = d_ggx_aniso(alpha, h);
	return G1 * D * glm::max(0.0f, glm::dot(wo, h)) / std::fabs(wo.z);
}
// :101-125 (spherical-caps branch)
inline vec3 sample_ggx_vndf_common(const vec2& alpha, const vec3& wo, const vec2& xi) {
	const vec3 wo_hemisphere = glm::normalize(vec3(alpha.x * wo.x, alpha.y * wo.y, wo.z));
	const float phi = 2.0f * PI * xi.x;
	const float z = ((1.0f - xi.y) * (1.0f + wo_hemisphere.z)) - wo_hemisphere.z;
	const float sin_theta = std::sqrt(glm::clamp(1.0f - z * z, 0.0f, 1.0f));
	float sp, cp;
	lmb_sincosf(phi, &sp, &cp);
	const float x = sin_theta * cp;
	const float y = sin_theta * sp;
	const vec3 n_h = vec3(x, y, z) + wo_hemisphere;
	return glm::normalize(vec3(alpha.x * n_h.x, alpha.y * n_h.y, glm::max(0.0f, n_h.z)));
}

struct BsdfSample {
	vec3 f{0};
	vec3 wi{0};
	float pdf = 0;
	float cos_theta = 0;
};

// ------------------------------------------------------------------------------------------------ diffuse.glsl (Lambertian mode)
// :13-21
inline BsdfSample sample_lambertian(const lmb_material& mat, const vec3& wo, const vec2& xi) {
	BsdfSample s;
	s.wi = sample_hemisphere(xi);
	s.cos_theta = s.wi.z;
	s.pdf = s.cos_theta * INV_PI;
	if (glm::min(s.wi.z, wo.z) <= 0.0f) {
		s.f = vec3(0);
		return s;
	}
	s.f = v3(mat.albedo) * INV_PI;
	return s;
}
// :58-65
inline vec3 eval_lambertian(const lmb_material& mat, const vec3& wo, const vec3& wi, float& pdf_w) {
	if (glm::min(wi.z, wo.z) <= 0.0f) return vec3(0);
	pdf_w = wi.z * INV_PI;
	return v3(mat.albedo) * INV_PI;
}
// :85-90
inline float lambertian_pdf(const vec3& wo, const vec3& wi) {
	if (glm::min(wi.z, wo.z) <= 0.0f) return 0.0f;
	return wi.z * INV_PI;
}

// ------------------------------------------------------------------------------------------------ mirror.glsl / glass.glsl
// mirror.glsl:3-14 with n_s = (0,0,1)
inline BsdfSample sample_mirror(const vec3& wo) {
	BsdfSample s;
	const vec3 n_s(0, 0, 1);
	if (glm::dot(wo, n_s) <= 0) return s;
	s.wi = glm::reflect(-wo, n_s);
	s.cos_theta = glm::dot(s.wi, n_s);
	s.pdf = 1.0f;
	s.f = vec3(1.0f) / s.cos_theta;
	return s;
}
// glass.glsl:4-16
inline BsdfSample sample_glass(const lmb_material& mat, const vec3& wo, uint mode, bool forward_facing) {
	BsdfSample s;
	const vec3 n_s(0, 0, 1);
	vec3 f;
	float unused;
	refract_dir(n_s, wo, forward_facing, mat.ior, mode, s.wi, f, unused);
	s.cos_theta = glm::dot(n_s, s.wi);
	s.pdf = 1.f;
	s.f = f / std::fabs(s.cos_theta);
	return s;
}

// ------------------------------------------------------------------------------------------------ dielectric.glsl
// :6-8
inline float modify_thin_roughness(float ior, float roughness) { return glm::clamp((0.65f * ior - 0.35f) * roughness, 0.0f, 1.0f); }

// :10-108
inline BsdfSample sample_dielectric(const lmb_material& mat, const vec3& wo, uint mode, bool forward_facing, const vec2& xi) {
	BsdfSample s;
	if (wo.z <= 0.0f) return s;
	const float roughness = mat.thin == 1 ? modify_thin_roughness(mat.ior, mat.roughness) : mat.roughness;
	const float alpha = roughness * roughness;
	const bool has_reflection = has_prop(mat.bsdf_props, LMB_FLAG_REFLECTION);
	const bool has_transmission = has_prop(mat.bsdf_props, LMB_FLAG_TRANSMISSION);
	if ((mat.ior == 1.0f && mat.thin == 0) || effectively_delta(alpha)) {
		const float F = fresnel_dielectric(wo.z, mat.ior, forward_facing);
		if (!has_reflection && !has_transmission) return s;
		float pr = F;
		float pt = 1.0f - F;
		if (!has_reflection) pr = 0.0f;
		if (!has_transmission) pt = 0.0f;
		const bool is_reflection = ((pr + pt) * xi.x) < pr;
		if (is_reflection) {
			s.wi = vec3(-wo.x, -wo.y, wo.z);
			s.cos_theta = s.wi.z;
			s.f = vec3(F) / std::fabs(s.cos_theta);
			s.pdf = pr / (pr + pt);
		} else {
			vec3 f;
			float unused;
			refract_dir(vec3(0, 0, 1), wo, forward_facing, mat.ior, mode, s.wi, f, unused);
			s.cos_theta = s.wi.z;
			s.f = f * (1.0f - F) / std::fabs(s.cos_theta);
			s.pdf = pt / (pr + pt);
		}
		return s;
	}
	if (!has_reflection && !has_transmission) return s;
	float D;
	const vec3 h = sample_ggx_vndf_common(vec2(alpha), wo, xi);
	float pdf_w = vndf_pdf_iso(alpha, wo, h, D);
	const float F = fresnel_dielectric(glm::dot(wo, h), mat.ior, forward_facing);
	const float pr = has_reflection ? F : 0.0f;
	const float pt = has_transmission ? (1.0f - F) : 0.0f;
	const bool is_reflection = ((pr + pt) * xi.x) < pr;
	vec3 f(0);
	vec3 wi(0);
	if (is_reflection) {
		wi = glm::reflect(-wo, h);
		if (wo.z * wi.z < 0) {
			// GLSL returns vec3(0) here with wi already written and pdf_w holding the VNDF pdf; cos_theta stays 0
			s.wi = wi;
			s.pdf = pdf_w;
			return s;
		}
		pdf_w = pdf_w * (pr / (pr + pt)) / (4.0f * std::fabs(glm::dot(wo, h)));
		f = vec3(0.25f * D * F * g_ggx_corr_iso(alpha, wo, wi) / (wi.z * wo.z));
	} else {
		vec3 base_col;
		float possibly_modified_inv_eta;
		if (mat.thin == 1) {
			wi = glm::reflect(-wo, h);
			base_col = g_sqrt(v3(mat.albedo));
			possibly_modified_inv_eta = mat.ior;
			// `f` is read uninitialised by the reference on this branch (dielectric.glsl:78,101); defined as vec3(0)
		} else {
			refract_dir(h, wo, forward_facing, mat.ior, mode, wi, f, possibly_modified_inv_eta);
			base_col = v3(mat.albedo);
		}
		if (wo.z * wi.z > 0 || wi.z > 0) {
			s.wi = wi;
			s.pdf = pdf_w;
			return s;
		}
		float jacobian_denom = glm::dot(wi, h) + glm::dot(wo, h) * possibly_modified_inv_eta;
		jacobian_denom = jacobian_denom * jacobian_denom;
		const float jacobian = std::fabs(glm::dot(wi, h)) / jacobian_denom;
		pdf_w = pdf_w * (pt / (pr + pt)) * jacobian;
		f = base_col * f * (1.0f - F) * D * g_ggx_corr_iso(alpha, wo, wi) *
			std::fabs(glm::dot(wi, h) * glm::dot(wo, h) / (wi.z * wo.z * jacobian_denom));
	}
	s.wi = wi;
	s.pdf = pdf_w;
	s.cos_theta = wi.z;
	s.f = f;
	return s;
}

// :110-189. Q1: the transmission branch declares a second `vec3 f`, so the function returns the outer one, which
// GLSL leaves undefined; the oracle freezes it to vec3(0). pdf_w is still produced on that branch.
inline vec3 eval_dielectric(const lmb_material& mat, const vec3& wo, const vec3& wi, float& pdf_w, bool forward_facing, uint /*mode*/) {
	pdf_w = 0.0f;
	const float roughness = mat.thin == 1 ? modify_thin_roughness(mat.ior, mat.roughness) : mat.roughness;
	const float alpha = roughness * roughness;
	if (alpha == 0 || mat.ior == 1) return vec3(0);
	const bool has_reflection = has_prop(mat.bsdf_props, LMB_FLAG_REFLECTION);
	const bool has_transmission = has_prop(mat.bsdf_props, LMB_FLAG_TRANSMISSION);
	if (!has_reflection && !has_transmission) return vec3(0);
	float eta = 1.0f;
	const bool is_reflection = wi.z * wo.z > 0.0f;
	if (!is_reflection) eta = forward_facing ? mat.ior : 1.0f / mat.ior;
	vec3 h = glm::normalize(wo + wi * eta);
	h *= g_sign(h.z);
	if (wi.z == 0 || wo.z == 0 || glm::dot(h, h) == 0) return vec3(0);
	if (glm::dot(wi, h) * wi.z < 0 || glm::dot(wo, h) * wo.z < 0) return vec3(0);
	const float F = fresnel_dielectric(glm::dot(wo, h), mat.ior, forward_facing);
	const float pr = has_reflection ? F : 0.0f;
	const float pt = has_transmission ? (1.0f - F) : 0.0f;
	float D;
	pdf_w = vndf_pdf_iso(alpha, wo, h, D);
	const float G = g_ggx_corr_iso(alpha, wo, wi);
	vec3 f(0);
	if (is_reflection) {
		const float jacobian = 1.0f / (4.0f * std::fabs(glm::dot(wo, h)));
		const float prob_reflection = pr / (pr + pt);
		pdf_w = pdf_w * jacobian * prob_reflection;
		f = vec3(0.25f * D * G * F / std::fabs(wo.z * wi.z));
	} else {
		float jacobian_denom = glm::dot(wi, h) + glm::dot(wo, h) / eta;
		jacobian_denom = jacobian_denom * jacobian_denom;
		const float jacobian = std::fabs(glm::dot(wi, h)) / jacobian_denom;
		const float prob_refraction = pt / (pr + pt);
		pdf_w = pdf_w * jacobian * prob_refraction;
	}
	return f;
}

// ------------------------------------------------------------------------------------------------ conductor.glsl
// :6-43
inline BsdfSample sample_conductor(const lmb_material& mat, const vec3& wo, const vec2& xi) {
	BsdfSample s;
	if (wo.z <= 0.0f) return s;
	const float alpha = mat.roughness * mat.roughness;
	if (!has_prop(mat.bsdf_props, LMB_FLAG_REFLECTION)) return s;
	if (effectively_delta(alpha)) {
		s.wi = vec3(-wo.x, -wo.y, wo.z);
		s.pdf = 1.0f;
		s.cos_theta = s.wi.z;
		const vec3 F = fresnel_conductor(s.cos_theta, v3(mat.albedo), v3(mat.k));
		s.f = F / std::fabs(s.cos_theta);
		return s;
	}
	float D;
	const vec3 h = sample_ggx_vndf_common(vec2(alpha), wo, xi);
	s.pdf = vndf_pdf_iso(alpha, wo, h, D);
	const vec3 F = fresnel_conductor(glm::dot(wo, h), v3(mat.albedo), v3(mat.k));
	s.wi = glm::reflect(-wo, h);
	if (wo.z * s.wi.z < 0) return s;  // f = 0, pdf keeps the VNDF value, cos_theta = 0
	s.pdf /= (4.0f * glm::dot(wo, h));
	s.cos_theta = s.wi.z;
	s.f = 0.25f * D * F * g_ggx_corr_iso(alpha, wo, s.wi) / (s.wi.z * wo.z);
	return s;
}
// :45-72
inline vec3 eval_conductor(const lmb_material& mat, const vec3& wo, const vec3& wi, float& pdf_w) {
	pdf_w = 0;
	const float alpha = mat.roughness * mat.roughness;
	if (effectively_delta(alpha)) return vec3(0);
	if (wo.z * wi.z < 0) return vec3(0);
	if (wo.z == 0 || wi.z == 0) return vec3(0);
	vec3 h = glm::normalize(wo + wi);
	h *= g_sign(h.z);
	const float jacobian = 1.0f / (4.0f * glm::dot(wo, h));
	float D;
	pdf_w = vndf_pdf_iso(alpha, wo, h, D) * jacobian;
	const vec3 F = fresnel_conductor(glm::dot(wo, h), v3(mat.albedo), v3(mat.k));
	return 0.25f * D * F * g_ggx_corr_iso(alpha, wo, wi) / (wi.z * wo.z);
}

// ------------------------------------------------------------------------------------------------ principled.glsl
// :20-24
inline vec2 calc_anisotropy(float roughness, float anisotropic) {
	const float aspect = std::sqrt(1.0f - 0.9f * anisotropic);
	const float roughness_sqr = roughness * roughness;
	return vec2(glm::max(0.001f, roughness_sqr / aspect), glm::max(0.001f, roughness_sqr * aspect));
}
struct LobeProbs {
	float spec, diff, clearcoat, spec_trans;
};
// :26-48
inline LobeProbs sampling_probs(const lmb_material& mat, float F_dielectric, bool forward_facing) {
	LobeProbs p;
	const float brdf_weight = (1.0f - mat.spec_trans) * (1.0f - mat.metallic);
	const float bsdf_weight = (1.0f - mat.metallic) * mat.spec_trans;
	p.spec = forward_facing ? (1.0f - bsdf_weight * (1.0f - F_dielectric)) : F_dielectric;
	p.spec_trans = forward_facing ? (bsdf_weight * (1.0f - F_dielectric)) : (1.0f - F_dielectric);
	p.diff = forward_facing ? brdf_weight : 0.0f;
	p.clearcoat = forward_facing ? 0.25f * glm::clamp(mat.clearcoat, 0.0f, 1.0f) : 0.0f;
	const float norm = 1.0f / (p.spec + p.spec_trans + p.diff + p.clearcoat);
	p.spec *= norm;
	p.diff *= norm;
	p.clearcoat *= norm;
	p.spec_trans *= norm;
	return p;
}
// :50-74
inline vec3 disney_diffuse_factor(const lmb_material& mat, const vec3& wo, const vec3& wi) {
	const vec3 h = glm::normalize(wi + wo);
	float f_wi, f_wo;
	disney_fresnel(wi, wo, mat.roughness, f_wi, f_wo);
	const float roughness_sqr = mat.roughness * mat.roughness;
	float ss = 0;
	const float rr = 2.0f * roughness_sqr * wi.z * wi.z;
	const float f_retro = rr * (f_wi + f_wo * f_wi * f_wo * (rr - 1.0f));
	const float f_diff = (1.0f - 0.5f * f_wi) * (1.0f - 0.5f * f_wo);
	if (mat.flatness > 0.0f) {
		const float fss90 = 0.5f * rr;
		const float f_ss = glm::mix(1.0f, fss90, f_wi) * glm::mix(1.0f, fss90, f_wo);
		ss = 1.25f * (f_ss * (1.0f / (wi.z + wo.z) - 0.5f) + 0.5f);
	}
	(void)h;
	const float ss_approx_and_diff = glm::mix(f_diff + f_retro, ss, mat.flatness);
	return v3(mat.albedo) * ss_approx_and_diff * INV_PI;
}
// :76-83
inline float clearcoat_factor(const lmb_material& mat, const vec3& wo, const vec3& wi, const vec3& h, float& D) {
	const float alpha_2 = 0.25f * 0.25f;
	D = d_ggx_iso(glm::mix(0.1f, 0.001f, mat.clearcoat_gloss), h.z);
	const float F = fresnel_schlick(0.04f, 1.0f, glm::dot(wi, h));
	const float G = g1_ggx_iso(alpha_2, wo.z) * g1_ggx_iso(alpha_2, wi.z);
	return 0.25f * mat.clearcoat * D * F * G;
}
// :86-94
inline BsdfSample sample_disney_diffuse(const lmb_material& mat, const vec3& wo, const vec2& xi) {
	BsdfSample s;
	s.wi = sample_hemisphere(xi);
	s.cos_theta = s.wi.z;
	s.pdf = s.cos_theta * INV_PI;
	if (glm::min(s.wi.z, wo.z) <= 0.0f) return s;
	s.f = disney_diffuse_factor(mat, wo, s.wi);
	return s;
}
// :96-134 (ANISOTROPIC == 1). wi / pdf_w / cos_theta are inout in the reference: untouched values stay 0.
inline BsdfSample sample_principled_brdf(const lmb_material& mat, const vec3& wo, const vec2& xi, float eta) {
	BsdfSample s;
	if (!has_prop(mat.bsdf_props, LMB_FLAG_REFLECTION)) return s;
	float D;
	const vec2 alpha = calc_anisotropy(mat.roughness, mat.anisotropy);
	if (effectively_delta(alpha)) {
		s.wi = vec3(-wo.x, -wo.y, wo.z);
		s.pdf = 1.0f;
		s.cos_theta = s.wi.z;
		const vec3 F = disney_fresnel(mat, wo, vec3(0, 0, 1), s.wi, eta);
		s.f = F / std::fabs(s.cos_theta);
		return s;
	}
	const vec3 h = sample_ggx_vndf_common(alpha, wo, xi);
	s.pdf = vndf_pdf_aniso(alpha, wo, h, D);
	s.wi = glm::reflect(-wo, h);
	if (wo.z * s.wi.z < 0) return s;
	const vec3 F = disney_fresnel(mat, wo, h, s.wi, eta);
	s.pdf /= (4.0f * glm::dot(wo, h));
	s.cos_theta = s.wi.z;
	s.f = 0.25f * D * F * g_ggx_corr_aniso(alpha, wo, s.wi) / (s.wi.z * wo.z);
	return s;
}
// :137-164. Q2: sin_t comes from the inout cos_theta, which is 0 on entry -> sin_t = 1.
inline BsdfSample sample_clearcoat(const lmb_material& mat, const vec3& wo, const vec2& xi) {
	BsdfSample s;
	const float alpha_2 = 0.25f * 0.25f;
	const float cos_t = std::sqrt(glm::max(0.0f, (1.0f - g_pow(alpha_2, 1.0f - xi.x)) / (1.0f - alpha_2)));
	const float sin_t = std::sqrt(glm::max(0.0f, 1.0f - s.cos_theta * s.cos_theta));
	const float phi = TWO_PI * xi.y;
	float sp, cp;
	lmb_sincosf(phi, &sp, &cp);
	vec3 h(sin_t * cp, sin_t * sp, cos_t);
	if (glm::dot(h, wo) < 0.0f) h *= -1.0f;
	s.wi = glm::reflect(-wo, h);
	if (glm::dot(s.wi, wo) < 0.0f) return s;
	float D;
	const float f_clearcoat = clearcoat_factor(mat, wo, s.wi, h, D);
	s.pdf = D / (4.0f * glm::dot(wo, h));
	s.cos_theta = s.wi.z;
	s.f = vec3(f_clearcoat);
	return s;
}
// :166-177
inline vec3 eval_clearcoat(const lmb_material& mat, const vec3& wo, const vec3& wi, float& pdf_w) {
	const vec3 h = glm::normalize(wo + wi);
	float D;
	const float f_clearcoat = clearcoat_factor(mat, wo, wi, h, D);
	pdf_w = D / (4.0f * glm::dot(wo, h));
	return vec3(f_clearcoat);
}
// :184-225
inline vec3 eval_principled_brdf(const lmb_material& mat, const vec3& wo, const vec3& wi, float& pdf_w, bool forward_facing) {
	pdf_w = 0;
	const vec2 alpha = calc_anisotropy(mat.roughness, mat.anisotropy);
	if (effectively_delta(alpha)) return vec3(0);
	if (wo.z * wi.z < 0) return vec3(0);
	if (wo.z == 0 || wi.z == 0) return vec3(0);
	const vec3 h = glm::normalize(wo + wi);
	const float jacobian = 1.0f / (4.0f * glm::dot(wo, h));
	float D;
	pdf_w = vndf_pdf_aniso(alpha, wo, h, D) * jacobian;
	const float eta = forward_facing ? mat.ior : 1.0f / mat.ior;
	const vec3 F = disney_fresnel(mat, wo, h, wi, eta);
	return 0.25f * D * F * g_ggx_corr_aniso(alpha, wo, wi) / (wi.z * wo.z);
}
// :258-287
inline BsdfSample sample_principled(const lmb_material& mat, const vec3& wo, uint mode, bool forward_facing, const vec3& xi) {
	BsdfSample s;
	if (wo.z <= 0.0f) return s;
	const float F = fresnel_dielectric(wo.z, mat.ior, forward_facing);
	const LobeProbs p = sampling_probs(mat, F, forward_facing);
	const float eta = forward_facing ? mat.ior : 1.0f / mat.ior;
	float p_lobe = 0.0f;
	const vec2 xy(xi.x, xi.y);
	if (xi.z < p.spec) {
		s = sample_principled_brdf(mat, wo, xy, eta);
		p_lobe = p.spec;
	} else if (xi.z > p.spec && xi.z <= (p.spec + p.clearcoat)) {
		s = sample_clearcoat(mat, wo, xy);
		p_lobe = p.clearcoat;
	} else if (xi.z > (p.spec + p.clearcoat) && xi.z <= (p.spec + p.clearcoat + p.diff)) {
		s = sample_disney_diffuse(mat, wo, xy);
		p_lobe = p.diff;
	} else if (p.spec_trans >= 0.0f && xi.z <= (p.spec + p.clearcoat + p.diff + p.spec_trans)) {
		s = sample_dielectric(mat, wo, mode, forward_facing, xy);
		p_lobe = p.spec_trans;
	}
	s.pdf *= p_lobe;
	return s;
}
// :289-337 (eval_reverse_pdf == false on the Path integrator's call sites)
inline vec3 eval_principled(const lmb_material& mat, const vec3& wo, const vec3& wi, float& pdf_w, bool forward_facing, uint mode) {
	pdf_w = 0.0f;
	const float F = fresnel_dielectric(wo.z, mat.ior, forward_facing);
	const LobeProbs p = sampling_probs(mat, F, forward_facing);
	vec3 f(0);
	float pdf = 0;
	const float brdf_weight = (1.0f - mat.spec_trans) * (1.0f - mat.metallic);
	const float bsdf_weight = (1.0f - mat.metallic) * mat.spec_trans;
	if (p.spec > 0) {
		f += eval_principled_brdf(mat, wo, wi, pdf, forward_facing);
		pdf *= p.spec;
		pdf_w += pdf;
	}
	const bool upper_hemisphere = glm::min(wi.z, wo.z) > 0;
	if (upper_hemisphere) {
		if (p.diff > 0) {
			pdf_w += p.diff * lambertian_pdf(wo, wi);
			f += brdf_weight * disney_diffuse_factor(mat, wo, wi);
		}
		if (p.clearcoat > 0) {
			f += eval_clearcoat(mat, wo, wi, pdf);
			pdf *= p.clearcoat;
			pdf_w += pdf;
		}
	}
	if (p.spec_trans > 0) {
		f += bsdf_weight * eval_dielectric(mat, wo, wi, pdf, forward_facing, mode);
		pdf *= p.spec_trans;
		pdf_w += pdf;
	}
	return f;
}

// ------------------------------------------------------------------------------------------------ bsdf_commons.glsl
// :68-121. `rands` are the three numbers sample_bsdf(seed) always draws (:117-121).
inline BsdfSample sample_bsdf(const vec3& n_s, const vec3& wo_world, const lmb_material& mat, uint mode, bool forward_facing,
							  const vec3& rands) {
	vec3 T, B;
	branchless_onb(n_s, T, B);
	const vec3 wo = to_local(wo_world, T, B, n_s);
	BsdfSample s;
	const vec2 xy(rands.x, rands.y);
	switch (mat.bsdf_type) {
		case LMB_BSDF_DIFFUSE:
			s = sample_lambertian(mat, wo, xy);
			break;
		case LMB_BSDF_MIRROR:
			s = sample_mirror(wo);
			break;
		case LMB_BSDF_GLASS:
			s = sample_glass(mat, wo, mode, forward_facing);
			break;
		case LMB_BSDF_DIELECTRIC:
			s = sample_dielectric(mat, wo, mode, forward_facing, xy);
			break;
		case LMB_BSDF_CONDUCTOR:
			s = sample_conductor(mat, wo, xy);
			break;
		case LMB_BSDF_PRINCIPLED:
			s = sample_principled(mat, wo, mode, forward_facing, rands);
			break;
		default:
			break;
	}
	s.wi = to_world(s.wi, T, B, n_s);
	return s;
}
// :123-183 with eval_reverse == false (the 7-argument overload used by pt_commons.glsl:18)
inline vec3 eval_bsdf(const vec3& n_s, const vec3& wo_world, const lmb_material& mat, uint mode, bool forward_facing,
					  const vec3& wi_world, float& pdf_w) {
	pdf_w = 0;
	vec3 T, B;
	branchless_onb(n_s, T, B);
	const vec3 wo = to_local(wo_world, T, B, n_s);
	const vec3 wi = to_local(wi_world, T, B, n_s);
	switch (mat.bsdf_type) {
		case LMB_BSDF_DIFFUSE:
			return eval_lambertian(mat, wo, wi, pdf_w);
		case LMB_BSDF_MIRROR:
		case LMB_BSDF_GLASS:
			return vec3(0);
		case LMB_BSDF_DIELECTRIC:
			return eval_dielectric(mat, wo, wi, pdf_w, forward_facing, mode);
		case LMB_BSDF_CONDUCTOR:
			return eval_conductor(mat, wo, wi, pdf_w);
		case LMB_BSDF_PRINCIPLED:
			return eval_principled(mat, wo, wi, pdf_w, forward_facing, mode);
		default:
			break;
	}
	return vec3(0);
}

// ------------------------------------------------------------------------------------------------ bsdf_pdf (BDPT side of the BSDFs)
// The stand-alone pdf functions of the reference (used by bdpt_commons.glsl, not by the Path integrator). Restated here as an
// INDEPENDENT check of the sampling code above: where the GLSL means sample_*'s pdf, eval_*'s pdf_w and *_pdf to be the same
// density, tests/test_oracle.py compares the three (SURVEY.md 8f-3 groundwork).
// diffuse.glsl:85-90
inline float lambertian_diffuse_pdf(const vec3& wo, const vec3& wi) {
	if (glm::min(wi.z, wo.z) <= 0.0f) return 0.0f;
	return wi.z * INV_PI;
}
// dielectric.glsl:191-240
inline float dielectric_pdf(const lmb_material& mat, const vec3& wo, const vec3& wi, bool forward_facing) {
	const float roughness = mat.thin == 1 ? modify_thin_roughness(mat.ior, mat.roughness) : mat.roughness;
	const float alpha = roughness * roughness;
	if (alpha == 0 || mat.ior == 1) return 0.0f;
	const bool has_reflection = has_prop(mat.bsdf_props, LMB_FLAG_REFLECTION);
	const bool has_transmission = has_prop(mat.bsdf_props, LMB_FLAG_TRANSMISSION);
	if (!has_reflection && !has_transmission) return 0.0f;
	const bool is_reflection = wi.z * wo.z > 0;
	float eta = 1.0f;
	if (!is_reflection) eta = forward_facing ? mat.ior : 1.0f / mat.ior;
	vec3 h = glm::normalize(wo + wi * eta);
	h *= float(glm::sign(h.z));
	if (wi.z == 0 || wo.z == 0 || glm::dot(h, h) == 0) return 0.0f;
	if (glm::dot(wi, h) * wi.z < 0 || glm::dot(wo, h) * wo.z < 0) return 0.0f;
	const float F = fresnel_dielectric(glm::dot(wo, h), mat.ior, forward_facing);
	const float pr = has_reflection ? F : 0.0f;
	const float pt = has_transmission ? (1.0f - F) : 0.0f;
	float D;
	float pdf_w = vndf_pdf_iso(alpha, wo, h, D);
	if (is_reflection) {
		const float jacobian = 1.0f / (4.0f * std::fabs(glm::dot(wo, h)));
		const float prob_reflection = pr / (pr + pt);
		pdf_w = pdf_w * jacobian * prob_reflection;
	} else {
		float jacobian_denom = glm::dot(wi, h) + glm::dot(wo, h) / eta;
		jacobian_denom = jacobian_denom * jacobian_denom;
		const float jacobian = std::fabs(glm::dot(wi, h)) / jacobian_denom;
		const float prob_refraction = pt / (pr + pt);
		pdf_w = pdf_w * jacobian * prob_refraction;
	}
	return pdf_w;
}
// conductor.glsl:74-90
inline float conductor_pdf(const lmb_material& mat, const vec3& wo, const vec3& wi) {
	const float alpha = mat.roughness * mat.roughness;
	if (effectively_delta(alpha)) return 0.0f;
	if (wo.z * wi.z < 0) return 0.0f;
	if (wo.z == 0 || wi.z == 0) return 0.0f;
	vec3 h = glm::normalize(wo + wi);
	h *= float(glm::sign(h.z));
	float D;
	return vndf_pdf_iso(alpha, wo, h, D) / (4.0f * glm::dot(wo, h));
}
// principled.glsl:177-181
inline float clearcoat_pdf(const lmb_material& mat, const vec3& wo, const vec3& wi) {
	const vec3 h = glm::normalize(wo + wi);
	const float D = d_ggx_iso(glm::mix(0.1f, 0.001f, mat.clearcoat_gloss), h.z);
	return D / (4.0f * glm::dot(wo, h));
}
// principled.glsl:226-251 (ANISOTROPIC == 1)
inline float principled_brdf_pdf(const lmb_material& mat, const vec3& wo, const vec3& wi) {
	const vec2 alpha = calc_anisotropy(mat.roughness, mat.anisotropy);
	if (effectively_delta(alpha)) return 0.0f;
	if (wo.z * wi.z < 0) return 0.0f;
	if (wo.z == 0 || wi.z == 0) return 0.0f;
	vec3 h = glm::normalize(wo + wi);
	h *= float(glm::sign(h.z));
	float D;
	return vndf_pdf_aniso(alpha, wo, h, D) / (4.0f * glm::dot(wo, h));
}
// principled.glsl:338-360
inline float principled_pdf(const lmb_material& mat, const vec3& wo, const vec3& wi, bool forward_facing) {
	float pdf = 0.0f;
	const float F = fresnel_dielectric(wo.z, mat.ior, forward_facing);
	const LobeProbs p = sampling_probs(mat, F, forward_facing);
	if (p.spec > 0) pdf += p.spec * principled_brdf_pdf(mat, wo, wi);
	const bool upper = glm::min(wi.z, wo.z) > 0;
	if (upper) {
		if (p.diff > 0) pdf += p.diff * lambertian_diffuse_pdf(wo, wi);
		if (p.clearcoat > 0) pdf += p.clearcoat * clearcoat_pdf(mat, wo, wi);
	}
	if (p.spec_trans > 0) pdf += p.spec_trans * dielectric_pdf(mat, wo, wi, forward_facing);
	return pdf;
}
// bsdf_commons.glsl:26-66
inline float bsdf_pdf(const lmb_material& mat, const vec3& n_s, const vec3& wo_world, const vec3& wi_world, bool forward_facing) {
	vec3 T, B;
	branchless_onb(n_s, T, B);
	const vec3 wo = to_local(wo_world, T, B, n_s);
	const vec3 wi = to_local(wi_world, T, B, n_s);
	switch (mat.bsdf_type) {
		case LMB_BSDF_DIFFUSE:
			return lambertian_diffuse_pdf(wo, wi);
		case LMB_BSDF_MIRROR:
		case LMB_BSDF_GLASS:
			return 0.0f;
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

// ------------------------------------------------------------------------------------------------ atmosphere.glsl
namespace atmo {
constexpr float PLANET_RADIUS = 6371000.0f;
constexpr float ATMOSPHERE_HEIGHT = 100000.0f;
constexpr float RAYLEIGH_HEIGHT = ATMOSPHERE_HEIGHT * 0.08f;
constexpr float MIE_HEIGHT = ATMOSPHERE_HEIGHT * 0.012f;
inline vec3 planet_center() { return vec3(0, -PLANET_RADIUS, 0); }
inline vec3 c_rayleigh() { return vec3(5.802f, 13.558f, 33.100f) * 1e-6f; }
inline vec3 c_mie() { return vec3(3.996f, 3.996f, 3.996f) * 1e-6f; }
inline vec3 c_ozone() { return vec3(0.650f, 1.881f, 0.085f) * 1e-6f; }

// :47-64
inline vec2 sphere_intersection(vec3 ray_start, const vec3& ray_dir, const vec3& center, float radius) {
	ray_start -= center;
	const float a = glm::dot(ray_dir, ray_dir);
	const float b = 2.0f * glm::dot(ray_start, ray_dir);
	const float c = glm::dot(ray_start, ray_start) - (radius * radius);
	float d = b * b - 4 * a * c;
	if (d < 0) return vec2(-1);
	d = std::sqrt(d);
	return vec2(-b - d, -b + d) / (2 * a);
}
// :65-72
inline vec2 planet_intersection(const vec3& s, const vec3& d) {
	return sphere_intersection(s, d, planet_center() * vec3(0, 1.00001f, 0), PLANET_RADIUS);
}
inline vec2 atmosphere_intersection(const vec3& s, const vec3& d) {
	return sphere_intersection(s, d, planet_center(), PLANET_RADIUS + ATMOSPHERE_HEIGHT);
}
// :76-87
inline float phase_rayleigh(float costh) { return 3 * (1 + costh * costh) / (16 * PI); }
inline float phase_mie(float costh, float g) {
	g = glm::min(g, 0.9381f);
	const float k = 1.55f * g - 0.55f * g * g * g;
	const float kcosth = k * costh;
	return (1 - k * k) / ((4 * PI) * (1 - kcosth) * (1 - kcosth));
}
// :91-112
inline float height(const vec3& p) { return glm::distance(p, planet_center()) - PLANET_RADIUS; }
inline vec3 density(float h) {
	return vec3(g_exp(-glm::max(0.0f, h / RAYLEIGH_HEIGHT)), g_exp(-glm::max(0.0f, h / MIE_HEIGHT)),
				glm::max(0.0f, 1 - std::fabs(h - 25000.0f) / 15000.0f));
}
// :120-137
inline vec3 optical_depth(const vec3& ray_start, const vec3& ray_dir) {
	const vec2 isect = atmosphere_intersection(ray_start, ray_dir);
	const float ray_length = isect.y;
	const int sample_count = 8;
	const float step = ray_length / sample_count;
	vec3 od(0.0f);
	for (int i = 0; i < sample_count; i++) {
		const vec3 local_pos = ray_start + ray_dir * (i + 0.5f) * step;
		const vec3 local_density = density(height(local_pos));
		od += local_density * step;
	}
	return od;
}
// :140-144
inline vec3 absorb(const vec3& od) {
	return g_exp(-(od.x * c_rayleigh() + od.y * c_mie() * 1.1f + od.z * c_ozone()) * 1.0f);
}
// :148-204
inline vec3 integrate_scattering(vec3 ray_start, const vec3& ray_dir, float ray_length, const vec3& light_dir, const vec3& light_color) {
	const float ray_height = height(ray_start);
	const float exponent = 1 + glm::clamp(1 - ray_height / ATMOSPHERE_HEIGHT, 0.0f, 1.0f) * 8;
	const vec2 isect = atmosphere_intersection(ray_start, ray_dir);
	ray_length = glm::min(ray_length, isect.y);
	if (isect.x > 0) {
		ray_start += ray_dir * isect.x;
		ray_length -= isect.x;
	}
	const float costh = glm::dot(ray_dir, light_dir);
	const float phase_r = phase_rayleigh(costh);
	const float phase_m = phase_mie(costh, 0.85f);
	const int sample_count = 64;
	vec3 od(0), rayleigh(0), mie(0);
	float prev_ray_time = 0;
	for (int i = 0; i < sample_count; i++) {
		const float ray_time = g_pow(float(i) / sample_count, exponent) * ray_length;
		const float step = (ray_time - prev_ray_time);
		const vec3 local_pos = ray_start + ray_dir * ray_time;
		const vec3 local_density = density(height(local_pos));
		od += local_density * step;
		const vec3 view_tr = absorb(od);
		const vec3 od_light = optical_depth(local_pos, light_dir);
		const vec3 light_tr = absorb(od_light);
		rayleigh += view_tr * light_tr * phase_r * local_density.x * step;
		mie += view_tr * light_tr * phase_m * local_density.y * step;
		prev_ray_time = ray_time;
	}
	return (rayleigh * c_rayleigh() + mie * c_mie()) * light_color * 20.0f;
}
}  // namespace atmo

}  // namespace orc
