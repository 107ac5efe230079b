// shading.cuh -- device-side BSDF / light / sky / RNG functions of the B200 path integrator.
//
// What each block computes and which reference lines define it (all under src/shaders/ of yuphin/Lumen):
//   RNG                      utils.glsl:121-154 (pcg4d, state = (x, y, frame, counter))
//   ray offsets              utils.glsl:73-97
//   Fresnel / refraction     bsdf/sampling_commons.glsl:16-174
//   GGX / VNDF               bsdf/microfacet_commons.glsl:5-137
//   diffuse, mirror, glass   bsdf/diffuse.glsl:13-21,58-65  bsdf/mirror.glsl:3-14  bsdf/glass.glsl:4-16
//   dielectric               bsdf/dielectric.glsl:6-189
//   conductor                bsdf/conductor.glsl:6-72
//   principled (Disney)      bsdf/principled.glsl:20-337
//   dispatch                 bsdf_commons.glsl:68-183
//   sky                      atmosphere/atmosphere.glsl:47-204
// Undefined-in-GLSL spots are frozen as documented in DESIGN.md ("Frozen quirks"): the uninitialised `f` of
// eval_dielectric's transmission branch and of the thin rough-transmission sample is vec3(0).
#pragma once
#include "lmb_detmath.h"
#include "lmb_types.h"
#include "vec.cuh"

namespace lmb {

#define LMB_PI 3.14159265359f
#define LMB_TWO_PI 6.28318530718f
#define LMB_INV_PI (1.0f / LMB_PI)
#define LMB_EPS 0.001f

// Out-of-line copies of the deterministic transcendentals and of the hash: the shade kernels are instruction-fetch bound
// (ncu: 83 % of stall samples were "no instruction" with everything inlined, 341 KB of SASS), so anything called from many
// sites is a real function on the device.
#define LMB_DN static __device__ __noinline__
LMB_DN float d_powf(float x, float y) { return lmb_powf(x, y); }
LMB_DN float d_expf(float x) { return lmb_expf(x); }
// the same function, bit for bit (lmb_detmath.h lmb_expf_fast), inlined: 15 FP32 / integer instructions in the window every exponent
// of the sky march falls into but the far tail, no call, nothing on the conversion pipe
LMB_D float d_expf_fast(float x) { return (x > -86.0f && x <= 88.0f) ? lmb_expf_window(x) : d_expf(x); }
LMB_DN float2 d_sincosf(float x) {
	float s, c;
	lmb_sincosf(x, &s, &c);
	return make_float2(s, c);
}

// ------------------------------------------------------------------------------------------------ RNG
struct Rng {
	uint32_t x, y, z, w;
};
LMB_DN uint32_t pcg4d_x(uint32_t vx, uint32_t vy, uint32_t vz, uint32_t vw) {
	vx = vx * 1664525u + 1013904223u;
	vy = vy * 1664525u + 1013904223u;
	vz = vz * 1664525u + 1013904223u;
	vw = vw * 1664525u + 1013904223u;
	vx += vy * vw;
	vy += vz * vx;
	vz += vx * vy;
	vw += vy * vz;
	vx ^= vx >> 16u;
	vy ^= vy >> 16u;
	vz ^= vz >> 16u;
	vw ^= vw >> 16u;
	vx += vy * vw;
	return vx;
}
LMB_D void pcg4d_full(uint32_t& vx, uint32_t& vy, uint32_t& vz, uint32_t& vw) {
	vx = vx * 1664525u + 1013904223u;
	vy = vy * 1664525u + 1013904223u;
	vz = vz * 1664525u + 1013904223u;
	vw = vw * 1664525u + 1013904223u;
	vx += vy * vw;
	vy += vz * vx;
	vz += vx * vy;
	vw += vy * vz;
	vx ^= vx >> 16u;
	vy ^= vy >> 16u;
	vz ^= vz >> 16u;
	vw ^= vw >> 16u;
	vx += vy * vw;
	vy += vz * vx;
	vz += vx * vy;
	vw += vy * vz;
}
LMB_D float uint_to_float(uint32_t x) { return __uint_as_float(0x3f800000u | (x >> 9)) - 1.0f; }
LMB_D float rand1(Rng& s) {
	s.w++;
	return uint_to_float(pcg4d_x(s.x, s.y, s.z, s.w));
}
LMB_D V3 rand3(Rng& s) {
	const float a = rand1(s);
	const float b = rand1(s);
	const float c = rand1(s);
	return V3{a, b, c};
}
LMB_D V4 rand4(Rng& s) {
	const float a = rand1(s);
	const float b = rand1(s);
	const float c = rand1(s);
	const float d = rand1(s);
	return V4{a, b, c, d};
}

// ------------------------------------------------------------------------------------------------ utils
LMB_D float bump_ulp(float v, int o) { return __int_as_float(__float_as_int(v) + ((v < 0) ? -o : o)); }
LMB_D V3 offset_ray(const V3& p, const V3& n) {
	const float origin = 1.0f / 32.0f;
	const float float_scale = 1.0f / 65536.0f;
	const float int_scale = 256.0f;
	const int ox = (int)(int_scale * n.x), oy = (int)(int_scale * n.y), oz = (int)(int_scale * n.z);
	const V3 p_i = v3(bump_ulp(p.x, ox), bump_ulp(p.y, oy), bump_ulp(p.z, oz));
	return v3(fabsf(p.x) < origin ? p.x + float_scale * n.x : p_i.x, fabsf(p.y) < origin ? p.y + float_scale * n.y : p_i.y,
			  fabsf(p.z) < origin ? p.z + float_scale * n.z : p_i.z);
}
LMB_D V3 offset_ray2(const V3& p, const V3& n) {
	const float float_scale = 2.0f / 65536.0f;
	return p + float_scale * n;
}
LMB_D float luminance(const V3& rgb) { return dot(rgb, v3(0.2126f, 0.7152f, 0.0722f)); }
LMB_D void branchless_onb(const V3& n, V3& b1, V3& b2) {
	const float sign = n.z >= 0.0f ? 1.0f : -1.0f;
	const float a = -1.0f / (sign + n.z);
	const float b = n.x * n.y * a;
	b1 = v3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
	b2 = v3(b, sign + n.y * n.y * a, -n.y);
}
LMB_D V3 to_world(const V3& v, const V3& T, const V3& B, const V3& N) { return v.x * T + v.y * B + v.z * N; }
LMB_D V3 to_local(const V3& v, const V3& T, const V3& B, const V3& N) { return v3(dot(v, T), dot(v, B), dot(v, N)); }
LMB_D V2 concentric_sample_disk(const V2& rands) {
	const V2 offset = 2.0f * rands - 1.0f;
	if (offset.x == 0 && offset.y == 0) return v2(0, 0);
	float theta, r;
	if (fabsf(offset.x) > fabsf(offset.y)) {
		r = offset.x;
		theta = 0.25f * LMB_PI * offset.y / offset.x;
	} else {
		r = offset.y;
		theta = LMB_PI * (0.5f - 0.25f * offset.x / offset.y);
	}
	const float2 sc = d_sincosf(theta);
	return r * v2(sc.y, sc.x);
}
LMB_D bool has_prop(uint32_t props, uint32_t flag) { return (props & flag) != 0; }

// ------------------------------------------------------------------------------------------------ sampling_commons
LMB_D bool effectively_delta(float alpha) { return alpha <= 0.0064f; }
LMB_D bool effectively_delta(const V2& alpha) { return gmin(alpha.x, alpha.y) <= 0.0064f; }

LMB_D void refract_dir(const V3& n_s, const V3& wo, bool forward_facing, float eta, uint32_t mode, V3& wi, V3& f, float& inv_eta) {
	const float cos_i = dot(n_s, wo);
	inv_eta = forward_facing ? 1.0f / eta : eta;
	const float sin2_t = inv_eta * inv_eta * (1.0f - cos_i * cos_i);
	if (sin2_t >= 1.0f) {
		wi = reflect(-wo, n_s);
	} else {
		const float cos_t = sqrtf(1 - sin2_t);
		wi = -inv_eta * wo + (inv_eta * cos_i - cos_t) * n_s;
	}
	f = mode == 1 ? v3(inv_eta * inv_eta) : v3(1.0f);
}

LMB_DN float fresnel_dielectric(float cos_i, float eta, bool forward_facing) {
	cos_i = gclamp(cos_i, -1.0f, 1.0f);
	if (!forward_facing) eta = 1.0f / eta;
	if (cos_i < 0) {
		eta = 1.0f / eta;
		cos_i = -cos_i;
	}
	const float sin2_i = 1 - cos_i * cos_i;
	const float sin2_t = sin2_i / (eta * eta);
	if (sin2_t >= 1) return 1.f;
	const float cos_t = sqrtf(1 - sin2_t);
	const float r_parallel = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
	const float r_perp = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
	return 0.5f * (r_parallel * r_parallel + r_perp * r_perp);
}

LMB_DN float fresnel_conductor(float cos_i, float eta, float k) {
	const float cos_sqr = cos_i * cos_i;
	const float sin_sqr = gmax(1.0f - cos_sqr, 0.0f);
	const float sin_4 = sin_sqr * sin_sqr;
	const float inner_term = eta * eta - k * k - sin_sqr;
	const float a_sq_p_b_sq = sqrtf(gmax(inner_term * inner_term + 4.0f * eta * eta * k * k, 0.0f));
	const float a = sqrtf(gmax((a_sq_p_b_sq + inner_term) * 0.5f, 0.0f));
	const float rs = ((a_sq_p_b_sq + cos_sqr) - (2.0f * a * cos_i)) / ((a_sq_p_b_sq + cos_sqr) + (2.0f * a * cos_i));
	const float rp = ((cos_sqr * a_sq_p_b_sq + sin_4) - (2.0f * a * cos_i * sin_sqr)) / ((cos_sqr * a_sq_p_b_sq + sin_4) + (2.0f * a * cos_i * sin_sqr));
	return 0.5f * (rs + rs * rp);
}
LMB_D V3 fresnel_conductor(float cos_i, const V3& eta, const V3& k) {
	return v3(fresnel_conductor(cos_i, eta.x, k.x), fresnel_conductor(cos_i, eta.y, k.y), fresnel_conductor(cos_i, eta.z, k.z));
}
LMB_D float fresnel_schlick(float f0, float f90, float ns) { return f0 + (f90 - f0) * d_powf(gmax(1.0f - ns, 0.0f), 5.0f); }
LMB_D V3 fresnel_schlick(const V3& f0, const V3& f90, float ns) { return f0 + (f90 - f0) * d_powf(gmax(1.0f - ns, 0.0f), 5.0f); }
LMB_D float eta_to_schlick_R0(float eta) {
	const float val = (eta - 1.0f) / (eta + 1.0f);
	return val * val;
}
LMB_D float disney_fresnel(const V3& wi, const V3& wo, float roughness, float& f_wi, float& f_wo) {
	const V3 h = normalize(wi + wo);
	const float wo_dot_h = dot(wo, h);
	const float fd90 = 0.5f + 2.0f * wo_dot_h * wo_dot_h * roughness;
	const float fd0 = 1.f;
	f_wi = fresnel_schlick(fd0, fd90, wi.z);
	f_wo = fresnel_schlick(fd0, fd90, wo.z);
	return f_wi * f_wo;
}
LMB_D V3 calc_tint(const V3& albedo) {
	const float lum = luminance(albedo);
	return lum > 0 ? albedo / lum : v3(1.0f);
}
LMB_D V3 disney_fresnel(const lmb_material& mat, const V3& wo, const V3& h, float eta) {
	const float wo_dot_h = dot(wo, h);
	const V3 albedo = v3(mat.albedo);
	V3 R0 = eta_to_schlick_R0(eta) * mix(v3(1.0f), calc_tint(albedo), mat.specular_tint);
	R0 = mix(R0, albedo, mat.metallic);
	const float fr_dielectric = fresnel_dielectric(wo_dot_h, eta, true);
	const V3 fr_metallic = fresnel_schlick(R0, v3(1.0f), wo_dot_h);
	return mix(v3(fr_dielectric), fr_metallic, mat.metallic);
}
LMB_DN V3 sample_hemisphere(const V2& xi) {
	const V2 d = concentric_sample_disk(xi);
	const float z = sqrtf(gmax(0.f, 1.f - dot(d, d)));
	return v3(d.x, d.y, z);
}

// ------------------------------------------------------------------------------------------------ microfacet_commons
LMB_D float smith_lambda_iso(float alpha_sqr, float cos_theta) {
	if (cos_theta == 0) return 0;
	const float cos_sqr = cos_theta * cos_theta;
	const float tan_sqr = gmax(1.0f - cos_sqr, 0.0f) / cos_sqr;
	return 0.5f * (sqrtf(1.0f + alpha_sqr * tan_sqr) - 1);
}
LMB_D float smith_lambda_aniso(const V3& w, const V2& alpha) {
	const float cos_sqr = w.z * w.z;
	const float sin_sqr = gmax(1.0f - cos_sqr, 0.0f);
	const float tan_sqr = sin_sqr / cos_sqr;
	if (isinf(tan_sqr)) return 0.0f;
	const V2 cos_phi_sqr = sin_sqr == 0.0f ? v2(1.0f, 0.0f) : gclamp(v2(w.x * w.x, w.y * w.y), 0.0f, 1.0f) / sin_sqr;
	const float alpha_sqr = dot(cos_phi_sqr, alpha * alpha);
	return 0.5f * (sqrtf(1.0f + alpha_sqr * tan_sqr) - 1);
}
LMB_D float g1_ggx_aniso(const V3& w, const V2& alpha) { return 1.0f / (1.0f + smith_lambda_aniso(w, alpha)); }
LMB_D float g_ggx_corr_iso(float alpha, const V3& wo, const V3& wi) {
	const float alpha_sqr = alpha * alpha;
	return 1.0f / (1.0f + smith_lambda_iso(alpha_sqr, wo.z) + smith_lambda_iso(alpha_sqr, wi.z));
}
LMB_D float g_ggx_corr_aniso(const V2& alpha, const V3& wo, const V3& wi) {
	return 1.0f / (1.0f + smith_lambda_aniso(wo, alpha) + smith_lambda_aniso(wi, alpha));
}
LMB_D float d_ggx_aniso(const V2& alpha, const V3& h) {
	const float cos_sqr = h.z * h.z;
	const float sin_sqr = gmax(1.0f - cos_sqr, 0.0f);
	const float tan_sqr = sin_sqr / cos_sqr;
	if (isinf(tan_sqr)) return 0.0f;
	const float cos_4 = cos_sqr * cos_sqr;
	if (cos_4 < 1e-16f) return 0.0f;
	const V2 phi_sqr = sin_sqr == 0.0f ? v2(1.0f, 0.0f) : gclamp(v2(h.x * h.x, h.y * h.y), 0.0f, 1.0f) / sin_sqr;
	const V2 alpha_sqr = phi_sqr / (alpha * alpha);
	const float e = tan_sqr * (alpha_sqr.x + alpha_sqr.y);
	return 1.0f / (LMB_PI * alpha.x * alpha.y * cos_4 * (1.0f + e) * (1.0f + e));
}
LMB_D float d_ggx_iso(float alpha_sqr, float cos_theta) {
	const float d = ((cos_theta * alpha_sqr - cos_theta) * cos_theta + 1);
	return alpha_sqr / (d * d * LMB_PI);
}
LMB_D float g1_ggx_iso(float alpha_sqr, float cos_theta) {
	const float cos_sqr = cos_theta * cos_theta;
	const float tan_sqr = gmax(1.0f - cos_sqr, 0.0f) / cos_sqr;
	return 2.0f / (1.0f + sqrtf(1.0f + alpha_sqr * tan_sqr));
}
LMB_D float vndf_pdf_iso(float alpha, const V3& wo, const V3& h, float& D) {
	D = 0.0f;
	if (wo.z <= 0) return 0.0f;
	const float alpha_sqr = alpha * alpha;
	const float G1 = g1_ggx_iso(alpha_sqr, wo.z);
	D = d_ggx_iso(alpha_sqr, h.z);
	return G1 * D * gmax(0.0f, dot(wo, h)) / fabsf(wo.z);
}
LMB_D float vndf_pdf_aniso(const V2& alpha, const V3& wo, const V3& h, float& D) {
	D = 0.0f;
	if (wo.z <= 0) return 0.0f;
	const float G1 = g1_ggx_aniso(wo, alpha);
	D = d_ggx_aniso(alpha, h);
	return G1 * D * gmax(0.0f, dot(wo, h)) / fabsf(wo.z);
}
LMB_DN V3 sample_ggx_vndf_common(const V2& alpha, const V3& wo, const V2& xi) {
	const V3 wo_hemisphere = normalize(v3(alpha.x * wo.x, alpha.y * wo.y, wo.z));
	const float phi = 2.0f * LMB_PI * xi.x;
	const float z = ((1.0f - xi.y) * (1.0f + wo_hemisphere.z)) - wo_hemisphere.z;
	const float sin_theta = sqrtf(gclamp(1.0f - z * z, 0.0f, 1.0f));
	const float2 scp = d_sincosf(phi);
	const float x = sin_theta * scp.y;
	const float y = sin_theta * scp.x;
	const V3 n_h = v3(x, y, z) + wo_hemisphere;
	return normalize(v3(alpha.x * n_h.x, alpha.y * n_h.y, gmax(0.0f, n_h.z)));
}

struct BsdfSample {
	V3 f;
	V3 wi;
	float pdf;
	float cos_theta;
};
LMB_D BsdfSample zero_sample() { return BsdfSample{v3(0.0f), v3(0.0f), 0.0f, 0.0f}; }

// ------------------------------------------------------------------------------------------------ diffuse / mirror / glass
LMB_D BsdfSample sample_lambertian(const lmb_material& mat, const V3& wo, const V2& xi) {
	BsdfSample s = zero_sample();
	s.wi = sample_hemisphere(xi);
	s.cos_theta = s.wi.z;
	s.pdf = s.cos_theta * LMB_INV_PI;
	if (gmin(s.wi.z, wo.z) <= 0.0f) return s;
	s.f = v3(mat.albedo) * LMB_INV_PI;
	return s;
}
LMB_D V3 eval_lambertian(const lmb_material& mat, const V3& wo, const V3& wi, float& pdf_w) {
	if (gmin(wi.z, wo.z) <= 0.0f) return v3(0.0f);
	pdf_w = wi.z * LMB_INV_PI;
	return v3(mat.albedo) * LMB_INV_PI;
}
LMB_D float lambertian_pdf(const V3& wo, const V3& wi) {
	if (gmin(wi.z, wo.z) <= 0.0f) return 0.0f;
	return wi.z * LMB_INV_PI;
}
LMB_D BsdfSample sample_mirror(const V3& wo) {
	BsdfSample s = zero_sample();
	const V3 n_s = v3(0, 0, 1);
	if (dot(wo, n_s) <= 0) return s;
	s.wi = reflect(-wo, n_s);
	s.cos_theta = dot(s.wi, n_s);
	s.pdf = 1.0f;
	s.f = v3(1.0f) / s.cos_theta;
	return s;
}
LMB_D BsdfSample sample_glass(const lmb_material& mat, const V3& wo, uint32_t mode, bool forward_facing) {
	BsdfSample s = zero_sample();
	const V3 n_s = v3(0, 0, 1);
	V3 f;
	float unused;
	refract_dir(n_s, wo, forward_facing, mat.ior, mode, s.wi, f, unused);
	s.cos_theta = dot(n_s, s.wi);
	s.pdf = 1.f;
	s.f = f / fabsf(s.cos_theta);
	return s;
}

// ------------------------------------------------------------------------------------------------ dielectric
LMB_D float modify_thin_roughness(float ior, float roughness) { return gclamp((0.65f * ior - 0.35f) * roughness, 0.0f, 1.0f); }

static __device__ __noinline__ BsdfSample sample_dielectric(const lmb_material& mat, const V3& wo, uint32_t mode, bool forward_facing, const V2& xi) {
	BsdfSample s = zero_sample();
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
			s.wi = v3(-wo.x, -wo.y, wo.z);
			s.cos_theta = s.wi.z;
			s.f = v3(F) / fabsf(s.cos_theta);
			s.pdf = pr / (pr + pt);
		} else {
			V3 f;
			float unused;
			refract_dir(v3(0, 0, 1), wo, forward_facing, mat.ior, mode, s.wi, f, unused);
			s.cos_theta = s.wi.z;
			s.f = f * (1.0f - F) / fabsf(s.cos_theta);
			s.pdf = pt / (pr + pt);
		}
		return s;
	}
	if (!has_reflection && !has_transmission) return s;
	float D;
	const V3 h = sample_ggx_vndf_common(v2(alpha, alpha), wo, xi);
	float pdf_w = vndf_pdf_iso(alpha, wo, h, D);
	const float F = fresnel_dielectric(dot(wo, h), mat.ior, forward_facing);
	const float pr = has_reflection ? F : 0.0f;
	const float pt = has_transmission ? (1.0f - F) : 0.0f;
	const bool is_reflection = ((pr + pt) * xi.x) < pr;
	V3 f = v3(0.0f);
	V3 wi = v3(0.0f);
	if (is_reflection) {
		wi = reflect(-wo, h);
		if (wo.z * wi.z < 0) {
			s.wi = wi;
			s.pdf = pdf_w;
			return s;
		}
		pdf_w = pdf_w * (pr / (pr + pt)) / (4.0f * fabsf(dot(wo, h)));
		f = v3(0.25f * D * F * g_ggx_corr_iso(alpha, wo, wi) / (wi.z * wo.z));
	} else {
		V3 base_col;
		float possibly_modified_inv_eta;
		if (mat.thin == 1) {
			wi = reflect(-wo, h);
			base_col = vsqrt(v3(mat.albedo));
			possibly_modified_inv_eta = mat.ior;
		} else {
			refract_dir(h, wo, forward_facing, mat.ior, mode, wi, f, possibly_modified_inv_eta);
			base_col = v3(mat.albedo);
		}
		if (wo.z * wi.z > 0 || wi.z > 0) {
			s.wi = wi;
			s.pdf = pdf_w;
			return s;
		}
		float jacobian_denom = dot(wi, h) + dot(wo, h) * possibly_modified_inv_eta;
		jacobian_denom = jacobian_denom * jacobian_denom;
		const float jacobian = fabsf(dot(wi, h)) / jacobian_denom;
		pdf_w = pdf_w * (pt / (pr + pt)) * jacobian;
		f = base_col * f * (1.0f - F) * D * g_ggx_corr_iso(alpha, wo, wi) * fabsf(dot(wi, h) * dot(wo, h) / (wi.z * wo.z * jacobian_denom));
	}
	s.wi = wi;
	s.pdf = pdf_w;
	s.cos_theta = wi.z;
	s.f = f;
	return s;
}

static __device__ __noinline__ V3 eval_dielectric(const lmb_material& mat, const V3& wo, const V3& wi, float& pdf_w, bool forward_facing) {
	pdf_w = 0.0f;
	const float roughness = mat.thin == 1 ? modify_thin_roughness(mat.ior, mat.roughness) : mat.roughness;
	const float alpha = roughness * roughness;
	if (alpha == 0 || mat.ior == 1) return v3(0.0f);
	const bool has_reflection = has_prop(mat.bsdf_props, LMB_FLAG_REFLECTION);
	const bool has_transmission = has_prop(mat.bsdf_props, LMB_FLAG_TRANSMISSION);
	if (!has_reflection && !has_transmission) return v3(0.0f);
	float eta = 1.0f;
	const bool is_reflection = wi.z * wo.z > 0.0f;
	if (!is_reflection) eta = forward_facing ? mat.ior : 1.0f / mat.ior;
	V3 h = normalize(wo + wi * eta);
	h *= gsign(h.z);
	if (wi.z == 0 || wo.z == 0 || dot(h, h) == 0) return v3(0.0f);
	if (dot(wi, h) * wi.z < 0 || dot(wo, h) * wo.z < 0) return v3(0.0f);
	const float F = fresnel_dielectric(dot(wo, h), mat.ior, forward_facing);
	const float pr = has_reflection ? F : 0.0f;
	const float pt = has_transmission ? (1.0f - F) : 0.0f;
	float D;
	pdf_w = vndf_pdf_iso(alpha, wo, h, D);
	const float G = g_ggx_corr_iso(alpha, wo, wi);
	V3 f = v3(0.0f);
	if (is_reflection) {
		const float jacobian = 1.0f / (4.0f * fabsf(dot(wo, h)));
		const float prob_reflection = pr / (pr + pt);
		pdf_w = pdf_w * jacobian * prob_reflection;
		f = v3(0.25f * D * G * F / fabsf(wo.z * wi.z));
	} else {
		float jacobian_denom = dot(wi, h) + dot(wo, h) / eta;
		jacobian_denom = jacobian_denom * jacobian_denom;
		const float jacobian = fabsf(dot(wi, h)) / jacobian_denom;
		const float prob_refraction = pt / (pr + pt);
		pdf_w = pdf_w * jacobian * prob_refraction;
	}
	return f;
}

// ------------------------------------------------------------------------------------------------ conductor
static __device__ __noinline__ BsdfSample sample_conductor(const lmb_material& mat, const V3& wo, const V2& xi) {
	BsdfSample s = zero_sample();
	if (wo.z <= 0.0f) return s;
	const float alpha = mat.roughness * mat.roughness;
	if (!has_prop(mat.bsdf_props, LMB_FLAG_REFLECTION)) return s;
	if (effectively_delta(alpha)) {
		s.wi = v3(-wo.x, -wo.y, wo.z);
		s.pdf = 1.0f;
		s.cos_theta = s.wi.z;
		const V3 F = fresnel_conductor(s.cos_theta, v3(mat.albedo), v3(mat.k));
		s.f = F / fabsf(s.cos_theta);
		return s;
	}
	float D;
	const V3 h = sample_ggx_vndf_common(v2(alpha, alpha), wo, xi);
	s.pdf = vndf_pdf_iso(alpha, wo, h, D);
	const V3 F = fresnel_conductor(dot(wo, h), v3(mat.albedo), v3(mat.k));
	s.wi = reflect(-wo, h);
	if (wo.z * s.wi.z < 0) return s;
	s.pdf /= (4.0f * dot(wo, h));
	s.cos_theta = s.wi.z;
	s.f = 0.25f * D * F * g_ggx_corr_iso(alpha, wo, s.wi) / (s.wi.z * wo.z);
	return s;
}
static __device__ __noinline__ V3 eval_conductor(const lmb_material& mat, const V3& wo, const V3& wi, float& pdf_w) {
	pdf_w = 0;
	const float alpha = mat.roughness * mat.roughness;
	if (effectively_delta(alpha)) return v3(0.0f);
	if (wo.z * wi.z < 0) return v3(0.0f);
	if (wo.z == 0 || wi.z == 0) return v3(0.0f);
	V3 h = normalize(wo + wi);
	h *= gsign(h.z);
	const float jacobian = 1.0f / (4.0f * dot(wo, h));
	float D;
	pdf_w = vndf_pdf_iso(alpha, wo, h, D) * jacobian;
	const V3 F = fresnel_conductor(dot(wo, h), v3(mat.albedo), v3(mat.k));
	return 0.25f * D * F * g_ggx_corr_iso(alpha, wo, wi) / (wi.z * wo.z);
}

// ------------------------------------------------------------------------------------------------ principled
LMB_D V2 calc_anisotropy(float roughness, float anisotropic) {
	const float aspect = sqrtf(1.0f - 0.9f * anisotropic);
	const float roughness_sqr = roughness * roughness;
	return v2(gmax(0.001f, roughness_sqr / aspect), gmax(0.001f, roughness_sqr * aspect));
}
struct LobeProbs {
	float spec, diff, clearcoat, spec_trans;
};
LMB_D LobeProbs sampling_probs(const lmb_material& mat, float F_dielectric, bool forward_facing) {
	LobeProbs p;
	const float brdf_weight = (1.0f - mat.spec_trans) * (1.0f - mat.metallic);
	const float bsdf_weight = (1.0f - mat.metallic) * mat.spec_trans;
	p.spec = forward_facing ? (1.0f - bsdf_weight * (1.0f - F_dielectric)) : F_dielectric;
	p.spec_trans = forward_facing ? (bsdf_weight * (1.0f - F_dielectric)) : (1.0f - F_dielectric);
	p.diff = forward_facing ? brdf_weight : 0.0f;
	p.clearcoat = forward_facing ? 0.25f * gclamp(mat.clearcoat, 0.0f, 1.0f) : 0.0f;
	const float norm = 1.0f / (p.spec + p.spec_trans + p.diff + p.clearcoat);
	p.spec *= norm;
	p.diff *= norm;
	p.clearcoat *= norm;
	p.spec_trans *= norm;
	return p;
}
LMB_D V3 disney_diffuse_factor(const lmb_material& mat, const V3& wo, const V3& wi) {
	float f_wi, f_wo;
	disney_fresnel(wi, wo, mat.roughness, f_wi, f_wo);
	const float roughness_sqr = mat.roughness * mat.roughness;
	float ss = 0;
	const float rr = 2.0f * roughness_sqr * wi.z * wi.z;
	const float f_retro = rr * (f_wi + f_wo * f_wi * f_wo * (rr - 1.0f));
	const float f_diff = (1.0f - 0.5f * f_wi) * (1.0f - 0.5f * f_wo);
	if (mat.flatness > 0.0f) {
		const float fss90 = 0.5f * rr;
		const float f_ss = mix(1.0f, fss90, f_wi) * mix(1.0f, fss90, f_wo);
		ss = 1.25f * (f_ss * (1.0f / (wi.z + wo.z) - 0.5f) + 0.5f);
	}
	const float ss_approx_and_diff = mix(f_diff + f_retro, ss, mat.flatness);
	return v3(mat.albedo) * ss_approx_and_diff * LMB_INV_PI;
}
LMB_D float clearcoat_factor(const lmb_material& mat, const V3& wo, const V3& wi, const V3& h, float& D) {
	const float alpha_2 = 0.25f * 0.25f;
	D = d_ggx_iso(mix(0.1f, 0.001f, mat.clearcoat_gloss), h.z);
	const float F = fresnel_schlick(0.04f, 1.0f, dot(wi, h));
	const float G = g1_ggx_iso(alpha_2, wo.z) * g1_ggx_iso(alpha_2, wi.z);
	return 0.25f * mat.clearcoat * D * F * G;
}
LMB_D BsdfSample sample_disney_diffuse(const lmb_material& mat, const V3& wo, const V2& xi) {
	BsdfSample s = zero_sample();
	s.wi = sample_hemisphere(xi);
	s.cos_theta = s.wi.z;
	s.pdf = s.cos_theta * LMB_INV_PI;
	if (gmin(s.wi.z, wo.z) <= 0.0f) return s;
	s.f = disney_diffuse_factor(mat, wo, s.wi);
	return s;
}
LMB_D BsdfSample sample_principled_brdf(const lmb_material& mat, const V3& wo, const V2& xi, float eta) {
	BsdfSample s = zero_sample();
	if (!has_prop(mat.bsdf_props, LMB_FLAG_REFLECTION)) return s;
	float D;
	const V2 alpha = calc_anisotropy(mat.roughness, mat.anisotropy);
	if (effectively_delta(alpha)) {
		s.wi = v3(-wo.x, -wo.y, wo.z);
		s.pdf = 1.0f;
		s.cos_theta = s.wi.z;
		const V3 F = disney_fresnel(mat, wo, v3(0, 0, 1), eta);
		s.f = F / fabsf(s.cos_theta);
		return s;
	}
	const V3 h = sample_ggx_vndf_common(alpha, wo, xi);
	s.pdf = vndf_pdf_aniso(alpha, wo, h, D);
	s.wi = reflect(-wo, h);
	if (wo.z * s.wi.z < 0) return s;
	const V3 F = disney_fresnel(mat, wo, h, eta);
	s.pdf /= (4.0f * dot(wo, h));
	s.cos_theta = s.wi.z;
	s.f = 0.25f * D * F * g_ggx_corr_aniso(alpha, wo, s.wi) / (s.wi.z * wo.z);
	return s;
}
LMB_D BsdfSample sample_clearcoat(const lmb_material& mat, const V3& wo, const V2& xi) {
	BsdfSample s = zero_sample();
	const float alpha_2 = 0.25f * 0.25f;
	const float cos_t = sqrtf(gmax(0.0f, (1.0f - d_powf(alpha_2, 1.0f - xi.x)) / (1.0f - alpha_2)));
	const float sin_t = sqrtf(gmax(0.0f, 1.0f - s.cos_theta * s.cos_theta));  // Q2: stale inout cos_theta (= 0)
	const float phi = LMB_TWO_PI * xi.y;
	const float2 scp = d_sincosf(phi);
	V3 h = v3(sin_t * scp.y, sin_t * scp.x, cos_t);
	if (dot(h, wo) < 0.0f) h *= -1.0f;
	s.wi = reflect(-wo, h);
	if (dot(s.wi, wo) < 0.0f) return s;
	float D;
	const float f_clearcoat = clearcoat_factor(mat, wo, s.wi, h, D);
	s.pdf = D / (4.0f * dot(wo, h));
	s.cos_theta = s.wi.z;
	s.f = v3(f_clearcoat);
	return s;
}
LMB_D V3 eval_clearcoat(const lmb_material& mat, const V3& wo, const V3& wi, float& pdf_w) {
	const V3 h = normalize(wo + wi);
	float D;
	const float f_clearcoat = clearcoat_factor(mat, wo, wi, h, D);
	pdf_w = D / (4.0f * dot(wo, h));
	return v3(f_clearcoat);
}
LMB_D V3 eval_principled_brdf(const lmb_material& mat, const V3& wo, const V3& wi, float& pdf_w, bool forward_facing) {
	pdf_w = 0;
	const V2 alpha = calc_anisotropy(mat.roughness, mat.anisotropy);
	if (effectively_delta(alpha)) return v3(0.0f);
	if (wo.z * wi.z < 0) return v3(0.0f);
	if (wo.z == 0 || wi.z == 0) return v3(0.0f);
	const V3 h = normalize(wo + wi);
	const float jacobian = 1.0f / (4.0f * dot(wo, h));
	float D;
	pdf_w = vndf_pdf_aniso(alpha, wo, h, D) * jacobian;
	const float eta = forward_facing ? mat.ior : 1.0f / mat.ior;
	const V3 F = disney_fresnel(mat, wo, h, eta);
	return 0.25f * D * F * g_ggx_corr_aniso(alpha, wo, wi) / (wi.z * wo.z);
}
static __device__ __noinline__ BsdfSample sample_principled(const lmb_material& mat, const V3& wo, uint32_t mode, bool forward_facing, const V3& xi) {
	BsdfSample s = zero_sample();
	if (wo.z <= 0.0f) return s;
	const float F = fresnel_dielectric(wo.z, mat.ior, forward_facing);
	const LobeProbs p = sampling_probs(mat, F, forward_facing);
	const float eta = forward_facing ? mat.ior : 1.0f / mat.ior;
	float p_lobe = 0.0f;
	const V2 xy = v2(xi.x, xi.y);
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
static __device__ __noinline__ V3 eval_principled(const lmb_material& mat, const V3& wo, const V3& wi, float& pdf_w, bool forward_facing) {
	pdf_w = 0.0f;
	const float F = fresnel_dielectric(wo.z, mat.ior, forward_facing);
	const LobeProbs p = sampling_probs(mat, F, forward_facing);
	V3 f = v3(0.0f);
	float pdf = 0;
	const float brdf_weight = (1.0f - mat.spec_trans) * (1.0f - mat.metallic);
	const float bsdf_weight = (1.0f - mat.metallic) * mat.spec_trans;
	if (p.spec > 0) {
		f += eval_principled_brdf(mat, wo, wi, pdf, forward_facing);
		pdf *= p.spec;
		pdf_w += pdf;
	}
	const bool upper_hemisphere = gmin(wi.z, wo.z) > 0;
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
		f += bsdf_weight * eval_dielectric(mat, wo, wi, pdf, forward_facing);
		pdf *= p.spec_trans;
		pdf_w += pdf;
	}
	return f;
}

// ------------------------------------------------------------------------------------------------ dispatch
// TYPE = LMB_BSDF_ANY dispatches on mat.bsdf_type at run time; a concrete LMB_BSDF_* constant folds the switch away so
// that the per-material shade kernels carry only their own lobe code (material-sorted queues, wavefront.cu).
#define LMB_BSDF_ANY 0xFFFFFFFFu
template <uint32_t TYPE>
LMB_D BsdfSample sample_bsdf_t(const V3& n_s, const V3& wo_world, const lmb_material& mat, uint32_t mode, bool forward_facing, const V3& rands) {
	V3 T, B;
	branchless_onb(n_s, T, B);
	const V3 wo = to_local(wo_world, T, B, n_s);
	BsdfSample s = zero_sample();
	const V2 xy = v2(rands.x, rands.y);
	switch (TYPE == LMB_BSDF_ANY ? mat.bsdf_type : TYPE) {
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
LMB_D BsdfSample sample_bsdf(const V3& n_s, const V3& wo_world, const lmb_material& mat, uint32_t mode, bool forward_facing, const V3& rands) {
	return sample_bsdf_t<LMB_BSDF_ANY>(n_s, wo_world, mat, mode, forward_facing, rands);
}
template <uint32_t TYPE>
LMB_D V3 eval_bsdf_t(const V3& n_s, const V3& wo_world, const lmb_material& mat, bool forward_facing, const V3& wi_world, float& pdf_w) {
	pdf_w = 0;
	V3 T, B;
	branchless_onb(n_s, T, B);
	const V3 wo = to_local(wo_world, T, B, n_s);
	const V3 wi = to_local(wi_world, T, B, n_s);
	switch (TYPE == LMB_BSDF_ANY ? mat.bsdf_type : TYPE) {
		case LMB_BSDF_DIFFUSE:
			return eval_lambertian(mat, wo, wi, pdf_w);
		case LMB_BSDF_DIELECTRIC:
			return eval_dielectric(mat, wo, wi, pdf_w, forward_facing);
		case LMB_BSDF_CONDUCTOR:
			return eval_conductor(mat, wo, wi, pdf_w);
		case LMB_BSDF_PRINCIPLED:
			return eval_principled(mat, wo, wi, pdf_w, forward_facing);
		default:
			break;
	}
	return v3(0.0f);
}

LMB_D V3 eval_bsdf(const V3& n_s, const V3& wo_world, const lmb_material& mat, bool forward_facing, const V3& wi_world, float& pdf_w) {
	return eval_bsdf_t<LMB_BSDF_ANY>(n_s, wo_world, mat, forward_facing, wi_world, pdf_w);
}

// ------------------------------------------------------------------------------------------------ sky
namespace atmo {
#define LMB_PLANET_RADIUS 6371000.0f
#define LMB_ATMOSPHERE_HEIGHT 100000.0f
#define LMB_RAYLEIGH_HEIGHT (LMB_ATMOSPHERE_HEIGHT * 0.08f)
#define LMB_MIE_HEIGHT (LMB_ATMOSPHERE_HEIGHT * 0.012f)
LMB_D V3 planet_center() { return v3(0, -LMB_PLANET_RADIUS, 0); }
LMB_D V3 c_rayleigh() { return v3(5.802f, 13.558f, 33.100f) * 1e-6f; }
LMB_D V3 c_mie() { return v3(3.996f, 3.996f, 3.996f) * 1e-6f; }
LMB_D V3 c_ozone() { return v3(0.650f, 1.881f, 0.085f) * 1e-6f; }
LMB_D V2 sphere_intersection(V3 ray_start, const V3& ray_dir, const V3& center, float radius) {
	ray_start -= center;
	const float a = dot(ray_dir, ray_dir);
	const float b = 2.0f * dot(ray_start, ray_dir);
	const float c = dot(ray_start, ray_start) - (radius * radius);
	float d = b * b - 4 * a * c;
	if (d < 0) return v2(-1.0f, -1.0f);
	d = sqrtf(d);
	return v2(-b - d, -b + d) / (2 * a);
}
LMB_D V2 planet_intersection(const V3& s, const V3& d) { return sphere_intersection(s, d, planet_center() * v3(0, 1.00001f, 0), LMB_PLANET_RADIUS); }
LMB_D V2 atmosphere_intersection(const V3& s, const V3& d) {
	return sphere_intersection(s, d, planet_center(), LMB_PLANET_RADIUS + LMB_ATMOSPHERE_HEIGHT);
}
LMB_D float phase_rayleigh(float costh) { return 3 * (1 + costh * costh) / (16 * LMB_PI); }
LMB_D float phase_mie(float costh, float g) {
	g = gmin(g, 0.9381f);
	const float k = 1.55f * g - 0.55f * g * g * g;
	const float kcosth = k * costh;
	return (1 - k * k) / ((4 * LMB_PI) * (1 - kcosth) * (1 - kcosth));
}
LMB_D float height(const V3& p) { return length(planet_center() - p) - LMB_PLANET_RADIUS; }
// x / C for a compile-time constant C, bit for bit the IEEE quotient the oracle computes: q = x * RN(1/C), one residual step
// q' = fma(fma(-q, C, x), RN(1/C), q) (Markstein). Checked exhaustively over all 2^32 floats for the three divisors below
// (tools/check_div_const.c): the only mismatches are |x| < 1e-36, where the quotient is subnormal (and the sign of a zero), so
// those take the division instruction. 3 FP instructions instead of the ~9 + FCHK of div.rn; density() runs 9 x 3 of them in
// each of the 64 march steps of an escaped ray.
template <typename C>
LMB_D float div_const(float x, C) {  // requires |x| > 1e-30 (density() tests once for its three divisions)
	constexpr float c = C::value();
	constexpr float rc = 1.0f / c;
	const float q = x * rc;
	return fmaf(fmaf(-q, c, x), rc, q);
}
struct CRayleighH { static constexpr float value() { return LMB_RAYLEIGH_HEIGHT; } };
struct CMieH { static constexpr float value() { return LMB_MIE_HEIGHT; } };
struct COzoneW { static constexpr float value() { return 15000.0f; } };
// The exponentials go through lmb_expf_window (bit for bit lmb_expf inside its window, lmb_detmath.h) with ONE range test per group:
// every argument of the march sits in the window (the most negative is -h / 1200 m at the top of the atmosphere, -83.3), so the
// d_expf calls below are the cold side of a branch that is there for exactness only.
LMB_D bool exp_in_window(float x) { return fabsf(x) < 86.0f; }
LMB_D V3 density(float h) {
	const float ho = fabsf(h - 25000.0f);
	float qr, qm, qo;
	if (fabsf(h) > 1e-30f && ho > 1e-30f) {  // one guard for the three constant divisions (div_const)
		qr = div_const(h, CRayleighH{}), qm = div_const(h, CMieH{}), qo = div_const(ho, COzoneW{});
	} else {
		qr = h / LMB_RAYLEIGH_HEIGHT, qm = h / LMB_MIE_HEIGHT, qo = ho / 15000.0f;
	}
	const float xr = -gmax(0.0f, qr), xm = -gmax(0.0f, qm);
	const float oz = gmax(0.0f, 1 - qo);
	if (exp_in_window(xr) && exp_in_window(xm)) return v3(lmb_expf_window(xr), lmb_expf_window(xm), oz);
	return v3(d_expf(xr), d_expf(xm), oz);
}
LMB_D V3 vexp(const V3& a) {
	if (exp_in_window(a.x) && exp_in_window(a.y) && exp_in_window(a.z)) return v3(lmb_expf_window(a.x), lmb_expf_window(a.y), lmb_expf_window(a.z));
	return v3(d_expf(a.x), d_expf(a.y), d_expf(a.z));
}
LMB_D V3 absorb(const V3& od) { return vexp(-(od.x * c_rayleigh() + od.y * c_mie() * 1.1f + od.z * c_ozone()) * 1.0f); }
LMB_D V3 optical_depth(const V3& ray_start, const V3& ray_dir) {
	const V2 isect = atmosphere_intersection(ray_start, ray_dir);
	const float ray_length = isect.y;
	const int sample_count = 8;
	const float step = ray_length / sample_count;
	V3 od = v3(0.0f);
	for (int i = 0; i < sample_count; i++) {
		const V3 local_pos = ray_start + ray_dir * (i + 0.5f) * step;
		const V3 local_density = density(height(local_pos));
		od += local_density * step;
	}
	return od;
}
static __device__ __noinline__ V3 integrate_scattering(V3 ray_start, const V3& ray_dir, float ray_length, const V3& light_dir, const V3& light_color) {
	const float ray_height = height(ray_start);
	const float exponent = 1 + gclamp(1 - ray_height / LMB_ATMOSPHERE_HEIGHT, 0.0f, 1.0f) * 8;
	const V2 isect = atmosphere_intersection(ray_start, ray_dir);
	ray_length = gmin(ray_length, isect.y);
	if (isect.x > 0) {
		ray_start += ray_dir * isect.x;
		ray_length -= isect.x;
	}
	const float costh = dot(ray_dir, light_dir);
	const float phase_r = phase_rayleigh(costh);
	const float phase_m = phase_mie(costh, 0.85f);
	const int sample_count = 64;
	V3 od = v3(0.0f), rayleigh = v3(0.0f), mie = v3(0.0f);
	float prev_ray_time = 0;
	for (int i = 0; i < sample_count; i++) {
		const float ray_time = d_powf(float(i) / sample_count, exponent) * ray_length;
		const float step = (ray_time - prev_ray_time);
		const V3 local_pos = ray_start + ray_dir * ray_time;
		const V3 local_density = density(height(local_pos));
		od += local_density * step;
		const V3 view_tr = absorb(od);
		const V3 od_light = optical_depth(local_pos, light_dir);
		const V3 light_tr = absorb(od_light);
		rayleigh += view_tr * light_tr * phase_r * local_density.x * step;
		mie += view_tr * light_tr * phase_m * local_density.y * step;
		prev_ray_time = ray_time;
	}
	return (rayleigh * c_rayleigh() + mie * c_mie()) * light_color * 20.0f;
}
}  // namespace atmo

}  // namespace lmb
