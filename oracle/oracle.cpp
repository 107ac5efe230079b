// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h). Not part of the product path.
//
// oracle.cpp: the per-pixel path loop of the reference (src/shaders/integrators/path/path.rgen:26-112), its NEE+MIS
// routine (src/shaders/integrators/pt_commons.glsl:3-42), hit-record construction (src/shaders/ray.rchit:24-84),
// light sampling (src/shaders/commons.glsl:112-300) and the film update, restated in C++/glm over the CPU LBVH of
// lbvh_cpu.h, parallelised with OpenMP over scanlines.
#include "oracle.h"

#include <omp.h>

#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "lbvh_cpu.h"
#include "shading.h"

using namespace orc;

struct orc_scene {
	lmb_scene_desc sd;
	Lbvh bvh;
	float srgb_lut[256];
};

namespace {

struct HitPayload {  // utils.glsl:15-26
	vec3 n_g, n_s, pos;
	vec2 uv;
	uint material_idx, triangle_idx, instance_idx;
	float area;
};

struct Counters {
	uint64_t closest = 0, shadow = 0, probe = 0, nan_px = 0;
	TraceStats ts;
};

// commons.glsl:9-17 sampler: LINEAR filter, REPEAT addressing, LOD 0 (raygen stage has no derivatives),
// VK_FORMAT_R8G8B8A8_SRGB (LumenScene.cpp:193-213). Texel-centre convention and fp32 weights are this build's
// definition of what the fixed-function sampler does (hardware uses 8-bit sub-texel weights; unpinned).
inline vec3 sample_texture(const orc_scene& s, uint32_t id, const vec2& uv) {
	const lmb_texture& t = s.sd.textures[id];
	const int W = (int)t.width, H = (int)t.height;
	float u = uv.x * (float)W - 0.5f, v = uv.y * (float)H - 0.5f;
	if (!(std::fabs(u) < 1e9f)) u = 0.0f;
	if (!(std::fabs(v) < 1e9f)) v = 0.0f;
	const float x0f = std::floor(u), y0f = std::floor(v);
	const float fx = u - x0f, fy = v - y0f;
	auto wrap = [](int i, int n) { return ((i % n) + n) % n; };
	const int x0 = wrap((int)x0f, W), x1 = wrap((int)x0f + 1, W);
	const int y0 = wrap((int)y0f, H), y1 = wrap((int)y0f + 1, H);
	auto texel = [&](int x, int y) {
		const uint8_t* p = t.rgba8 + 4 * ((size_t)y * W + x);
		return vec3(s.srgb_lut[p[0]], s.srgb_lut[p[1]], s.srgb_lut[p[2]]);
	};
	const vec3 top = texel(x0, y0) * (1.0f - fx) + texel(x1, y0) * fx;
	const vec3 bot = texel(x0, y1) * (1.0f - fx) + texel(x1, y1) * fx;
	return top * (1.0f - fy) + bot * fy;
}

// bsdf_commons.glsl:16-22
inline lmb_material load_material(const orc_scene& s, uint material_idx, const vec2& uv) {
	lmb_material m = s.sd.materials[material_idx];
	if (m.texture_id > -1) {
		const vec3 a = v3(m.albedo) * sample_texture(s, (uint32_t)m.texture_id, uv);
		m.albedo[0] = a.x, m.albedo[1] = a.y, m.albedo[2] = a.z;
	}
	return m;
}

inline vec3 mul_dir_w2o(const vec3& v, const mat4& w2o) {  // vec3 * mat4x3 (row vector times matrix)
	return vec3(glm::dot(v, vec3(w2o[0])), glm::dot(v, vec3(w2o[1])), glm::dot(v, vec3(w2o[2])));
}

// ray.rchit:24-84
inline HitPayload build_hit(const orc_scene& s, const Hit& h) {
	HitPayload p;
	const Lbvh& b = s.bvh;
	const uint mesh = b.tri_mesh[h.prim], prim = b.tri_local[h.prim];
	const lmb_prim_mesh_info& pinfo = s.sd.prim_infos[mesh];
	const uint index_offset = pinfo.index_offset + 3 * prim;
	const lmb_vertex& a0 = s.sd.vertices[s.sd.indices[index_offset + 0] + pinfo.vertex_offset];
	const lmb_vertex& a1 = s.sd.vertices[s.sd.indices[index_offset + 1] + pinfo.vertex_offset];
	const lmb_vertex& a2 = s.sd.vertices[s.sd.indices[index_offset + 2] + pinfo.vertex_offset];
	const vec3 v0 = v3(a0.pos), v1 = v3(a1.pos), v2 = v3(a2.pos);
	const vec3 n0 = v3(a0.normal), n1 = v3(a1.normal), n2 = v3(a2.normal);
	const vec2 uv0(a0.uv0[0], a0.uv0[1]), uv1(a1.uv0[0], a1.uv0[1]), uv2(a2.uv0[0], a2.uv0[1]);
	const vec3 bary(1.0f - h.b1 - h.b2, h.b1, h.b2);
	const mat4 o2w = m4(s.sd.world_matrices + 16 * mesh);
	const mat4 w2o = m4(s.sd.inv_world_matrices + 16 * mesh);
	const vec3 pos = v0 * bary.x + v1 * bary.y + v2 * bary.z;
	p.pos = vec3(o2w * vec4(pos, 1.0f));
	const vec3 nrm = glm::normalize(n0 * bary.x + n1 * bary.y + n2 * bary.z);
	p.n_s = glm::normalize(mul_dir_w2o(nrm, w2o));
	p.uv = uv0 * bary.x + uv1 * bary.y + uv2 * bary.z;
	const vec3 e0 = v2 - v0;
	const vec3 e1 = v1 - v0;
	const vec3 e0t = vec3(o2w * vec4(e0, 0.0f));
	const vec3 e1t = vec3(o2w * vec4(e1, 0.0f));
	p.n_g = glm::normalize(mul_dir_w2o(glm::cross(e0, e1), w2o));
	p.material_idx = pinfo.material_index;
	p.triangle_idx = prim;
	p.instance_idx = mesh;
	p.area = 0.5f * glm::length(glm::cross(e0t, e1t));
	return p;
}

struct TriangleRecord {
	vec3 pos, n_s;
	float triangle_pdf;
};

// commons.glsl:112-149 (Q7: w = 1 on edge vectors and on the normal)
inline TriangleRecord sample_triangle(const orc_scene& s, const lmb_prim_mesh_info& pinfo, const vec2& rands, uint triangle_idx,
									  const mat4& world_matrix, const mat4& inv_world, vec2& uv) {
	TriangleRecord r;
	const uint index_offset = pinfo.index_offset + 3 * triangle_idx;
	const lmb_vertex& a0 = s.sd.vertices[s.sd.indices[index_offset + 0] + pinfo.vertex_offset];
	const lmb_vertex& a1 = s.sd.vertices[s.sd.indices[index_offset + 1] + pinfo.vertex_offset];
	const lmb_vertex& a2 = s.sd.vertices[s.sd.indices[index_offset + 2] + pinfo.vertex_offset];
	const vec3 v0 = v3(a0.pos), v1 = v3(a1.pos), v2 = v3(a2.pos);
	const vec3 n0 = v3(a0.normal), n1 = v3(a1.normal), n2 = v3(a2.normal);
	const mat4 inv_tr_mat = glm::transpose(inv_world);
	const float sq = std::sqrt(rands.x);
	uv = vec2(1 - sq, rands.y * sq);
	const vec3 bary(1.0f - uv.x - uv.y, uv.x, uv.y);
	const vec4 etmp0 = world_matrix * vec4(v1 - v0, 1.0f);
	const vec4 etmp1 = world_matrix * vec4(v2 - v0, 1.0f);
	const vec3 pos = v0 * bary.x + v1 * bary.y + v2 * bary.z;
	const vec3 nrm = glm::normalize(n0 * bary.x + n1 * bary.y + n2 * bary.z);
	const vec4 world_pos = world_matrix * vec4(pos, 1.0f);
	r.n_s = glm::normalize(vec3(inv_tr_mat * vec4(nrm, 1.0f)));
	r.triangle_pdf = 2.0f / glm::length(glm::cross(vec3(etmp0), vec3(etmp1)));
	r.pos = vec3(world_pos);
	return r;
}

struct LightSample {
	vec3 Le{0};
	vec3 wi{0};
	float wi_len = 0, pdf_w = 0, pdf_a = 0, cos_from_light = 0;
	uint light_idx = 0, flags = 0, triangle_idx = 0, instance_idx = 0;
	vec2 bary{0};
	vec3 n{0}, pos{0};  // `out vec3 n, out vec3 pos` of commons.glsl:224-226 (read by bdpt_commons.glsl only)
};

// commons.glsl:224-300
inline LightSample sample_light_Li(const orc_scene& s, const vec4& rands, const vec3& p, int num_lights) {
	LightSample o;
	o.light_idx = (uint)(rands.x * (float)num_lights);
	// a scene without lights has no light buffer in the reference (LumenScene.cpp:134); both sides read one all-zero Light
	static const lmb_light zero_light{};
	const lmb_light& light = s.sd.n_lights ? s.sd.lights[o.light_idx] : zero_light;
	const uint type = light.light_flags & 0x7u;
	o.flags = light.light_flags;
	switch (type) {
		case LMB_LIGHT_AREA: {
			const lmb_prim_mesh_info& pinfo = s.sd.prim_infos[light.prim_mesh_idx];
			const uint material_idx = pinfo.material_index;
			o.triangle_idx = (uint)(rands.y * (float)light.num_triangles);
			// light.world_matrix == prim mesh world matrix (LumenScene.cpp:118); its inverse is the mesh's
			const mat4 wm = m4(light.world_matrix);
			const mat4 iwm = m4(s.sd.inv_world_matrices + 16 * light.prim_mesh_idx);
			const TriangleRecord rec = sample_triangle(s, pinfo, vec2(rands.z, rands.w), o.triangle_idx, wm, iwm, o.bary);
			const lmb_material light_mat = load_material(s, material_idx, o.bary);
			o.wi = rec.pos - p;
			const float wi_len_sqr = glm::dot(o.wi, o.wi);
			o.wi_len = std::sqrt(wi_len_sqr);
			o.wi /= o.wi_len;
			o.cos_from_light = std::fabs(glm::dot(rec.n_s, -o.wi));
			o.Le = v3(light_mat.emissive_factor);
			o.pdf_a = rec.triangle_pdf;
			o.pdf_w = o.pdf_a * wi_len_sqr / o.cos_from_light;
			o.instance_idx = light.prim_mesh_idx;
			o.n = rec.n_s;
			o.pos = rec.pos;
		} break;
		case LMB_LIGHT_SPOT: {
			o.wi = v3(light.pos) - p;
			const float wi_len_sqr = glm::dot(o.wi, o.wi);
			o.wi_len = std::sqrt(wi_len_sqr);
			o.wi /= o.wi_len;
			const vec3 light_dir = glm::normalize(v3(light.to) - v3(light.pos));
			o.cos_from_light = glm::dot(-o.wi, light_dir);
			const float cos_width = g_cos(PI / 6);
			const float cos_faloff = g_cos(25 * PI / 180);
			float faloff;
			if (o.cos_from_light < cos_width) {
				faloff = 0;
			} else if (o.cos_from_light >= cos_faloff) {
				faloff = 1;
			} else {
				const float d = (o.cos_from_light - cos_width) / (cos_faloff - cos_width);
				faloff = (d * d) * (d * d);
			}
			o.pdf_a = 1;
			o.pdf_w = wi_len_sqr;
			o.Le = v3(light.L) * faloff;
			o.n = -o.wi;
			o.pos = v3(light.pos);
		} break;
		case LMB_LIGHT_DIRECTIONAL: {
			const vec3 dir = glm::normalize(v3(light.pos) - v3(light.to));
			const vec3 light_p = p + dir * (2 * light.world_radius);
			o.wi = light_p - p;
			o.wi_len = glm::length(o.wi);
			o.wi /= o.wi_len;
			o.pdf_a = 1;
			o.pdf_w = 1;
			o.Le = v3(light.L);
			o.cos_from_light = 1.0f;
			o.n = -o.wi;
			o.pos = light_p;
		} break;
		default:
			break;
	}
	return o;
}

// commons.glsl:156-168
inline vec3 shade_atmosphere(const orc_scene& s, uint dir_light_idx, const vec3& sky_col, const vec3& ray_origin, const vec3& ray_dir,
							 float ray_length) {
	if (dir_light_idx == 0xFFFFFFFFu) return sky_col;
	const lmb_light& light = s.sd.lights[dir_light_idx];
	const vec3 light_dir = -glm::normalize(v3(light.to) - v3(light.pos));
	const vec2 planet_isect = atmo::planet_intersection(ray_origin, ray_dir);
	if (planet_isect.x > 0) ray_length = glm::min(ray_length, planet_isect.x);
	return atmo::integrate_scattering(ray_origin, ray_dir, ray_length, light_dir, v3(light.L));
}

constexpr float T_MIN = 0.001f;   // path.rgen:19
constexpr float T_MAX = 10000.0f; // path.rgen:20

// pt_commons.glsl:3-42
inline vec3 uniform_sample_light(const orc_scene& s, const lmb_pc_path& pc, uvec4& seed, const lmb_material& mat, const HitPayload& cur,
								 bool side, const vec3& n_s, const vec3& wo, Counters& c) {
	const vec3 pos = cur.pos;
	vec3 res(0);
	const vec4 r4 = rand4(seed);
	const LightSample ls = sample_light_Li(s, r4, pos, pc.num_lights);
	const vec3 p = offset_ray2(pos, n_s);
	float bsdf_pdf;
	float cos_x = glm::dot(n_s, ls.wi);
	vec3 f = eval_bsdf(n_s, wo, mat, 1, side, ls.wi, bsdf_pdf);
	c.shadow++;
	const Hit sh = trace<true>(s.bvh, p, ls.wi, 0.0f, ls.wi_len - EPS, &c.ts);
	const bool visible = sh.prim == 0xFFFFFFFFu;
	if (visible && ls.pdf_w > 0) {
		const float mis_weight = ((ls.flags >> 5) & 1u) ? 1.0f : 1.0f / (1.0f + bsdf_pdf / ls.pdf_w);
		res += mis_weight * f * std::fabs(cos_x) * ls.Le / ls.pdf_w;
	}
	if ((ls.flags & 0x7u) == LMB_LIGHT_AREA) {
		const vec3 r3 = rand3(seed);
		const BsdfSample bs = sample_bsdf(n_s, wo, mat, 1, side, r3);
		f = bs.f;
		bsdf_pdf = bs.pdf;
		cos_x = bs.cos_theta;
		if (bsdf_pdf != 0) {
			c.probe++;
			const Hit ph = trace<false>(s.bvh, p, bs.wi, T_MIN, T_MAX, &c.ts);
			// ray.rmiss only writes material_idx, so after a miss the payload still holds the record of the surface
			// being shaded (same rayPayloadEXT location 0); the reference compares those stale ids. Reproduced.
			const HitPayload pl = (ph.prim != 0xFFFFFFFFu) ? build_hit(s, ph) : cur;
			if (pl.triangle_idx == ls.triangle_idx && pl.instance_idx == ls.instance_idx) {
				const float wi_len = glm::length(pl.pos - pos);
				const float g = std::fabs(glm::dot(pl.n_s, -bs.wi)) / (wi_len * wi_len);
				const float mis_weight = 1.0f / (1 + ls.pdf_a / (g * bsdf_pdf));
				res += f * mis_weight * std::fabs(cos_x) * ls.Le / bsdf_pdf;
			}
		}
	}
	return res;
}

// path.rgen:26-104: radiance of one (pixel, frame) sample
inline vec3 trace_pixel(const orc_scene& s, const lmb_pc_path& pc, const lmb_scene_ubo& ubo, uint px, uint py, uint frame, Counters& c) {
	uvec4 seed(px, py, frame, 0);
	const mat4 inv_view = m4(ubo.inv_view), inv_proj = m4(ubo.inv_projection);
	const vec2 pixel = vec2((float)px, (float)py) + vec2(0.5f);
	const float j0 = rand1(seed);
	const float j1 = rand1(seed);
	const vec2 rands = vec2(j0, j1) - 0.5f;
	const vec2 in_uv = (pixel + rands) / vec2((float)pc.size_x, (float)pc.size_y);
	const vec2 d = in_uv * 2.0f - 1.0f;
	vec3 origin = vec3(inv_view * vec4(0, 0, 0, 1));
	const vec4 target = inv_proj * vec4(d.x, d.y, 1, 1);
	vec3 direction = vec3(inv_view * vec4(glm::normalize(vec3(target)), 0));  // commons.glsl:30-33

	vec3 col(0);
	bool last_specular = false;
	vec3 throughput(1);
	const vec3 sky_col = v3(pc.sky_col);
	for (int depth = 0;; depth++) {
		c.closest++;
		const Hit h = trace<false>(s.bvh, origin, direction, T_MIN, T_MAX, &c.ts);
		if (h.prim == 0xFFFFFFFFu) {
			if (depth > 0 || pc.direct_lighting == 1) {
				col += throughput * shade_atmosphere(s, pc.dir_light_idx, sky_col, origin, direction, T_MAX);
			}
			break;
		}
		const HitPayload payload = build_hit(s, h);
		const lmb_material hit_mat = load_material(s, payload.material_idx, payload.uv);
		if ((depth == 0 && pc.direct_lighting == 1) || last_specular) {
			col += throughput * v3(hit_mat.emissive_factor);
		}
		if (depth >= pc.max_depth - 1) break;
		const vec3 wo = -direction;
		vec3 n_s = payload.n_s;
		bool side = true;
		vec3 n_g = payload.n_g;
		if (glm::dot(payload.n_g, wo) < 0.0f) n_g = -n_g;
		if (glm::dot(n_g, payload.n_s) < 0) {
			n_s = -n_s;
			side = false;
		}
		origin = offset_ray(payload.pos, n_g);
		last_specular = (hit_mat.bsdf_props & LMB_FLAG_SPECULAR) != 0;
		if (!last_specular) {
			const float light_pick_pdf = 1.0f / (float)pc.light_triangle_count;
			if (depth > 0 || pc.direct_lighting == 1) {
				col += throughput * uniform_sample_light(s, pc, seed, hit_mat, payload, side, n_s, wo, c) / light_pick_pdf;
			}
		}
		const vec3 r3 = rand3(seed);
		const BsdfSample bs = sample_bsdf(n_s, wo, hit_mat, 1, side, r3);
		direction = bs.wi;
		if (bs.pdf == 0) break;
		throughput *= bs.f * std::fabs(bs.cos_theta) / bs.pdf;
		float rr_scale = 1.0f;
		if (has_prop(hit_mat.bsdf_props, LMB_FLAG_TRANSMISSION)) {
			rr_scale *= side ? 1.0f / hit_mat.ior : hit_mat.ior;
		}
		if (depth > 3) {  // RR_MIN_DEPTH, path.rgen:21,93
			const float rr_prob = glm::min(0.95f, luminance(throughput) * rr_scale);
			if (rr_prob == 0 || rr_prob < rand1(seed))
				break;
			else
				throughput /= rr_prob;
		}
	}
	return col;
}

#include "bdpt.h"

void add_counters(orc_stats* st, const Counters& c) {
	if (!st) return;
	st->rays_closest += c.closest;
	st->rays_shadow += c.shadow;
	st->rays_probe += c.probe;
	st->nodes_visited += c.ts.nodes;
	st->tris_tested += c.ts.tris;
	st->nan_pixels += c.nan_px;
}

int pick_threads(int n) { return n > 0 ? n : omp_get_max_threads(); }

}  // namespace

extern "C" {

int orc_max_threads(void) { return omp_get_max_threads(); }

int orc_scene_create(const lmb_scene_desc* sd, orc_scene** out) {
	if (!sd || !out) return -1;
	orc_scene* s = new orc_scene();
	s->sd = *sd;
	lbvh_build(*sd, s->bvh);
	for (int i = 0; i < 256; i++) {
		const double c = i / 255.0;
		s->srgb_lut[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
	}
	*out = s;
	return 0;
}
void orc_scene_destroy(orc_scene* s) { delete s; }

uint32_t orc_lbvh_num_tris(const orc_scene* s) { return s->bvh.n_tris; }
const uint32_t* orc_lbvh_left(const orc_scene* s) { return s->bvh.left.data(); }
const uint32_t* orc_lbvh_right(const orc_scene* s) { return s->bvh.right.data(); }
const uint32_t* orc_lbvh_parent(const orc_scene* s) { return s->bvh.parent.data(); }
const uint32_t* orc_lbvh_leaf_prim(const orc_scene* s) { return s->bvh.leaf_prim.data(); }
const uint32_t* orc_lbvh_morton(const orc_scene* s) { return s->bvh.morton.data(); }
const uint64_t* orc_lbvh_keys(const orc_scene* s) { return s->bvh.sorted_keys.data(); }
const float* orc_lbvh_aabb(const orc_scene* s) { return s->bvh.aabb.data(); }

int orc_render_frame_raw(const orc_scene* s, const lmb_pc_path* pc, const lmb_scene_ubo* ubo, uint32_t frame, float* rgb, orc_stats* stats,
						 int n_threads) {
	const uint W = pc->size_x, H = pc->size_y;
	const int nt = pick_threads(n_threads);
	const auto t0 = std::chrono::steady_clock::now();
	Counters total;
#pragma omp parallel num_threads(nt)
	{
		Counters c;
#pragma omp for schedule(dynamic, 1)
		for (int y = 0; y < (int)H; y++) {
			for (uint x = 0; x < W; x++) {
				const vec3 col = trace_pixel(*s, *pc, *ubo, x, (uint)y, frame, c);
				float* o = rgb + 3 * ((size_t)y * W + x);
				o[0] = col.x, o[1] = col.y, o[2] = col.z;
			}
		}
#pragma omp critical
		{
			total.closest += c.closest, total.shadow += c.shadow, total.probe += c.probe;
			total.ts.nodes += c.ts.nodes, total.ts.tris += c.ts.tris;
		}
	}
	if (stats) {
		add_counters(stats, total);
		stats->seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		stats->threads = nt;
	}
	return 0;
}

// Row subset of orc_render / orc_render_frame_raw (the pixel shards of lmb_set_pixel_shard): rows row_first, row_first + stride, ...
static uint32_t g_row_first = 0, g_row_stride = 1;
void orc_set_row_shard(uint32_t row_first, uint32_t row_stride) { g_row_first = row_first, g_row_stride = row_stride ? row_stride : 1; }

int orc_render(const orc_scene* s, const lmb_pc_path* pc, const lmb_scene_ubo* ubo, uint32_t first_frame, uint32_t n_frames, float* rgba,
			   orc_stats* stats, int n_threads) {
	const uint W = pc->size_x, H = pc->size_y;
	const int nt = pick_threads(n_threads);
	const auto t0 = std::chrono::steady_clock::now();
	Counters total;
#pragma omp parallel num_threads(nt)
	{
		Counters c;
		for (uint32_t frame = first_frame; frame < first_frame + n_frames; frame++) {
#pragma omp for schedule(dynamic, 1)
			for (int y = 0; y < (int)H; y++) {
				if ((uint32_t)y % g_row_stride != g_row_first) continue;
				for (uint x = 0; x < W; x++) {
					const vec3 col = trace_pixel(*s, *pc, *ubo, x, (uint)y, frame, c);
					// path.rgen:102-112
					if (g_isnan(luminance(col))) {
						c.nan_px++;
						continue;
					}
					float* o = rgba + 4 * ((size_t)y * W + x);
					if (frame > 0) {
						const float w = 1.0f / float(frame + 1);
						const vec3 old_col(o[0], o[1], o[2]);
						const vec3 m = glm::mix(old_col, col, w);
						o[0] = m.x, o[1] = m.y, o[2] = m.z, o[3] = 1.0f;
					} else {
						o[0] = col.x, o[1] = col.y, o[2] = col.z, o[3] = 1.0f;
					}
				}
			}
		}
#pragma omp critical
		{
			total.closest += c.closest, total.shadow += c.shadow, total.probe += c.probe, total.nan_px += c.nan_px;
			total.ts.nodes += c.ts.nodes, total.ts.tris += c.ts.tris;
		}
	}
	if (stats) {
		add_counters(stats, total);
		stats->seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		stats->threads = nt;
	}
	return 0;
}

// bdpt.rgen:39-89 for one frame, split in the two passes of quirk B2 (oracle/bdpt.h): col_rgb = the pixel's own strategies,
// splat_rgb = the light-tracer image of the same frame (splats applied in source-pixel order, rows then columns).
int orc_render_bdpt_frame_raw(const orc_scene* s, const lmb_pc_bdpt* pc, const lmb_scene_ubo* ubo, uint32_t frame, float* col_rgb,
							  float* splat_rgb, orc_stats* stats, int n_threads) {
	if (!s || !pc || !ubo || !col_rgb || !splat_rgb) return -1;
	if (pc->max_depth < 1 || pc->max_depth + 1 > BDPT_MAX_VERTS) return -2;
	const uint W = pc->size_x, H = pc->size_y;
	const int nt = pick_threads(n_threads);
	const auto t0 = std::chrono::steady_clock::now();
	struct SplatRec {
		uint x, y;
		vec3 c;
	};
	std::vector<std::vector<SplatRec>> rows(H);
	Counters total;
#pragma omp parallel num_threads(nt)
	{
		Counters c;
#pragma omp for schedule(dynamic, 1)
		for (int y = 0; y < (int)H; y++) {
			for (uint x = 0; x < W; x++) {
				const vec3 col = bdpt_pixel(*s, *pc, *ubo, x, (uint)y, frame, c, [&](uint cx, uint cy, const vec3& v) { rows[y].push_back({cx, cy, v}); });
				float* o = col_rgb + 3 * ((size_t)y * W + x);
				o[0] = col.x, o[1] = col.y, o[2] = col.z;
			}
		}
#pragma omp critical
		{
			total.closest += c.closest, total.shadow += c.shadow;
			total.ts.nodes += c.ts.nodes, total.ts.tris += c.ts.tris;
		}
	}
	std::memset(splat_rgb, 0, (size_t)W * H * 3 * sizeof(float));
	for (uint y = 0; y < H; y++)
		for (const SplatRec& r : rows[y]) {
			float* o = splat_rgb + 3 * ((size_t)r.y * W + r.x);
			o[0] += r.c.x, o[1] += r.c.y, o[2] += r.c.z;
		}
	if (stats) {
		add_counters(stats, total);
		stats->seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		stats->threads = nt;
	}
	return 0;
}

void orc_bdpt_set_only_s(int s) { g_bdpt_only_s = s; }
void orc_bdpt_set_check_restore(int on) { g_bdpt_check_restore = on; }
long long orc_bdpt_restore_violations(void) { return g_bdpt_restore_violations; }

int orc_render_bdpt(const orc_scene* s, const lmb_pc_bdpt* pc, const lmb_scene_ubo* ubo, uint32_t first_frame, uint32_t n_frames, float* rgba,
					orc_stats* stats, int n_threads) {
	if (!s || !pc || !ubo || !rgba) return -1;
	const uint W = pc->size_x, H = pc->size_y;
	std::vector<float> col((size_t)W * H * 3), splat((size_t)W * H * 3);
	for (uint32_t frame = first_frame; frame < first_frame + n_frames; frame++) {
		const int rc = orc_render_bdpt_frame_raw(s, pc, ubo, frame, col.data(), splat.data(), stats, n_threads);
		if (rc) return rc;
		for (size_t i = 0; i < (size_t)W * H; i++) {  // bdpt.rgen:76-89
			vec3 c(col[3 * i], col[3 * i + 1], col[3 * i + 2]);
			c += vec3(splat[3 * i], splat[3 * i + 1], splat[3 * i + 2]);
			if (g_isnan(luminance(c))) {
				if (stats) stats->nan_pixels++;
				continue;
			}
			float* o = rgba + 4 * i;
			if (frame > 0) {
				const float w = 1.0f / float(frame + 1);
				const vec3 m = glm::mix(vec3(o[0], o[1], o[2]), c, w);
				o[0] = m.x, o[1] = m.y, o[2] = m.z, o[3] = 1.0f;
			} else {
				o[0] = c.x, o[1] = c.y, o[2] = c.z, o[3] = 1.0f;
			}
		}
	}
	return 0;
}

int orc_trace_closest(const orc_scene* s, const float* rays, uint32_t n, orc_hit* hits, orc_stats* stats, int n_threads) {
	const int nt = pick_threads(n_threads);
	uint64_t nodes = 0, tris = 0;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1024) reduction(+ : nodes, tris)
	for (int64_t i = 0; i < (int64_t)n; i++) {
		const float* r = rays + 8 * i;
		TraceStats ts;
		const Hit h = trace<false>(s->bvh, vec3(r[0], r[1], r[2]), vec3(r[4], r[5], r[6]), r[3], r[7], &ts);
		hits[i] = {h.t, h.b1, h.b2, h.prim};
		nodes += ts.nodes, tris += ts.tris;
	}
	if (stats) stats->nodes_visited += nodes, stats->tris_tested += tris, stats->rays_closest += n;
	return 0;
}

// One ray, no OpenMP: the intersection callback of the mechanically translated reference shaders (oracle/glslref), whose
// traceRayEXT is outside the shader source. mesh / local = gl_InstanceCustomIndexEXT / gl_PrimitiveID of the hit.
int orc_trace1(const orc_scene* s, const float* r, int any_hit, orc_hit* hit, uint32_t* mesh, uint32_t* local) {
	const vec3 o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
	const Hit h = any_hit ? trace<true>(s->bvh, o, d, r[3], r[7], nullptr) : trace<false>(s->bvh, o, d, r[3], r[7], nullptr);
	*hit = {h.t, h.b1, h.b2, h.prim};
	if (h.prim != 0xFFFFFFFFu) *mesh = s->bvh.tri_mesh[h.prim], *local = s->bvh.tri_local[h.prim];
	return h.prim != 0xFFFFFFFFu;
}

int orc_trace_closest_brute(const orc_scene* s, const float* rays, uint32_t n, orc_hit* hits, int n_threads) {
	const int nt = pick_threads(n_threads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 64)
	for (int64_t i = 0; i < (int64_t)n; i++) {
		const float* r = rays + 8 * i;
		const Hit h = trace_brute(s->bvh, vec3(r[0], r[1], r[2]), vec3(r[4], r[5], r[6]), r[3], r[7]);
		hits[i] = {h.t, h.b1, h.b2, h.prim};
	}
	return 0;
}

int orc_trace_any(const orc_scene* s, const float* rays, uint32_t n, uint8_t* occluded, orc_stats* stats, int n_threads) {
	const int nt = pick_threads(n_threads);
	uint64_t nodes = 0, tris = 0;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1024) reduction(+ : nodes, tris)
	for (int64_t i = 0; i < (int64_t)n; i++) {
		const float* r = rays + 8 * i;
		TraceStats ts;
		const Hit h = trace<true>(s->bvh, vec3(r[0], r[1], r[2]), vec3(r[4], r[5], r[6]), r[3], r[7], &ts);
		occluded[i] = h.prim != 0xFFFFFFFFu;
		nodes += ts.nodes, tris += ts.tris;
	}
	if (stats) stats->nodes_visited += nodes, stats->tris_tested += tris, stats->rays_shadow += n;
	return 0;
}

// ------------------------------------------------------------------------------------------------ KAT probes
void orc_kat_pcg4d(const uint32_t* in4, uint32_t n, uint32_t* out4) {
	for (uint32_t i = 0; i < n; i++) {
		const uvec4 r = pcg4d(uvec4(in4[4 * i], in4[4 * i + 1], in4[4 * i + 2], in4[4 * i + 3]));
		out4[4 * i] = r.x, out4[4 * i + 1] = r.y, out4[4 * i + 2] = r.z, out4[4 * i + 3] = r.w;
	}
}
void orc_kat_rand(const uint32_t* seed4, uint32_t n, uint32_t draws, float* out) {
	for (uint32_t i = 0; i < n; i++) {
		uvec4 s(seed4[4 * i], seed4[4 * i + 1], seed4[4 * i + 2], seed4[4 * i + 3]);
		for (uint32_t k = 0; k < draws; k++) out[(size_t)i * draws + k] = rand1(s);
	}
}
void orc_kat_detmath(const float* x, const float* y, uint32_t n, float* out_sin, float* out_cos, float* out_exp, float* out_pow) {
	for (uint32_t i = 0; i < n; i++) {
		lmb_sincosf(x[i], &out_sin[i], &out_cos[i]);
		out_exp[i] = lmb_expf(x[i]);
		out_pow[i] = lmb_powf(std::fabs(x[i]), y[i]);
	}
}
void orc_kat_offset_ray(const float* p3, const float* n3, uint32_t n, float* out3, float* out3_b) {
	for (uint32_t i = 0; i < n; i++) {
		const vec3 a = offset_ray(v3(p3 + 3 * i), v3(n3 + 3 * i));
		const vec3 b = offset_ray2(v3(p3 + 3 * i), v3(n3 + 3 * i));
		for (int k = 0; k < 3; k++) out3[3 * i + k] = a[k], out3_b[3 * i + k] = b[k];
	}
}
void orc_kat_sample_bsdf(const lmb_material* mat, const float* n_s3, const float* wo3, const float* rands3, const uint8_t* side, uint32_t n,
						 float* out8) {
	for (uint32_t i = 0; i < n; i++) {
		const BsdfSample s = sample_bsdf(v3(n_s3 + 3 * i), v3(wo3 + 3 * i), *mat, 1, side[i] != 0, v3(rands3 + 3 * i));
		float* o = out8 + 8 * (size_t)i;
		o[0] = s.f.x, o[1] = s.f.y, o[2] = s.f.z, o[3] = s.wi.x, o[4] = s.wi.y, o[5] = s.wi.z, o[6] = s.pdf, o[7] = s.cos_theta;
	}
}
void orc_kat_eval_bsdf(const lmb_material* mat, const float* n_s3, const float* wo3, const float* wi3, const uint8_t* side, uint32_t n,
					   float* out4) {
	for (uint32_t i = 0; i < n; i++) {
		float pdf;
		const vec3 f = eval_bsdf(v3(n_s3 + 3 * i), v3(wo3 + 3 * i), *mat, 1, side[i] != 0, v3(wi3 + 3 * i), pdf);
		float* o = out4 + 4 * (size_t)i;
		o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = pdf;
	}
}
void orc_kat_bsdf_pdf(const lmb_material* mat, const float* n_s3, const float* wo3, const float* wi3, const uint8_t* side, uint32_t n, float* out) {
	for (uint32_t i = 0; i < n; i++) out[i] = bsdf_pdf(*mat, v3(n_s3 + 3 * i), v3(wo3 + 3 * i), v3(wi3 + 3 * i), side[i] != 0);
}
void orc_kat_atmosphere(const float* origin3, const float* dir3, const float* light_dir3, const float* light_L3, uint32_t n, float* out3) {
#pragma omp parallel for schedule(dynamic, 16)
	for (int64_t i = 0; i < (int64_t)n; i++) {
		const vec3 o = v3(origin3 + 3 * i), d = v3(dir3 + 3 * i);
		float ray_length = T_MAX;
		const vec2 pi = atmo::planet_intersection(o, d);
		if (pi.x > 0) ray_length = glm::min(ray_length, pi.x);
		const vec3 r = atmo::integrate_scattering(o, d, ray_length, v3(light_dir3), v3(light_L3));
		out3[3 * i] = r.x, out3[3 * i + 1] = r.y, out3[3 * i + 2] = r.z;
	}
}
void orc_kat_sample_light(const orc_scene* s, int32_t num_lights, const float* rands4, const float* p3, uint32_t n, float* out16) {
	for (uint32_t i = 0; i < n; i++) {
		const float* r = rands4 + 4 * i;
		const LightSample l = sample_light_Li(*s, vec4(r[0], r[1], r[2], r[3]), v3(p3 + 3 * i), num_lights);
		float* o = out16 + 16 * (size_t)i;
		o[0] = l.Le.x, o[1] = l.Le.y, o[2] = l.Le.z, o[3] = l.wi.x, o[4] = l.wi.y, o[5] = l.wi.z;
		o[6] = l.wi_len, o[7] = l.pdf_w, o[8] = l.pdf_a, o[9] = l.cos_from_light;
		o[10] = (float)l.light_idx, o[11] = (float)l.flags, o[12] = (float)l.triangle_idx, o[13] = (float)l.instance_idx;
		o[14] = l.bary.x, o[15] = l.bary.y;
	}
}
void orc_kat_light_Le(const orc_scene* s, int32_t num_lights, int32_t total_light, const float* rands6, uint32_t n, float* out16) {
	for (uint32_t i = 0; i < n; i++) {
		const float* r = rands6 + 6 * i;
		const LightEmission l = sample_light_Le(*s, vec4(r[0], r[1], r[2], r[3]), vec2(r[4], r[5]), num_lights, total_light);
		float* o = out16 + 16 * (size_t)i;
		o[0] = l.L.x, o[1] = l.L.y, o[2] = l.L.z, o[3] = l.pos.x, o[4] = l.pos.y, o[5] = l.pos.z, o[6] = l.wi.x, o[7] = l.wi.y, o[8] = l.wi.z;
		o[9] = l.n.x, o[10] = l.n.y, o[11] = l.n.z, o[12] = l.cos_from_light, o[13] = l.pdf_pos_a, o[14] = l.pdf_dir_w, o[15] = (float)l.flags;
	}
}
void orc_kat_texture(const orc_scene* s, uint32_t tex, const float* uv2, uint32_t n, float* out3) {
	for (uint32_t i = 0; i < n; i++) {
		const vec3 c = sample_texture(*s, tex, vec2(uv2[2 * i], uv2[2 * i + 1]));
		out3[3 * i] = c.x, out3[3 * i + 1] = c.y, out3[3 * i + 2] = c.z;
	}
}

// ------------------------------------------------------------------------------------------------ RMSE
// src/shaders/rmse/calc_rmse.comp:21-42, reduce_rmse.comp:22-49, output_rmse.comp:21-24 and the dispatch loop in
// src/RayTracer/RayTracer.cpp:215-241. Frozen choices where GLSL is undefined: inactive lanes of the tail workgroup
// contribute nothing, shared slots of wholly inactive subgroups read 0, subgroup sums run in lane order.
float orc_rmse_literal(const float* a, const float* b, uint32_t n_pixels) {
	const uint32_t WG = 1024, SG = 32;
	std::vector<float> res((n_pixels + WG - 1) / WG);
	for (uint32_t wg = 0; wg < res.size(); wg++) {
		float data[32];
		for (int k = 0; k < 32; k++) data[k] = 0.0f;
		for (uint32_t sg = 0; sg < WG / SG; sg++) {
			float sum = 0.0f;
			bool any = false;
			for (uint32_t l = 0; l < SG; l++) {
				const uint32_t idx = wg * WG + sg * SG + l;
				if (idx >= n_pixels) break;
				const vec3 diff(a[4 * idx] - b[4 * idx], a[4 * idx + 1] - b[4 * idx + 1], a[4 * idx + 2] - b[4 * idx + 2]);
				sum += glm::dot(diff, diff);
				any = true;
			}
			if (any) data[sg] = sum;
		}
		float mn = data[0];
		for (int k = 1; k < 32; k++) mn = std::min(mn, data[k]);  // calc_rmse.comp:37 subgroupMin
		res[wg] = mn;
	}
	size_t live = res.size();
	while (live > 1) {
		const size_t groups = (live + WG - 1) / WG;
		for (size_t g2 = 0; g2 < groups; g2++) {
			float sum = 0.0f;
			for (size_t k = g2 * WG; k < std::min(live, (g2 + 1) * (size_t)WG); k++) sum += res[k];
			res[g2] = sum;
		}
		live = groups;
	}
	return std::sqrt(res.empty() ? 0.0f : res[0]) / ((float)n_pixels * 3.0f);
}
// ------------------------------------------------------------------------------------------------ EXR half conversion
// ImageUtils::save_exr (src/Framework/ImageUtils.cpp:72-76) asks tinyexr for HALF channels; tinyexr converts with
// float_to_half_full (libs/tinyexr.h:898-934): mantissa truncated to 10 bits plus one when the first dropped bit is set
// (round half up in magnitude), NaN -> quiet NaN 0x200, overflow -> inf, fp32 subnormals -> signed zero, half subnormals
// by shifting the mantissa with the hidden bit. Pinned against the real tinyexr in tests/test_post_cpu.py (EXR round trip).
void orc_float_to_half(const float* in, uint32_t n, uint16_t* out) {
	for (uint32_t i = 0; i < n; i++) {
		uint32_t u;
		memcpy(&u, &in[i], 4);
		const uint32_t sign = u >> 31, exponent = (u >> 23) & 0xFFu, mantissa = u & 0x7FFFFFu;
		uint32_t o = 0;  // exponent << 10 | mantissa of the half, carries of the rounding run into the exponent on purpose
		if (exponent == 0) {
			o = 0;
		} else if (exponent == 255) {
			o = (31u << 10) | (mantissa ? 0x200u : 0u);
		} else {
			const int newexp = (int)exponent - 127 + 15;
			if (newexp >= 31) {
				o = 31u << 10;
			} else if (newexp <= 0) {
				if ((14 - newexp) <= 24) {
					const uint32_t mant = mantissa | 0x800000u;
					o = mant >> (14 - newexp);
					if ((mant >> (13 - newexp)) & 1u) o++;
				}
			} else {
				o = ((uint32_t)newexp << 10) | (mantissa >> 13);
				if (mantissa & 0x1000u) o++;
			}
		}
		out[i] = (uint16_t)((sign << 15) | (o & 0x7FFFu));
	}
}

double orc_rmse_true(const float* a, const float* b, uint32_t n_pixels) {
	double acc = 0;
	for (uint32_t i = 0; i < n_pixels; i++)
		for (int k = 0; k < 3; k++) {
			const double d = (double)a[4 * i + k] - (double)b[4 * i + k];
			acc += d * d;
		}
	return std::sqrt(acc / (3.0 * (double)n_pixels));
}

}  // extern "C"
