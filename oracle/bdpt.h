// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl_compat.h). Included by oracle.cpp inside its anonymous namespace.
//
// bdpt.h: CPU restatement of the reference's bidirectional path tracer:
//   src/shaders/integrators/bdpt/bdpt.rgen:27-90            per-pixel driver: sub-path generation, (s, t) connection loop, film
//   src/shaders/integrators/bdpt_commons.glsl:13-641        random walks, calc_mis_weight, bdpt_connect_cam, bdpt_connect
//   src/shaders/commons.glsl:40-111, 217-222, 335-416       light pdfs, uniform_sample_cone, sample_light_Le
//   src/shaders/utils.glsl:156-182                          quaternions, make_coord_system
//   src/RayTracer/BDPT.cpp:55-95                            push constants; both vertex buffers are ZEROED before every frame
//
// What the reference leaves undefined, and what this restatement (and the CUDA path) define instead -- quirks B1-B6:
//   B1 RNG seed = (x, y, frame_num ^ pc.time, 0) (bdpt.rgen:36-37). BDPT.cpp:57 draws pc.time = rand() % UINT_MAX on the host
//      every frame; here `time` is an INPUT (lmb_pc_bdpt.time), so a render is a pure function of its arguments.
//   B2 Light-tracer splats (t == 1): the reference does a non-atomic `tmp_col.d[idx] += splat` from every pixel's invocation
//      and reads + clears its own entry in the SAME dispatch (bdpt.rgen:69-78), so a splat lands in whichever frame happens
//      to read it (or is lost to the race). Defined here: every splat of frame f is added to frame f (two passes). The sum
//      is taken in pixel order x-major here and by float atomics on the GPU: equal up to summation order.
//   B3 Vertices the walk never writes are all-zero (the render graph zeroes both buffers per frame, BDPT.cpp:79-80). The
//      camera walk's "escaped" vertex (bdpt_commons.glsl:137-142) therefore has material_idx 0, and bdpt_connect's s == 0
//      strategy multiplies its throughput by MATERIAL 0's emissive_factor (`mat_idx != -1` is always true). Reproduced.
//   B4 `PathVertex sampled` is uninitialised where no strategy writes it (bdpt_connect_cam never sets pdf_fwd): zero here; the
//      values are written into vertex 0 and restored without being read in between.
//   B5 ivec2(float) of a NaN / out-of-range splat coordinate is undefined in GLSL: such a splat is dropped.
//   B6 reads of light_vtx(s - 2) with s == 1 (bdpt_commons.glsl:332) index the previous pixel's vertices; the value is dead
//      (overwritten or unused for s == 1), so it is not read here.
//   bdpt_generate_camera_subpath writes `light_verts.d[..].mode = 1` (typo for camera_verts, :272); `mode` only matters to
//   sample_bsdf, which receives a literal, so the slip has no effect; reproduced anyway.

struct PathVertex {  // bdpt_commons.h:18-33
	vec3 dir{0}, n_s{0}, pos{0};
	vec2 uv{0};
	vec3 throughput{0};
	uint light_flags = 0, light_idx = 0, material_idx = 0, delta = 0, side = 0, mode = 0;
	float area = 0, pdf_fwd = 0, pdf_rev = 0;
};

constexpr int BDPT_MAX_VERTS = 34;  // max_depth + 1 <= 34

struct Bdpt {
	const orc_scene& s;
	const lmb_pc_bdpt& pc;
	const lmb_scene_ubo& ubo;
	Counters& c;
	uvec4 seed;
	uint screen_size;
	float light_pdf_pos = 0;
	uint rng_after_walks = 0;  // seed.w when the connection loop starts (diagnostic: orc_bdpt_set_check_restore)
	PathVertex light_verts[BDPT_MAX_VERTS];
	PathVertex camera_verts[BDPT_MAX_VERTS];
};

inline bool is_light_finite(uint f) { return ((f >> 4) & 1u) != 0; }  // commons.glsl:42
inline bool is_light_delta(uint f) { return ((f >> 5) & 1u) != 0; }   // commons.glsl:44
inline float uniform_cone_pdf(float cos_max) { return 1.0f / (TWO_PI * (1 - cos_max)); }  // commons.glsl:40
inline bool same_hemisphere(const vec3& wi, const vec3& wo, const vec3& n) { return glm::dot(wi, n) > 0 && glm::dot(wo, n) > 0; }  // bsdf_commons.glsl:24
inline bool is_zero(const vec3& f) { return f.x == 0 && f.y == 0 && f.z == 0; }  // `f == vec3(0)`

// commons.glsl:81-95 (the uint-flags overload). Unknown light type: GLSL falls off the end (undefined) -> 0.
inline float light_pdf(uint light_flags, const vec3& n_s, const vec3& wi) {
	const float cos_width = g_cos(30 * PI / 180);
	switch (light_flags & 0x7u) {
		case LMB_LIGHT_AREA:
			return glm::max(glm::dot(n_s, wi) / PI, 0.0f);
		case LMB_LIGHT_SPOT:
			return uniform_cone_pdf(cos_width);
		case LMB_LIGHT_DIRECTIONAL:
			return 0;
	}
	return 0;
}
// commons.glsl:64-79
inline float light_pdf_a_to_w(uint light_flags, float pdf_a, float wi_len_sqr, float cos_from_light) {
	switch (light_flags & 0x7u) {
		case LMB_LIGHT_AREA:
			return pdf_a * wi_len_sqr / cos_from_light;
		case LMB_LIGHT_SPOT:
			return wi_len_sqr / cos_from_light;
		case LMB_LIGHT_DIRECTIONAL:
			return 1;
	}
	return 0;
}
// utils.glsl:157-173. normalize(vec4) written out: v * (1 / sqrt((x*x + y*y) + (z*z + w*w))) (glm's dot(vec4) pairing).
inline vec4 to_local_quat(const vec3& v) {
	if (v.z < -0.99999f) return vec4(1, 0, 0, 0);
	const vec4 q(v.y, -v.x, 0.0f, 1.0f + v.z);
	const float inv = 1.0f / std::sqrt((q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w));
	return vec4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
}
inline vec4 invert_quat(const vec4& q) { return vec4(-q.x, -q.y, -q.z, q.w); }
inline vec3 rot_quat(const vec4& q, const vec3& v) {
	const vec3 q_axis(q.x, q.y, q.z);
	return (2.0f * glm::dot(q_axis, v)) * q_axis + (q.w * q.w - glm::dot(q_axis, q_axis)) * v + (2.0f * q.w) * glm::cross(q_axis, v);
}
// utils.glsl:175-182
inline void make_coord_system(const vec3& v1, vec3& v2, vec3& v3o) {
	if (std::fabs(v1.x) > std::fabs(v1.y))
		v2 = glm::normalize(vec3(-v1.z, 0, v1.x));
	else
		v2 = glm::normalize(vec3(0, v1.z, -v1.y));
	v3o = glm::cross(v1, v2);
}
// commons.glsl:217-222
inline vec3 uniform_sample_cone(const vec2& uv, float cos_max) {
	const float cos_theta = (1.0f - uv.x) + uv.x * cos_max;
	const float sin_theta = std::sqrt(1 - cos_theta * cos_theta);
	const float phi = uv.y * TWO_PI;
	return vec3(g_cos(phi) * sin_theta, g_sin(phi) * sin_theta, cos_theta);
}
// sampling_commons.glsl:147-161 (concentric-disk mode), explicit normal
inline vec3 sample_hemisphere_n(const vec2& xi, const vec3& n) {
	vec3 T, B;
	branchless_onb(n, T, B);
	const vec2 d = concentric_sample_disk(xi);
	const float z = std::sqrt(glm::max(0.f, 1.f - glm::dot(d, d)));
	return to_world(vec3(d, z), T, B, n);
}

struct LightEmission {
	vec3 L{0}, pos{0}, wi{0}, n{0};
	float cos_from_light = 0, pdf_pos_a = 0, pdf_dir_w = 0;
	uint flags = 0;
};
// commons.glsl:335-406
inline LightEmission sample_light_Le(const orc_scene& s, const vec4& rands_pos, const vec2& rands_dir, int num_lights, int total_light) {
	LightEmission o;
	const uint light_idx = (uint)(rands_pos.x * (float)num_lights);
	static const lmb_light zero_light{};
	const lmb_light& light = s.sd.n_lights ? s.sd.lights[light_idx] : zero_light;
	o.flags = light.light_flags;
	switch (light.light_flags & 0x7u) {
		case LMB_LIGHT_AREA: {
			const lmb_prim_mesh_info& pinfo = s.sd.prim_infos[light.prim_mesh_idx];
			const uint triangle_idx = (uint)(rands_pos.y * (float)light.num_triangles);
			vec2 bary;
			const TriangleRecord rec = sample_triangle(s, pinfo, vec2(rands_pos.z, rands_pos.w), triangle_idx, m4(light.world_matrix),
													   m4(s.sd.inv_world_matrices + 16 * light.prim_mesh_idx), bary);
			const lmb_material light_mat = load_material(s, pinfo.material_index, bary);
			o.pos = rec.pos;
			o.wi = sample_hemisphere_n(rands_dir, rec.n_s);
			o.L = v3(light_mat.emissive_factor);
			o.cos_from_light = glm::max(glm::dot(rec.n_s, o.wi), 0.0f);
			o.pdf_pos_a = rec.triangle_pdf;
			o.pdf_dir_w = glm::dot(o.wi, rec.n_s) / PI;
			o.n = rec.n_s;
		} break;
		case LMB_LIGHT_SPOT: {
			const float cos_width = g_cos(30 * PI / 180);
			const float cos_faloff = g_cos(25 * PI / 180);
			const vec3 light_dir = glm::normalize(v3(light.to) - v3(light.pos));
			const vec4 local_quat = to_local_quat(light_dir);
			o.wi = rot_quat(invert_quat(local_quat), uniform_sample_cone(rands_dir, cos_width));
			o.pos = v3(light.pos);
			o.cos_from_light = glm::dot(o.wi, light_dir);
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
			const vec3 dir = -glm::normalize(v3(light.to) - v3(light.pos));
			vec3 v1, v2;
			make_coord_system(dir, v1, v2);
			const vec2 uv = concentric_sample_disk(rands_dir);
			const vec3 l_pos = v3(light.world_center) + light.world_radius * (uv.x * v1 + uv.y * v2);
			o.pos = l_pos + dir * light.world_radius;
			o.wi = -dir;
			o.L = v3(light.L);
			o.pdf_pos_a = 1.0f / (PI * light.world_radius * light.world_radius);
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

constexpr float BDPT_T_MIN = 0.001f;  // bdpt_commons.glsl:24-25, 124-125
constexpr float BDPT_T_MAX = 1e6f;

// bdpt_commons.glsl:13-111 (EYE = false) and :113-216 (EYE = true). verts[i + 1] is the GLSL's vtx(i).
template <bool EYE>
inline int bdpt_random_walk(Bdpt& k, PathVertex* verts, int max_depth, vec3 throughput, float pdf) {
	auto vtx = [&](int i) -> PathVertex& { return verts[i + 1]; };
	if (max_depth == 0) return 0;
	int b = 0;
	int prev = 0;
	vec3 ray_pos = vtx(-1).pos;
	float pdf_fwd = pdf;
	float pdf_rev = 0.0f;
	vec3 wi = vtx(-1).dir;
	const bool finite_light = is_light_finite(vtx(-1).light_flags);
	while (true) {
		prev = b - 1;
		k.c.closest++;
		const Hit h = trace<false>(k.s.bvh, ray_pos, wi, BDPT_T_MIN, BDPT_T_MAX, &k.c.ts);
		if (h.prim == 0xFFFFFFFFu) {
			if (EYE) {
				vtx(b).throughput = throughput;
				vtx(b).pdf_fwd = pdf_fwd;
				b++;
			}
			break;
		}
		const HitPayload payload = build_hit(k.s, h);
		vec3 wo = vtx(prev).pos - payload.pos;
		const float wo_len = glm::length(wo);
		wo /= wo_len;
		vec3 n_s = payload.n_s;
		bool side = true;
		vec3 n_g = payload.n_g;
		if (glm::dot(payload.n_g, wo) < 0.0f) n_g = -n_g;
		if (glm::dot(n_g, n_s) < 0) {
			n_s *= -1.0f;
			side = false;
		}
		vtx(b).pdf_fwd = pdf_fwd * std::fabs(glm::dot(wo, n_s)) / (wo_len * wo_len);
		vtx(b).n_s = n_s;
		vtx(b).area = payload.area;
		vtx(b).pos = payload.pos;
		vtx(b).uv = payload.uv;
		vtx(b).material_idx = payload.material_idx;
		vtx(b).throughput = throughput;
		vtx(b).side = (uint)side;
		vtx(b).mode = EYE ? 1 : 0;
		const lmb_material mat = load_material(k.s, payload.material_idx, payload.uv);
		const bool mat_specular = (mat.bsdf_props & LMB_FLAG_SPECULAR) == LMB_FLAG_SPECULAR;
		const bool mat_transmissive = (mat.bsdf_props & LMB_FLAG_TRANSMISSION) == LMB_FLAG_TRANSMISSION;
		vtx(b).delta = (uint)mat_specular;
		if (++b >= max_depth) break;
		const vec3 r3 = rand3(k.seed);
		const BsdfSample bs = sample_bsdf(n_s, wo, mat, EYE ? 1 : 0, side, r3);
		wi = bs.wi;
		pdf_fwd = bs.pdf;
		const bool same_hem = same_hemisphere(wi, wo, n_s);
		if (is_zero(bs.f) || pdf_fwd == 0 || (!same_hem && !mat_transmissive)) break;
		throughput *= bs.f * std::fabs(bs.cos_theta) / pdf_fwd;
		pdf_rev = pdf_fwd;
		if (!mat_specular) pdf_rev = bsdf_pdf(mat, n_s, wi, wo, side);
		const bool g_term = EYE ? true : (prev > -1 || finite_light);
		if (g_term) pdf_rev *= std::fabs(glm::dot(vtx(prev).n_s, wo)) / (wo_len * wo_len);
		vtx(prev).pdf_rev = pdf_rev;
		ray_pos = offset_ray(payload.pos, n_g);
	}
	return b;
}

// bdpt_commons.glsl:218-260
inline int bdpt_generate_light_subpath(Bdpt& k, int max_depth) {
	const vec4 rands_pos = rand4(k.seed);
	const vec2 rands_dir = rand2(k.seed);
	const LightEmission le = sample_light_Le(k.s, rands_pos, rands_dir, k.pc.num_lights, k.pc.light_triangle_count);
	if (le.pdf_dir_w <= 0) return 0;
	k.light_pdf_pos = le.pdf_pos_a;
	PathVertex* lv = k.light_verts;
	lv[0].pos = le.pos;
	lv[0].light_flags = le.flags;
	lv[0].delta = 0;
	lv[0].dir = le.wi;
	lv[0].pdf_fwd = le.pdf_pos_a;
	lv[0].n_s = le.n;
	lv[0].side = 1;
	lv[0].mode = 0;
	const vec3 throughput = le.L * le.cos_from_light / (le.pdf_dir_w * lv[0].pdf_fwd);
	lv[0].throughput = le.L;
	const int num_light_verts = bdpt_random_walk<false>(k, lv, max_depth - 1, throughput, le.pdf_dir_w) + 1;
	if (!is_light_finite(le.flags)) lv[1].pdf_fwd = le.pdf_pos_a * std::fabs(glm::dot(le.wi, lv[1].n_s));
	if (is_light_delta(le.flags)) lv[0].pdf_fwd = 0;
	return num_light_verts;
}

inline vec3 sample_camera(const mat4& inv_view, const mat4& inv_proj, const vec2& d) {  // commons.glsl:30-33
	const vec4 target = inv_proj * vec4(d.x, d.y, 1, 1);
	return vec3(inv_view * vec4(glm::normalize(vec3(target)), 0));
}

// bdpt_commons.glsl:262-286
inline int bdpt_generate_camera_subpath(Bdpt& k, const vec2& d, const vec3& origin, int max_depth, float cam_area) {
	const mat4 inv_view = m4(k.ubo.inv_view), inv_proj = m4(k.ubo.inv_projection);
	PathVertex* cv = k.camera_verts;
	cv[0].pos = origin;
	cv[0].dir = sample_camera(inv_view, inv_proj, d);
	cv[0].area = cam_area;
	cv[0].throughput = vec3(1.0f);
	cv[0].delta = 0;
	cv[0].n_s = vec3((-inv_view) * vec4(0, 0, 1, 0));
	cv[0].side = 1;
	k.light_verts[0].mode = 1;  // sic, :272
	const float cos_theta = glm::dot(cv[0].dir, cv[0].n_s);
	const float pdf = 1 / (cam_area * (float)k.screen_size * cos_theta * cos_theta * cos_theta);
	return bdpt_random_walk<true>(k, cv, max_depth - 1, vec3(1), pdf) + 1;
}

inline float remap0(float v) { return v != 0.0f ? v : 1.0f; }

// Diagnostics of the restatement itself (tests/test_oracle_bdpt.py): orc_bdpt_set_only_s(s) keeps only the strategies with that
// many light vertices and gives them weight 1 (s = 0: emission found by the camera walk; s = 1: next-event estimation -- each an
// unbiased estimator on its own, so their images must converge to each other where both apply). -1 (default): the reference's weights.
static int g_bdpt_only_s = -1;
inline int bdpt_only_s() { return g_bdpt_only_s; }
// orc_bdpt_set_check_restore(1): after every calc_mis_weight both sub-paths are compared byte for byte with their state before the
// call. The GLSL patches vertices in place and restores them (bdpt_commons.glsl:311-340, 441-466); the CUDA path's pair-parallel
// kernels evaluate the pairs of a pixel side by side and therefore RELY on that restore being complete. Violations are counted.
static int g_bdpt_check_restore = 0;
static long long g_bdpt_restore_violations = 0;

// bdpt_commons.glsl:288-470
inline float calc_mis_weight_impl(Bdpt& k, int s, int t, const PathVertex& sampled);
inline float calc_mis_weight(Bdpt& k, int s, int t, const PathVertex& sampled) {
	if (!g_bdpt_check_restore) return calc_mis_weight_impl(k, s, t, sampled);
	PathVertex before_l[BDPT_MAX_VERTS], before_c[BDPT_MAX_VERTS];
	std::memcpy(before_l, k.light_verts, sizeof(before_l));
	std::memcpy(before_c, k.camera_verts, sizeof(before_c));
	const float w = calc_mis_weight_impl(k, s, t, sampled);
	if (std::memcmp(before_l, k.light_verts, sizeof(before_l)) != 0 || std::memcmp(before_c, k.camera_verts, sizeof(before_c)) != 0) {
#pragma omp atomic
		g_bdpt_restore_violations++;
	}
	return w;
}
inline float calc_mis_weight_impl(Bdpt& k, int s, int t, const PathVertex& sampled) {
	if (bdpt_only_s() >= 0) return s == bdpt_only_s() ? 1.0f : 0.0f;
	PathVertex* cam = k.camera_verts;
	PathVertex* lig = k.light_verts;
	bool s_0_changed = false;
	float s_0_pdf = 0;
	vec3 s_0_pdf_pos(0), s_0_pdf_nrm(0);
	bool t_0_changed = false;
	uint idx_1 = 0xFFFFFFFFu, idx_2 = 0xFFFFFFFFu, idx_3 = 0xFFFFFFFFu, idx_4 = 0xFFFFFFFFu;
	float idx_1_val = 0, idx_2_val = 0, idx_3_val = 0, idx_4_val = 0;
	uint delta_t_old = 0, delta_s_old = 0;
	if (s + t == 2) return 1.0f;
	if (s == 1) {
		s_0_pdf = lig[0].pdf_fwd, s_0_pdf_pos = lig[0].pos, s_0_pdf_nrm = lig[0].n_s;
		lig[0].pdf_fwd = sampled.pdf_fwd, lig[0].pos = sampled.pos, lig[0].n_s = sampled.n_s;
		s_0_changed = true;
	}
	if (t == 1) {
		s_0_pdf = cam[0].pdf_fwd, s_0_pdf_pos = cam[0].pos, s_0_pdf_nrm = cam[0].n_s;
		cam[0].pdf_fwd = sampled.pdf_fwd, cam[0].pos = sampled.pos, cam[0].n_s = sampled.n_s;
		t_0_changed = true;
	}
	if (t > 0) {
		delta_t_old = cam[t - 1].delta;
		cam[t - 1].delta = 0;
	}
	if (s > 0) {
		delta_s_old = lig[s - 1].delta;
		lig[s - 1].delta = 0;
	}
	if (t > 0) {
		idx_1_val = cam[t - 1].pdf_rev;
		idx_1 = (uint)t;
		if (s > 0) {
			vec3 dir = cam[t - 1].pos - lig[s - 1].pos;
			const float dir_len = glm::length(dir);
			dir /= dir_len;
			float pdf_rev = 0;  // GLSL leaves it unassigned only for s < 1, excluded here
			if (s >= 2) {
				const lmb_material mat = load_material(k.s, lig[s - 1].material_idx, lig[s - 1].uv);
				const vec3 wo = glm::normalize(lig[s - 2].pos - lig[s - 1].pos);
				pdf_rev = bsdf_pdf(mat, lig[s - 1].n_s, wo, dir, lig[s - 1].side == 1);
				pdf_rev *= std::fabs(glm::dot(dir, cam[t - 1].n_s)) / (dir_len * dir_len);
			} else if (s == 1) {
				if (!is_light_finite(lig[0].light_flags)) {
					pdf_rev = k.light_pdf_pos;
					pdf_rev *= std::fabs(glm::dot(dir, cam[t - 1].n_s));
				} else {
					pdf_rev = light_pdf(lig[0].light_flags, lig[0].n_s, dir);
					pdf_rev *= std::fabs(glm::dot(dir, cam[t - 1].n_s)) / (dir_len * dir_len);
				}
			}
			cam[t - 1].pdf_rev = pdf_rev;
		} else {
			cam[t - 1].pdf_rev = 1.0f / ((float)k.pc.light_triangle_count * cam[t - 1].area);
		}
	}
	if (t > 1) {
		idx_2_val = cam[t - 2].pdf_rev;
		idx_2 = (uint)t;
		vec3 dir = cam[t - 2].pos - cam[t - 1].pos;
		const float dir_len = glm::length(dir);
		dir /= dir_len;
		if (s > 0) {
			const lmb_material mat = load_material(k.s, cam[t - 1].material_idx, cam[t - 1].uv);
			const vec3 wo = glm::normalize(lig[s - 1].pos - cam[t - 1].pos);
			cam[t - 2].pdf_rev = bsdf_pdf(mat, cam[t - 1].n_s, wo, dir, cam[t - 1].side == 1);
			if (cam[t - 2].pdf_rev != 0) cam[t - 2].pdf_rev *= std::fabs(glm::dot(dir, cam[t - 2].n_s)) / (dir_len * dir_len);
		} else {
			const float cos_x = glm::dot(cam[t - 1].n_s, dir);
			const float cos_y = glm::dot(cam[t - 2].n_s, dir);
			cam[t - 2].pdf_rev = std::fabs(cos_x * cos_y) / (PI * dir_len * dir_len);
		}
	}
	if (s > 0) {
		idx_3_val = lig[s - 1].pdf_rev;
		idx_3 = (uint)s;
		vec3 dir = lig[s - 1].pos - cam[t - 1].pos;
		const float dir_len = glm::length(dir);
		dir /= dir_len;
		if (t == 1) {
			const float cos_theta = glm::dot(cam[0].n_s, dir);
			float pdf = 1.0f / (cam[0].area * (float)k.screen_size * cos_theta * cos_theta * cos_theta);
			pdf *= std::fabs(glm::dot(dir, lig[s - 1].n_s)) / (dir_len * dir_len);
			lig[s - 1].pdf_rev = pdf;
		} else {
			const vec3 wo = glm::normalize(cam[t - 2].pos - cam[t - 1].pos);
			const lmb_material mat = load_material(k.s, cam[t - 1].material_idx, cam[t - 1].uv);
			lig[s - 1].pdf_rev = bsdf_pdf(mat, cam[t - 1].n_s, wo, dir, cam[t - 1].side == 1);
			if ((s == 1 && is_light_finite(lig[0].light_flags)) || s > 1)
				lig[s - 1].pdf_rev *= std::fabs(glm::dot(dir, lig[s - 1].n_s)) / (dir_len * dir_len);
		}
	}
	if (s > 1) {
		idx_4_val = lig[s - 2].pdf_rev;
		idx_4 = (uint)s;
		vec3 dir = lig[s - 2].pos - lig[s - 1].pos;
		const vec3 wo = glm::normalize(cam[t - 1].pos - lig[s - 1].pos);
		const float dir_len = glm::length(dir);
		dir /= dir_len;
		const lmb_material mat = load_material(k.s, lig[s - 1].material_idx, lig[s - 1].uv);
		lig[s - 2].pdf_rev = bsdf_pdf(mat, lig[s - 1].n_s, wo, dir, lig[s - 1].side == 1);
		if ((s == 2 && is_light_finite(lig[0].light_flags)) || s > 2)
			lig[s - 2].pdf_rev *= std::fabs(glm::dot(dir, lig[s - 2].n_s)) / (dir_len * dir_len);
	}
	float sum_ri = 0.0f;
	float weight = 1.0f;
	for (int i = t - 1; i > 0; i--) {
		weight *= remap0(cam[i].pdf_rev) / remap0(cam[i].pdf_fwd);
		if (cam[i].delta == 0 && cam[i - 1].delta == 0) sum_ri += weight;
	}
	weight = 1.0f;
	for (int i = s - 1; i >= 0; i--) {
		weight *= remap0(lig[i].pdf_rev) / remap0(lig[i].pdf_fwd);
		const bool delta_prev = i > 0 ? lig[i - 1].delta == 1 : is_light_delta(lig[0].light_flags);
		if (lig[i].delta == 0 && !delta_prev) sum_ri += weight;
	}
	if (s_0_changed) lig[0].pdf_fwd = s_0_pdf, lig[0].pos = s_0_pdf_pos, lig[0].n_s = s_0_pdf_nrm;
	if (t_0_changed) cam[0].pdf_fwd = s_0_pdf, cam[0].pos = s_0_pdf_pos, cam[0].n_s = s_0_pdf_nrm;
	if (idx_1 != 0xFFFFFFFFu) {
		cam[idx_1 - 1].pdf_rev = idx_1_val;
		cam[idx_1 - 1].delta = delta_t_old;
	}
	if (idx_2 != 0xFFFFFFFFu) cam[idx_2 - 2].pdf_rev = idx_2_val;
	if (idx_3 != 0xFFFFFFFFu) {
		lig[idx_3 - 1].pdf_rev = idx_3_val;
		lig[idx_3 - 1].delta = delta_s_old;
	}
	if (idx_4 != 0xFFFFFFFFu) lig[idx_4 - 2].pdf_rev = idx_4_val;
	return 1 / (1 + sum_ri);
}

// float -> int of `ivec2(...)`, quirk B5: false when the value has no int
inline bool splat_coord(float v, int& out) {
	if (!(std::fabs(v) < 1e9f)) return false;
	out = (int)v;
	return true;
}

// bdpt_commons.glsl:472-530. Returns the splat radiance; ok = false when there is no pixel to splat to.
inline vec3 bdpt_connect_cam(Bdpt& k, int s, int& cx, int& cy) {
	PathVertex* cam = k.camera_verts;
	PathVertex* lig = k.light_verts;
	PathVertex sampled;
	vec3 L(0);
	cx = cy = -1;
	vec3 dir = cam[0].pos - lig[s - 1].pos;
	const float len = glm::length(dir);
	dir /= len;
	const float cos_y = glm::dot(dir, lig[s - 1].n_s);
	const float cos_theta = glm::dot(cam[0].n_s, -dir);
	if (cos_theta <= 0.0f) return vec3(0);
	const float cos_3_theta = cos_theta * cos_theta * cos_theta;
	const float cam_pdf_ratio = std::fabs(cos_y) / (cam[0].area * cos_3_theta * len * len);
	const vec3 ray_origin = offset_ray2(lig[s - 1].pos, lig[s - 1].n_s);
	const lmb_material mat = load_material(k.s, lig[s - 1].material_idx, lig[s - 1].uv);
	const vec3 wo = glm::normalize(lig[s - 2].pos - lig[s - 1].pos);
	float unused_pdf;
	const vec3 f = eval_bsdf(lig[s - 1].n_s, wo, mat, lig[s - 1].mode, lig[s - 1].side == 1, dir, unused_pdf);
	if (is_zero(f)) return L;
	if (cam_pdf_ratio > 0.0f) {
		k.c.shadow++;
		const Hit sh = trace<true>(k.s.bvh, ray_origin, dir, 0.0f, len - EPS, &k.c.ts);
		if (sh.prim == 0xFFFFFFFFu) {
			sampled.pos = cam[0].pos;
			sampled.n_s = cam[0].n_s;
			L = lig[s - 1].throughput * cam_pdf_ratio * f / (float)k.screen_size;
		}
	}
	dir = -dir;
	const mat4 view = m4(k.ubo.view), proj = m4(k.ubo.projection);
	vec4 target = view * vec4(dir.x, dir.y, dir.z, 0);
	target /= target.z;
	target = (-proj) * target;
	const vec2 cf = 0.5f * (1.0f + vec2(target)) * vec2((float)k.pc.size_x, (float)k.pc.size_y) - 0.5f;
	if (!splat_coord(cf.x, cx) || !splat_coord(cf.y, cy)) {
		cx = cy = -1;
		return vec3(0);
	}
	if (cx < 0 || (uint)cx >= k.pc.size_x || cy < 0 || (uint)cy >= k.pc.size_y || glm::dot(dir, cam[0].n_s) < 0) return vec3(0);
	float mis_weight = 1.0f;
	if (luminance(L) != 0.0f) mis_weight = calc_mis_weight(k, s, 1, sampled);
	return mis_weight * L;
}

// bdpt_commons.glsl:532-641
inline vec3 bdpt_connect(Bdpt& k, int s, int t) {
	PathVertex* cam = k.camera_verts;
	PathVertex* lig = k.light_verts;
	vec3 L(0);
	PathVertex sampled;
	if (s == 0) {
		const uint mat_idx = cam[t - 1].material_idx;
		const lmb_material& mat = k.s.sd.materials[mat_idx];  // un-textured read (`materials.m[mat_idx]`), quirk B3
		L = v3(mat.emissive_factor) * cam[t - 1].throughput;
	} else if (s == 1) {
		// second assumption of the pair-parallel kernels: this is the (t - 2)-th rand4 after the walks
		if (g_bdpt_check_restore && k.seed.w != k.rng_after_walks + 4u * (uint)(t - 2)) {
#pragma omp atomic
			g_bdpt_restore_violations++;
		}
		const vec4 r4 = rand4(k.seed);
		const LightSample ls = sample_light_Li(k.s, r4, cam[t - 1].pos, k.pc.num_lights);
		const float cos_x = std::fabs(glm::dot(ls.wi, cam[t - 1].n_s));
		const vec3 ray_origin = offset_ray2(cam[t - 1].pos, cam[t - 1].n_s);
		const vec3 wo = glm::normalize(cam[t - 2].pos - cam[t - 1].pos);
		const lmb_material mat = load_material(k.s, cam[t - 1].material_idx, cam[t - 1].uv);
		float unused_pdf;
		const vec3 f = eval_bsdf(cam[t - 1].n_s, wo, mat, cam[t - 1].mode, cam[t - 1].side == 1, ls.wi, unused_pdf);
		if (!is_zero(f)) {
			k.c.shadow++;
			const Hit sh = trace<true>(k.s.bvh, ray_origin, ls.wi, 0.0f, ls.wi_len - EPS, &k.c.ts);
			if (sh.prim == 0xFFFFFFFFu) {
				const float pdf_light_w =
					light_pdf_a_to_w(ls.flags, ls.pdf_a, ls.wi_len * ls.wi_len, ls.cos_from_light) / (float)k.pc.light_triangle_count;
				sampled.pdf_fwd = ls.pdf_a / (float)k.pc.light_triangle_count;
				sampled.pos = ls.pos;
				sampled.n_s = ls.n;
				sampled.delta = (uint)is_light_delta(ls.flags);
				L = cam[t - 1].throughput * f * std::fabs(cos_x) * ls.Le / pdf_light_w;
			}
		}
	} else {
		const vec3 n_s = lig[s - 1].n_s;
		const vec3 n_t = cam[t - 1].n_s;
		vec3 d = lig[s - 1].pos - cam[t - 1].pos;
		const float len = glm::length(d);
		d /= len;
		const float G = glm::dot(n_s, -d) * glm::dot(n_t, d) / (len * len);
		if (G > 0) {
			const lmb_material mat_1 = load_material(k.s, cam[t - 1].material_idx, cam[t - 1].uv);
			const lmb_material mat_2 = load_material(k.s, lig[s - 1].material_idx, lig[s - 1].uv);
			const vec3 wo_1 = glm::normalize(cam[t - 2].pos - cam[t - 1].pos);
			const vec3 wo_2 = glm::normalize(lig[s - 2].pos - lig[s - 1].pos);
			float unused_pdf;
			const vec3 brdf1 = eval_bsdf(cam[t - 1].n_s, wo_1, mat_1, cam[t - 1].mode, cam[t - 1].side == 1, d, unused_pdf);
			const vec3 brdf2 = eval_bsdf(lig[s - 1].n_s, wo_2, mat_2, lig[s - 1].mode, lig[s - 1].side == 1, -d, unused_pdf);
			if (!is_zero(brdf1) && !is_zero(brdf2)) {
				const vec3 ray_origin = offset_ray2(cam[t - 1].pos, cam[t - 1].n_s);
				k.c.shadow++;
				const Hit sh = trace<true>(k.s.bvh, ray_origin, d, 0.0f, len - EPS, &k.c.ts);
				if (sh.prim == 0xFFFFFFFFu) L = lig[s - 1].throughput * G * brdf1 * brdf2 * cam[t - 1].throughput;
			}
		}
	}
	if (luminance(L) != 0.0f) {
		const float mis_weight = calc_mis_weight(k, s, t, sampled);
		L *= mis_weight;
	}
	return L;
}

// bdpt.rgen:39-78 without the film: `col` of the pixel's own strategies (t >= 2) is returned, the t == 1 splats are handed to
// `splat(x, y, rgb)` (quirk B2).
template <class Splat>
inline vec3 bdpt_pixel(const orc_scene& s, const lmb_pc_bdpt& pc, const lmb_scene_ubo& ubo, uint px, uint py, uint frame, Counters& c, Splat&& splat) {
	Bdpt k{s, pc, ubo, c, uvec4(px, py, frame ^ pc.time, 0), pc.size_x * pc.size_y};
	const mat4 inv_view = m4(ubo.inv_view), inv_proj = m4(ubo.inv_projection);
	const vec2 size((float)pc.size_x, (float)pc.size_y);
	const vec2 pixel = vec2((float)px, (float)py) + vec2(0.5f);
	const vec2 in_uv = pixel / size;
	const vec2 d = in_uv * 2.0f - 1.0f;
	const vec4 origin = inv_view * vec4(0, 0, 0, 1);
	vec3 col(0);
	vec4 area_int = inv_proj * vec4(2.0f / (float)pc.size_x, 2.0f / (float)pc.size_y, 0, 1);
	area_int /= area_int.w;
	const float cam_area = std::fabs(area_int.x * area_int.y);
	const int num_light_paths = bdpt_generate_light_subpath(k, pc.max_depth + 1);
	const int num_cam_paths = bdpt_generate_camera_subpath(k, d, vec3(origin), pc.max_depth + 1, cam_area);
	k.rng_after_walks = k.seed.w;
	for (int t = 1; t <= num_cam_paths; t++) {
		for (int sl = 0; sl <= num_light_paths; sl++) {
			const int depth = sl + t - 2;
			if (depth > (pc.max_depth - 1) || depth < 0 || (sl == 1 && t == 1)) continue;
			if (t == 1) {
				int cx, cy;
				const vec3 splat_col = bdpt_connect_cam(k, sl, cx, cy);
				if (luminance(splat_col) > 0) splat((uint)cx, (uint)cy, splat_col);
			} else {
				col += bdpt_connect(k, sl, t);
			}
		}
	}
	return col;
}
