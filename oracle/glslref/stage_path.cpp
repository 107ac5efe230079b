// ORACLE -- TEST INFRASTRUCTURE ONLY. src/shaders/integrators/path/path.rgen (+ commons.glsl, utils.glsl, bsdf_commons.glsl,
// bsdf/*.glsl, atmosphere/atmosphere.glsl, integrators/pt_commons.glsl), translated by glsl2cpp.py; plus the emulation of
// what Path::render sets up around it (descriptor bindings, push constants, the shader binding table) and the probes.
#include <omp.h>
#include "harness.h"

namespace glslref {
struct PathRgen : Stage {
	using Stage::Stage;
#include "gen/integrators/path/path.rgen.inc"
};
}  // namespace glslref

using namespace glslref;

namespace glslref {
thread_local uint64_t t_rays[3];
}
namespace {

void cb_intersect(const Env* env, const float* ray8, int first, Intersection* out) {
	const ref_scene* S = (const ref_scene*)env->user;
	struct {
		float t, b1, b2;
		uint32_t prim;
	} h;
	uint32_t mesh = 0, local = 0;
	out->hit = (uint32_t)S->trace1(S->user, ray8, first, &h, &mesh, &local);
	out->t = h.t, out->b1 = h.b1, out->b2 = h.b2, out->instance_custom_index = mesh, out->primitive_id = local;
}
void cb_texture(const Env* env, const sampler2D* s, const float* uv, float* rgba) {
	const ref_scene* S = (const ref_scene*)env->user;
	S->texture(S->user, s->id, uv, 1, rgba);
	rgba[3] = 1.0f;  // the alpha channel is not read on this path (bsdf_commons.glsl:19 takes .xyz)
}
mat4 load_mat4(const float* p) {
	mat4 m;
	for (int c = 0; c < 4; c++)
		for (int r = 0; r < 4; r++) m[c][r] = p[4 * c + r];  // column-major on both sides
	return m;
}
// The shader binding table of Path::render (Path.cpp:42-46): raygen = path.rgen, miss 0 = ray.rmiss, miss 1 = ray_shadow.rmiss,
// one hit group = ray.rchit + ray.rahit. ray.rahit (terminateRayEXT) only ever runs for rays that already carry
// gl_RayFlagsTerminateOnFirstHitEXT, so it changes nothing and is not modelled.
void cb_trace_ray(const Env* env, uint flags, uint cull_mask, uint, uint, uint miss_index, const vec3& o, float tmin, const vec3& d, float tmax,
				  void* payload) {
	const ref_scene* S = (const ref_scene*)env->user;
	const float ray[8] = {o.x, o.y, o.z, tmin, d.x, d.y, d.z, tmax};
	const bool first = (flags & Stage::gl_RayFlagsTerminateOnFirstHitEXT) != 0;
	t_rays[first ? 1 : (cull_mask == 0xFFu ? 0 : 2)]++;
	Intersection is;
	env->intersect(env, ray, first, &is);
	Stage::Inputs in;
	in.env = env;
	in.incoming_payload = payload;
	if (is.hit) {
		if (flags & Stage::gl_RayFlagsSkipClosestHitShaderEXT) return;
		in.hit = &is;
		in.ray_tmin = tmin;
		in.object_to_world = load_mat4(S->sd.world_matrices + 16 * is.instance_custom_index);
		in.world_to_object = load_mat4(S->sd.inv_world_matrices + 16 * is.instance_custom_index);
		run_rchit(in);
	} else if (miss_index == 0) {
		run_rmiss(in);
	} else {
		run_shadow_rmiss(in);
	}
}
Stage::Inputs probe_inputs(const ref_scene* s) {
	Stage::Inputs in;
	in.env = &s->env;
	return in;
}
inline vec3 v3(const float* p) { return vec3(p[0], p[1], p[2]); }
}  // namespace

extern "C" {

int ref_scene_create(const lmb_scene_desc* sd, const void* user, ref_trace1_fn trace1, ref_texture_fn texture, ref_scene** out) {
	ref_scene* s = new ref_scene();
	s->sd = *sd;
	s->user = user, s->trace1 = trace1, s->texture = texture;
	std::memset(&s->scene_desc, 0, sizeof(s->scene_desc));
	std::memset(&s->ubo, 0, sizeof(s->ubo));
	std::memset(&s->pc, 0, sizeof(s->pc));
	// Path.cpp:6-11
	s->scene_desc.index_addr = (uint64_t)(uintptr_t)sd->indices;
	s->scene_desc.material_addr = (uint64_t)(uintptr_t)sd->materials;
	s->scene_desc.prim_info_addr = (uint64_t)(uintptr_t)sd->prim_infos;
	s->scene_desc.compact_vertices_addr = (uint64_t)(uintptr_t)sd->vertices;
	BufferRegistry& reg = BufferRegistry::get();
	reg.add(sd->indices, (size_t)sd->n_indices * 4), reg.add(sd->materials, (size_t)sd->n_materials * sizeof(::Material));
	reg.add(sd->prim_infos, (size_t)sd->n_prim_meshes * sizeof(::PrimMeshInfo)), reg.add(sd->vertices, (size_t)sd->n_vertices * sizeof(::Vertex));
	s->samplers.resize(sd->n_textures ? sd->n_textures : 1);
	for (uint32_t i = 0; i < sd->n_textures; i++) s->samplers[i] = sampler2D{s, i};
	Env& e = s->env;
	e.user = s;
	e.push_constants = &s->pc;
	// Path.cpp:49-57 + commons.glsl:11-15: 0 image, 1 SceneUBO, 2 SceneDesc, 3 lights, 4 textures; set 1 binding 0 = TLAS
	e.sets[0][1] = &s->ubo;
	e.sets[0][2] = &s->scene_desc;
	e.sets[0][3] = (void*)sd->lights;
	e.sampler_arrays[4] = s->samplers.data();
	e.intersect = cb_intersect;
	e.texture = cb_texture;
	e.trace_ray = cb_trace_ray;
	*out = s;
	return 0;
}
void ref_scene_destroy(ref_scene* s) {
	if (!s) return;
	BufferRegistry& reg = BufferRegistry::get();
	reg.remove(s->sd.indices), reg.remove(s->sd.materials), reg.remove(s->sd.prim_infos), reg.remove(s->sd.vertices);
	delete s;
}

int ref_render_path(ref_scene* s, const lmb_pc_path* pc_in, const lmb_scene_ubo* ubo, uint32_t first_frame, uint32_t n_frames, float* rgba,
					uint64_t* rays3, int n_threads) {
	std::memcpy(&s->ubo, ubo, sizeof(s->ubo));
	const int W = (int)pc_in->size_x, H = (int)pc_in->size_y;
	s->env.images[0] = image2D{rgba, W, H};
	const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
	uint64_t r0 = 0, r1 = 0, r2 = 0;
	for (uint32_t f = first_frame; f < first_frame + n_frames; f++) {
		std::memcpy(&s->pc, pc_in, sizeof(s->pc));
		s->pc.frame_num = f;  // Path.cpp:32
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1) reduction(+ : r0, r1, r2)
		for (int y = 0; y < H; y++) {
			t_rays[0] = t_rays[1] = t_rays[2] = 0;
			for (int x = 0; x < W; x++) {
				Stage::Inputs in;
				in.env = &s->env;
				in.launch_id = uvec3(x, y, 0);
				in.launch_size = uvec3(W, H, 1);
				PathRgen inv(in);
				inv.main();
			}
			r0 += t_rays[0], r1 += t_rays[1], r2 += t_rays[2];
		}
	}
	if (rays3) rays3[0] += r0, rays3[1] += r1, rays3[2] += r2;
	return 0;
}

void ref_kat_pcg4d(ref_scene* s, const uint32_t* in4, uint32_t n, uint32_t* out4) {
	PathRgen st(probe_inputs(s));
	for (uint32_t i = 0; i < n; i++) {
		const uvec4 r = st.pcg4d(uvec4(in4[4 * i], in4[4 * i + 1], in4[4 * i + 2], in4[4 * i + 3]));
		out4[4 * i] = r.x, out4[4 * i + 1] = r.y, out4[4 * i + 2] = r.z, out4[4 * i + 3] = r.w;
	}
}
void ref_kat_rand(ref_scene* s, const uint32_t* seed4, uint32_t n, uint32_t draws, float* out) {
	PathRgen st(probe_inputs(s));
	for (uint32_t i = 0; i < n; i++) {
		uvec4 sd(seed4[4 * i], seed4[4 * i + 1], seed4[4 * i + 2], seed4[4 * i + 3]);
		// draws come in the groupings the shaders use: rand4 (NEE), rand3 (BSDF), rand (RR), rand2
		uint32_t k = 0;
		while (k < draws) {
			const uint32_t left = draws - k;
			if (left >= 4 && (k % 10) == 0) {
				const vec4 r = st.rand4(sd);
				for (int c = 0; c < 4; c++) out[(size_t)i * draws + k++] = r[c];
			} else if (left >= 3 && (k % 10) == 4) {
				const vec3 r = st.rand3(sd);
				for (int c = 0; c < 3; c++) out[(size_t)i * draws + k++] = r[c];
			} else if (left >= 2 && (k % 10) == 7) {
				const vec2 r = st.rand2(sd);
				for (int c = 0; c < 2; c++) out[(size_t)i * draws + k++] = r[c];
			} else {
				out[(size_t)i * draws + k++] = st.rand(sd);
			}
		}
	}
}
void ref_kat_offset_ray(ref_scene* s, const float* p3, const float* n3, uint32_t n, float* out3, float* out3_b) {
	PathRgen st(probe_inputs(s));
	for (uint32_t i = 0; i < n; i++) {
		const vec3 a = st.offset_ray(v3(p3 + 3 * i), v3(n3 + 3 * i));
		const vec3 b = st.offset_ray2(v3(p3 + 3 * i), v3(n3 + 3 * i));
		for (int k = 0; k < 3; k++) out3[3 * i + k] = a[k], out3_b[3 * i + k] = b[k];
	}
}
void ref_kat_sample_bsdf(ref_scene* s, const lmb_material* mat, const float* n_s3, const float* wo3, const float* rands3, const uint8_t* side,
						 uint32_t n, float* out8) {
	PathRgen st(probe_inputs(s));
	const ::Material m = *reinterpret_cast<const ::Material*>(mat);
	for (uint32_t i = 0; i < n; i++) {
		vec3 wi;
		float pdf, cos_theta;
		const vec3 f = st.sample_bsdf(v3(n_s3 + 3 * i), v3(wo3 + 3 * i), m, 1, side[i] != 0, wi, pdf, cos_theta, v3(rands3 + 3 * i));
		float* o = out8 + 8 * (size_t)i;
		o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = wi.x, o[4] = wi.y, o[5] = wi.z, o[6] = pdf, o[7] = cos_theta;
	}
}
void ref_kat_eval_bsdf(ref_scene* s, const lmb_material* mat, const float* n_s3, const float* wo3, const float* wi3, const uint8_t* side,
					   uint32_t n, float* out4) {
	PathRgen st(probe_inputs(s));
	const ::Material m = *reinterpret_cast<const ::Material*>(mat);
	for (uint32_t i = 0; i < n; i++) {
		float pdf;
		// the 7-argument overload pt_commons.glsl:19 calls (bsdf_commons.glsl:172-176)
		const vec3 f = st.eval_bsdf(v3(n_s3 + 3 * i), v3(wo3 + 3 * i), m, 1, side[i] != 0, v3(wi3 + 3 * i), pdf);
		float* o = out4 + 4 * (size_t)i;
		o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = pdf;
	}
}
void ref_kat_bsdf_pdf(ref_scene* s, const lmb_material* mat, const float* n_s3, const float* wo3, const float* wi3, const uint8_t* side,
					  uint32_t n, float* out) {
	PathRgen st(probe_inputs(s));
	const ::Material m = *reinterpret_cast<const ::Material*>(mat);
	for (uint32_t i = 0; i < n; i++) out[i] = st.bsdf_pdf(m, v3(n_s3 + 3 * i), v3(wo3 + 3 * i), v3(wi3 + 3 * i), side[i] != 0);
}
void ref_kat_atmosphere(ref_scene* s, const float* origin3, const float* dir3, const float* light_dir3, const float* light_L3, uint32_t n,
						float* out3) {
	// shade_atmosphere (commons.glsl:156-168) with a directional light whose -normalize(to - pos) is light_dir3 and L = light_L3
	::Light light;
	std::memset(&light, 0, sizeof(light));
	light.pos = v3(light_dir3);  // to = 0, pos = light_dir  =>  -normalize(to - pos) = normalize(light_dir)
	light.to = vec3(0);
	light.L = v3(light_L3);
	const void* saved = s->env.sets[0][3];
	s->env.sets[0][3] = &light;
#pragma omp parallel for schedule(dynamic, 16)
	for (int64_t i = 0; i < (int64_t)n; i++) {
		PathRgen st(probe_inputs(s));
		const vec3 r = st.shade_atmosphere(0u, vec3(0), v3(origin3 + 3 * i), v3(dir3 + 3 * i), 10000.0f);
		out3[3 * i] = r.x, out3[3 * i + 1] = r.y, out3[3 * i + 2] = r.z;
	}
	s->env.sets[0][3] = const_cast<void*>(saved);
}
void ref_kat_sample_light(ref_scene* s, int32_t num_lights, const float* rands4, const float* p3, uint32_t n, float* out16) {
	PathRgen st(probe_inputs(s));
	for (uint32_t i = 0; i < n; i++) {
		const float* r = rands4 + 4 * i;
		vec3 wi;
		float wi_len, pdf_w, pdf_a, cos_from_light;
		PathRgen::LightRecord rec;
		std::memset(&rec, 0, sizeof(rec));
		// the 9-argument overload pt_commons.glsl:12-13 calls (commons.glsl:309-315)
		const vec3 Le = st.sample_light_Li(vec4(r[0], r[1], r[2], r[3]), v3(p3 + 3 * i), num_lights, pdf_w, wi, wi_len, pdf_a, cos_from_light, rec);
		float* o = out16 + 16 * (size_t)i;
		o[0] = Le.x, o[1] = Le.y, o[2] = Le.z, o[3] = wi.x, o[4] = wi.y, o[5] = wi.z;
		o[6] = wi_len, o[7] = pdf_w, o[8] = pdf_a, o[9] = cos_from_light;
		o[10] = (float)rec.light_idx, o[11] = (float)rec.flags, o[12] = (float)rec.triangle_idx, o[13] = (float)rec.instance_idx;
		o[14] = rec.bary.x, o[15] = rec.bary.y;
	}
}
void ref_kat_light_Le(ref_scene* s, int32_t num_lights, int32_t total_light, const float* rands6, uint32_t n, float* out16) {
	PathRgen st(probe_inputs(s));
	for (uint32_t i = 0; i < n; i++) {
		const float* r = rands6 + 6 * i;
		float cos_from_light = 0, pdf_pos_a = 0, pdf_dir_w = 0;
		PathRgen::LightRecord rec;
		vec3 pos(0), wi(0), nn(0);
		// the 11-argument overload bdpt_commons.glsl calls (commons.glsl:408-415)
		const vec3 L = st.sample_light_Le(vec4(r[0], r[1], r[2], r[3]), vec2(r[4], r[5]), num_lights, total_light, cos_from_light, rec, pos, wi, nn,
										  pdf_pos_a, pdf_dir_w);
		float* o = out16 + 16 * (size_t)i;
		o[0] = L.x, o[1] = L.y, o[2] = L.z, o[3] = pos.x, o[4] = pos.y, o[5] = pos.z, o[6] = wi.x, o[7] = wi.y, o[8] = wi.z;
		o[9] = nn.x, o[10] = nn.y, o[11] = nn.z, o[12] = cos_from_light, o[13] = pdf_pos_a, o[14] = pdf_dir_w, o[15] = (float)rec.flags;
	}
}
void ref_kat_load_material(ref_scene* s, const uint32_t* material_idx, const float* uv2, uint32_t n, lmb_material* out) {
	PathRgen st(probe_inputs(s));
	for (uint32_t i = 0; i < n; i++) {
		const ::Material m = st.load_material(material_idx[i], vec2(uv2[2 * i], uv2[2 * i + 1]));
		std::memcpy(&out[i], &m, sizeof(m));
	}
}
}  // extern "C"
