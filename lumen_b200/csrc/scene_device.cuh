// scene_device.cuh -- device functions that read the uploaded scene: hit-record construction, material / texture fetch
// and light sampling.
//   build_hit          src/shaders/ray.rchit:24-84          (HitPayload, utils.glsl:15-26)
//   load_material      src/shaders/bsdf_commons.glsl:16-22  (+ sampler state LumenScene.cpp:193-213)
//   sample_triangle    src/shaders/commons.glsl:112-149
//   sample_light_Li    src/shaders/commons.glsl:224-300
//   shade_atmosphere   src/shaders/commons.glsl:156-168
#pragma once
#include "context.h"
#include "shading.cuh"

namespace lmb {

struct HitPayload {
	V3 n_g, n_s, pos;
	V2 uv;
	uint32_t material_idx, triangle_idx, instance_idx;
};

// Bilinear, REPEAT, LOD 0, sRGB-decoded texel fetch; texel centres at integer + 0.5, fp32 weights (same definition as the
// CPU oracle; the fixed-function sampler's sub-texel precision is not specified by Vulkan).
LMB_DN V3 sample_texture(const DeviceScene& sc, uint32_t id, const V2& uv) {
	const uint2 dim = sc.tex_dims[id];
	const uint8_t* __restrict__ px = sc.tex_data[id];
	const int W = (int)dim.x, H = (int)dim.y;
	float u = uv.x * (float)W - 0.5f, v = uv.y * (float)H - 0.5f;
	if (!(fabsf(u) < 1e9f)) u = 0.0f;
	if (!(fabsf(v) < 1e9f)) v = 0.0f;
	const float x0f = floorf(u), y0f = floorf(v);
	const float fx = u - x0f, fy = v - y0f;
	const int xi = (int)x0f, yi = (int)y0f;
	const int x0 = ((xi % W) + W) % W, x1 = (((xi + 1) % W) + W) % W;
	const int y0 = ((yi % H) + H) % H, y1 = (((yi + 1) % H) + H) % H;
	auto texel = [&](int x, int y) {
		const uchar4 t = reinterpret_cast<const uchar4*>(px)[(size_t)y * W + x];
		return v3(sc.srgb_lut[t.x], sc.srgb_lut[t.y], sc.srgb_lut[t.z]);
	};
	const V3 top = texel(x0, y0) * (1.0f - fx) + texel(x1, y0) * fx;
	const V3 bot = texel(x0, y1) * (1.0f - fx) + texel(x1, y1) * fx;
	return top * (1.0f - fy) + bot * fy;
}

LMB_D lmb_material load_material(const DeviceScene& sc, uint32_t material_idx, const V2& uv) {
	lmb_material m = sc.materials[material_idx];
	if (m.texture_id > -1) {
		const V3 a = v3(m.albedo) * sample_texture(sc, (uint32_t)m.texture_id, uv);
		m.albedo[0] = a.x, m.albedo[1] = a.y, m.albedo[2] = a.z;
	}
	return m;
}

LMB_D V3 vtx_pos(const lmb_vertex& v) { return v3(v.pos[0], v.pos[1], v.pos[2]); }
LMB_D V3 vtx_nrm(const lmb_vertex& v) { return v3(v.normal[0], v.normal[1], v.normal[2]); }

// `tri_rec[prim]` = (absolute vertex index of each corner, prim mesh index): the PrimMeshInfo -> index buffer -> vertex chain
// of ray.rchit:27-33 resolved once at upload, so the hit record costs one dependent fetch before the vertices instead of three.
// payload.triangle_idx (gl_PrimitiveID) is only compared by the area-light MIS probe: callers read sc.tri_local[prim] there.
//
// LMB_HIT_PACKED (default): `tri_shade[prim]` holds everything ray.rchit reads for one triangle -- the three corners' positions,
// normals and uvs, the mesh and its material index -- as ONE 128-byte-aligned record (7 x float4 used), gathered once at upload
// (post.cu k_ingest_triangles). The hit record then costs a single dependent 128-byte line instead of tri_rec (16 B) followed by three
// 32-byte vertex fetches at scattered addresses and the PrimMeshInfo: one level less in k_shade's dependent gather chain, which is
// what that kernel waits on (profiles/r01i: 24 % occupancy, long-scoreboard stalls 4-6 warps per issue). Same values, same
// arithmetic, same results.
#ifndef LMB_HIT_PACKED
#define LMB_HIT_PACKED 1
#endif
LMB_DN HitPayload build_hit(const DeviceScene& sc, uint32_t prim_global, float b1, float b2) {
	HitPayload p;
#if LMB_HIT_PACKED
	const float4* __restrict__ r = sc.tri_shade + 8 * (size_t)prim_global;
	const float4 r0 = __ldg(r + 0), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3), r4 = __ldg(r + 4), r5 = __ldg(r + 5), r6 = __ldg(r + 6);
	const uint32_t mesh = __float_as_uint(r6.x);
	const V3 q0 = v3(r0.x, r0.y, r0.z), q1 = v3(r1.x, r1.y, r1.z), q2 = v3(r2.x, r2.y, r2.z);
	const V3 n0 = v3(r3.x, r3.y, r3.z), n1 = v3(r4.x, r4.y, r4.z), n2 = v3(r5.x, r5.y, r5.z);
	const V2 t0 = v2(r0.w, r1.w), t1 = v2(r2.w, r3.w), t2 = v2(r4.w, r5.w);
	p.material_idx = __float_as_uint(r6.y);
#else
	const uint4 rec = __ldg(&sc.tri_rec[prim_global]);
	const uint32_t mesh = rec.w;
	const lmb_vertex a0 = sc.vertices[rec.x];
	const lmb_vertex a1 = sc.vertices[rec.y];
	const lmb_vertex a2 = sc.vertices[rec.z];
	const V3 q0 = vtx_pos(a0), q1 = vtx_pos(a1), q2 = vtx_pos(a2);
	const V3 n0 = vtx_nrm(a0), n1 = vtx_nrm(a1), n2 = vtx_nrm(a2);
	const V2 t0 = v2(a0.uv0[0], a0.uv0[1]), t1 = v2(a1.uv0[0], a1.uv0[1]), t2 = v2(a2.uv0[0], a2.uv0[1]);
	p.material_idx = sc.prim_infos[mesh].material_index;
#endif
	const V3 bary = v3(1.0f - b1 - b2, b1, b2);
	const M4 o2w = load_m4(sc.world_matrices + 16 * mesh);
	const M4 w2o = load_m4(sc.inv_world_matrices + 16 * mesh);
	const V3 pos = q0 * bary.x + q1 * bary.y + q2 * bary.z;
	p.pos = xyz(mul(o2w, v4(pos, 1.0f)));
	const V3 nrm = normalize(n0 * bary.x + n1 * bary.y + n2 * bary.z);
	p.n_s = normalize(mul_row(nrm, w2o));
	p.uv = t0 * bary.x + t1 * bary.y + t2 * bary.z;
	const V3 e0 = q2 - q0;
	const V3 e1 = q1 - q0;
	p.n_g = normalize(mul_row(cross(e0, e1), w2o));
	p.triangle_idx = 0xFFFFFFFFu;
	p.instance_idx = mesh;
	return p;
}

struct LightSample {
	V3 Le, wi;
	float wi_len, pdf_w, pdf_a, cos_from_light;
	uint32_t flags, triangle_idx, instance_idx;
};

LMB_DN LightSample sample_light_Li(const DeviceScene& sc, const V4& rands, const V3& p, int num_lights) {
	LightSample o;
	o.Le = v3(0.0f), o.wi = v3(0.0f);
	o.wi_len = 0, o.pdf_w = 0, o.pdf_a = 0, o.cos_from_light = 0;
	o.triangle_idx = 0, o.instance_idx = 0;
	const uint32_t light_idx = (uint32_t)(rands.x * (float)num_lights);
	const lmb_light& light = sc.lights[light_idx];
	const uint32_t type = light.light_flags & 0x7u;
	o.flags = light.light_flags;
	switch (type) {
		case LMB_LIGHT_AREA: {
			const uint32_t pm = light.prim_mesh_idx;
			const lmb_prim_mesh_info& pinfo = sc.prim_infos[pm];
			const uint32_t material_idx = pinfo.material_index;
			o.triangle_idx = (uint32_t)(rands.y * (float)light.num_triangles);
			const M4 wm = load_m4(light.world_matrix);
			const M4 inv_tr = transpose(load_m4(sc.inv_world_matrices + 16 * pm));
			// sample_triangle, commons.glsl:112-149
			const uint32_t index_offset = pinfo.index_offset + 3 * o.triangle_idx;
			const uint32_t vo = pinfo.vertex_offset;
			const lmb_vertex a0 = sc.vertices[sc.indices[index_offset + 0] + vo];
			const lmb_vertex a1 = sc.vertices[sc.indices[index_offset + 1] + vo];
			const lmb_vertex a2 = sc.vertices[sc.indices[index_offset + 2] + vo];
			const V3 q0 = vtx_pos(a0), q1 = vtx_pos(a1), q2 = vtx_pos(a2);
			const V3 n0 = vtx_nrm(a0), n1 = vtx_nrm(a1), n2 = vtx_nrm(a2);
			const float sq = sqrtf(rands.z);
			const V2 uv = v2(1 - sq, rands.w * sq);
			const V3 bary = v3(1.0f - uv.x - uv.y, uv.x, uv.y);
			const V4 etmp0 = mul(wm, v4(q1 - q0, 1.0f));
			const V4 etmp1 = mul(wm, v4(q2 - q0, 1.0f));
			const V3 pos = q0 * bary.x + q1 * bary.y + q2 * bary.z;
			const V3 nrm = normalize(n0 * bary.x + n1 * bary.y + n2 * bary.z);
			const V4 world_pos = mul(wm, v4(pos, 1.0f));
			const V3 rec_n_s = normalize(xyz(mul(inv_tr, v4(nrm, 1.0f))));
			const float triangle_pdf = 2.0f / length(cross(xyz(etmp0), xyz(etmp1)));
			const V3 rec_pos = xyz(world_pos);
			const lmb_material light_mat = load_material(sc, material_idx, uv);
			o.wi = rec_pos - p;
			const float wi_len_sqr = dot(o.wi, o.wi);
			o.wi_len = sqrtf(wi_len_sqr);
			o.wi /= o.wi_len;
			o.cos_from_light = fabsf(dot(rec_n_s, -o.wi));
			o.Le = v3(light_mat.emissive_factor);
			o.pdf_a = triangle_pdf;
			o.pdf_w = o.pdf_a * wi_len_sqr / o.cos_from_light;
			o.instance_idx = pm;
		} break;
		case LMB_LIGHT_SPOT: {
			o.wi = v3(light.pos) - p;
			const float wi_len_sqr = dot(o.wi, o.wi);
			o.wi_len = sqrtf(wi_len_sqr);
			o.wi /= o.wi_len;
			const V3 light_dir = normalize(v3(light.to) - v3(light.pos));
			o.cos_from_light = dot(-o.wi, light_dir);
			const float cos_width = lmb_cosf(LMB_PI / 6);
			const float cos_faloff = lmb_cosf(25 * LMB_PI / 180);
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
		} break;
		case LMB_LIGHT_DIRECTIONAL: {
			const V3 dir = normalize(v3(light.pos) - v3(light.to));
			const V3 light_p = p + dir * (2 * light.world_radius);
			o.wi = light_p - p;
			o.wi_len = length(o.wi);
			o.wi /= o.wi_len;
			o.pdf_a = 1;
			o.pdf_w = 1;
			o.Le = v3(light.L);
			o.cos_from_light = 1.0f;
		} break;
		default:
			break;
	}
	return o;
}

LMB_D V3 shade_atmosphere(const DeviceScene& sc, uint32_t dir_light_idx, const V3& sky_col, const V3& ray_origin, const V3& ray_dir, float ray_length) {
	if (dir_light_idx == 0xFFFFFFFFu) return sky_col;
	const lmb_light& light = sc.lights[dir_light_idx];
	const V3 light_dir = -normalize(v3(light.to) - v3(light.pos));
	const V2 planet_isect = atmo::planet_intersection(ray_origin, ray_dir);
	if (planet_isect.x > 0) ray_length = gmin(ray_length, planet_isect.x);
	return atmo::integrate_scattering(ray_origin, ray_dir, ray_length, light_dir, v3(light.L));
}

}  // namespace lmb
