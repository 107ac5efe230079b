// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl_compat.h). CPU restatement of the acceleration structure and of
// traceRayEXT.
//
// The reference has no BVH or intersection source: both live in the Vulkan driver / RT cores
// (call sites: src/shaders/integrators/path/path.rgen:48, src/shaders/integrators/pt_commons.glsl:21,32;
// build sites: src/Framework/AccelerationStructure.cpp:171-315, src/RayTracer/Integrator.cpp:137-160).
// The Vulkan spec only promises a watertight test and a hit inside (tmin, tmax); tie-breaking is unspecified, and
// the reference has no test that pins a result at that boundary -> PARITY UNPINNED for this piece. This file
// therefore *defines* it:
//   * canonical LBVH: world-space triangles, 30-bit Morton code of the triangle-AABB centre, key = morton<<32 | id,
//     Karras 2012 radix tree, bottom-up min/max refit;
//   * watertight ray/triangle test (Woop, Benthin, Wald 2013) with a double-precision fallback on zero edge values;
//   * closest hit = smallest t in (tmin, tmax), ties broken by the lower global triangle id, so the answer does not
//     depend on traversal order or on the BVH that is traversed; any-hit = "some t in (tmin, tmax) exists".
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "glsl_compat.h"

namespace orc {

struct Lbvh {
	uint32_t n_tris = 0;
	std::vector<vec3> tri_world;       // 3 per global triangle id
	std::vector<uint32_t> tri_mesh;    // prim-mesh index (gl_InstanceCustomIndexEXT) per global id
	std::vector<uint32_t> tri_local;   // mesh-local triangle number (gl_PrimitiveID) per global id
	vec3 scene_min{0}, scene_max{0};
	std::vector<uint32_t> morton;      // per global id
	std::vector<uint64_t> sorted_keys; // per leaf position
	std::vector<uint32_t> leaf_prim;   // leaf position -> global id
	std::vector<uint32_t> left, right; // per internal node (N-1); ids >= N-1 are leaves (N-1 + leaf position)
	std::vector<uint32_t> parent;      // per node (2N-1); root = 0xFFFFFFFF
	std::vector<float> aabb;           // per node: min.xyz, max.xyz (6 floats)
};

struct TraceStats {
	uint64_t nodes = 0;  // internal nodes popped (each costs one 64-byte node fetch on the GPU layout)
	uint64_t tris = 0;   // triangle tests
};

inline uint32_t expand_bits10(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}

inline uint32_t quantize10(float c, float lo, float hi) {
	const float ext = hi - lo;
	if (!(ext > 0.0f)) return 0u;
	float q = (c - lo) / ext * 1024.0f;
	q = std::min(std::max(q, 0.0f), 1023.0f);
	return (uint32_t)q;
}

inline uint32_t morton30(const vec3& c, const vec3& lo, const vec3& hi) {
	const uint32_t x = quantize10(c.x, lo.x, hi.x), y = quantize10(c.y, lo.y, hi.y), z = quantize10(c.z, lo.z, hi.z);
	return (expand_bits10(x) << 2) | (expand_bits10(y) << 1) | expand_bits10(z);
}

inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }

inline void lbvh_build(const lmb_scene_desc& sd, Lbvh& b) {
	// 1. flatten to world space (Integrator.cpp:148-158: one instance per prim mesh, transform = world_matrix)
	uint32_t total = 0;
	for (uint32_t m = 0; m < sd.n_prim_meshes; m++) total += sd.prim_idx_counts[m] / 3;
	b.n_tris = total;
	b.tri_world.resize(3 * (size_t)total);
	b.tri_mesh.resize(total);
	b.tri_local.resize(total);
	uint32_t g = 0;
	for (uint32_t m = 0; m < sd.n_prim_meshes; m++) {
		const mat4 M = m4(sd.world_matrices + 16 * m);
		const lmb_prim_mesh_info& pi = sd.prim_infos[m];
		const uint32_t nt = sd.prim_idx_counts[m] / 3;
		for (uint32_t t = 0; t < nt; t++, g++) {
			for (int k = 0; k < 3; k++) {
				const uint32_t vi = sd.indices[pi.index_offset + 3 * t + k] + pi.vertex_offset;
				b.tri_world[3 * (size_t)g + k] = vec3(M * vec4(v3(sd.vertices[vi].pos), 1.0f));
			}
			b.tri_mesh[g] = m;
			b.tri_local[g] = t;
		}
	}
	const uint32_t N = total;
	b.morton.assign(N, 0);
	b.sorted_keys.assign(N, 0);
	b.leaf_prim.assign(N, 0);
	b.left.assign(N > 0 ? N - 1 : 0, 0);
	b.right.assign(N > 0 ? N - 1 : 0, 0);
	b.parent.assign(N > 0 ? 2 * N - 1 : 0, 0xFFFFFFFFu);
	b.aabb.assign(N > 0 ? 6 * (size_t)(2 * N - 1) : 0, 0.0f);
	if (N == 0) return;
	// 2. scene bounds
	vec3 lo(3.402823466e+38f), hi(-3.402823466e+38f);
	for (size_t i = 0; i < b.tri_world.size(); i++) {
		lo = glm::min(lo, b.tri_world[i]);
		hi = glm::max(hi, b.tri_world[i]);
	}
	b.scene_min = lo;
	b.scene_max = hi;
	// 3. morton keys
	std::vector<uint64_t> keys(N);
	for (uint32_t i = 0; i < N; i++) {
		const vec3 &a = b.tri_world[3 * (size_t)i], &bb = b.tri_world[3 * (size_t)i + 1], &c = b.tri_world[3 * (size_t)i + 2];
		const vec3 tmin = glm::min(glm::min(a, bb), c), tmax = glm::max(glm::max(a, bb), c);
		const vec3 cen = (tmin + tmax) * 0.5f;
		b.morton[i] = morton30(cen, lo, hi);
		keys[i] = ((uint64_t)b.morton[i] << 32) | i;
	}
	// 4. sort (keys unique -> any correct sort gives the same permutation)
	std::sort(keys.begin(), keys.end());
	for (uint32_t i = 0; i < N; i++) {
		b.sorted_keys[i] = keys[i];
		b.leaf_prim[i] = (uint32_t)(keys[i] & 0xFFFFFFFFu);
	}
	// 5. Karras radix tree
	auto delta = [&](int i, int j) -> int {
		if (j < 0 || j >= (int)N) return -1;
		return clz64(keys[i] ^ keys[j]);
	};
	for (int i = 0; i < (int)N - 1; i++) {
		const int d = (delta(i, i + 1) - delta(i, i - 1)) >= 0 ? 1 : -1;
		const int dmin = delta(i, i - d);
		int lmax = 2;
		while (delta(i, i + lmax * d) > dmin) lmax *= 2;
		int l = 0;
		for (int t = lmax / 2; t >= 1; t /= 2)
			if (delta(i, i + (l + t) * d) > dmin) l += t;
		const int j = i + l * d;
		const int dnode = delta(i, j);
		int s = 0, t = l;
		do {
			t = (t + 1) >> 1;
			if (delta(i, i + (s + t) * d) > dnode) s += t;
		} while (t > 1);
		const int gamma = i + s * d + std::min(d, 0);
		const uint32_t lc = (std::min(i, j) == gamma) ? (N - 1 + gamma) : (uint32_t)gamma;
		const uint32_t rc = (std::max(i, j) == gamma + 1) ? (N - 1 + gamma + 1) : (uint32_t)(gamma + 1);
		b.left[i] = lc;
		b.right[i] = rc;
		b.parent[lc] = i;
		b.parent[rc] = i;
	}
	// 6. refit: leaves, then internal nodes bottom-up via arrival counters (serial here)
	for (uint32_t i = 0; i < N; i++) {
		const uint32_t p = b.leaf_prim[i];
		const vec3 &a = b.tri_world[3 * (size_t)p], &bb = b.tri_world[3 * (size_t)p + 1], &c = b.tri_world[3 * (size_t)p + 2];
		const vec3 tmin = glm::min(glm::min(a, bb), c), tmax = glm::max(glm::max(a, bb), c);
		float* o = &b.aabb[6 * (size_t)(N - 1 + i)];
		o[0] = tmin.x, o[1] = tmin.y, o[2] = tmin.z, o[3] = tmax.x, o[4] = tmax.y, o[5] = tmax.z;
	}
	if (N > 1) {
		std::vector<uint8_t> arrived(N - 1, 0);
		for (uint32_t i = 0; i < N; i++) {
			uint32_t n = b.parent[N - 1 + i];
			while (n != 0xFFFFFFFFu) {
				if (!arrived[n]) {
					arrived[n] = 1;
					break;
				}
				const float* l = &b.aabb[6 * (size_t)b.left[n]];
				const float* r = &b.aabb[6 * (size_t)b.right[n]];
				float* o = &b.aabb[6 * (size_t)n];
				for (int k = 0; k < 3; k++) {
					o[k] = std::min(l[k], r[k]);
					o[3 + k] = std::max(l[3 + k], r[3 + k]);
				}
				n = b.parent[n];
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Ray / triangle and ray / box definitions (shared, expression for expression, with lumen_b200/csrc/trace.cuh)
// ---------------------------------------------------------------------------------------------------------------
struct RayPre {
	vec3 o;
	int kx, ky, kz;
	float Sx, Sy, Sz;
	vec3 inv;  // guarded reciprocal direction for slab tests
};

// Part of the hit definition: a ray with a NaN or infinite component in its origin or direction hits nothing (NaN in the slab and
// edge tests would make the result depend on the nodes a particular walk visits; the Vulkan spec leaves such rays undefined).
inline bool ray_finite(const vec3& o, const vec3& d) { return (o.x * 0.0f + o.y * 0.0f + o.z * 0.0f + d.x * 0.0f + d.y * 0.0f + d.z * 0.0f) == 0.0f; }

inline RayPre ray_prepare(const vec3& o, const vec3& d) {
	RayPre r;
	r.o = o;
	const float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
	int kz = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
	int kx = kz + 1;
	if (kx == 3) kx = 0;
	int ky = kx + 1;
	if (ky == 3) ky = 0;
	if (d[kz] < 0.0f) std::swap(kx, ky);
	r.kx = kx, r.ky = ky, r.kz = kz;
	r.Sx = d[kx] / d[kz];
	r.Sy = d[ky] / d[kz];
	r.Sz = 1.0f / d[kz];
	for (int k = 0; k < 3; k++) {
		const float dk = (std::fabs(d[k]) > 1e-20f) ? d[k] : std::copysign(1e-20f, d[k]);
		r.inv[k] = 1.0f / dk;
	}
	return r;
}

// Returns true and (t, b1, b2) when the triangle is hit at some t; range checks are the caller's.
inline bool tri_intersect(const RayPre& r, const vec3& v0, const vec3& v1, const vec3& v2, float& t, float& b1, float& b2) {
	const vec3 A = v0 - r.o, B = v1 - r.o, C = v2 - r.o;
	const float Ax = std::fmaf(-r.Sx, A[r.kz], A[r.kx]);
	const float Ay = std::fmaf(-r.Sy, A[r.kz], A[r.ky]);
	const float Bx = std::fmaf(-r.Sx, B[r.kz], B[r.kx]);
	const float By = std::fmaf(-r.Sy, B[r.kz], B[r.ky]);
	const float Cx = std::fmaf(-r.Sx, C[r.kz], C[r.kx]);
	const float Cy = std::fmaf(-r.Sy, C[r.kz], C[r.ky]);
	float U = Cx * By - Cy * Bx;
	float V = Ax * Cy - Ay * Cx;
	float W = Bx * Ay - By * Ax;
	if (U == 0.0f || V == 0.0f || W == 0.0f) {
		U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
		V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
		W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
	}
	if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
	const float det = U + V + W;
	if (det == 0.0f) return false;
	const float Az = r.Sz * A[r.kz], Bz = r.Sz * B[r.kz], Cz = r.Sz * C[r.kz];
	const float T = U * Az + V * Bz + W * Cz;
	t = T / det;
	b1 = V / det;
	b2 = W / det;
	return true;
}

// Third part of the hit definition. The watertight test decides WHETHER the ray hits the triangle; its t carries rounding error
// of the order eps * (extent of the triangle along the ray), which can put t a few ulp outside the slab interval of the
// triangle's own bounding box -- and then a walk that already holds a slightly larger t culls that box while another walk,
// arriving in a different order, finds the triangle: the closest hit would depend on the tree. So an accepted hit must
// also pass the slab test of the triangle's exact fp32 bounding box, and its t is clamped into [near, far * pad] of that
// box, evaluated with box_intersect's own arithmetic. Every box that encloses the triangle then has near_B <= near_T <= t
// and t <= far_T * pad <= far_B * pad (fp subtraction, multiplication by a fixed reciprocal, min and max are monotone),
// so no box test can cull a triangle whose t is inside (tmin, best]: the result is independent of tree and order.
inline bool tri_clamp_t(const RayPre& r, const vec3& v0, const vec3& v1, const vec3& v2, float& t) {
	const vec3 lo = glm::min(glm::min(v0, v1), v2), hi = glm::max(glm::max(v0, v1), v2);
	const float t0x = (lo.x - r.o.x) * r.inv.x, t1x = (hi.x - r.o.x) * r.inv.x;
	const float t0y = (lo.y - r.o.y) * r.inv.y, t1y = (hi.y - r.o.y) * r.inv.y;
	const float t0z = (lo.z - r.o.z) * r.inv.z, t1z = (hi.z - r.o.z) * r.inv.z;
	const float n = std::max(std::max(std::min(t0x, t1x), std::min(t0y, t1y)), std::min(t0z, t1z));
	const float f = std::min(std::min(std::max(t0x, t1x), std::max(t0y, t1y)), std::max(t0z, t1z)) * 1.0000004f;
	if (!(n <= f)) return false;
	t = std::min(std::max(t, n), f);
	return true;
}

inline bool box_intersect(const RayPre& r, const float* bb, float tmin, float tmax, float& tnear) {
	const float t0x = (bb[0] - r.o.x) * r.inv.x, t1x = (bb[3] - r.o.x) * r.inv.x;
	const float t0y = (bb[1] - r.o.y) * r.inv.y, t1y = (bb[4] - r.o.y) * r.inv.y;
	const float t0z = (bb[2] - r.o.z) * r.inv.z, t1z = (bb[5] - r.o.z) * r.inv.z;
	const float n = std::max(std::max(std::min(t0x, t1x), std::min(t0y, t1y)), std::max(std::min(t0z, t1z), tmin));
	float f = std::min(std::min(std::max(t0x, t1x), std::max(t0y, t1y)), std::min(std::max(t0z, t1z), tmax));
	f = f * 1.0000004f;
	tnear = n;
	return n <= f;
}

struct Hit {
	float t;
	float b1, b2;   // barycentric weights of v1, v2 (hitAttributeEXT vec2 attribs)
	uint32_t prim;  // global triangle id, 0xFFFFFFFF = miss
};

template <bool ANY>
inline Hit trace(const Lbvh& b, const vec3& o, const vec3& d, float tmin, float tmax, TraceStats* st) {
	Hit h{tmax, 0, 0, 0xFFFFFFFFu};
	const uint32_t N = b.n_tris;
	if (N == 0 || !ray_finite(o, d)) return h;
	const RayPre r = ray_prepare(o, d);
	auto leaf_test = [&](uint32_t leafpos) -> bool {
		const uint32_t p = b.leaf_prim[leafpos];
		if (st) st->tris++;
		float t, b1, b2;
		if (!tri_intersect(r, b.tri_world[3 * (size_t)p], b.tri_world[3 * (size_t)p + 1], b.tri_world[3 * (size_t)p + 2], t, b1, b2))
			return false;
		if (!tri_clamp_t(r, b.tri_world[3 * (size_t)p], b.tri_world[3 * (size_t)p + 1], b.tri_world[3 * (size_t)p + 2], t)) return false;
		if (!(t > tmin)) return false;
		if (t < h.t || (t == h.t && p < h.prim && h.prim != 0xFFFFFFFFu)) {
			h.t = t, h.b1 = b1, h.b2 = b2, h.prim = p;
			return true;
		}
		return false;
	};
	if (N == 1) {
		leaf_test(0);
		return h;
	}
	uint32_t stack[128];
	int sp = 0;
	uint32_t node = 0;
	for (;;) {
		if (st) st->nodes++;
		const uint32_t lc = b.left[node], rc = b.right[node];
		float tl, tr;
		// closest hit keeps boxes with tnear == best t alive (equal-t tie-break needs them)
		bool hl = box_intersect(r, &b.aabb[6 * (size_t)lc], tmin, h.t, tl);
		bool hr = box_intersect(r, &b.aabb[6 * (size_t)rc], tmin, h.t, tr);
		uint32_t next = 0xFFFFFFFFu;
		if (hl && lc >= N - 1) {
			if (leaf_test(lc - (N - 1)) && ANY) return h;
			hl = false;
		}
		if (hr && rc >= N - 1) {
			if (leaf_test(rc - (N - 1)) && ANY) return h;
			hr = false;
		}
		if (hl && hr) {
			if (tr < tl) {
				stack[sp++] = lc;
				next = rc;
			} else {
				stack[sp++] = rc;
				next = lc;
			}
		} else if (hl) {
			next = lc;
		} else if (hr) {
			next = rc;
		}
		if (next == 0xFFFFFFFFu) {
			if (sp == 0) break;
			next = stack[--sp];
		}
		node = next;
	}
	return h;
}

// The hit definition evaluated with no tree at all: every triangle, same tests, same acceptance rule. trace<false> must return
// exactly this (tests/test_oracle.py); it is what makes hits independent of tree shape and traversal order.
inline Hit trace_brute(const Lbvh& b, const vec3& o, const vec3& d, float tmin, float tmax) {
	Hit h{tmax, 0, 0, 0xFFFFFFFFu};
	if (b.n_tris == 0 || !ray_finite(o, d)) return h;
	const RayPre r = ray_prepare(o, d);
	for (uint32_t p = 0; p < b.n_tris; p++) {
		const vec3 &v0 = b.tri_world[3 * (size_t)p], &v1 = b.tri_world[3 * (size_t)p + 1], &v2 = b.tri_world[3 * (size_t)p + 2];
		float t, b1, b2;
		if (!tri_intersect(r, v0, v1, v2, t, b1, b2) || !tri_clamp_t(r, v0, v1, v2, t) || !(t > tmin)) continue;
		if (t < h.t || (t == h.t && p < h.prim && h.prim != 0xFFFFFFFFu)) h.t = t, h.b1 = b1, h.b2 = b2, h.prim = p;
	}
	return h;
}

}  // namespace orc
