// trace.cuh -- software replacement for traceRayEXT on sm_100a (B200 has no RT cores).
//
// Replaces: the closest-hit query of src/shaders/integrators/path/path.rgen:48 and pt_commons.glsl:32, the
// terminate-on-first-hit shadow query of pt_commons.glsl:21-22, and the sentinel writes of ray.rmiss / ray_shadow.rmiss /
// ray.rahit. Definition of a hit (shared with the CPU oracle, oracle/lbvh_cpu.h): watertight ray/triangle test (Woop et
// al. 2013) with fp64 fallback on zero edge functions; closest = min t in (tmin, tmax), equal t -> lower global triangle
// id; any-hit = exists t in (tmin, tmax). Box tests are conservative (slab test on (b - o) * inv_d with the far side
// padded by 3 ulp), so results do not depend on traversal order or tree shape.
//
// Node layout in HBM/L2 (64 B, four 128-bit loads): both child AABBs + two child references.
//   n0 = (l.min.x, l.min.y, l.min.z, l.max.x)   n1 = (l.max.y, l.max.z, r.min.x, r.min.y)
//   n2 = (r.min.z, r.max.x, r.max.y, r.max.z)   n3 = (left ref, right ref, -, -) ; ref >= 0 internal, ref < 0 leaf ~pos
// Leaf triangles: 3 x float4 (48 B) world-space vertices in leaf order, .w of v0 = global triangle id.
#pragma once
#include "vec.cuh"

namespace lmb {

struct BvhView {
	const float4* nodes;  // 4 per internal node
	const float4* tris;   // 3 per leaf position
	uint32_t n_tris;
};

struct RayPre {
	V3 o;
	int kx, ky, kz;
	float Sx, Sy, Sz;
	V3 inv;
};

LMB_D float guard_inv(float d) {
	const float dk = (fabsf(d) > 1e-20f) ? d : copysignf(1e-20f, d);
	return 1.0f / dk;
}

// Part of the hit definition (oracle/lbvh_cpu.h ray_finite): a ray with a NaN or infinite component in its origin or direction
// hits nothing. Slab and edge tests on NaN would otherwise make the result depend on which nodes a walk happens to visit.
LMB_D bool ray_finite(const V3& o, const V3& d) { return (o.x * 0.0f + o.y * 0.0f + o.z * 0.0f + d.x * 0.0f + d.y * 0.0f + d.z * 0.0f) == 0.0f; }

LMB_D RayPre ray_prepare(const V3& o, const V3& d) {
	RayPre r;
	r.o = o;
	const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
	int kz = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
	int kx = kz + 1;
	if (kx == 3) kx = 0;
	int ky = kx + 1;
	if (ky == 3) ky = 0;
	if (comp(d, kz) < 0.0f) {
		const int t = kx;
		kx = ky;
		ky = t;
	}
	r.kx = kx, r.ky = ky, r.kz = kz;
	const float dz = comp(d, kz);
	r.Sx = comp(d, kx) / dz;
	r.Sy = comp(d, ky) / dz;
	r.Sz = 1.0f / dz;
	r.inv = v3(guard_inv(d.x), guard_inv(d.y), guard_inv(d.z));
	return r;
}

LMB_D bool tri_intersect(const RayPre& r, const V3& v0, const V3& v1, const V3& v2, float& t, float& b1, float& b2) {
	const V3 A = v0 - r.o, B = v1 - r.o, C = v2 - r.o;
	const float Akz = comp(A, r.kz), Bkz = comp(B, r.kz), Ckz = comp(C, r.kz);
	const float Ax = fmaf(-r.Sx, Akz, comp(A, r.kx));
	const float Ay = fmaf(-r.Sy, Akz, comp(A, r.ky));
	const float Bx = fmaf(-r.Sx, Bkz, comp(B, r.kx));
	const float By = fmaf(-r.Sy, Bkz, comp(B, r.ky));
	const float Cx = fmaf(-r.Sx, Ckz, comp(C, r.kx));
	const float Cy = fmaf(-r.Sy, Ckz, comp(C, r.ky));
	float U = Cx * By - Cy * Bx;
	float V = Ax * Cy - Ay * Cx;
	float W = Bx * Ay - By * Ax;
	if (U == 0.0f || V == 0.0f || W == 0.0f) {
		U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
		V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
		W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
	}
	if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
	const float det = U + V + W;
	if (det == 0.0f) return false;
	const float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
	const float T = U * Az + V * Bz + W * Cz;
	t = T / det;
	b1 = V / det;
	b2 = W / det;
	return true;
}

// Same arithmetic as tri_intersect up to t; the barycentric divisions are left to the caller (b1 = V / det, b2 = W / det),
// who needs them only for an accepted closest hit.
LMB_D bool tri_intersect_t(const RayPre& r, const float4& p0, const float4& p1, const float4& p2, float& t, float& V_out, float& W_out, float& det_out) {
	const bool x0 = r.kx == 0, x1 = r.kx == 1, y0 = r.ky == 0, y1 = r.ky == 1, z0 = r.kz == 0, z1 = r.kz == 1;
	const float Ax_ = p0.x - r.o.x, Ay_ = p0.y - r.o.y, Az_ = p0.z - r.o.z;
	const float Bx_ = p1.x - r.o.x, By_ = p1.y - r.o.y, Bz_ = p1.z - r.o.z;
	const float Cx_ = p2.x - r.o.x, Cy_ = p2.y - r.o.y, Cz_ = p2.z - r.o.z;
	const float Akz = z0 ? Ax_ : (z1 ? Ay_ : Az_), Bkz = z0 ? Bx_ : (z1 ? By_ : Bz_), Ckz = z0 ? Cx_ : (z1 ? Cy_ : Cz_);
	const float Akx = x0 ? Ax_ : (x1 ? Ay_ : Az_), Bkx = x0 ? Bx_ : (x1 ? By_ : Bz_), Ckx = x0 ? Cx_ : (x1 ? Cy_ : Cz_);
	const float Aky = y0 ? Ax_ : (y1 ? Ay_ : Az_), Bky = y0 ? Bx_ : (y1 ? By_ : Bz_), Cky = y0 ? Cx_ : (y1 ? Cy_ : Cz_);
	const float Ax = fmaf(-r.Sx, Akz, Akx), Ay = fmaf(-r.Sy, Akz, Aky);
	const float Bx = fmaf(-r.Sx, Bkz, Bkx), By = fmaf(-r.Sy, Bkz, Bky);
	const float Cx = fmaf(-r.Sx, Ckz, Ckx), Cy = fmaf(-r.Sy, Ckz, Cky);
	float U = Cx * By - Cy * Bx;
	float V = Ax * Cy - Ay * Cx;
	float W = Bx * Ay - By * Ax;
	if (U == 0.0f || V == 0.0f || W == 0.0f) {
		U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
		V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
		W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
	}
	if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
	const float det = U + V + W;
	if (det == 0.0f) return false;
	const float T = U * (r.Sz * Akz) + V * (r.Sz * Bkz) + W * (r.Sz * Ckz);
	t = T / det;
	V_out = V, W_out = W, det_out = det;
	return true;
}

// Third part of the hit definition (oracle/lbvh_cpu.h tri_clamp_t, expression for expression): an accepted hit passes the slab
// test of the triangle's exact bounding box and its t is clamped into [near, far * pad] of that box, so that no box test on
// an enclosing box can cull a triangle that would beat the current best -- the closest hit does not depend on tree or order.
LMB_D bool tri_clamp_t(const V3& o, const V3& inv, const float4& p0, const float4& p1, const float4& p2, float& t) {
	const float lox = fminf(fminf(p0.x, p1.x), p2.x), hix = fmaxf(fmaxf(p0.x, p1.x), p2.x);
	const float loy = fminf(fminf(p0.y, p1.y), p2.y), hiy = fmaxf(fmaxf(p0.y, p1.y), p2.y);
	const float loz = fminf(fminf(p0.z, p1.z), p2.z), hiz = fmaxf(fmaxf(p0.z, p1.z), p2.z);
	const float t0x = (lox - o.x) * inv.x, t1x = (hix - o.x) * inv.x;
	const float t0y = (loy - o.y) * inv.y, t1y = (hiy - o.y) * inv.y;
	const float t0z = (loz - o.z) * inv.z, t1z = (hiz - o.z) * inv.z;
	const float n = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
	const float f = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z)) * 1.0000004f;
	if (!(n <= f)) return false;
	t = fminf(fmaxf(t, n), f);
	return true;
}

LMB_D bool box_intersect(const RayPre& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float tmin, float tmax, float& tnear) {
	const float t0x = (lox - r.o.x) * r.inv.x, t1x = (hix - r.o.x) * r.inv.x;
	const float t0y = (loy - r.o.y) * r.inv.y, t1y = (hiy - r.o.y) * r.inv.y;
	const float t0z = (loz - r.o.z) * r.inv.z, t1z = (hiz - r.o.z) * r.inv.z;
	const float n = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
	float f = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax));
	f = f * 1.0000004f;
	tnear = n;
	return n <= f;
}

struct Hit {
	float t, b1, b2;
	uint32_t prim;
};

#define LMB_STACK_SIZE 64

// One ray, one thread, explicit stack. ANY = true returns on the first accepted triangle.
// `nodes_visited` / `tris_tested` feed lmb_stats (algorithmic bytes per ray for the roofline).
template <bool ANY>
LMB_D Hit trace_ray(const BvhView& bvh, const V3& o, const V3& d, float tmin, float tmax, uint32_t& nodes_visited, uint32_t& tris_tested) {
	Hit h{tmax, 0.0f, 0.0f, 0xFFFFFFFFu};
	if (bvh.n_tris == 0 || !ray_finite(o, d)) return h;
	const RayPre r = ray_prepare(o, d);
	auto leaf_test = [&](int leafpos) -> bool {
		const float4 a = __ldg(&bvh.tris[3 * leafpos + 0]);
		const float4 b = __ldg(&bvh.tris[3 * leafpos + 1]);
		const float4 c = __ldg(&bvh.tris[3 * leafpos + 2]);
		tris_tested++;
		float t, b1, b2;
		if (!tri_intersect(r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, b1, b2)) return false;
		if (!tri_clamp_t(r.o, r.inv, a, b, c, t)) return false;
		if (!(t > tmin)) return false;
		const uint32_t p = __float_as_uint(a.w);
		if (t < h.t || (t == h.t && p < h.prim && h.prim != 0xFFFFFFFFu)) {
			h.t = t, h.b1 = b1, h.b2 = b2, h.prim = p;
			return true;
		}
		return false;
	};
	if (bvh.n_tris == 1) {
		leaf_test(0);
		return h;
	}
	int stack[LMB_STACK_SIZE];
	int sp = 0;
	int node = 0;
	for (;;) {
		nodes_visited++;
		const float4 n0 = __ldg(&bvh.nodes[4 * node + 0]);
		const float4 n1 = __ldg(&bvh.nodes[4 * node + 1]);
		const float4 n2 = __ldg(&bvh.nodes[4 * node + 2]);
		const float4 n3 = __ldg(&bvh.nodes[4 * node + 3]);
		const int lc = __float_as_int(n3.x), rc = __float_as_int(n3.y);
		float tl, tr;
		bool hl = box_intersect(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, tmin, h.t, tl);
		bool hr = box_intersect(r, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, tmin, h.t, tr);
		if (hl && lc < 0) {
			if (leaf_test(~lc) && ANY) return h;
			hl = false;
		}
		if (hr && rc < 0) {
			if (leaf_test(~rc) && ANY) return h;
			hr = false;
		}
		int next = -1;
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
		if (next < 0) {
			if (sp == 0) break;
			next = stack[--sp];
		}
		node = next;
	}
	return h;
}

}  // namespace lmb
