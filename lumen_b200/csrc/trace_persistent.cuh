// trace_persistent.cuh -- persistent-thread, while-while BVH traversal with warp-level dynamic ray fetch.
//
// Why this shape (ncu on the first version, profiles/r01_first_light.md): a one-ray-per-thread loop ran at 5.3 active
// threads per warp instruction because (a) a warp waits for its slowest ray and (b) triangle tests interrupt node loops.
// Here every warp keeps pulling rays from the shared queue with one atomic per refill as soon as fewer than
// LMB_REFILL_LANES lanes are busy, node steps and triangle tests run in separate inner loops (Aila & Laine 2009), and
// the traversal stack lives in shared memory laid out [entry][thread] (bank = thread -> conflict free for any depth mix).
// Closest-hit, any-hit and MIS-probe rays share the kernel; results are identical to trace_ray() (trace.cuh) because
// hits are defined independently of traversal order.
#pragma once
#include "trace.cuh"

namespace lmb {

#define LMB_TRACE_THREADS 128
#define LMB_REFILL_LANES 22
#define LMB_SENTINEL ((int)0x80000000)

// A Source provides: bool load(uint32_t i, V3& o, V3& d, float& tmin, float& tmax, bool& any)  (false -> skip entry)
//                    void store(uint32_t i, const Hit& h, bool any)
template <typename Source>
__device__ __forceinline__ void trace_persistent(const BvhView& bvh, Source& src, uint32_t count, uint32_t* cursor, unsigned long long* stats,
												 int stat_closest, int stat_any) {
	__shared__ int s_stack[LMB_STACK_SIZE][LMB_TRACE_THREADS];
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;

	bool has = false;        // this lane owns a ray
	bool exhausted = false;  // warp-uniform: the queue ran dry
	uint32_t item = 0;
	bool any = false;
	RayPre r;
	float tmin = 0.0f;
	Hit h{0.0f, 0.0f, 0.0f, 0xFFFFFFFFu};
	int cur = LMB_SENTINEL, sp = 0;
	uint32_t n_nodes = 0, n_tris = 0, n_closest = 0, n_any = 0;

	for (;;) {
		// ---- refill: lanes without a ray take consecutive queue entries (one atomic per warp)
		if (!exhausted) {
			const uint32_t need = __ballot_sync(0xFFFFFFFFu, !has);
			if (need) {
				const int leader = __ffs(need) - 1;
				uint32_t base = 0;
				if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(need));
				base = __shfl_sync(0xFFFFFFFFu, base, leader);
				if (!has) {
					const uint32_t i = base + __popc(need & lt_mask);
					if (i < count) {
						V3 o, d;
						float tmax;
						src.load(i, o, d, tmin, tmax, item);  // item = the source's tag of the ray
						any = src.is_any(item);
						r = ray_prepare(o, d);
						h = Hit{tmax, 0.0f, 0.0f, 0xFFFFFFFFu};
						sp = 0;
						cur = (bvh.n_tris == 0 || !ray_finite(o, d)) ? LMB_SENTINEL : (bvh.n_tris == 1 ? ~0 : 0);
						has = true;
						if (any) n_any++;
						else n_closest++;
					}
				}
				if (base + (uint32_t)__popc(need) >= count) exhausted = true;
			}
		}
		if (__ballot_sync(0xFFFFFFFFu, has) == 0) break;

		// ---- traverse until too few lanes are busy
		for (;;) {
			// inner loop 1: internal nodes
			while (has && cur >= 0) {
				n_nodes++;
				const float4 n0 = __ldg(&bvh.nodes[4 * cur + 0]);
				const float4 n1 = __ldg(&bvh.nodes[4 * cur + 1]);
				const float4 n2 = __ldg(&bvh.nodes[4 * cur + 2]);
				const float4 n3 = __ldg(&bvh.nodes[4 * cur + 3]);
				const int lc = __float_as_int(n3.x), rc = __float_as_int(n3.y);
				float tl, tr;
				const bool hl = box_intersect(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, tmin, h.t, tl);
				const bool hr = box_intersect(r, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, tmin, h.t, tr);
				if (hl && hr) {
					const bool right_first = tr < tl;
					s_stack[sp++][tid] = right_first ? lc : rc;
					cur = right_first ? rc : lc;
				} else if (hl) {
					cur = lc;
				} else if (hr) {
					cur = rc;
				} else {
					cur = sp > 0 ? s_stack[--sp][tid] : LMB_SENTINEL;
				}
			}
			// inner loop 2: one postponed leaf (one triangle per leaf)
			if (has && cur != LMB_SENTINEL) {
				const int leafpos = ~cur;
				const float4 a = __ldg(&bvh.tris[3 * leafpos + 0]);
				const float4 b = __ldg(&bvh.tris[3 * leafpos + 1]);
				const float4 c = __ldg(&bvh.tris[3 * leafpos + 2]);
				n_tris++;
				float t, b1, b2;
				bool accepted = false;
				if (tri_intersect(r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, b1, b2) && tri_clamp_t(r.o, r.inv, a, b, c, t) && t > tmin) {
					const uint32_t p = __float_as_uint(a.w);
					if (t < h.t || (t == h.t && p < h.prim && h.prim != 0xFFFFFFFFu)) {
						h.t = t, h.b1 = b1, h.b2 = b2, h.prim = p;
						accepted = true;
					}
				}
				cur = (accepted && any) ? LMB_SENTINEL : (sp > 0 ? s_stack[--sp][tid] : LMB_SENTINEL);
			}
			if (has && cur == LMB_SENTINEL) {
				src.store(item, h);
				has = false;
			}
			const int busy = __popc(__ballot_sync(0xFFFFFFFFu, has));
			if (busy == 0 || (!exhausted && busy < LMB_REFILL_LANES)) break;
		}
	}
	// ---- statistics (one atomic per warp and counter)
	for (int o = 16; o > 0; o >>= 1) {
		n_nodes += __shfl_xor_sync(0xFFFFFFFFu, n_nodes, o);
		n_tris += __shfl_xor_sync(0xFFFFFFFFu, n_tris, o);
		n_closest += __shfl_xor_sync(0xFFFFFFFFu, n_closest, o);
		n_any += __shfl_xor_sync(0xFFFFFFFFu, n_any, o);
	}
	if (lane == 0 && stats) {
		if (n_nodes) atomicAdd(&stats[ST_NODES], (unsigned long long)n_nodes);
		if (n_tris) atomicAdd(&stats[ST_TRIS], (unsigned long long)n_tris);
		if (n_closest && stat_closest >= 0) atomicAdd(&stats[stat_closest], (unsigned long long)n_closest);
		if (n_any && stat_any >= 0) atomicAdd(&stats[stat_any], (unsigned long long)n_any);
	}
}

}  // namespace lmb
