// trace_wide.cuh -- persistent-thread traversal of the compressed 8-wide BVH (wide_bvh.cu) with warp-level dynamic ray
// fetch. Replaces traceRayEXT for all three ray kinds of the Path integrator (closest: path.rgen:48; any-hit shadow:
// pt_commons.glsl:21-22; MIS probe: pt_commons.glsl:32).
//
// One traversal step = one 80-byte node (five 128-bit loads) and eight slab tests done on the quantised planes directly:
//   t = q * 2^e * inv_d + (p - o) * inv_d      evaluated as  fma(1 + q * 2^-15, A, B),  A = 2^(e+15) * inv_d,
//   B = fma(p - o, inv_d, -A);  "1 + q * 2^-15" is one PRMT (the byte dropped into the mantissa of 1.0f), so a plane costs
//   PRMT + FFMA and no int->float conversion. Rounding of this form is < 1/256 grid step; the builder rounds boxes out by
//   >= 1/32 step on every side and the far side is padded by 8 ulp (the binary walk of trace.cuh pads by 3), so boxes stay
//   conservative.
// Hit children are collected in one 32-bit mask: bits 24..31 internal children in front-to-back order for this ray's
// octant (bit 24 + (slot ^ octant)), bits 0..23 the node's leaf triangles. The traversal stack holds (child_base, mask)
// groups: <= 1 push per level, LMB_WSTACK_SM entries per thread in shared memory laid out [entry][thread] (conflict free),
// deeper levels spill to local memory. Triangle tests, the definition of a hit and the tie-break are those of trace.cuh,
// bit for bit -- results are identical to the binary LBVH walk and to the CPU oracle.
#pragma once
#include "trace.cuh"

namespace lmb {

struct WideBvhView {
	const float4* nodes;  // 5 per node
	const float4* tris;   // 3 per triangle
	uint32_t n_tris;
};

#ifndef LMB_TRACE_THREADS
#define LMB_TRACE_THREADS 128
#endif
#define LMB_WSTACK_SM 12
#define LMB_WSTACK_LOCAL 52
#define LMB_WIDE_REFILL_LANES 20
#define LMB_WIDE_BLOCKS_PER_SM 6

template <typename Source>
__device__ __forceinline__ void trace_wide_persistent(const WideBvhView& bvh, Source& src, uint32_t count, uint32_t* cursor, unsigned long long* stats,
													  int stat_closest, int stat_any) {
	__shared__ uint2 s_stack[LMB_WSTACK_SM][LMB_TRACE_THREADS];
	uint2 l_stack[LMB_WSTACK_LOCAL];
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;

	bool has = false;        // this lane owns a ray
	bool exhausted = false;  // warp-uniform: the queue ran dry
	bool any = false;
	uint32_t item = 0;
	RayPre r;
	float tmin = 0.0f;
	Hit h{0.0f, 0.0f, 0.0f, 0xFFFFFFFFu};
	uint2 ng = make_uint2(0u, 0u);  // node group: (child_base, hits << 24 | imask)
	uint2 tg = make_uint2(0u, 0u);  // triangle group: (tri_base, mask)
	uint32_t oct_inv4 = 0;
	int sp = 0;
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
						src.load(i, o, d, tmin, tmax, any);
						item = i;
						r = ray_prepare(o, d);
						h = Hit{tmax, 0.0f, 0.0f, 0xFFFFFFFFu};
						sp = 0;
						oct_inv4 = ((r.inv.x < 0.0f ? 0u : 4u) | (r.inv.y < 0.0f ? 0u : 2u) | (r.inv.z < 0.0f ? 0u : 1u)) * 0x01010101u;
						ng = make_uint2(0u, bvh.n_tris ? 0x80000000u : 0u);
						tg = make_uint2(0u, 0u);
						has = true;
						if (any) n_any++;
						else n_closest++;
					}
				}
				if (base + (uint32_t)__popc(need) >= count) exhausted = true;
			}
		}
		if (__ballot_sync(0xFFFFFFFFu, has) == 0) break;

		for (;;) {
			// ---- one node step
			if (has && ng.y > 0x00FFFFFFu) {
				const uint32_t hits = ng.y;
				const int bit = 31 - __clz(hits);
				ng.y = hits & ~(1u << bit);
				const uint32_t slot = (uint32_t)(bit - 24) ^ (oct_inv4 & 7u);
				const uint32_t node = ng.x + __popc(hits & 0xFFu & ((1u << slot) - 1u));
				if (ng.y > 0x00FFFFFFu) {  // siblings still to visit
					if (sp < LMB_WSTACK_SM) s_stack[sp][tid] = ng;
					else l_stack[sp - LMB_WSTACK_SM] = ng;
					sp++;
				}
				n_nodes++;
				const float4* np = bvh.nodes + 5 * (size_t)node;
				const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
				const uint32_t ew = __float_as_uint(n0.w);
				const float ax = __uint_as_float(((ew & 0xFFu) + 15u) << 23) * r.inv.x;
				const float ay = __uint_as_float((((ew >> 8) & 0xFFu) + 15u) << 23) * r.inv.y;
				const float az = __uint_as_float((((ew >> 16) & 0xFFu) + 15u) << 23) * r.inv.z;
				const float bx = fmaf(n0.x - r.o.x, r.inv.x, -ax);
				const float by = fmaf(n0.y - r.o.y, r.inv.y, -ay);
				const float bz = fmaf(n0.z - r.o.z, r.inv.z, -az);
				const bool sx = r.inv.x < 0.0f, sy = r.inv.y < 0.0f, sz = r.inv.z < 0.0f;
				const uint32_t qlx[2] = {__float_as_uint(n2.x), __float_as_uint(n2.y)}, qly[2] = {__float_as_uint(n2.z), __float_as_uint(n2.w)};
				const uint32_t qlz[2] = {__float_as_uint(n3.x), __float_as_uint(n3.y)}, qhx[2] = {__float_as_uint(n3.z), __float_as_uint(n3.w)};
				const uint32_t qhy[2] = {__float_as_uint(n4.x), __float_as_uint(n4.y)}, qhz[2] = {__float_as_uint(n4.z), __float_as_uint(n4.w)};
				uint32_t hitmask = 0;
#pragma unroll
				for (int hf = 0; hf < 2; hf++) {
					const uint32_t meta = __float_as_uint(hf ? n1.w : n1.z);
					const uint32_t is_inner = (meta & (meta << 1)) & 0x10101010u;
					const uint32_t inner_mask = (is_inner >> 4) * 0xFFu;
					const uint32_t bit_index = (meta ^ (oct_inv4 & inner_mask)) & 0x1F1F1F1Fu;
					const uint32_t child_bits = (meta >> 5) & 0x07070707u;
					const uint32_t nx = sx ? qhx[hf] : qlx[hf], fx = sx ? qlx[hf] : qhx[hf];
					const uint32_t ny = sy ? qhy[hf] : qly[hf], fy = sy ? qly[hf] : qhy[hf];
					const uint32_t nz = sz ? qhz[hf] : qlz[hf], fz = sz ? qlz[hf] : qhz[hf];
#pragma unroll
					for (int j = 0; j < 4; j++) {
						const uint32_t sel = 0x7604u + (uint32_t)(j << 4);  // bytes (3F, 80, q_j, 00) = 1 + q_j * 2^-15
						const float tx0 = fmaf(__uint_as_float(__byte_perm(nx, 0x3F800000u, sel)), ax, bx);
						const float ty0 = fmaf(__uint_as_float(__byte_perm(ny, 0x3F800000u, sel)), ay, by);
						const float tz0 = fmaf(__uint_as_float(__byte_perm(nz, 0x3F800000u, sel)), az, bz);
						const float tx1 = fmaf(__uint_as_float(__byte_perm(fx, 0x3F800000u, sel)), ax, bx);
						const float ty1 = fmaf(__uint_as_float(__byte_perm(fy, 0x3F800000u, sel)), ay, by);
						const float tz1 = fmaf(__uint_as_float(__byte_perm(fz, 0x3F800000u, sel)), az, bz);
						const float tn = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, tmin));
						const float tf = fminf(fminf(tx1, ty1), fminf(tz1, h.t)) * 1.000001f;
						if (tn <= tf) hitmask |= ((child_bits >> (8 * j)) & 0xFFu) << ((bit_index >> (8 * j)) & 0xFFu);
					}
				}
				ng = make_uint2(__float_as_uint(n1.x), (hitmask & 0xFF000000u) | (ew >> 24));
				tg = make_uint2(__float_as_uint(n1.y), hitmask & 0x00FFFFFFu);
			}
			// ---- the node's leaf triangles
			while (tg.y != 0u) {
				const int bit = __ffs((int)tg.y) - 1;
				tg.y &= tg.y - 1u;
				const float4* tp = bvh.tris + 3 * (size_t)(tg.x + (uint32_t)bit);
				const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
				n_tris++;
				float t, b1, b2;
				if (tri_intersect(r, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t, b1, b2) && t > tmin) {
					const uint32_t p = __float_as_uint(a.w);
					if (t < h.t || (t == h.t && p < h.prim && h.prim != 0xFFFFFFFFu)) {
						h.t = t, h.b1 = b1, h.b2 = b2, h.prim = p;
						if (any) tg.y = 0u, ng.y = 0u, sp = 0;  // first accepted hit ends a shadow ray
					}
				}
			}
			// ---- next group, or done
			if (has && ng.y <= 0x00FFFFFFu) {
				if (sp > 0) {
					sp--;
					ng = sp < LMB_WSTACK_SM ? s_stack[sp][tid] : l_stack[sp - LMB_WSTACK_SM];
				} else {
					src.store(item, h, any);
					has = false;
				}
			}
			const int busy = __popc(__ballot_sync(0xFFFFFFFFu, has));
			if (busy == 0 || (!exhausted && busy < LMB_WIDE_REFILL_LANES)) break;
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
