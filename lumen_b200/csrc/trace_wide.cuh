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
// A child is entered iff tn <= tf * pad; the sign bit of fma(tf, pad, -tn) says exactly that (the box test's tmin is kept strictly
// positive, so the result is never -0) and a funnel shift per child collects the eight sign bits: FFMA + SHF instead of FMUL +
// FSETP + SEL + IADD3. The kernel sits on the ALU pipe (PRMT, FMNMX, LOP, SEL: ncu 71 % against 24 % on the FMA pipe), so work
// is moved to the FMA pipe where the arithmetic allows it.
// The eight box tests give one 8-bit mask in slot order; a handful of bit operations per NODE (not per child) turn it into
// the internal-children group (bit 24 + (slot ^ octant): highest bit = nearest child) and the 24-bit leaf-triangle mask. The traversal stack holds (child_base, mask)
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
	uint32_t one_bits;  // 0x3F800000, passed as DATA: keeps the PRMT byte selectors of the slab test immediates (no MOV per plane)
	uint32_t refill_lanes;     // lanes pop new rays once fewer than this many are busy (LMB_WIDE_REFILL_LANES)
	uint32_t tri_round_lanes;  // a triangle round starts once this many lanes hold triangles (LMB_TRI_ROUND_LANES)
};

#ifndef LMB_TRACE_THREADS
#define LMB_TRACE_THREADS 128
#endif
// Entries of the traversal stack kept in shared memory (the rest, up to 64, spill to local memory). Shared memory is taken from the
// L1: with 11 entries, 8 blocks needed the 228 KB carve-out and left the L1 28 KB -- the 10 M-triangle grid ran at 1483 Mrays/s;
// 6 entries fit the 196 KB step (L1 60 KB): 2239 Mrays/s, classroom stand-in trace 59.0 -> 58.5 ms (profiles/r02b_summary.md).
// The unpinned instantiation (7 blocks per SM, wavefront.cu) keeps 4 and lands on the 164 KB step.
#ifndef LMB_WSTACK_SM
#define LMB_WSTACK_SM 6
#endif
#ifndef LMB_WSTACK_SM_UNPINNED
#define LMB_WSTACK_SM_UNPINNED 4
#endif
#define LMB_WSTACK_TOTAL 64
#ifndef LMB_WIDE_REFILL_LANES
#define LMB_WIDE_REFILL_LANES 28
#endif
#ifndef LMB_WIDE_BLOCKS_PER_SM
#define LMB_WIDE_BLOCKS_PER_SM 8
#endif
#ifndef LMB_WIDE_BLOCKS_PER_SM_UNPINNED
#define LMB_WIDE_BLOCKS_PER_SM_UNPINNED 7
#endif
#ifndef LMB_TRI_ROUND_LANES
#define LMB_TRI_ROUND_LANES 10  // pinned walker (BVH in the L2): 8 / 9 / 10 / 11 / 12 = trace 58.56 / 58.22 / 58.28 / 58.65 / 59.35 ms
#endif
#ifndef LMB_TRI_ROUND_LANES_UNPINNED
#define LMB_TRI_ROUND_LANES_UNPINNED 8  // 10 M-triangle grid: 2338 Mrays/s with 8, 2309 with 10
#endif
#ifndef LMB_DEFER_STORE
#define LMB_DEFER_STORE 1
#endif
#ifndef LMB_TRACE_WCOUNT
#define LMB_TRACE_WCOUNT 1  // warp-level scheduling counters kept by the walker: 0 none, 1 loop trips, 3 + triangle rounds and refills
#endif
#ifndef LMB_TRACE_FFMA2
#define LMB_TRACE_FFMA2 1
#endif
#ifndef LMB_PREFETCH
#define LMB_PREFETCH 0  // 1: leaf triangles, 2: + entered children, of the unpinned (BVH beyond the L2) instantiation
#endif

// Ray constants of the triangle test as one lane publishes them for the whole warp (see the triangle phase below).
struct TriRay {
	V3 o;
	float tmin, Sx, Sy, Sz;
	uint32_t k;  // kx | ky << 2 | kz << 4
	V3 inv;      // guarded 1/d (tri_clamp_t)
};

// tri_intersect + tri_clamp_t (trace.cuh) on a TriRay: same expressions, same order; returns t only, the caller divides V / det,
// W / det. The slab interval of the triangle's box is taken from the origin-relative vertices the edge test needs anyway:
// min(a - o, b - o, c - o) == min(a, b, c) - o bit for bit (fp subtraction is monotone), so only (near, far) stay live.
LMB_D bool tri_test(const TriRay& r, const float4& p0, const float4& p1, const float4& p2, float& t, float& V_out, float& W_out, float& det_out) {
	const uint32_t kx = r.k & 3u, ky = (r.k >> 2) & 3u, kz = r.k >> 4;
	const bool x0 = kx == 0, x1 = kx == 1, y0 = ky == 0, y1 = ky == 1, z0 = kz == 0, z1 = kz == 1;
	const float Ax_ = p0.x - r.o.x, Ay_ = p0.y - r.o.y, Az_ = p0.z - r.o.z;
	const float Bx_ = p1.x - r.o.x, By_ = p1.y - r.o.y, Bz_ = p1.z - r.o.z;
	const float Cx_ = p2.x - r.o.x, Cy_ = p2.y - r.o.y, Cz_ = p2.z - r.o.z;
	const float t0x = fminf(fminf(Ax_, Bx_), Cx_) * r.inv.x, t1x = fmaxf(fmaxf(Ax_, Bx_), Cx_) * r.inv.x;
	const float t0y = fminf(fminf(Ay_, By_), Cy_) * r.inv.y, t1y = fmaxf(fmaxf(Ay_, By_), Cy_) * r.inv.y;
	const float t0z = fminf(fminf(Az_, Bz_), Cz_) * r.inv.z, t1z = fmaxf(fmaxf(Az_, Bz_), Cz_) * r.inv.z;
	const float box_n = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
	const float box_f = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z)) * 1.0000004f;
	if (!(box_n <= box_f)) return false;
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
	t = fminf(fmaxf(T / det, box_n), box_f);
	V_out = V, W_out = W, det_out = det;
	return true;
}

// Shared memory of one traversal block.
template <int WS>
struct TraceSmem {
	uint2 stack[WS][LMB_TRACE_THREADS];
	float4 ray_a[LMB_TRACE_THREADS];  // o.xyz, tmin            } TriRay of the ray each lane owns, written once per ray
	float4 ray_b[LMB_TRACE_THREADS];  // Sx, Sy, Sz, k (bits)   }
	float4 ray_c[LMB_TRACE_THREADS];  // guarded 1/d.xyz (tri_clamp_t)
	uint32_t pair[LMB_TRACE_THREADS]; // owner lane << 27 | triangle index: one triangle test of this round
	float4 res[LMB_TRACE_THREADS];    // t (or -1), V, W, det
	uint32_t res_prim[LMB_TRACE_THREADS];
	// per-warp ready queue: 32 rays fetched and prepared by the whole warp at once, handed to lanes as they run dry
	float4 rq_a[LMB_TRACE_THREADS];   // o.xyz, tmin
	float4 rq_b[LMB_TRACE_THREADS];   // Sx, Sy, Sz, k (bits)
	float4 rq_c[LMB_TRACE_THREADS];   // 1/d.xyz, tmax
	uint32_t rq_i[LMB_TRACE_THREADS]; // the source's tag of the ray
};

template <bool PIN, typename Source>
__device__ __forceinline__ void trace_wide_persistent(const WideBvhView& bvh, Source& src, uint32_t count, uint32_t* cursor, unsigned long long* stats,
													  int stat_closest, int stat_any) {
	constexpr int WS = PIN ? LMB_WSTACK_SM : LMB_WSTACK_SM_UNPINNED;
	__shared__ TraceSmem<WS> sm;
	uint2 l_stack[LMB_WSTACK_TOTAL - WS];
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int wbase = tid & ~31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const uint32_t one_bits = bvh.one_bits;
	// PIN: the shared-memory address of this thread's stack column is pinned in a register; left to itself the compiler rebuilds it at
	// every push and pop from two special-register reads (S2R SR_CgaCtaId, SR_TID.X) to save that register. Pinning takes ~10
	// instructions off a node step, which pays when the kernel is issue bound (BVH resident in L2: classroom stand-in trace -2 %), and
	// costs 11 % when it is bound by memory latency (10 M-triangle torus grid, 590 MB of nodes + triangles: 1455 -> 1290 Mrays/s with
	// any way of pinning, inline PTX or an opaque C++ pointer -- the register it takes is then missed by loads in flight). The host
	// picks the instantiation by the footprint of the BVH against the L2 (wavefront.cu trace_pinned).
	uint32_t stack_base = 0;
	if (PIN) asm volatile("mov.u32 %0, %1;" : "=r"(stack_base) : "r"((uint32_t)__cvta_generic_to_shared(&sm.stack[0][tid])));

	bool has = false;        // this lane owns a ray that is still being traced
	bool done = false;       // this lane's ray is finished and waits for its store (LMB_DEFER_STORE)
	bool exhausted = false;  // warp-uniform: the global queue ran dry
	uint32_t rq_head = 0, rq_count = 0;  // warp-uniform: the ready queue holds entries [rq_head, rq_count)
	bool any = false;
	uint32_t item = 0;
	V3 ro = v3(0.0f), rinv = v3(1.0f);
	float tmin_box = 0.0f;  // tmin of the box test
	Hit h{0.0f, 0.0f, 0.0f, 0xFFFFFFFFu};
	float det = 1.0f;  // h.b1, h.b2 hold V, W of the current best hit until the ray is done
	uint2 ng = make_uint2(0u, 0u);  // node group: (child_base, hits << 24 | imask)
	uint2 tg = make_uint2(0u, 0u);  // triangle group: (tri_base, mask over the node's 24 leaf-triangle bits)
	uint32_t tl = 0;                // leaf24 of the node tg came from: triangle of bit b = tg.x + popc(tl & ((1 << b) - 1))
	uint32_t oct_inv4 = 0;
	int sp = 0;
	uint32_t n_nodes = 0, n_tris = 0, n_closest = 0, n_any = 0;
	uint32_t w_iters = 0, w_rounds = 0, w_refills = 0;  // warp-uniform scheduling counters (lmb_stats.trace_*): one IADD each
#ifdef LMB_TRACE_PROFILE
	uint32_t p_iters = 0, p_node_trips = 0, p_node_lanes = 0, p_has = 0, p_parked = 0, p_rounds = 0, p_pairs = 0, p_refills = 0;  // lane 0 only
#endif

	for (;;) {
#if LMB_DEFER_STORE
		// ---- finished rays: the two IEEE divisions of the barycentrics (~40 instructions with their slow path) and the store ran at
		// 2.3 of 32 lanes in 60 % of the loop trips when every lane finished its ray on its own (profiles/r02a: 7 % of the kernel's warp
		// instructions); here the lanes that finished since the last refill (>= 32 - refill_lanes of them, or all at the end) do it together
		if (done) {
			if (!any && h.prim != 0xFFFFFFFFu) h.b1 = h.b1 / det, h.b2 = h.b2 / det;
			src.store(item, h);
			done = false;
		}
#endif
		// ---- refill. Rays come out of the global queue 32 at a time: the whole warp fetches and prepares them (coalesced
		// loads, ray_prepare at 32 of 32 lanes, one cursor atomic per 32 rays) into the warp's ready queue in shared memory;
		// lanes that ran dry pop from it for the price of a few shared loads, so a refill is worth doing for a handful of
		// idle lanes instead of a dozen (profiles/r01f: 20 of 32 lanes active with refill-in-place below 20 busy lanes).
		for (;;) {
			const uint32_t need = __ballot_sync(0xFFFFFFFFu, !has);
			if (need == 0u) break;
			if (rq_head == rq_count) {
				if (exhausted) break;
				uint32_t base = 0;
#ifdef LMB_TRACE_PROFILE
				p_refills++;
#endif
#if LMB_TRACE_WCOUNT > 1
				w_refills++;
#endif
				if (lane == 0) base = atomicAdd(cursor, 32u);
				base = __shfl_sync(0xFFFFFFFFu, base, 0);
				rq_head = 0, rq_count = base < count ? min(32u, count - base) : 0u;
				if (base + 32u >= count) exhausted = true;
				if ((uint32_t)lane < rq_count) {
					V3 o, d;
					float tmin_, tmax_;
					uint32_t tag_;  // what the source needs to store the result (the queue ENTRY, not its index: no reload at the end)
					src.load(base + lane, o, d, tmin_, tmax_, tag_);
					const RayPre r = ray_prepare(o, d);
					sm.rq_a[tid] = make_float4(r.o.x, r.o.y, r.o.z, tmin_);
					// bit 6 of the axis word: non-finite ray, hits nothing by definition (ray_finite, trace.cuh)
					sm.rq_b[tid] = make_float4(r.Sx, r.Sy, r.Sz, __uint_as_float((uint32_t)r.kx | ((uint32_t)r.ky << 2) | ((uint32_t)r.kz << 4) | (ray_finite(o, d) ? 0u : 64u)));
					sm.rq_c[tid] = make_float4(r.inv.x, r.inv.y, r.inv.z, tmax_);
					sm.rq_i[tid] = tag_;
				}
				__syncwarp();
				if (rq_count == 0u) break;
			}
			const uint32_t avail = rq_count - rq_head;
			const uint32_t rank = (uint32_t)__popc(need & lt_mask);
			if (!has && rank < avail) {
				const int q = wbase + (int)(rq_head + rank);
				const float4 qa = sm.rq_a[q], qb = sm.rq_b[q], qc = sm.rq_c[q];
				const uint32_t qi = sm.rq_i[q];
				item = qi, any = src.is_any(qi);  // the source's tag travels with the ray
				ro = v3(qa.x, qa.y, qa.z), rinv = v3(qc.x, qc.y, qc.z);
				tmin_box = fmaxf(qa.w, __uint_as_float(1u));  // strictly positive (smallest denormal): any accepted t > tmin >= 0 is at least that
				sm.ray_a[tid] = qa, sm.ray_b[tid] = qb, sm.ray_c[tid] = qc;
				h = Hit{qc.w, 0.0f, 0.0f, 0xFFFFFFFFu};
				sp = 0;
				oct_inv4 = ((rinv.x < 0.0f ? 0u : 4u) | (rinv.y < 0.0f ? 0u : 2u) | (rinv.z < 0.0f ? 0u : 1u)) * 0x01010101u;
				ng = make_uint2(0u, (bvh.n_tris && !(__float_as_uint(qb.w) & 64u)) ? 0x80000000u : 0u);
				tg = make_uint2(0u, 0u);
				has = true;
				if (any) n_any++;
				else n_closest++;
			}
			rq_head += min((uint32_t)__popc(need), avail);
			__syncwarp();
		}
		if (__ballot_sync(0xFFFFFFFFu, has) == 0) break;

		for (;;) {
#ifdef LMB_TRACE_PROFILE
			{
				const uint32_t st = __ballot_sync(0xFFFFFFFFu, has && tg.y == 0u && ng.y > 0x00FFFFFFu);
				p_iters++, p_node_trips += st != 0u, p_node_lanes += __popc(st), p_has += __popc(__ballot_sync(0xFFFFFFFFu, has));
			}
#endif
			// ---- one node step
#if LMB_TRACE_WCOUNT > 0
			w_iters++;
#endif
			if (has && tg.y == 0u && ng.y > 0x00FFFFFFu) {
				const uint32_t hits = ng.y;
				const int bit = 31 - __clz(hits);
				ng.y = hits & ~(1u << bit);
				const uint32_t slot = (uint32_t)(bit - 24) ^ (oct_inv4 & 7u);
				const uint32_t node = ng.x + __popc(hits & 0xFFu & ((1u << slot) - 1u));
				if (ng.y > 0x00FFFFFFu) {  // siblings still to visit
					if (sp < WS) {
						if (PIN) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(stack_base + (uint32_t)sp * (LMB_TRACE_THREADS * 8u)), "r"(ng.x), "r"(ng.y));
						else sm.stack[sp][tid] = ng;
					}
					else l_stack[sp - WS] = ng;
					sp++;
				}
#ifndef LMB_TRACE_NO_STATS
				n_nodes++;
#endif
				const float4* np = bvh.nodes + 5 * (size_t)node;
				const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
				const uint32_t ew = __float_as_uint(n0.w);  // (ex + 15) | (ey + 15) << 8 | (ez + 15) << 16
				const float ax = __uint_as_float((ew & 0xFFu) << 23) * rinv.x;  // the exponent bytes carry the + 15
				const float ay = __uint_as_float(((ew >> 8) & 0xFFu) << 23) * rinv.y;
				const float az = __uint_as_float(((ew >> 16) & 0xFFu) << 23) * rinv.z;
				const float bx = fmaf(n0.x - ro.x, rinv.x, -ax);
				const float by = fmaf(n0.y - ro.y, rinv.y, -ay);
				const float bz = fmaf(n0.z - ro.z, rinv.z, -az);
				const bool sx = rinv.x < 0.0f, sy = rinv.y < 0.0f, sz = rinv.z < 0.0f;
				const uint32_t qlx[2] = {__float_as_uint(n2.x), __float_as_uint(n2.y)}, qly[2] = {__float_as_uint(n2.z), __float_as_uint(n2.w)};
				const uint32_t qlz[2] = {__float_as_uint(n3.x), __float_as_uint(n3.y)}, qhx[2] = {__float_as_uint(n3.z), __float_as_uint(n3.w)};
				const uint32_t qhy[2] = {__float_as_uint(n4.x), __float_as_uint(n4.y)}, qhz[2] = {__float_as_uint(n4.z), __float_as_uint(n4.w)};
				uint32_t hits8 = 0;  // bit s: the ray enters the box of slot s
#pragma unroll
				for (int hf = 1; hf >= 0; hf--) {
					const uint32_t nx = sx ? qhx[hf] : qlx[hf], fx = sx ? qlx[hf] : qhx[hf];
					const uint32_t ny = sy ? qhy[hf] : qly[hf], fy = sy ? qly[hf] : qhy[hf];
					const uint32_t nz = sz ? qhz[hf] : qlz[hf], fz = sz ? qlz[hf] : qhz[hf];
#pragma unroll
					for (int j = 3; j >= 0; j--) {
						const uint32_t sel = 0x7604u + (uint32_t)(j << 4);  // bytes (3F, 80, q_j, 00) = 1 + q_j * 2^-15
#if LMB_TRACE_FFMA2
						// near and far plane of one axis share A and B: one packed FFMA2 (fma.rn.f32x2, sm_100; each half an IEEE fma, same values)
						// per axis and child instead of two FFMA. The FMA pipe does the same work (FFMA2 = 2 FFMA of pipe time, measured:
						// tools/micro/ffma2_rate.cu); what is saved is the issue slot, which this kernel is short of: trace 58.24 -> 57.75 ms.
						// (Packing the entry test of two children as well, with negated near planes, spilled at 64 registers: 64.6 ms.)
						const float2 tx = __ffma2_rn(make_float2(__uint_as_float(__byte_perm(nx, one_bits, sel)), __uint_as_float(__byte_perm(fx, one_bits, sel))), make_float2(ax, ax), make_float2(bx, bx));
						const float2 ty = __ffma2_rn(make_float2(__uint_as_float(__byte_perm(ny, one_bits, sel)), __uint_as_float(__byte_perm(fy, one_bits, sel))), make_float2(ay, ay), make_float2(by, by));
						const float2 tz = __ffma2_rn(make_float2(__uint_as_float(__byte_perm(nz, one_bits, sel)), __uint_as_float(__byte_perm(fz, one_bits, sel))), make_float2(az, az), make_float2(bz, bz));
						const float tx0 = tx.x, tx1 = tx.y, ty0 = ty.x, ty1 = ty.y, tz0 = tz.x, tz1 = tz.y;
#else
						const float tx0 = fmaf(__uint_as_float(__byte_perm(nx, one_bits, sel)), ax, bx);
						const float ty0 = fmaf(__uint_as_float(__byte_perm(ny, one_bits, sel)), ay, by);
						const float tz0 = fmaf(__uint_as_float(__byte_perm(nz, one_bits, sel)), az, bz);
						const float tx1 = fmaf(__uint_as_float(__byte_perm(fx, one_bits, sel)), ax, bx);
						const float ty1 = fmaf(__uint_as_float(__byte_perm(fy, one_bits, sel)), ay, by);
						const float tz1 = fmaf(__uint_as_float(__byte_perm(fz, one_bits, sel)), az, bz);
#endif
						const float tn = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, tmin_box));
						const float tf = fminf(fminf(tx1, ty1), fminf(tz1, h.t));
						// entered <=> tn <= tf * pad. The sign of fma(tf, pad, -tn) says exactly that (tn > 0, so the result is never -0), one
						// FFMA on the FMA pipe instead of FMUL + FSETP, and a funnel shift collects the eight sign bits (slot 7 first) instead
						// of SEL + IADD3 per child: 12 fewer instructions on the ALU pipe per node, the pipe this kernel sits on.
						hits8 = __funnelshift_l(__float_as_uint(fmaf(tf, 1.000001f, -tn)), hits8, 1);
					}
				}
				hits8 = ~hits8 & 0xFFu;
				// empty slots carry an inverted box (qlo 255, qhi 0) and can never be entered
				const uint32_t mw = __float_as_uint(n1.z);
				const uint32_t imask = mw >> 24;
				// internal children: bit (slot ^ octant) so that the highest set bit is the nearest child (front to back)
				uint32_t ih = hits8 & imask;
				if (oct_inv4 & 1u) ih = ((ih & 0xAAu) >> 1) | ((ih & 0x55u) << 1);
				if (oct_inv4 & 2u) ih = ((ih & 0xCCu) >> 2) | ((ih & 0x33u) << 2);
				if (oct_inv4 & 4u) ih = ((ih & 0xF0u) >> 4) | ((ih & 0x0Fu) << 4);
				// leaf children: bit s -> bits 3s..3s+2, masked by the node's per-slot triangle counts
				uint32_t lh = hits8 & ~imask;
				lh = (lh | (lh << 8)) & 0x0000F00Fu;
				lh = (lh | (lh << 4)) & 0x000C30C3u;
				lh = (lh | (lh << 2)) & 0x00249249u;
				tl = mw & 0x00FFFFFFu;
				ng = make_uint2(__float_as_uint(n1.x), (ih << 24) | imask);
				tg = make_uint2(__float_as_uint(n1.y), (lh * 7u) & tl);
#if LMB_PREFETCH
				if (!PIN) {
					// The unpinned instantiation runs when nodes + triangles exceed the L2 (wavefront.cu trace_pinned): the walk then waits on
					// DRAM at every level (10 M-triangle grid: long-scoreboard stalls 6 warps per issue, DRAM 14 % of peak). What this node step
					// just found out is fetched ahead: the leaf triangles of the coming round (contiguous per node: one or two lines) and the
					// entered children (the 8 children of a node are contiguous, 80 B each), which otherwise miss one after the other as the
					// ray comes back to them.
					if (tg.y) {
						const char* tp = (const char*)(bvh.tris + 3 * (size_t)tg.x);
						asm volatile("prefetch.global.L2 [%0];" ::"l"(tp));
						asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + 128));
					}
#if LMB_PREFETCH > 1
					uint32_t pm = hits8 & imask;
					while (pm) {
						const int slot = __ffs((int)pm) - 1;
						pm &= pm - 1u;
						const char* cp = (const char*)(bvh.nodes + 5 * (size_t)(ng.x + __popc(imask & ((1u << slot) - 1u))));
						asm volatile("prefetch.global.L2 [%0];" ::"l"(cp));
						asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + 79));
					}
#endif
				}
#endif
			}
			// ---- triangle phase, warp-cooperative: the (owner lane, triangle) pairs of all lanes are spread over the 32 lanes,
			// tested once each (any lane tests for any owner: the owner's ray constants sit in shared memory), and every owner
			// folds the results of its own pairs into its hit. A node step leaves only a few lanes with 1..3 triangles each;
			// testing them lane-by-owner ran the ~150-instruction test at 3 of 32 lanes (profiles/r01b).
			// A round is worth its fixed cost only with enough pairs: lanes holding triangles sit out node steps until
			// LMB_TRI_ROUND_LANES lanes hold some, or nobody else can step.
			uint32_t tri_lanes = __ballot_sync(0xFFFFFFFFu, tg.y != 0u);
			if ((uint32_t)__popc(tri_lanes) < bvh.tri_round_lanes && __ballot_sync(0xFFFFFFFFu, has && tg.y == 0u && ng.y > 0x00FFFFFFu) != 0u) {
#ifdef LMB_TRACE_PROFILE
				p_parked += __popc(tri_lanes);
#endif
				tri_lanes = 0u;
			}
			while (tri_lanes) {
#if LMB_TRACE_WCOUNT > 1
				w_rounds++;
#endif
				const uint32_t cnt = (uint32_t)__popc(tg.y);
				uint32_t incl = cnt;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
					if (lane >= o) incl += v;
				}
				const uint32_t excl = incl - cnt;
				const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
#ifdef LMB_TRACE_PROFILE
				p_rounds++, p_pairs += min(total, 32u);
#endif
				uint32_t wrote = 0;
				while (tg.y != 0u && excl + wrote < 32u) {
					const int bit = __ffs((int)tg.y) - 1;
					tg.y &= tg.y - 1u;
					sm.pair[wbase + excl + wrote] = ((uint32_t)lane << 27) | (tg.x + (uint32_t)__popc(tl & ((1u << bit) - 1u)));
					wrote++;
				}
				__syncwarp();
				if ((uint32_t)lane < min(total, 32u)) {
					const uint32_t pr = sm.pair[wbase + lane];
					const int owner = wbase + (int)(pr >> 27);
					const float4 ra = sm.ray_a[owner], rb = sm.ray_b[owner];
					const float4* tp = bvh.tris + 3 * (size_t)(pr & 0x07FFFFFFu);
					const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
					const float4 rc = sm.ray_c[owner];
					const TriRay tr{v3(ra.x, ra.y, ra.z), ra.w, rb.x, rb.y, rb.z, __float_as_uint(rb.w), v3(rc.x, rc.y, rc.z)};
#ifndef LMB_TRACE_NO_STATS
					n_tris++;
#endif
					float t, V = 0.0f, W = 0.0f, det = 1.0f;
					if (!(tri_test(tr, a, b, c, t, V, W, det) && t > tr.tmin)) t = -1.0f;
					sm.res[wbase + lane] = make_float4(t, V, W, det);
					sm.res_prim[wbase + lane] = __float_as_uint(a.w);
				}
				__syncwarp();
				for (uint32_t k = 0; k < wrote; k++) {
					const float4 rs = sm.res[wbase + excl + k];
					const uint32_t p = sm.res_prim[wbase + excl + k];
					if (rs.x >= 0.0f && (rs.x < h.t || (rs.x == h.t && p < h.prim && h.prim != 0xFFFFFFFFu))) {
						h.t = rs.x, h.prim = p, h.b1 = rs.y, h.b2 = rs.z, det = rs.w;  // b1 = V / det, b2 = W / det at the end
					}
				}
				if (any && h.prim != 0xFFFFFFFFu) tg.y = 0u, ng.y = 0u, sp = 0;  // first accepted hit ends a shadow ray
				__syncwarp();
				tri_lanes = __ballot_sync(0xFFFFFFFFu, tg.y != 0u);
			}
			// ---- next group, or done
			if (has && tg.y == 0u && ng.y <= 0x00FFFFFFu) {
				if (sp > 0) {
					sp--;
					if (sp >= WS) ng = l_stack[sp - WS];
					else if (PIN) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ng.x), "=r"(ng.y) : "r"(stack_base + (uint32_t)sp * (LMB_TRACE_THREADS * 8u)));
					else ng = sm.stack[sp][tid];
				} else {
#if LMB_DEFER_STORE
					has = false, done = true;  // barycentrics and the store wait for the refill, where several lanes finish together
#else
					if (!any && h.prim != 0xFFFFFFFFu) h.b1 = h.b1 / det, h.b2 = h.b2 / det;
					src.store(item, h);
					has = false;
#endif
				}
			}
			const int busy = __popc(__ballot_sync(0xFFFFFFFFu, has));
			if (busy == 0 || ((!exhausted || rq_head != rq_count) && (uint32_t)busy < bvh.refill_lanes)) break;
		}
	}
	// ---- statistics (one atomic per warp and counter)
	for (int o = 16; o > 0; o >>= 1) {
		n_nodes += __shfl_xor_sync(0xFFFFFFFFu, n_nodes, o);
		n_tris += __shfl_xor_sync(0xFFFFFFFFu, n_tris, o);
		n_closest += __shfl_xor_sync(0xFFFFFFFFu, n_closest, o);
		n_any += __shfl_xor_sync(0xFFFFFFFFu, n_any, o);
	}
#ifdef LMB_TRACE_PROFILE
	if (lane == 0 && stats) {
		atomicAdd(&stats[ST_P_ITERS], (unsigned long long)p_iters), atomicAdd(&stats[ST_P_NODE_TRIPS], (unsigned long long)p_node_trips);
		atomicAdd(&stats[ST_P_NODE_LANES], (unsigned long long)p_node_lanes), atomicAdd(&stats[ST_P_HAS_LANES], (unsigned long long)p_has);
		atomicAdd(&stats[ST_P_PARKED_LANES], (unsigned long long)p_parked), atomicAdd(&stats[ST_P_ROUNDS], (unsigned long long)p_rounds);
		atomicAdd(&stats[ST_P_PAIRS], (unsigned long long)p_pairs), atomicAdd(&stats[ST_P_REFILLS], (unsigned long long)p_refills);
	}
#endif
	if (lane == 0 && stats) {
		if (w_iters) atomicAdd(&stats[ST_W_ITERS], (unsigned long long)w_iters);
		if (w_rounds) atomicAdd(&stats[ST_W_ROUNDS], (unsigned long long)w_rounds);
		if (w_refills) atomicAdd(&stats[ST_W_REFILLS], (unsigned long long)w_refills);
		if (n_nodes) atomicAdd(&stats[ST_NODES], (unsigned long long)n_nodes);
		if (n_tris) atomicAdd(&stats[ST_TRIS], (unsigned long long)n_tris);
		if (n_closest && stat_closest >= 0) atomicAdd(&stats[stat_closest], (unsigned long long)n_closest);
		if (n_any && stat_any >= 0) atomicAdd(&stats[stat_any], (unsigned long long)n_any);
	}
}

}  // namespace lmb
