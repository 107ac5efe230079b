// wide_bvh.cu -- collapses the canonical LBVH (lbvh.cu) into the compressed 8-wide BVH the traversal kernel walks.
//
// Why: the binary LBVH costs ~33 dependent node fetches per ray on the classroom stand-in and the warp waits for its
// slowest lane at every one of them (profiles/r01a_bvh2_baseline.md: 10.2 of 32 lanes active, long-scoreboard bound). An
// 8-wide node with quantised child boxes (layout after Ylitie, Karras, Laine, "Efficient Incoherent Ray Traversal on GPUs
// Through Compressed Wide BVHs", HPG 2017) cuts the dependent chain to ~1/3 and makes every step the same 8-box test.
// The wide tree is a deterministic function of the canonical LBVH (last sentence): which BVH2
// nodes become wide nodes, their children and slots depend only on left/right/aabb; only the ARRAY POSITIONS of nodes and
// triangles depend on atomic allocation order. Hits do not depend on either (trace.cuh: conservative boxes, closest =
// min t then lowest triangle id), so the bit-exact LBVH stays the parity surface and this is a pure traversal layout.
//
// Collapse (one thread per wide node, one launch per tree level):
//   children = {left, right} of the BVH2 node; repeatedly open the child with the largest surface area that still holds
//   more than LMB_WIDE_LEAF_TRIS triangles until 8 children; spare slots then open the remaining multi-triangle leaves.
//   Children are placed in slots by a greedy assignment on  dot(child centre - node centre, octant direction of slot),
//   so that visiting slots in order of (slot XOR ray octant) is front to back.
// Node, 80 bytes = 5 x 128-bit loads:
//   n0 = (p.x, p.y, p.z, ex | ey << 8 | ez << 16)                    p = node box min - step/16, e* = biased exponents of the grid step + 15
//   n1 = (child_base, tri_base, imask << 24 | leaf24, 0)              imask bit s: slot s is an internal child; leaf24 bits 3s..3s+2 =
//                                                                     1/3/7 for a leaf child with 1/2/3 triangles (0: internal or empty).
//                                                                     Leaf triangles are stored compactly in slot order, so triangle
//                                                                     (s, k) sits at tri_base + popc(leaf24 & ((1 << (3s + k)) - 1))
//   n2 = (qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7])        child box = p + q * 2^(e-127), q in 0..255,
//   n3 = (qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7])        rounded outwards by at least 1/32 grid step on every side
//   n4 = (qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7])
// Internal children of a node are contiguous from child_base in slot order; leaf triangles are copied next to each other
// from tri_base (3 x float4 world-space vertices, v0.w = global triangle id), at most 24 per node.
// HBM traffic (algorithmic): per wide node 8 x (24 B box + 8 B links) rd + 80 B wr; per triangle 48 B rd + 48 B wr.
#include <stdio.h>

#include "context.h"
#include "vec.cuh"

namespace lmb {

namespace {

constexpr uint32_t LMB_WIDE_LEAF_TRIS = 3;
enum { WC_ITEMS_OUT = 0, WC_NODES = 1, WC_TRIS = 2, WC_COUNT = 4 };

__device__ __forceinline__ float half_area(const float* a) {
	const float dx = a[3] - a[0], dy = a[4] - a[1], dz = a[5] - a[2];
	return dx * dy + dy * dz + dz * dx;
}

// smallest biased exponent E (1..254) with  ext <= 254 * 2^(E-127). Flat or very thin boxes get a grid step of at least
// ~8 ulp of their coordinates so that the grid origin below the box and the plane offsets stay representable.
__device__ __forceinline__ uint32_t grid_exponent(float lo, float hi) {
	const float ext = fmaxf(fmaxf(hi - lo, fmaxf(fabsf(lo), fabsf(hi)) * 0.000244140625f), 1e-30f);
	int k;
	frexpf(ext, &k);        // ext = m * 2^k, m in [0.5, 1)  =>  ext < 2^k
	int E = k - 8 + 127;    // step 2^(k-8): ext / step < 256
	E = max(E, 1);
	while (E < 254 && ext > 254.0f * __uint_as_float((uint32_t)E << 23)) E++;
	return (uint32_t)E;
}

__global__ void __launch_bounds__(128) k_collapse(uint32_t n, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
													  const uint32_t* __restrict__ count, const float* __restrict__ aabb, const float4* __restrict__ leaf_tris, const uint2* __restrict__ items_in,
													  uint32_t n_in, uint2* __restrict__ items_out, uint32_t* __restrict__ counters, float4* __restrict__ wnodes,
													  float4* __restrict__ wtris) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_in) return;
	const uint32_t w = items_in[i].x, b = items_in[i].y;
	const uint32_t first_leaf = n - 1;
	uint32_t ch[8];
	int cnt;
	const float* nb;  // node box
	if (n == 1) {
		ch[0] = 0, cnt = 1, nb = aabb;
	} else {
		ch[0] = left[b], ch[1] = right[b], cnt = 2, nb = aabb + 6 * (size_t)b;
	}
	auto ntris = [&](uint32_t c) { return c >= first_leaf ? 1u : count[c]; };
	for (int phase = 0; phase < 2 && n > 1; phase++) {
		while (cnt < 8) {
			int best = -1;
			float best_area = -1.0f;
			for (int k = 0; k < cnt; k++) {
				const uint32_t c = ch[k];
				if (c >= first_leaf) continue;
				if (phase == 0 && ntris(c) <= LMB_WIDE_LEAF_TRIS) continue;
				const float a = half_area(aabb + 6 * (size_t)c);
				if (a > best_area) best_area = a, best = k;
			}
			if (best < 0) break;
			const uint32_t c = ch[best];
			ch[best] = left[c];
			ch[cnt++] = right[c];
		}
	}
	// ---- slots: greedy assignment, largest  dot(centre offset, octant direction)  first
	const float ncx = (nb[0] + nb[3]) * 0.5f, ncy = (nb[1] + nb[4]) * 0.5f, ncz = (nb[2] + nb[5]) * 0.5f;
	float ox[8], oy[8], oz[8];
	for (int k = 0; k < cnt; k++) {
		const float* a = aabb + 6 * (size_t)ch[k];
		ox[k] = (a[0] + a[3]) * 0.5f - ncx, oy[k] = (a[1] + a[4]) * 0.5f - ncy, oz[k] = (a[2] + a[5]) * 0.5f - ncz;
	}
	int who[8];  // slot -> child index in ch[], -1 = empty
	for (int s = 0; s < 8; s++) who[s] = -1;
	uint32_t placed = 0;
	for (int it = 0; it < cnt; it++) {
		int bk = -1, bs = -1;
		float bc = 0.0f;
		for (int k = 0; k < cnt; k++) {
			if (placed & (1u << k)) continue;
			for (int s = 0; s < 8; s++) {
				if (who[s] >= 0) continue;
				const float c = ((s & 4) ? ox[k] : -ox[k]) + ((s & 2) ? oy[k] : -oy[k]) + ((s & 1) ? oz[k] : -oz[k]);
				if (bk < 0 || c > bc) bc = c, bk = k, bs = s;
			}
		}
		who[bs] = bk;
		placed |= 1u << bk;
	}
	// ---- allocation
	uint32_t imask = 0, n_internal = 0, n_leaf_tris = 0;
	for (int s = 0; s < 8; s++) {
		if (who[s] < 0) continue;
		const uint32_t c = ch[who[s]];
		const uint32_t nt = ntris(c);
		if (c < first_leaf && nt > LMB_WIDE_LEAF_TRIS)
			imask |= 1u << s, n_internal++;
		else
			n_leaf_tris += nt;
	}
	const uint32_t child_base = n_internal ? atomicAdd(&counters[WC_NODES], n_internal) : 0u;
	const uint32_t item_base = n_internal ? atomicAdd(&counters[WC_ITEMS_OUT], n_internal) : 0u;
	const uint32_t tri_base = n_leaf_tris ? atomicAdd(&counters[WC_TRIS], n_leaf_tris) : 0u;
	// ---- quantisation grid
	// The grid origin sits step/16 below the node box so that children touching the node's lower faces get the same outward
	// slack as everything else (the traversal's fused slab arithmetic is only good to ~1/512 step).
	const uint32_t E[3] = {grid_exponent(nb[0], nb[3]), grid_exponent(nb[1], nb[4]), grid_exponent(nb[2], nb[5])};
	float P[3];
	for (int a = 0; a < 3; a++) {
		const float step = __uint_as_float(E[a] << 23);
		P[a] = nb[a] - step * 0.0625f;
		if (!(P[a] < nb[a])) P[a] = nextafterf(nb[a], -3.402823466e+38f);  // step/16 below the ulp of the coordinate
	}
	uint32_t leaf24 = 0, qlo[3][8], qhi[3][8];
	uint32_t rank = 0, tri_off = 0;
	for (int s = 0; s < 8; s++) {
		if (who[s] < 0) {
			for (int a = 0; a < 3; a++) qlo[a][s] = 255u, qhi[a][s] = 0u;
			continue;
		}
		const uint32_t c = ch[who[s]];
		const float* cb = aabb + 6 * (size_t)c;
		for (int a = 0; a < 3; a++) {
			const float step = __uint_as_float(E[a] << 23);
			// outward rounding with 1/32 step of slack (absorbs the rounding of the traversal's fused slab arithmetic)
			const float lo = floorf((cb[a] - P[a]) / step - 0.03125f);
			const float hi = ceilf((cb[3 + a] - P[a]) / step + 0.03125f);
			qlo[a][s] = (uint32_t)fminf(fmaxf(lo, 0.0f), 255.0f);
			qhi[a][s] = (uint32_t)fminf(fmaxf(hi, 0.0f), 255.0f);
		}
		const uint32_t nt = ntris(c);
		if (imask & (1u << s)) {
			items_out[item_base + rank] = make_uint2(child_base + rank, c);
			rank++;
		} else {
			leaf24 |= ((1u << nt) - 1u) << (3 * s);
			// the <= LMB_WIDE_LEAF_TRIS leaves below c, left to right
			uint32_t todo[4] = {c, 0, 0, 0};
			int sp2 = 1;
			while (sp2 > 0) {
				const uint32_t x = todo[--sp2];
				if (x >= first_leaf) {
					const size_t d = 3 * (size_t)(tri_base + tri_off), q = 3 * (size_t)(x - first_leaf);
					wtris[d] = leaf_tris[q], wtris[d + 1] = leaf_tris[q + 1], wtris[d + 2] = leaf_tris[q + 2];
					tri_off++;
				} else {
					todo[sp2++] = right[x];
					todo[sp2++] = left[x];
				}
			}
		}
	}
	auto pack4 = [](const uint32_t* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); };
	auto f = [](uint32_t u) { return __uint_as_float(u); };
	float4* o = wnodes + 5 * (size_t)w;
	// stored with the + 15 the slab test needs (A = 2^(e + 15) / d, trace_wide.cuh); E <= 239 for any scene below 1e36 in size
	o[0] = make_float4(P[0], P[1], P[2], f(min(E[0] + 15u, 255u) | (min(E[1] + 15u, 255u) << 8) | (min(E[2] + 15u, 255u) << 16)));
	o[1] = make_float4(f(child_base), f(tri_base), f((imask << 24) | leaf24), 0.0f);
	o[2] = make_float4(f(pack4(qlo[0])), f(pack4(qlo[0] + 4)), f(pack4(qlo[1])), f(pack4(qlo[1] + 4)));
	o[3] = make_float4(f(pack4(qlo[2])), f(pack4(qlo[2] + 4)), f(pack4(qhi[0])), f(pack4(qhi[0] + 4)));
	o[4] = make_float4(f(pack4(qhi[1])), f(pack4(qhi[1] + 4)), f(pack4(qhi[2])), f(pack4(qhi[2] + 4)));
}

}  // namespace

void free_wide_bvh(lmb_ctx* ctx) {
	DeviceWideBvh& wb = ctx->wide;
	cudaFree(wb.nodes), cudaFree(wb.tris), cudaFree(wb.items[0]), cudaFree(wb.items[1]), cudaFree(wb.counters);
	wb = DeviceWideBvh{};
}

int build_wide_bvh(lmb_ctx* ctx) {
	free_wide_bvh(ctx);
	const DeviceBvh& b = ctx->bvh;
	DeviceWideBvh& wb = ctx->wide;
	const uint32_t n = b.n;
	if (n == 0) return 0;
	cudaStream_t st = ctx->stream;
	const size_t max_nodes = std::max<size_t>(n - 1, 1);  // every wide node is a distinct BVH2 internal node
	LMB_CUDA(ctx, cudaMalloc((void**)&wb.nodes, max_nodes * 80));
	LMB_CUDA(ctx, cudaMalloc((void**)&wb.tris, (size_t)n * 48));
	LMB_CUDA(ctx, cudaMalloc((void**)&wb.items[0], max_nodes * sizeof(uint2)));
	LMB_CUDA(ctx, cudaMalloc((void**)&wb.items[1], max_nodes * sizeof(uint2)));
	LMB_CUDA(ctx, cudaMalloc((void**)&wb.counters, WC_COUNT * sizeof(uint32_t)));
	cudaEventRecord(ctx->ev[0], st);
	const uint2 root = make_uint2(0u, 0u);
	const uint32_t init[WC_COUNT] = {0u, 1u, 0u, 0u};
	LMB_CUDA(ctx, cudaMemcpyAsync(wb.items[0], &root, sizeof(root), cudaMemcpyHostToDevice, st));
	LMB_CUDA(ctx, cudaMemcpyAsync(wb.counters, init, sizeof(init), cudaMemcpyHostToDevice, st));
	uint32_t n_in = 1, levels = 0;
	int cur = 0;
	while (n_in > 0) {
		if (b.q_left)  // quality tree (ploc.cu)
			k_collapse<<<(n_in + 127) / 128, 128, 0, st>>>(n, b.q_left, b.q_right, b.q_count, b.q_aabb, b.tris, wb.items[cur], n_in, wb.items[cur ^ 1],
														  wb.counters, wb.nodes, wb.tris);
		else  // canonical Karras tree
			k_collapse<<<(n_in + 127) / 128, 128, 0, st>>>(n, b.left, b.right, b.span_count, b.aabb, b.tris, wb.items[cur], n_in, wb.items[cur ^ 1],
														  wb.counters, wb.nodes, wb.tris);
		LMB_CUDA(ctx, cudaMemcpyAsync(&n_in, wb.counters + WC_ITEMS_OUT, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
		LMB_CUDA(ctx, cudaStreamSynchronize(st));
		LMB_CUDA(ctx, cudaMemsetAsync(wb.counters + WC_ITEMS_OUT, 0, sizeof(uint32_t), st));
		cur ^= 1;
		if (++levels > 4096) return set_error(ctx, LMB_ERR_INVALID, "build_wide_bvh: runaway collapse");
	}
	uint32_t fin[WC_COUNT];
	LMB_CUDA(ctx, cudaMemcpyAsync(fin, wb.counters, sizeof(fin), cudaMemcpyDeviceToHost, st));
	cudaEventRecord(ctx->ev[1], st);
	LMB_CUDA(ctx, cudaStreamSynchronize(st));
	LMB_CUDA(ctx, cudaGetLastError());
	wb.n_nodes = fin[WC_NODES], wb.n_tris = fin[WC_TRIS], wb.levels = levels;
	if (wb.n_tris != n) return set_error(ctx, LMB_ERR_INVALID, "build_wide_bvh: triangle count mismatch after collapse");
	cudaEventElapsedTime(&ctx->stats.ms_build_wide, ctx->ev[0], ctx->ev[1]);
	cudaFree(wb.items[0]), cudaFree(wb.items[1]);
	wb.items[0] = wb.items[1] = nullptr;
	return 0;
}

}  // namespace lmb
