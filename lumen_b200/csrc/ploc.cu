// ploc.cu -- quality binary tree for TRAVERSAL, built from the canonical LBVH's Morton-sorted leaf order by parallel
// locally-ordered clustering (Meister & Bittner, "Parallel Locally-Ordered Clustering for Bounding Volume Hierarchy
// Construction", TVCG 2018): every cluster looks LMB_PLOC_RADIUS places left and right for the neighbour whose union box has
// the smallest surface area; mutual nearest neighbours merge; the cluster array is compacted in order; repeat until one
// cluster is left. The Karras tree splits on Morton bits and ignores surface area -- on the classroom stand-in a ray visits
// 10.9 eight-wide nodes; agglomerative clustering over the same leaf order gives flatter, tighter subtrees.
// The canonical LBVH (lbvh.cu) stays the parity surface; this tree is a deterministic function of its
// sorted leaf order and leaf boxes: distances are symmetric floats, ties go to the lower index, compaction keeps order.
// Only node ids depend on atomic order. Numbering follows the canonical arrays so that wide_bvh.cu takes either tree:
// internal nodes 0 .. n-2 with the root at 0, leaf k (sorted position) at n-1+k; aabb[6 * id]; count[id] = triangles below.
// Per iteration and live cluster: 2 * 32 B box reads for the window (shared-memory tile), 48 B cluster record rd + wr.
#include <stdio.h>

#include "context.h"
#include "vec.cuh"

namespace lmb {

namespace {

#ifndef LMB_PLOC_RADIUS
#define LMB_PLOC_RADIUS 16
#endif
constexpr int NN_THREADS = 256;
constexpr int SCAN_THREADS = 1024;

struct Cluster {
	float lo[3];
	uint32_t node;
	float hi[3];
	uint32_t count;
};

__device__ __forceinline__ float union_half_area(const Cluster& a, const Cluster& b) {
	const float dx = fmaxf(a.hi[0], b.hi[0]) - fminf(a.lo[0], b.lo[0]);
	const float dy = fmaxf(a.hi[1], b.hi[1]) - fminf(a.lo[1], b.lo[1]);
	const float dz = fmaxf(a.hi[2], b.hi[2]) - fminf(a.lo[2], b.lo[2]);
	return dx * dy + dy * dz + dz * dx;
}

__global__ void __launch_bounds__(256) k_ploc_init(uint32_t n, const float* __restrict__ aabb, Cluster* __restrict__ cl, uint32_t* __restrict__ count) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const float* b = aabb + 6 * (size_t)(n - 1 + k);
	Cluster c;
	c.lo[0] = b[0], c.lo[1] = b[1], c.lo[2] = b[2], c.hi[0] = b[3], c.hi[1] = b[4], c.hi[2] = b[5];
	c.node = n - 1 + k, c.count = 1;
	cl[k] = c;
	count[n - 1 + k] = 1;
}

// nearest neighbour within the window, smallest union area, lowest index on ties
__global__ void __launch_bounds__(NN_THREADS) k_ploc_nn(uint32_t n_cur, const Cluster* __restrict__ cl, uint32_t* __restrict__ nn) {
	__shared__ Cluster tile[NN_THREADS + 2 * LMB_PLOC_RADIUS];
	const int first = (int)(blockIdx.x * NN_THREADS) - LMB_PLOC_RADIUS;
	for (int t = threadIdx.x; t < NN_THREADS + 2 * LMB_PLOC_RADIUS; t += NN_THREADS) {
		const int g = first + t;
		if (g >= 0 && g < (int)n_cur) tile[t] = cl[g];
	}
	__syncthreads();
	const int i = (int)(blockIdx.x * NN_THREADS + threadIdx.x);
	if (i >= (int)n_cur) return;
	const Cluster me = tile[threadIdx.x + LMB_PLOC_RADIUS];
	float best = 3.402823466e+38f;
	int bj = -1;
	for (int d = -LMB_PLOC_RADIUS; d <= LMB_PLOC_RADIUS; d++) {
		const int j = i + d;
		if (d == 0 || j < 0 || j >= (int)n_cur) continue;
		const float a = union_half_area(me, tile[threadIdx.x + LMB_PLOC_RADIUS + d]);
		if (a < best || bj < 0) best = a, bj = j;
	}
	nn[i] = (uint32_t)bj;
}

// mutual nearest neighbours merge into a new internal node at the lower index; the higher index is dropped
__global__ void __launch_bounds__(256) k_ploc_merge(uint32_t n, uint32_t n_cur, Cluster* __restrict__ cl, const uint32_t* __restrict__ nn,
													 uint32_t* __restrict__ valid, uint32_t* __restrict__ merge_counter, uint32_t* __restrict__ left,
													 uint32_t* __restrict__ right, float* __restrict__ aabb, uint32_t* __restrict__ count) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_cur) return;
	const uint32_t j = nn[i];
	uint32_t keep = 1;
	if (nn[j] == i) {
		if (i < j) {
			const Cluster a = cl[i], b = cl[j];
			const uint32_t id = n - 2u - atomicAdd(merge_counter, 1u);  // the last merge (the root) gets id 0
			Cluster m;
			for (int k = 0; k < 3; k++) m.lo[k] = fminf(a.lo[k], b.lo[k]), m.hi[k] = fmaxf(a.hi[k], b.hi[k]);
			m.node = id, m.count = a.count + b.count;
			left[id] = a.node, right[id] = b.node, count[id] = m.count;
			float* o = aabb + 6 * (size_t)id;
			o[0] = m.lo[0], o[1] = m.lo[1], o[2] = m.lo[2], o[3] = m.hi[0], o[4] = m.hi[1], o[5] = m.hi[2];
			cl[i] = m;  // (j only reads cl[i] through nn, never the record: no race)
		} else {
			keep = 0;
		}
	}
	valid[i] = keep;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_block_sums(const uint32_t* __restrict__ valid, uint32_t n_cur, uint32_t* __restrict__ sums) {
	__shared__ uint32_t warp_sums[32];
	const uint32_t i = blockIdx.x * SCAN_THREADS + threadIdx.x;
	uint32_t v = i < n_cur ? valid[i] : 0u;
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
	if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x < 32) {
		uint32_t w = warp_sums[threadIdx.x];
		for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xFFFFFFFFu, w, o);
		if (threadIdx.x == 0) sums[blockIdx.x] = w;
	}
}

// exclusive scan of the block sums in place (single block), total -> *total_out
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(uint32_t* sums, uint32_t count, uint32_t* total_out) {
	__shared__ uint32_t warp_sums[32];
	__shared__ uint32_t carry_s;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (uint32_t base = 0; base < count; base += SCAN_THREADS) {
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = i < count ? sums[i] : 0u;
		uint32_t s = v;
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
			if (lane >= o) s += t;
		}
		if (lane == 31) warp_sums[warp] = s;
		__syncthreads();
		if (warp == 0) {
			uint32_t w = warp_sums[lane];
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o);
				if (lane >= o) w += t;
			}
			warp_sums[lane] = w;
		}
		__syncthreads();
		const uint32_t carry = carry_s;
		if (i < count) sums[i] = carry + (warp > 0 ? warp_sums[warp - 1] : 0u) + (s - v);
		__syncthreads();
		if (threadIdx.x == SCAN_THREADS - 1) carry_s = carry + warp_sums[31];
		__syncthreads();
	}
	if (threadIdx.x == 0) *total_out = carry_s;
}

// order-preserving compaction of the surviving clusters
__global__ void __launch_bounds__(SCAN_THREADS) k_compact(const uint32_t* __restrict__ valid, uint32_t n_cur, const uint32_t* __restrict__ sums,
														 const Cluster* __restrict__ in, Cluster* __restrict__ out) {
	__shared__ uint32_t warp_sums[32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t i = blockIdx.x * SCAN_THREADS + threadIdx.x;
	const uint32_t v = i < n_cur ? valid[i] : 0u;
	uint32_t s = v;
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
		if (lane >= o) s += t;
	}
	if (lane == 31) warp_sums[warp] = s;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = warp_sums[lane];
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o);
			if (lane >= o) w += t;
		}
		warp_sums[lane] = w;
	}
	__syncthreads();
	if (v) out[sums[blockIdx.x] + (warp > 0 ? warp_sums[warp - 1] : 0u) + (s - v)] = in[i];
}

}  // namespace

void free_ploc(lmb_ctx* ctx) {
	DeviceBvh& b = ctx->bvh;
	cudaFree(b.q_left), cudaFree(b.q_right), cudaFree(b.q_aabb), cudaFree(b.q_count);
	b.q_left = b.q_right = b.q_count = nullptr;
	b.q_aabb = nullptr;
}

// Builds b.q_left / q_right / q_aabb / q_count. Needs the canonical leaf boxes (b.aabb) of build_lbvh.
int build_ploc(lmb_ctx* ctx) {
	free_ploc(ctx);
	DeviceBvh& b = ctx->bvh;
	const uint32_t n = b.n;
	if (n == 0) return 0;
	cudaStream_t st = ctx->stream;
	Cluster* cl[2] = {nullptr, nullptr};
	uint32_t *nn = nullptr, *valid = nullptr, *sums = nullptr, *ctr = nullptr;
	const uint32_t max_blocks = (n + SCAN_THREADS - 1) / SCAN_THREADS;
	LMB_CUDA(ctx, cudaMalloc((void**)&b.q_left, std::max<size_t>(n, 1) * 4));
	LMB_CUDA(ctx, cudaMalloc((void**)&b.q_right, std::max<size_t>(n, 1) * 4));
	LMB_CUDA(ctx, cudaMalloc((void**)&b.q_count, 2 * (size_t)n * 4));
	LMB_CUDA(ctx, cudaMalloc((void**)&b.q_aabb, 12 * (size_t)n * 4));
	LMB_CUDA(ctx, cudaMalloc((void**)&cl[0], (size_t)n * sizeof(Cluster)));
	LMB_CUDA(ctx, cudaMalloc((void**)&cl[1], (size_t)n * sizeof(Cluster)));
	LMB_CUDA(ctx, cudaMalloc((void**)&nn, (size_t)n * 4));
	LMB_CUDA(ctx, cudaMalloc((void**)&valid, (size_t)n * 4));
	LMB_CUDA(ctx, cudaMalloc((void**)&sums, (size_t)max_blocks * 4));
	LMB_CUDA(ctx, cudaMalloc((void**)&ctr, 8));
	cudaEventRecord(ctx->ev[0], st);
	LMB_CUDA(ctx, cudaMemsetAsync(ctr, 0, 8, st));
	LMB_CUDA(ctx, cudaMemcpyAsync(b.q_aabb + 6 * (size_t)(n - 1), b.aabb + 6 * (size_t)(n - 1), 6 * (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
	k_ploc_init<<<(n + 255) / 256, 256, 0, st>>>(n, b.aabb, cl[0], b.q_count);
	uint32_t n_cur = n, iters = 0;
	int cur = 0;
	int rc = 0;
	while (n_cur > 1) {
		const uint32_t sblocks = (n_cur + SCAN_THREADS - 1) / SCAN_THREADS;
		k_ploc_nn<<<(n_cur + NN_THREADS - 1) / NN_THREADS, NN_THREADS, 0, st>>>(n_cur, cl[cur], nn);
		k_ploc_merge<<<(n_cur + 255) / 256, 256, 0, st>>>(n, n_cur, cl[cur], nn, valid, ctr, b.q_left, b.q_right, b.q_aabb, b.q_count);
		k_block_sums<<<sblocks, SCAN_THREADS, 0, st>>>(valid, n_cur, sums);
		k_scan_sums<<<1, SCAN_THREADS, 0, st>>>(sums, sblocks, ctr + 1);
		k_compact<<<sblocks, SCAN_THREADS, 0, st>>>(valid, n_cur, sums, cl[cur], cl[cur ^ 1]);
		uint32_t n_next = 0;
		if ((rc = check_cuda(ctx, cudaMemcpyAsync(&n_next, ctr + 1, 4, cudaMemcpyDeviceToHost, st), "ploc readback"))) break;
		if ((rc = check_cuda(ctx, cudaStreamSynchronize(st), "ploc iteration"))) break;
		if (n_next >= n_cur || n_next == 0) {
			rc = set_error(ctx, LMB_ERR_INVALID, "build_ploc: clustering made no progress");
			break;
		}
		n_cur = n_next;
		cur ^= 1;
		iters++;
	}
	cudaEventRecord(ctx->ev[1], st);
	cudaStreamSynchronize(st);
	if (!rc) rc = check_cuda(ctx, cudaGetLastError(), "build_ploc");
	cudaEventElapsedTime(&ctx->stats.ms_build_ploc, ctx->ev[0], ctx->ev[1]);
	ctx->stats.ploc_iterations = iters;
	cudaFree(cl[0]), cudaFree(cl[1]), cudaFree(nn), cudaFree(valid), cudaFree(sums), cudaFree(ctr);
	return rc;
}

}  // namespace lmb
