// lbvh.cu -- GPU LBVH builder for sm_100a.
//
// Replaces the driver-built BLAS/TLAS of Integrator::create_accel (src/RayTracer/Integrator.cpp:137-160,
// src/Framework/AccelerationStructure.cpp:171-315): one world-space tree over every triangle of every prim mesh
// (instance transform = LumenPrimMesh::world_matrix, applied once here), keeping (prim mesh index, mesh-local triangle)
// per triangle because the integrator compares those ids (pt_commons.glsl:33, ray.rchit:79-80).
//
// Canonical construction (bit-exact against oracle/lbvh_cpu.h, checked by tests/test_lbvh.py):
//   1. k_flatten     world-space vertices (glm mat4*vec4 operation order), scene bounds by ordered-uint atomics
//   2. k_morton      30-bit Morton code of the triangle-AABB centre, key = morton << 32 | global id
//   3. radix sort    stable LSD over the 30 Morton bits (4 x 8-bit passes): histogram / scan / ranked scatter
//   4. k_karras      Karras 2012 radix tree, delta = clz64(key_i ^ key_j)
//   5. k_refit       bottom-up min/max refit with per-node arrival counters
//   6. k_pack        64-byte traversal nodes (both child boxes + refs) and leaf-ordered triangles
// All steps are integer or min/max arithmetic -> independent of thread scheduling.
//
// HBM traffic per triangle (algorithmic): flatten 108 B rd + 48 B wr, morton 48 rd + 12 wr, sort 4 passes x (8 rd hist +
// 8 rd + 8 wr scatter), karras ~8 x 8 rd + 12 wr, refit 48 rd + 24 wr (+24 rd/24 wr per internal node), pack 48+48 rd,
// 64 + 48 wr  => ~0.7 KB per triangle.
#include <stdio.h>

#include "context.h"
#include "vec.cuh"

namespace lmb {

namespace {

constexpr int RADIX_THREADS = 256;
constexpr int RADIX_ITEMS = 16;
constexpr int RADIX_TILE = RADIX_THREADS * RADIX_ITEMS;

__device__ __forceinline__ uint32_t enc_float(float f) {
	const uint32_t b = __float_as_uint(f);
	return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_float(uint32_t e) { return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e); }

__global__ void k_init_bounds(uint32_t* bounds_enc) {
	if (threadIdx.x < 3) bounds_enc[threadIdx.x] = 0xFFFFFFFFu;  // min
	else if (threadIdx.x < 6) bounds_enc[threadIdx.x] = 0u;      // max
}

__global__ void __launch_bounds__(256) k_flatten(DeviceScene sc, float4* __restrict__ tri_world, uint32_t* __restrict__ bounds_enc) {
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	float lo[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
	float hi[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
	if (g < sc.n_tris) {
		const uint32_t mesh = sc.tri_mesh[g], local = sc.tri_local[g];
		const lmb_prim_mesh_info pi = sc.prim_infos[mesh];
		const M4 M = load_m4(sc.world_matrices + 16 * mesh);
#pragma unroll
		for (int k = 0; k < 3; k++) {
			const uint32_t vi = sc.indices[pi.index_offset + 3 * local + k] + pi.vertex_offset;
			const V4 w = mul(M, v4(v3(sc.vertices[vi].pos), 1.0f));
			tri_world[3 * (size_t)g + k] = make_float4(w.x, w.y, w.z, 0.0f);
			lo[0] = fminf(lo[0], w.x), lo[1] = fminf(lo[1], w.y), lo[2] = fminf(lo[2], w.z);
			hi[0] = fmaxf(hi[0], w.x), hi[1] = fmaxf(hi[1], w.y), hi[2] = fmaxf(hi[2], w.z);
		}
	}
#pragma unroll
	for (int k = 0; k < 3; k++) {
		for (int o = 16; o > 0; o >>= 1) {
			lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
			hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
		}
	}
	if ((threadIdx.x & 31) == 0) {
#pragma unroll
		for (int k = 0; k < 3; k++) {
			atomicMin(&bounds_enc[k], enc_float(lo[k]));
			atomicMax(&bounds_enc[3 + k], enc_float(hi[k]));
		}
	}
}

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
__device__ __forceinline__ uint32_t quantize10(float c, float lo, float hi) {
	const float ext = hi - lo;
	if (!(ext > 0.0f)) return 0u;
	float q = (c - lo) / ext * 1024.0f;
	q = gmin(gmax(q, 0.0f), 1023.0f);
	return (uint32_t)q;
}

__global__ void __launch_bounds__(256) k_morton(uint32_t n, const float4* __restrict__ tri_world, const uint32_t* __restrict__ bounds_enc,
												 uint32_t* __restrict__ morton, uint64_t* __restrict__ keys) {
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n) return;
	const float4 a = tri_world[3 * (size_t)g], b = tri_world[3 * (size_t)g + 1], c = tri_world[3 * (size_t)g + 2];
	const V3 tmin = v3(gmin(gmin(a.x, b.x), c.x), gmin(gmin(a.y, b.y), c.y), gmin(gmin(a.z, b.z), c.z));
	const V3 tmax = v3(gmax(gmax(a.x, b.x), c.x), gmax(gmax(a.y, b.y), c.y), gmax(gmax(a.z, b.z), c.z));
	const V3 cen = (tmin + tmax) * 0.5f;
	const V3 lo = v3(dec_float(bounds_enc[0]), dec_float(bounds_enc[1]), dec_float(bounds_enc[2]));
	const V3 hi = v3(dec_float(bounds_enc[3]), dec_float(bounds_enc[4]), dec_float(bounds_enc[5]));
	const uint32_t x = quantize10(cen.x, lo.x, hi.x), y = quantize10(cen.y, lo.y, hi.y), z = quantize10(cen.z, lo.z, hi.z);
	const uint32_t m = (expand_bits10(x) << 2) | (expand_bits10(y) << 1) | expand_bits10(z);
	morton[g] = m;
	keys[g] = ((uint64_t)m << 32) | g;
}

// ---- radix sort: one 8-bit digit per pass ---------------------------------------------------------------------
__global__ void __launch_bounds__(RADIX_THREADS) k_radix_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ hist,
															   uint32_t n_blocks) {
	__shared__ uint32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t base = blockIdx.x * RADIX_TILE;
#pragma unroll 4
	for (int r = 0; r < RADIX_ITEMS; r++) {
		const uint32_t i = base + r * RADIX_THREADS + threadIdx.x;
		if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
	}
	__syncthreads();
	hist[threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `count` uint32 in place, single block of 1024 threads
__global__ void __launch_bounds__(1024) k_scan_exclusive(uint32_t* data, uint32_t count) {
	__shared__ uint32_t warp_sums[32];
	__shared__ uint32_t carry_s;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (uint32_t base = 0; base < count; base += 1024) {
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = (i < count) ? data[i] : 0u;
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
		const uint32_t prefix = carry + (warp > 0 ? warp_sums[warp - 1] : 0u) + (s - v);
		if (i < count) data[i] = prefix;
		__syncthreads();
		if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
		__syncthreads();
	}
}

__global__ void __launch_bounds__(RADIX_THREADS) k_radix_scatter(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint32_t n, int shift,
																  const uint32_t* __restrict__ hist, uint32_t n_blocks) {
	__shared__ uint32_t wc[RADIX_THREADS / 32][256];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t run = hist[threadIdx.x * n_blocks + blockIdx.x];  // thread d owns digit d: next free slot for this block
	const uint32_t base = blockIdx.x * RADIX_TILE;
	for (int r = 0; r < RADIX_ITEMS; r++) {
		const uint32_t i = base + r * RADIX_THREADS + threadIdx.x;
		if (base + r * RADIX_THREADS >= n) break;
#pragma unroll
		for (int w = 0; w < RADIX_THREADS / 32; w++) wc[w][threadIdx.x] = 0;
		__syncthreads();
		const bool active = i < n;
		const uint64_t key = active ? in[i] : 0ull;
		const uint32_t digit = active ? ((uint32_t)(key >> shift) & 255u) : 0xFFFFFFFFu;
		const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
		const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
		if (active && rank == 0) wc[warp][digit] = __popc(peers);
		__syncthreads();
		{
			uint32_t acc = run;
#pragma unroll
			for (int w = 0; w < RADIX_THREADS / 32; w++) {
				const uint32_t c = wc[w][threadIdx.x];
				wc[w][threadIdx.x] = acc;
				acc += c;
			}
			run = acc;
		}
		__syncthreads();
		if (active) out[wc[warp][digit] + rank] = key;
		__syncthreads();
	}
}

// ---- ray ordering for array queries (lmb_trace_closest_device_ex, sort_rays): key = 12-bit Morton code of the origin's cell in the
// scene box (4 bits per axis, outside origins clamp) followed by the direction octant (3 bits). A counting sort over the 2^15 bins
// (histogram by global atomics, one scan, scatter by atomics on the bin cursors: three light kernels, 0.4 ms for 2^24 rays; an LSD
// radix sort of wider keys cost 3 ms and ate the gain; 2^18 bins with a finer origin grid or extra direction bits trace no faster and
// pay 0.2 ms more for the scan). Rays that start in the same region and head the same way become neighbours in the persistent
// walker's 32-ray fetches: their node and triangle fetches hit the same L1 / L2 lines. The order inside a bin is whatever the
// atomics give; hits land at the rays' own indices and do not depend on it.
constexpr uint32_t RAY_BINS = 1u << 15;
__global__ void __launch_bounds__(256) k_ray_keys(const float4* __restrict__ rays, uint32_t n, const uint32_t* __restrict__ bounds_enc, uint32_t* __restrict__ keys,
												   uint32_t* __restrict__ hist) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4 o = rays[2 * (size_t)i], d = rays[2 * (size_t)i + 1];
	const V3 lo = v3(dec_float(bounds_enc[0]), dec_float(bounds_enc[1]), dec_float(bounds_enc[2]));
	const V3 hi = v3(dec_float(bounds_enc[3]), dec_float(bounds_enc[4]), dec_float(bounds_enc[5]));
	auto cell = [](float v, float a, float b) {
		const float t = (v - a) / fmaxf(b - a, 1e-30f) * 16.0f;
		return (uint32_t)fminf(fmaxf(t, 0.0f), 15.0f);  // NaN -> 0
	};
	const uint32_t om = (expand_bits10(cell(o.x, lo.x, hi.x)) << 2) | (expand_bits10(cell(o.y, lo.y, hi.y)) << 1) | expand_bits10(cell(o.z, lo.z, hi.z));
	const uint32_t oct = (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u);
	const uint32_t key = ((om & 0xFFFu) << 3) | oct;
	keys[i] = key;
	atomicAdd(&hist[key], 1u);
}
__global__ void __launch_bounds__(256) k_ray_scatter(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	order[atomicAdd(&cursor[keys[i]], 1u)] = i;
}

// ---- Karras 2012 ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, uint64_t ki, int j) {
	if (j < 0 || j >= n) return -1;
	return __clzll((long long)(ki ^ keys[j]));
}

__global__ void __launch_bounds__(256) k_karras(const uint64_t* __restrict__ keys, int n, uint32_t* __restrict__ left, uint32_t* __restrict__ right,
												 uint32_t* __restrict__ parent, uint32_t* __restrict__ leaf_prim, uint32_t* __restrict__ span_count) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) leaf_prim[i] = (uint32_t)(keys[i] & 0xFFFFFFFFull);
	if (i == 0) parent[0] = 0xFFFFFFFFu;
	if (i >= n - 1) return;
	const uint64_t ki = keys[i];
	const int d = (delta(keys, n, ki, i + 1) - delta(keys, n, ki, i - 1)) >= 0 ? 1 : -1;
	const int dmin = delta(keys, n, ki, i - d);
	int lmax = 2;
	while (delta(keys, n, ki, i + lmax * d) > dmin) lmax *= 2;
	int l = 0;
	for (int t = lmax / 2; t >= 1; t /= 2)
		if (delta(keys, n, ki, i + (l + t) * d) > dmin) l += t;
	const int j = i + l * d;
	const int dnode = delta(keys, n, ki, j);
	int s = 0, t = l;
	do {
		t = (t + 1) >> 1;
		if (delta(keys, n, ki, i + (s + t) * d) > dnode) s += t;
	} while (t > 1);
	const int gamma = i + s * d + min(d, 0);
	const uint32_t lc = (min(i, j) == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
	const uint32_t rc = (max(i, j) == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
	left[i] = lc;
	right[i] = rc;
	span_count[i] = (uint32_t)(max(i, j) - min(i, j) + 1);
	parent[lc] = (uint32_t)i;
	parent[rc] = (uint32_t)i;
}

// ---- refit ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_refit(uint32_t n, const float4* __restrict__ tri_world, const uint32_t* __restrict__ leaf_prim,
												const uint32_t* __restrict__ left, const uint32_t* __restrict__ right, const uint32_t* __restrict__ parent,
												float* aabb, uint32_t* arrive) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t p = leaf_prim[i];
	const float4 a = tri_world[3 * (size_t)p], b = tri_world[3 * (size_t)p + 1], c = tri_world[3 * (size_t)p + 2];
	float* o = aabb + 6 * (size_t)(n - 1 + i);
	o[0] = gmin(gmin(a.x, b.x), c.x), o[1] = gmin(gmin(a.y, b.y), c.y), o[2] = gmin(gmin(a.z, b.z), c.z);
	o[3] = gmax(gmax(a.x, b.x), c.x), o[4] = gmax(gmax(a.y, b.y), c.y), o[5] = gmax(gmax(a.z, b.z), c.z);
	if (n == 1) return;
	uint32_t node = parent[n - 1 + i];
	while (node != 0xFFFFFFFFu) {
		__threadfence();
		if (atomicAdd(&arrive[node], 1u) == 0u) return;  // first child to arrive: the sibling finishes the node
		__threadfence();
		const volatile float* l = aabb + 6 * (size_t)left[node];
		const volatile float* r = aabb + 6 * (size_t)right[node];
		volatile float* w = aabb + 6 * (size_t)node;
#pragma unroll
		for (int k = 0; k < 3; k++) {
			w[k] = gmin(l[k], r[k]);
			w[3 + k] = gmax(l[3 + k], r[3 + k]);
		}
		node = parent[node];
	}
}

__global__ void __launch_bounds__(256) k_pack(uint32_t n, const float4* __restrict__ tri_world, const uint32_t* __restrict__ leaf_prim,
											   const uint32_t* __restrict__ left, const uint32_t* __restrict__ right, const float* __restrict__ aabb,
											   float4* __restrict__ nodes, float4* __restrict__ tris) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		const uint32_t p = leaf_prim[i];
		float4 a = tri_world[3 * (size_t)p];
		a.w = __uint_as_float(p);
		tris[3 * (size_t)i] = a;
		tris[3 * (size_t)i + 1] = tri_world[3 * (size_t)p + 1];
		tris[3 * (size_t)i + 2] = tri_world[3 * (size_t)p + 2];
	}
	if (i + 1 < n) {
		const uint32_t l = left[i], r = right[i];
		const float* bl = aabb + 6 * (size_t)l;
		const float* br = aabb + 6 * (size_t)r;
		const int lref = (l >= n - 1) ? ~(int)(l - (n - 1)) : (int)l;
		const int rref = (r >= n - 1) ? ~(int)(r - (n - 1)) : (int)r;
		nodes[4 * (size_t)i + 0] = make_float4(bl[0], bl[1], bl[2], bl[3]);
		nodes[4 * (size_t)i + 1] = make_float4(bl[4], bl[5], br[0], br[1]);
		nodes[4 * (size_t)i + 2] = make_float4(br[2], br[3], br[4], br[5]);
		nodes[4 * (size_t)i + 3] = make_float4(__int_as_float(lref), __int_as_float(rref), 0.0f, 0.0f);
	}
}

template <typename T>
int dmalloc(lmb_ctx* ctx, T** p, size_t count) {
	return check_cuda(ctx, cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)), "cudaMalloc(bvh)");
}

}  // namespace

void free_bvh(lmb_ctx* ctx) {
	DeviceBvh& b = ctx->bvh;
	cudaFree(b.morton), cudaFree(b.keys), cudaFree(b.keys_tmp), cudaFree(b.leaf_prim), cudaFree(b.left), cudaFree(b.right);
	cudaFree(b.parent), cudaFree(b.aabb), cudaFree(b.arrive), cudaFree(b.bounds_enc), cudaFree(b.tri_world), cudaFree(b.nodes);
	cudaFree(b.tris), cudaFree(b.radix_hist), cudaFree(b.span_count);
	b = DeviceBvh{};
	free_wide_bvh(ctx);
	free_ploc(ctx);
}

// Order of `n` rays (device array of 2 float4 each) for the array walker: (*order_out)[i] = index of the i-th ray to trace. Scratch is
// kept in the context and grows on demand.
int sort_rays(lmb_ctx* ctx, const float4* d_rays, uint32_t n, const uint32_t** order_out) {
	if (!ctx->bvh.bounds_enc) return set_error(ctx, LMB_ERR_INVALID, "sort_rays: no accel built");
	cudaStream_t st = ctx->stream;
	if (ctx->ray_sort_cap < n) {
		cudaFree(ctx->ray_keys), cudaFree(ctx->ray_order), cudaFree(ctx->ray_hist);
		ctx->ray_keys = ctx->ray_order = ctx->ray_hist = nullptr, ctx->ray_sort_cap = 0;
		int rc;
		if ((rc = dmalloc(ctx, &ctx->ray_keys, n)) || (rc = dmalloc(ctx, &ctx->ray_order, n)) || (rc = dmalloc(ctx, &ctx->ray_hist, RAY_BINS))) return rc;
		ctx->ray_sort_cap = n;
	}
	cudaMemsetAsync(ctx->ray_hist, 0, RAY_BINS * sizeof(uint32_t), st);
	k_ray_keys<<<(n + 255) / 256, 256, 0, st>>>(d_rays, n, ctx->bvh.bounds_enc, ctx->ray_keys, ctx->ray_hist);
	k_scan_exclusive<<<1, 1024, 0, st>>>(ctx->ray_hist, RAY_BINS);
	k_ray_scatter<<<(n + 255) / 256, 256, 0, st>>>(ctx->ray_keys, n, ctx->ray_hist, ctx->ray_order);
	*order_out = ctx->ray_order;
	return check_cuda(ctx, cudaGetLastError(), "sort_rays");
}

int build_lbvh(lmb_ctx* ctx) {
	free_bvh(ctx);
	DeviceBvh& b = ctx->bvh;
	const uint32_t n = ctx->scene.n_tris;
	b.n = n;
	cudaStream_t st = ctx->stream;
	const uint32_t radix_blocks = (n + RADIX_TILE - 1) / RADIX_TILE;
	int rc;
	if ((rc = dmalloc(ctx, &b.morton, n))) return rc;
	if ((rc = dmalloc(ctx, &b.keys, n))) return rc;
	if ((rc = dmalloc(ctx, &b.keys_tmp, n))) return rc;
	if ((rc = dmalloc(ctx, &b.leaf_prim, n))) return rc;
	if ((rc = dmalloc(ctx, &b.left, n))) return rc;
	if ((rc = dmalloc(ctx, &b.right, n))) return rc;
	if ((rc = dmalloc(ctx, &b.parent, 2 * (size_t)n))) return rc;
	if ((rc = dmalloc(ctx, &b.aabb, 12 * (size_t)n))) return rc;
	if ((rc = dmalloc(ctx, &b.arrive, n))) return rc;
	if ((rc = dmalloc(ctx, &b.bounds_enc, 8))) return rc;
	if ((rc = dmalloc(ctx, &b.tri_world, 3 * (size_t)n))) return rc;
	if ((rc = dmalloc(ctx, &b.nodes, 4 * (size_t)n))) return rc;
	if ((rc = dmalloc(ctx, &b.tris, 3 * (size_t)n))) return rc;
	if ((rc = dmalloc(ctx, &b.radix_hist, 256 * (size_t)std::max(radix_blocks, 1u)))) return rc;
	if ((rc = dmalloc(ctx, &b.span_count, n))) return rc;
	if (n == 0) {
		b.built = true;
		return 0;
	}
	const uint32_t grid = (n + 255) / 256;
	cudaEventRecord(ctx->ev[0], st);
	k_init_bounds<<<1, 32, 0, st>>>(b.bounds_enc);
	k_flatten<<<grid, 256, 0, st>>>(ctx->scene, b.tri_world, b.bounds_enc);
	k_morton<<<grid, 256, 0, st>>>(n, b.tri_world, b.bounds_enc, b.morton, b.keys);
	cudaEventRecord(ctx->ev[1], st);
	uint64_t* src = b.keys;
	uint64_t* dst = b.keys_tmp;
	for (int pass = 0; pass < 4; pass++) {
		const int shift = 32 + 8 * pass;
		k_radix_hist<<<radix_blocks, RADIX_THREADS, 0, st>>>(src, n, shift, b.radix_hist, radix_blocks);
		k_scan_exclusive<<<1, 1024, 0, st>>>(b.radix_hist, 256 * radix_blocks);
		k_radix_scatter<<<radix_blocks, RADIX_THREADS, 0, st>>>(src, dst, n, shift, b.radix_hist, radix_blocks);
		std::swap(src, dst);
	}
	// 4 passes: result is back in b.keys
	cudaEventRecord(ctx->ev[2], st);
	cudaMemsetAsync(b.arrive, 0, sizeof(uint32_t) * n, st);
	k_karras<<<grid, 256, 0, st>>>(b.keys, (int)n, b.left, b.right, b.parent, b.leaf_prim, b.span_count);
	cudaEventRecord(ctx->ev[3], st);
	k_refit<<<grid, 256, 0, st>>>(n, b.tri_world, b.leaf_prim, b.left, b.right, b.parent, b.aabb, b.arrive);
	k_pack<<<grid, 256, 0, st>>>(n, b.tri_world, b.leaf_prim, b.left, b.right, b.aabb, b.nodes, b.tris);
	cudaEventRecord(ctx->ev[4], st);
	LMB_CUDA(ctx, cudaStreamSynchronize(st));
	LMB_CUDA(ctx, cudaGetLastError());
	cudaEventElapsedTime(&ctx->stats.ms_build_morton, ctx->ev[0], ctx->ev[1]);
	cudaEventElapsedTime(&ctx->stats.ms_build_sort, ctx->ev[1], ctx->ev[2]);
	cudaEventElapsedTime(&ctx->stats.ms_build_tree, ctx->ev[2], ctx->ev[3]);
	cudaEventElapsedTime(&ctx->stats.ms_build_refit, ctx->ev[3], ctx->ev[4]);
	cudaEventElapsedTime(&ctx->stats.ms_build_accel, ctx->ev[0], ctx->ev[4]);
	ctx->stats.ms_build_ploc = 0.0f, ctx->stats.ploc_iterations = 0;
	if (ctx->use_ploc && n > 1 && (rc = build_ploc(ctx))) return rc;
	ctx->stats.tree_cost_ratio = 0.0f;
	if (b.q_left && ctx->tree_auto) {
		// Which binary tree to walk is MEASURED on the two collapsed 8-wide trees: 64 K probe rays that leave random surface points in
		// random directions (wavefront.cu k_probe_rays), node steps + triangle tests counted by the traversal kernel itself. Clustering
		// usually wins (classroom stand-in: 9.5 vs 11.5 nodes per ray); on a regular grid of small closed objects the Morton splits are
		// already object-aligned and the Karras tree is the better one (10 M-triangle torus grid: 19.4 vs 21.4 nodes per ray) -- and
		// neither the surface-area cost of the binary trees (0.90) nor that of the wide trees (0.95) sees that: both favour clustering
		// there, because they price rays that never stop, while a closest-hit ray ends at the first surface.
		double cost_ploc = 0.0, cost_karras = 0.0;
		if ((rc = build_wide_bvh(ctx))) return rc;
		// the traversal stack holds one entry per level (shared + local = 64, trace_wide.cuh): a clustering deeper than 56 levels is not walked
		// at all, not even by the probe (it loses below and the canonical tree, whose depth the 62 key bits bound, is taken)
		const bool ploc_walkable = ctx->wide.levels <= 56;
		if (ploc_walkable && (rc = probe_wide_tree(ctx, 1u << 16, &cost_ploc))) return rc;
		const DeviceWideBvh wide_ploc = ctx->wide;
		const float ms_ploc_wide = ctx->stats.ms_build_wide;
		ctx->wide = DeviceWideBvh{};
		uint32_t* const ql = b.q_left;
		b.q_left = nullptr;  // build_wide_bvh collapses the canonical tree
		rc = build_wide_bvh(ctx);
		if (!rc) rc = probe_wide_tree(ctx, 1u << 16, &cost_karras);
		b.q_left = ql;
		ctx->stats.ms_build_wide += ms_ploc_wide;
		ctx->stats.tree_cost_ratio = cost_karras > 0.0 ? (float)(cost_ploc / cost_karras) : 0.0f;
		if (getenv("LMB_VERBOSE")) fprintf(stderr, "lumen_b200: probe steps per ray: karras %.3f clustered %.3f ratio %.4f\n", cost_karras, cost_ploc, ctx->stats.tree_cost_ratio);
		// the probe is a stand-in for the real ray distribution, where clustering tends to do better than on the probe (classroom stand-in:
		// probe 0.99, path tracing 0.875): the Karras tree has to win by 3 % to be taken
		const bool keep_ploc = !rc && ploc_walkable && cost_ploc < 1.03 * cost_karras;
		if (keep_ploc) {
			free_wide_bvh(ctx);
			ctx->wide = wide_ploc;
		} else {
			DeviceWideBvh loser = wide_ploc;
			cudaFree(loser.nodes), cudaFree(loser.tris), cudaFree(loser.counters);
			free_ploc(ctx);
			ctx->stats.ploc_iterations = 0;
		}
		if (rc) return rc;
	} else {
		if ((rc = build_wide_bvh(ctx))) return rc;
	}
	if (b.q_left && ctx->wide.levels > 56) {
		// the traversal stack holds one entry per level (64 in total): an adversarially deep clustering falls back to the
		// canonical tree, whose depth is bounded by the 62 key bits
		free_ploc(ctx);
		if ((rc = build_wide_bvh(ctx))) return rc;
	}
	ctx->stats.ms_build_accel += ctx->stats.ms_build_wide + ctx->stats.ms_build_ploc;
	ctx->stats.wide_nodes = ctx->wide.n_nodes, ctx->stats.wide_levels = ctx->wide.levels;
	b.built = true;
	return 0;
}

}  // namespace lmb
