// wavefront.cu -- the path integrator as STREAMING wavefront kernels (replaces the single vkCmdTraceRaysKHR(W, H, 1)
// dispatch of src/RayTracer/Path.cpp:39-58 / path.rgen main()).
//
// Path state travels with the path: every bounce has a DENSE list of live paths (position i = 0 .. n-1) whose ray,
// throughput, radiance and pixel id sit at position i of the planes of that bounce's parity; shading writes the survivors
// to consecutive positions of the other parity. A slot-indexed layout (state at pixel slot, queues of slot numbers) loses
// coalescing bounce by bounce as paths die and are regrouped by material -- DRAM bytes per shaded path grew 2.3x from
// bounce 0 to bounce 2 (profiles/r01e_summary.md) -- here every plane access is a run of consecutive positions.
// Per bounce d (parity p = d & 1), in stream order:
//   k_trace     ONE persistent traversal kernel for every ray in flight: continuation rays of list d (path.rgen:48) and
//               the shadow any-hit + MIS-probe rays of the light samples taken at bounce d-1 (pt_commons.glsl:21-22, 32)
//   k_connect   MIS weights of bounce d-1's light samples, added to the radiance of their path at its position in list d
//               (pt_commons.glsl:23-40, path.rgen:78) -- before bounce d adds emission, as in the reference
//   k_classify  retires paths that only waited for that light sample, handles escaped rays (path.rgen:49-55), sorts the rest
//               by the BSDF type they hit into one queue per type
//   k_shade<T>  per BSDF type: hit record, material (+texture), emission, NEE sample + eval, continuation sample,
//               throughput, Russian roulette (path.rgen:56-100); appends survivors to list d+1, light samples to the NEE list
// then k_miss evaluates the sky for the escaped rays (commons.glsl:156-168) and k_film applies the running-mean / sum film
// update in frame order (path.rgen:102-112). A path's radiance reaches acc[pixel slot] exactly once, when it ends.
// RNG state is a pure function of (x, y, frame, draw counter), so reordering paths cannot change any sample; every float
// expression keeps the order of the GLSL (see vec.cuh), including the order in which terms are added to a path's radiance.
#include <stdio.h>
#include <stdlib.h>

#include "context.h"
#include "scene_device.cuh"
#include "trace_persistent.cuh"
#include "trace_wide.cuh"

namespace lmb {

namespace {

#ifndef LMB_SHADE_MIN_BLOCKS
#define LMB_SHADE_MIN_BLOCKS 4
#endif
#ifndef LMB_SHADE_PREFETCH
#define LMB_SHADE_PREFETCH 1
#endif
#ifndef LMB_SHADE_PHASED_TYPES
#define LMB_SHADE_PHASED_TYPES 56u  // mask of LMB_BSDF_* whose k_shade runs its stages block-synchronously: dielectric, conductor, principled
#endif
#ifndef LMB_SHADE_PHASED_THREADS
#define LMB_SHADE_PHASED_THREADS 256  // block size of those kernels: 8 warps in step (128: shade 21.9 ms, 256: 21.3, 512: 21.4)
#endif
#ifndef LMB_SHADE_MIN_BLOCKS_DIFFUSE
#define LMB_SHADE_MIN_BLOCKS_DIFFUSE LMB_SHADE_MIN_BLOCKS
#endif
constexpr float T_MIN = 0.001f;    // path.rgen:19
constexpr float T_MAX = 10000.0f;  // path.rgen:20

struct RenderParams {
	M4 inv_view, inv_proj;
	V3 sky_col;
	uint32_t width, height, n_pix;       // n_pix = pixels THIS context renders = width * (rows of its shard)
	uint32_t row_first, row_stride;      // pixel shard (lmb_set_pixel_shard): local row r is image row row_first + r * row_stride
	uint32_t first_frame, frame_stride;  // frame of batch slot fb = first_frame + fb * frame_stride
	uint32_t n_active;                    // live slots this batch = n_batch_frames * n_pix
	int32_t num_lights, max_depth, light_triangle_count;
	uint32_t dir_light_idx, direct_lighting;
};

// Path state planes of one parity (position-indexed) and the lists shared by both.
struct PathPlanes {
	float4* ray_o;  // origin.xyz, tmin
	float4* ray_d;  // direction.xyz, tmax
	float4* thr;    // throughput.xyz, rng draw counter (bits)
	float4* col;    // radiance.xyz, flags (bits)
	uint32_t* pix;  // pixel slot = frame_in_batch * W*H + y*W + x
};
constexpr uint32_t COL_LAST_SPECULAR = 1u;  // path.rgen:73
constexpr uint32_t COL_ENDED = 2u;          // no continuation ray: the path only waits for the light sample of its last bounce

// Work counters, double-buffered by bounce parity: launch d reads [d & 1] while k_shade(d) fills [(d + 1) & 1], which
// k_trace(d) zeroed when it started (together with the per-material counters). A shading warp appends to three lists per
// trip -- the typed ray queue, the light-sample (NEE) list and the next path list -- with one atomic each: the NEE slot early (its
// latency hides behind the light sample), ray-queue space and the path position back to back at the end.
// Every counter sits on its own 128-byte line (CNT_LINE words apart): each takes one atomic per warp trip, and atomics on one line
// serialise in a single L2 slice. Measured effect over 20 counters packed into 80 bytes: small (shade 7.17 -> 7.11 ms) -- the shading
// kernels are bound by their dependent scene gathers, not by these atomics.
constexpr int CNT_LINE = 32;
enum Counter { CNT_TRACE = 0, CNT_CURSOR = 2 * CNT_LINE, CNT_MISS = 4 * CNT_LINE, CNT_PAIR = 6 * CNT_LINE /* + 0, 2: NEE count; + 1, 3: path count (x CNT_LINE) */,
			   CNT_MAT = 10 * CNT_LINE, CNT_COUNT = 20 * CNT_LINE };
__device__ __forceinline__ uint32_t nee_count(const uint32_t* counters, int parity) { return counters[CNT_PAIR + 2 * parity * CNT_LINE]; }
__device__ __forceinline__ uint32_t path_count(const uint32_t* counters, int parity) { return counters[CNT_PAIR + (2 * parity + 1) * CNT_LINE]; }

// material-sorted shade queues: index = log2(bsdf_type) for the six known types, 6 = unknown type (bsdf_type 0, quirk Q8)
constexpr int N_MAT_QUEUES = 7;

// trace queue entry = index | type << 30: CONTINUE -> path position in the current list, SHADOW / PROBE -> NEE record
constexpr uint32_t RAY_CONTINUE = 0u, RAY_SHADOW = 1u, RAY_PROBE = 2u, IDX_MASK = 0x3FFFFFFFu;

// NEE record: float4 planes of n_slots each, indexed by position in the NEE list
enum NeePlane { NEE_P = 0, NEE_WI, NEE_LDIR, NEE_PROBE_WI, NEE_F2, NEE_LE, NEE_T, NEE_POS, NEE_PLANES };
constexpr uint32_t NEE_FLAG_SHADOW_CONTRIB = 1u;  // pdf_light_w > 0
constexpr uint32_t NEE_FLAG_PROBE = 2u;           // area light and bsdf_pdf != 0
constexpr uint32_t NEE_FLAG_STALE_MATCH = 4u;     // the surface being shaded IS the sampled light triangle (see k_connect)

__device__ __forceinline__ float4 f4(const V3& v, float w) { return make_float4(v.x, v.y, v.z, w); }
__device__ __forceinline__ V3 xyz(const float4& v) { return V3{v.x, v.y, v.z}; }

__device__ __forceinline__ void flush_stats(unsigned long long* stats, int slot, uint32_t v) {
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
	if ((threadIdx.x & 31) == 0 && v) atomicAdd(&stats[slot], (unsigned long long)v);
}

__global__ void k_begin_batch(uint32_t* counters, uint32_t n_active) {
	counters[CNT_TRACE] = n_active, counters[CNT_TRACE + CNT_LINE] = 0;
	counters[CNT_CURSOR] = 0, counters[CNT_CURSOR + CNT_LINE] = 0;
	counters[CNT_MISS] = 0;
	counters[CNT_PAIR] = 0, counters[CNT_PAIR + CNT_LINE] = n_active;
	counters[CNT_PAIR + 2 * CNT_LINE] = 0, counters[CNT_PAIR + 3 * CNT_LINE] = 0;
}
// first thing k_trace(d) does: the lists k_classify(d) / k_shade(d) are about to fill start empty
__device__ __forceinline__ void reset_next_counters(uint32_t* counters, int parity) {
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		counters[CNT_TRACE + (parity ^ 1) * CNT_LINE] = 0;
		counters[CNT_CURSOR + (parity ^ 1) * CNT_LINE] = 0;
		counters[CNT_PAIR + 2 * (parity ^ 1) * CNT_LINE] = 0, counters[CNT_PAIR + (2 * (parity ^ 1) + 1) * CNT_LINE] = 0;
		for (int m = 0; m < N_MAT_QUEUES; m++) counters[CNT_MAT + m * CNT_LINE] = 0;
	}
}

// path.rgen:23-45 + sample_camera (commons.glsl:30-33): list 0 is the pixel slots in order
__global__ void __launch_bounds__(256) k_raygen(RenderParams rp, PathPlanes pl, uint32_t* __restrict__ trace_queue, unsigned long long* stats) {
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[ST_CLOSEST], (unsigned long long)rp.n_active);
	for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < rp.n_active; slot += gridDim.x * blockDim.x) {
		const uint32_t fb = slot / rp.n_pix, pix = slot - fb * rp.n_pix;
		const uint32_t px = pix % rp.width, py = rp.row_first + (pix / rp.width) * rp.row_stride;
		Rng seed{px, py, rp.first_frame + fb * rp.frame_stride, 0u};
		const float j0 = rand1(seed);
		const float j1 = rand1(seed);
		const V2 pixel = v2((float)px, (float)py) + 0.5f;
		const V2 rands = v2(j0, j1) - 0.5f;
		const V2 in_uv = (pixel + rands) / v2((float)rp.width, (float)rp.height);
		const V2 d = in_uv * 2.0f - 1.0f;
		const V3 origin = lmb::xyz(mul(rp.inv_view, v4(0, 0, 0, 1)));
		const V4 target = mul(rp.inv_proj, v4(d.x, d.y, 1, 1));
		const V3 direction = lmb::xyz(mul(rp.inv_view, v4(normalize(lmb::xyz(target)), 0)));
		pl.ray_o[slot] = f4(origin, T_MIN);
		pl.ray_d[slot] = f4(direction, T_MAX);
		pl.thr[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(seed.w));
		pl.col[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
		pl.pix[slot] = slot;
		trace_queue[slot] = slot | (RAY_CONTINUE << 30);
	}
}

// Ray source of the wavefront: typed queue entries over the path planes of the current parity and the NEE planes.
struct WavefrontSource {
	const uint32_t* __restrict__ queue;
	const float4* __restrict__ ray_o;
	const float4* __restrict__ ray_d;
	const float4* __restrict__ nee;
	float4* __restrict__ hit;
	float4* __restrict__ probe_hit;
	uint32_t* __restrict__ shadow_occ;
	uint32_t n_slots;
	// the tag of a ray is its queue entry (index | type << 30)
	__device__ __forceinline__ bool is_any(uint32_t e) const { return (e >> 30) == RAY_SHADOW; }
	__device__ __forceinline__ void load(uint32_t i, V3& o, V3& d, float& tmin, float& tmax, uint32_t& e) const {
		e = queue[i];
		const uint32_t type = e >> 30, idx = e & IDX_MASK;
		if (type == RAY_CONTINUE) {
			const float4 o4 = ray_o[idx], d4 = ray_d[idx];
			o = xyz(o4), d = xyz(d4), tmin = o4.w, tmax = d4.w;
		} else if (type == RAY_SHADOW) {  // pt_commons.glsl:21-22: tmin 0, tmax = wi_len - EPS, terminate on first hit
			const float4 p4 = nee[NEE_P * (size_t)n_slots + idx];
			o = xyz(p4), d = xyz(nee[NEE_WI * (size_t)n_slots + idx]), tmin = 0.0f, tmax = p4.w;
		} else {  // pt_commons.glsl:32
			o = xyz(nee[NEE_P * (size_t)n_slots + idx]), d = xyz(nee[NEE_PROBE_WI * (size_t)n_slots + idx]), tmin = T_MIN, tmax = T_MAX;
		}
	}
	__device__ __forceinline__ void store(uint32_t e, const Hit& h) const {
		const uint32_t type = e >> 30, idx = e & IDX_MASK;
		if (type == RAY_SHADOW)
			shadow_occ[idx] = h.prim != 0xFFFFFFFFu;
		else
			(type == RAY_CONTINUE ? hit : probe_hit)[idx] = make_float4(h.t, h.b1, h.b2, __uint_as_float(h.prim));
	}
};

// Blocks per SM of the wide walker. Issue bound (BVH in the L2, pinned instantiation): 8 x 128 threads at 64 registers = the whole register
// file. Latency bound (BVH beyond the L2, unpinned instantiation): 7 -- one block's shared memory less puts the carve-out one step down
// (164 KB) and leaves the L1 92 KB instead of 60 for nodes and triangles: 10 M-triangle grid 2239 -> 2338 Mrays/s; the classroom stand-in
// loses 2.7 % with 7 (profiles/r02b_summary.md).
constexpr int wide_blocks_per_sm(bool pin) { return pin ? LMB_WIDE_BLOCKS_PER_SM : LMB_WIDE_BLOCKS_PER_SM_UNPINNED; }
template <bool PIN>
__global__ void __launch_bounds__(LMB_TRACE_THREADS, wide_blocks_per_sm(PIN)) k_trace(WideBvhView bvh, WavefrontSource src, uint32_t* counters, int parity, unsigned long long* stats) {
	reset_next_counters(counters, parity);
	trace_wide_persistent<PIN>(bvh, src, counters[CNT_TRACE + parity * CNT_LINE], &counters[CNT_CURSOR + parity * CNT_LINE], stats, -1, -1);
}
// the same rays over the binary LBVH (LMB_TRAVERSAL=bvh2; A/B measurements and the canonical node counts)
__global__ void __launch_bounds__(LMB_TRACE_THREADS) k_trace_bvh2(BvhView bvh, WavefrontSource src, uint32_t* counters, int parity, unsigned long long* stats) {
	reset_next_counters(counters, parity);
	trace_persistent(bvh, src, counters[CNT_TRACE + parity * CNT_LINE], &counters[CNT_CURSOR + parity * CNT_LINE], stats, -1, -1);
}

// Escaped rays of a sun + sky scene wait for k_miss in their own dense record (the path planes are recycled two bounces on).
struct MissPlanes {
	float4* ray_o;
	float4* ray_d;
	float4* thr;
	float4* col;
	uint32_t* pix;
};

// k_classify: walks list d. Paths that ended at bounce d-1 and only waited for k_connect hand their radiance to acc; escaped
// rays get the constant sky or a miss record (path.rgen:49-55); the rest are sorted by the BSDF type of the surface they hit
// (one byte per triangle, L2-resident) into one queue of POSITIONS per type, so that k_shade<TYPE> runs a single lobe's code.
__global__ void __launch_bounds__(256) k_classify(const __grid_constant__ RenderParams rp, const __grid_constant__ DeviceScene sc, int depth, uint32_t* __restrict__ counters, int parity, PathPlanes pl,
													   const float4* __restrict__ hit, uint32_t* __restrict__ mat_queues, MissPlanes ms, float4* __restrict__ acc,
													   uint32_t n_slots) {
	const uint32_t count = path_count(counters, parity);
	const int lane = threadIdx.x & 31;
	constexpr int U = 4;  // positions per lane and trip: independent col -> hit -> table chains in flight
	// Queue space is claimed per BLOCK trip (256 x U positions): warps add their per-destination counts in shared memory and
	// one thread per destination does the global atomic. One atomic per warp and destination serialised on a handful of
	// L2 addresses (~0.85 cycles each, 0.55 M of them per launch) and WAS the kernel's run time (profiles/r01f).
	__shared__ uint32_t s_cnt[8], s_base[8];
	const uint32_t per_block = 256u * U;
	for (uint32_t bbase = blockIdx.x * per_block; bbase < count; bbase += gridDim.x * per_block) {  // block-uniform trip count
		if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
		__syncthreads();
		const uint32_t base = bbase + (threadIdx.x >> 5) * (32u * U);
		uint32_t pos[U], flags[U], prim[U], loc[U];
		int dest[U];  // 0..6 material queue, 7 miss record, -1 retired
#pragma unroll
		for (int u = 0; u < U; u++) {
			pos[u] = base + u * 32 + lane;
			flags[u] = pos[u] < count ? __float_as_uint(pl.col[pos[u]].w) : COL_ENDED;
		}
#pragma unroll
		for (int u = 0; u < U; u++) prim[u] = (pos[u] < count && !(flags[u] & COL_ENDED)) ? __float_as_uint(hit[pos[u]].w) : 0xFFFFFFFFu;
#pragma unroll
		for (int u = 0; u < U; u++) {
			dest[u] = -1;
			if (pos[u] >= count) continue;
			if (flags[u] & COL_ENDED) {
				acc[pl.pix[pos[u]]] = pl.col[pos[u]];
			} else if (prim[u] == 0xFFFFFFFFu) {  // path.rgen:49-55
				if ((depth > 0 || rp.direct_lighting == 1) && rp.dir_light_idx != 0xFFFFFFFFu) {
					dest[u] = 7;  // 64 x 8 step sky march: deferred to k_miss
				} else {
					const float4 c4 = pl.col[pos[u]];
					V3 col = xyz(c4);
					if (depth > 0 || rp.direct_lighting == 1) col += xyz(pl.thr[pos[u]]) * rp.sky_col;  // constant sky (commons.glsl:157-159)
					acc[pl.pix[pos[u]]] = f4(col, 0.0f);
				}
			} else {
				dest[u] = sc.tri_matq[prim[u]];
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dest[u]);
			loc[u] = 0;
			if (dest[u] >= 0) {
				const int leader = __ffs(peers) - 1;
				uint32_t at = 0;
				if (lane == leader) at = atomicAdd(&s_cnt[dest[u]], (uint32_t)__popc(peers));
				loc[u] = __shfl_sync(peers, at, leader) + __popc(peers & ((1u << lane) - 1u));
			}
		}
		__syncthreads();
		if (threadIdx.x < 8 && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(threadIdx.x == 7 ? &counters[CNT_MISS] : &counters[CNT_MAT + threadIdx.x * CNT_LINE], s_cnt[threadIdx.x]);
		__syncthreads();
#pragma unroll
		for (int u = 0; u < U; u++) {
			if (dest[u] < 0) continue;
			const uint32_t at = s_base[dest[u]] + loc[u];
			if (dest[u] == 7) {
				ms.ray_o[at] = pl.ray_o[pos[u]], ms.ray_d[at] = pl.ray_d[pos[u]], ms.thr[at] = pl.thr[pos[u]], ms.col[at] = pl.col[pos[u]];
				ms.pix[at] = pl.pix[pos[u]];
			} else {
				mat_queues[(size_t)dest[u] * n_slots + at] = pos[u];
			}
		}
	}
}

// k_shade<TYPE>: one whole bounce for the paths that hit a surface of BSDF type TYPE (path.rgen:56-100): hit record
// (ray.rchit), material (+texture), emission, depth cut, normal orientation; NEE up to its two traceRayEXT calls
// (pt_commons.glsl:3-20, 28-30); continuation sample, throughput update and Russian roulette. The path state stays in
// registers across the three stages: per path-bounce the kernel reads hit + ray direction + throughput + radiance + pixel id
// at the path's position (68 B, plus L2-resident scene data) and writes the survivor's state to the next free position of the
// other parity, the NEE record to the next free NEE position and <= 3 typed ray-queue entries.
// LAST = true is the bounce at depth max_depth - 1, which only collects emission (path.rgen:57-62) for any BSDF type and
// walks list d directly.
constexpr bool shade_phased(uint32_t type, bool last) { return !last && ((LMB_SHADE_PHASED_TYPES) & type) != 0; }
constexpr int shade_threads(uint32_t type, bool last) { return shade_phased(type, last) ? LMB_SHADE_PHASED_THREADS : 128; }
constexpr int shade_min_blocks(uint32_t type, bool last) {
	return last ? 8 : (type == LMB_BSDF_DIFFUSE ? LMB_SHADE_MIN_BLOCKS_DIFFUSE : LMB_SHADE_MIN_BLOCKS) * 128 / shade_threads(type, last);
}
template <uint32_t TYPE, bool LAST>
__global__ void __launch_bounds__(shade_threads(TYPE, LAST), shade_min_blocks(TYPE, LAST)) k_shade(const __grid_constant__ RenderParams rp, const __grid_constant__ DeviceScene sc, int depth, uint32_t* __restrict__ counters, int parity, int count_idx,
													const uint32_t* __restrict__ queue, PathPlanes pl, PathPlanes nx, const float4* __restrict__ hit,
													uint32_t* __restrict__ trace_queue, float4* __restrict__ nee, uint32_t* __restrict__ nee_path,
													float4* __restrict__ acc, uint32_t n_slots, unsigned long long* stats) {
	const uint32_t count = LAST ? path_count(counters, parity) : counters[count_idx];
	const int lane = threadIdx.x & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint32_t n_shadow = 0, n_probe = 0, n_cont = 0;
	const uint32_t stride = gridDim.x * blockDim.x;
#if LMB_SHADE_PREFETCH
	// The dependent chain queue -> path state / hit -> 128-byte shading record is where this kernel waits (ncu r02b: a third of its
	// stall samples sit on those three loads, at 4 warps per scheduler). The NEXT trip's chain is walked ahead in three short hops
	// spread over the current trip: its queue entry (one register), then its hit's triangle id + L2 prefetches of its state lines,
	// then an L1 prefetch of its shading record.
	uint32_t at_ahead = 0xFFFFFFFFu;
	{
		const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
		if (!LAST && i0 < count) at_ahead = queue[i0];
	}
#endif
	// PHASED (per BSDF type, LMB_SHADE_PHASED_TYPES): the warps of a block take the three stages of a trip in step, with a barrier between the
	// stages. k_shade<principled> is 110 KB of code against a 32 KB instruction cache and its 16 warps per SM drift through it
	// independently: ncu put 47 % of its stall time on no_instruction (5.5 warps per issue). In step, a block has ONE stage's code
	// resident at a time: shade stage 23.15 -> 21.3 ms with dielectric / conductor / principled phased in 256-thread blocks; the diffuse
	// kernel (45 KB, no_instruction 0.4) loses 0.9 ms to the barriers and stays as it was. The trip count is block-uniform here.
	constexpr bool PHASED = shade_phased(TYPE, LAST);
	for (uint32_t base = blockIdx.x * blockDim.x + (PHASED ? 0u : (threadIdx.x & ~31u)); base < count; base += stride) {  // warp- (block-) uniform trip count
		const uint32_t i = base + (PHASED ? threadIdx.x : (uint32_t)lane);
#if LMB_SHADE_PREFETCH
		const uint32_t at_now = at_ahead;
		at_ahead = (!LAST && i + stride < count) ? queue[i + stride] : 0xFFFFFFFFu;
#endif
		// ---- stage 1: hit record, material, emission, depth cut, normal orientation (path.rgen:56-74)
		bool active = i < count;
		uint32_t slot = 0, prim = 0xFFFFFFFFu, out_flags = 0, rng_w = 0;
		V3 col = v3(0.0f), throughput = v3(0.0f), origin = v3(0.0f), wo = v3(0.0f), n_s = v3(0.0f), pos = v3(0.0f), payload_n_s = v3(0.0f);
		bool side = true, want_nee = false;
		uint32_t payload_instance = 0;
#ifdef LMB_SHADE_MATERIAL_REF
		// A/B variant (tools/build_variant.sh matref -DLMB_SHADE_MATERIAL_REF=1): an untextured material is read through
		// a pointer to the scene's record instead of a 104-byte local copy. Off by default: not measured yet, same values either way.
		lmb_material hit_mat_tex;
		const lmb_material* hit_mat_p = &hit_mat_tex;
#define hit_mat (*hit_mat_p)
#else
		lmb_material hit_mat;
#endif
		if (active) {
#if LMB_SHADE_PREFETCH
			const uint32_t at = LAST ? i : at_now;
#else
			const uint32_t at = LAST ? i : queue[i];
#endif
			const float4 c4 = pl.col[at];
			const float4 h4 = hit[at];
			prim = __float_as_uint(h4.w);
			// LAST walks list d itself: paths k_classify already handed to acc (ended, escaped) are skipped
			if (LAST && ((__float_as_uint(c4.w) & COL_ENDED) || prim == 0xFFFFFFFFu)) active = false;
			if (active) {
				const float4 t4 = pl.thr[at];
				const float4 d4 = pl.ray_d[at];
				slot = pl.pix[at];
				throughput = xyz(t4), rng_w = __float_as_uint(t4.w);
				col = xyz(c4);
				const bool last_specular_in = (__float_as_uint(c4.w) & COL_LAST_SPECULAR) != 0;
				const HitPayload payload = build_hit(sc, prim, h4.y, h4.z);
#ifdef LMB_SHADE_MATERIAL_REF
				{
					const lmb_material& m0 = sc.materials[payload.material_idx];
					if (m0.texture_id > -1) {
						hit_mat_tex = load_material(sc, payload.material_idx, payload.uv);
						hit_mat_p = &hit_mat_tex;
					} else {
						hit_mat_p = &m0;
					}
				}
#else
				hit_mat = load_material(sc, payload.material_idx, payload.uv);
#endif
				if ((depth == 0 && rp.direct_lighting == 1) || last_specular_in) col += throughput * v3(hit_mat.emissive_factor);
				if (LAST || depth >= rp.max_depth - 1) {
					acc[slot] = f4(col, 0.0f);
					active = false;
				} else {
					wo = -xyz(d4);
					n_s = payload.n_s, payload_n_s = payload.n_s, payload_instance = payload.instance_idx;
					V3 n_g = payload.n_g;
					if (dot(payload.n_g, wo) < 0.0f) n_g = -n_g;
					if (dot(n_g, payload.n_s) < 0) {
						n_s = -n_s;
						side = false;
					}
					origin = offset_ray(payload.pos, n_g);
					const bool last_specular = (hit_mat.bsdf_props & LMB_FLAG_SPECULAR) != 0;
					out_flags = last_specular ? COL_LAST_SPECULAR : 0u;
					pos = payload.pos;
					want_nee = !last_specular && (depth > 0 || rp.direct_lighting == 1);
				}
			}
		}
		if (LAST) continue;
#if LMB_SHADE_PREFETCH
		uint32_t prim_ahead = 0xFFFFFFFFu;
		if (at_ahead != 0xFFFFFFFFu) {
			prim_ahead = __float_as_uint(hit[at_ahead].w);
			asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.col + at_ahead));
			asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.thr + at_ahead));
			asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.ray_d + at_ahead));
			asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.pix + at_ahead));
		}
#endif
		if (PHASED) __syncthreads();
		const uint32_t fb = slot / rp.n_pix, pix = slot - fb * rp.n_pix;
		Rng seed{pix % rp.width, rp.row_first + (pix / rp.width) * rp.row_stride, rp.first_frame + fb * rp.frame_stride, rng_w};
		// ---- stage 2: next-event estimation (path.rgen:75-80, pt_commons.glsl:3-20, 28-30). The NEE record goes straight to
		// its position in the NEE list (one atomic per warp, its latency hidden behind the light sample)
		const uint32_t b_sh = __ballot_sync(0xFFFFFFFFu, want_nee);
		uint32_t k = 0;
		if (b_sh) {
			if (lane == 0) k = atomicAdd(&counters[CNT_PAIR + 2 * (parity ^ 1) * CNT_LINE], (uint32_t)__popc(b_sh));
			k = __shfl_sync(0xFFFFFFFFu, k, 0) + __popc(b_sh & lt_mask);
		}
		bool do_probe = false;
		if (want_nee) {
			const V4 r4 = rand4(seed);
			const LightSample ls = sample_light_Li(sc, r4, pos, rp.num_lights);
			const V3 p = offset_ray2(pos, n_s);
			float bsdf_pdf;
			const float cos_x = dot(n_s, ls.wi);
			const V3 f = eval_bsdf_t<TYPE>(n_s, wo, hit_mat, side, ls.wi, bsdf_pdf);
			uint32_t flags = 0;
			V3 ldir = v3(0.0f);
			if (ls.pdf_w > 0) {
				flags |= NEE_FLAG_SHADOW_CONTRIB;
				const float mis_weight = ((ls.flags >> 5) & 1u) ? 1.0f : 1.0f / (1.0f + bsdf_pdf / ls.pdf_w);
				ldir = mis_weight * f * fabsf(cos_x) * ls.Le / ls.pdf_w;
			}
			n_shadow++;
			nee[NEE_P * (size_t)n_slots + k] = f4(p, ls.wi_len - LMB_EPS);
			nee[NEE_LDIR * (size_t)n_slots + k] = f4(ldir, ls.pdf_a);
			nee[NEE_T * (size_t)n_slots + k] = f4(throughput, __uint_as_float(ls.instance_idx));
			if ((ls.flags & 0x7u) == LMB_LIGHT_AREA) {
				const V3 r3 = rand3(seed);
				const BsdfSample bs = sample_bsdf_t<TYPE>(n_s, wo, hit_mat, 1, side, r3);
				if (bs.pdf != 0) {
					flags |= NEE_FLAG_PROBE;
					// ray.rmiss writes only material_idx: when the probe misses, the reference still compares the ids of the
					// payload left by the previous trace, i.e. of the surface being shaded, and then uses its pos / n_s
					// (wi_len = |pos - pos| = 0). That can only match when this surface is the sampled light triangle.
					float g_stale = 0.0f;
					if (payload_instance == ls.instance_idx && sc.tri_local[prim] == ls.triangle_idx) {
						flags |= NEE_FLAG_STALE_MATCH;
						const float wi_len = length(pos - pos);
						g_stale = fabsf(dot(payload_n_s, -bs.wi)) / (wi_len * wi_len);  // ray.rchit's un-flipped shading normal
					}
					nee[NEE_PROBE_WI * (size_t)n_slots + k] = f4(bs.wi, bs.pdf);
					nee[NEE_F2 * (size_t)n_slots + k] = f4(bs.f, fabsf(bs.cos_theta));
					nee[NEE_LE * (size_t)n_slots + k] = f4(ls.Le, __uint_as_float(ls.triangle_idx));
					nee[NEE_POS * (size_t)n_slots + k] = f4(pos, g_stale);
					do_probe = true;
					n_probe++;
				}
			}
			nee[NEE_WI * (size_t)n_slots + k] = f4(ls.wi, __uint_as_float(flags));
		}
#if LMB_SHADE_PREFETCH
		if (prim_ahead != 0xFFFFFFFFu) asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.tri_shade + 8 * (size_t)prim_ahead));
#endif
		// ---- stage 3: continuation sample, throughput, Russian roulette (path.rgen:81-100)
		if (PHASED) __syncthreads();
		bool alive = false;
		V3 wi_next = v3(0.0f);
		if (active) {
			const V3 r3 = rand3(seed);
			const BsdfSample bs = sample_bsdf_t<TYPE>(n_s, wo, hit_mat, 1, side, r3);
			alive = bs.pdf != 0;
			if (alive) {
				throughput *= bs.f * fabsf(bs.cos_theta) / bs.pdf;
				float rr_scale = 1.0f;
				if (has_prop(hit_mat.bsdf_props, LMB_FLAG_TRANSMISSION)) rr_scale *= side ? 1.0f / hit_mat.ior : hit_mat.ior;
				if (depth > 3) {
					const float rr_prob = gmin(0.95f, luminance(throughput) * rr_scale);
					if (rr_prob == 0 || rr_prob < rand1(seed))
						alive = false;
					else
						throughput /= rr_prob;
				}
			}
			wi_next = bs.wi;
			if (alive) n_cont++;
			else if (!want_nee) acc[slot] = f4(col, 0.0f);  // the path ends here with nothing pending
		}
		// ---- one position in list d+1 per path that goes on (or waits for its light sample) and <= 3 typed ray-queue entries
		const bool go_on = alive || want_nee;
		const uint32_t b_pr = __ballot_sync(0xFFFFFFFFu, do_probe), b_ct = __ballot_sync(0xFFFFFFFFu, alive), b_go = __ballot_sync(0xFFFFFFFFu, go_on);
		if (b_go) {
			const uint32_t n_sh = __popc(b_sh), n_pr = __popc(b_pr);
			uint32_t at = 0, at_path = 0;
			if (lane == 0) {
				at = atomicAdd(&counters[CNT_TRACE + (parity ^ 1) * CNT_LINE], n_sh + n_pr + (uint32_t)__popc(b_ct));
				at_path = atomicAdd(&counters[CNT_PAIR + (2 * (parity ^ 1) + 1) * CNT_LINE], (uint32_t)__popc(b_go));
			}
			at = __shfl_sync(0xFFFFFFFFu, at, 0), at_path = __shfl_sync(0xFFFFFFFFu, at_path, 0);
			if (go_on) {
				const uint32_t j = at_path + __popc(b_go & lt_mask);
				nx.col[j] = f4(col, __uint_as_float(out_flags | (alive ? 0u : COL_ENDED)));
				nx.pix[j] = slot;
				if (alive) {
					nx.thr[j] = f4(throughput, __uint_as_float(seed.w));
					nx.ray_o[j] = f4(origin, T_MIN);
					nx.ray_d[j] = f4(wi_next, T_MAX);
					trace_queue[at + n_sh + n_pr + __popc(b_ct & lt_mask)] = j | (RAY_CONTINUE << 30);
				}
				if (want_nee) {
					nee_path[k] = j;
					trace_queue[at + __popc(b_sh & lt_mask)] = k | (RAY_SHADOW << 30);
					if (do_probe) trace_queue[at + n_sh + __popc(b_pr & lt_mask)] = k | (RAY_PROBE << 30);
				}
			}
		}
	}
	flush_stats(stats, ST_SHADOW, n_shadow);
	flush_stats(stats, ST_PROBE, n_probe);
	flush_stats(stats, ST_CLOSEST, n_cont);
}
#ifdef LMB_SHADE_MATERIAL_REF
#undef hit_mat
#endif

// pt_commons.glsl:23-27, 33-39 and the accumulation of path.rgen:78, once both rays of the light sample are traced: the
// result goes to the radiance of the path at its position in the current list (before this bounce adds emission)
__global__ void __launch_bounds__(128) k_connect(const __grid_constant__ RenderParams rp, const __grid_constant__ DeviceScene sc, const uint32_t* __restrict__ counters, int parity,
												  const uint32_t* __restrict__ nee_path, const float4* __restrict__ nee, const float4* __restrict__ probe_hit,
												  const uint32_t* __restrict__ shadow_occ, float4* __restrict__ colb, uint32_t n_slots) {
	const uint32_t count = nee_count(counters, parity);
	const float light_pick_pdf = 1.0f / (float)rp.light_triangle_count;
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const float4 wi4 = nee[NEE_WI * (size_t)n_slots + k];
		const float4 l4 = nee[NEE_LDIR * (size_t)n_slots + k];
		const float4 t4 = nee[NEE_T * (size_t)n_slots + k];
		const uint32_t flags = __float_as_uint(wi4.w);
		V3 res = v3(0.0f);
		if (shadow_occ[k] == 0u && (flags & NEE_FLAG_SHADOW_CONTRIB)) res += xyz(l4);
		if (flags & NEE_FLAG_PROBE) {
			const float4 pw4 = nee[NEE_PROBE_WI * (size_t)n_slots + k];
			const float4 f4v = nee[NEE_F2 * (size_t)n_slots + k];
			const float4 le4 = nee[NEE_LE * (size_t)n_slots + k];
			const float4 pos4 = nee[NEE_POS * (size_t)n_slots + k];
			const V3 wi = xyz(pw4);
			const float bsdf_pdf = pw4.w;
			const float4 ph = probe_hit[k];
			const uint32_t prim = __float_as_uint(ph.w);
			bool match = false;
			float g = 0.0f;
			if (prim != 0xFFFFFFFFu) {
				if (sc.tri_local[prim] == __float_as_uint(le4.w) && sc.tri_mesh[prim] == __float_as_uint(t4.w)) {
					const HitPayload pl = build_hit(sc, prim, ph.y, ph.z);
					const float wi_len = length(pl.pos - xyz(pos4));
					g = fabsf(dot(pl.n_s, -wi)) / (wi_len * wi_len);
					match = true;
				}
			} else if (flags & NEE_FLAG_STALE_MATCH) {
				g = pos4.w;
				match = true;
			}
			if (match) {
				const float mis_weight = 1.0f / (1 + l4.w / (g * bsdf_pdf));
				res += xyz(f4v) * mis_weight * f4v.w * xyz(le4) / bsdf_pdf;
			}
		}
		const uint32_t j = nee_path[k];
		const float4 c4 = colb[j];
		const V3 col = xyz(c4) + xyz(t4) * res / light_pick_pdf;
		colb[j] = f4(col, c4.w);
	}
}

// Escaped rays with a sun + sky light: col += throughput * shade_atmosphere(...) (path.rgen:50-53, commons.glsl:156-168).
// A path misses at most once and nothing is added to its radiance afterwards, so the march may run any time between the k_classify
// that found the miss and the film update, and it keeps the order of the float additions.
// The march (64 x 9 density evaluations, ~1500 exponentials per ray) is pure FP32 issue with no memory traffic -- 13.6 ms of a 100 ms
// step when it ran after the bounce loop -- while k_trace leaves a quarter of the issue slots idle and k_shade two thirds. So it runs
// BESIDE the bounce loop: after k_classify(d) a one-thread kernel notes how many miss records exist (marks[d + 1]) and a small
// persistent grid on a second stream marches records [marks[d], marks[d + 1]) while the render stream goes on with bounce d + 1.
// Work is handed out 32 records at a time from a cursor per range; when the bounce loop is over the render stream runs the same
// kernel over ALL ranges with a full grid, so whatever the side grid has not reached is finished at full width, then waits for it.
constexpr int MISS_RANGES = 64;  // the first 63 bounces have a range of their own; deeper bounces share the last one
__global__ void k_miss_mark(const uint32_t* __restrict__ counters, uint32_t* __restrict__ marks, int range, bool first) {
	const uint32_t begin = first ? 0u : marks[range];
	if (first) marks[range] = 0u;
	marks[range + 1] = counters[CNT_MISS];
	marks[MISS_RANGES + 1 + range] = begin;
}
__global__ void __launch_bounds__(128, 8) k_miss(const __grid_constant__ RenderParams rp, const __grid_constant__ DeviceScene sc, uint32_t* __restrict__ marks, int range_first, int range_last, MissPlanes ms,
													 float4* __restrict__ acc) {
	const int lane = threadIdx.x & 31;
	for (int range = range_first; range <= range_last; range++) {
		const uint32_t end = marks[range + 1];
		uint32_t* cursor = &marks[MISS_RANGES + 1 + range];
		if (*(volatile uint32_t*)cursor >= end) continue;
		for (;;) {
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(cursor, 32u);
			base = __shfl_sync(0xFFFFFFFFu, base, 0);
			if (base >= end) break;
			const uint32_t i = base + lane;
			if (i < end) {
				const float4 c4 = ms.col[i];
				const V3 sky = shade_atmosphere(sc, rp.dir_light_idx, rp.sky_col, xyz(ms.ray_o[i]), xyz(ms.ray_d[i]), T_MAX);
				const V3 col = xyz(c4) + xyz(ms.thr[i]) * sky;
				acc[ms.pix[i]] = f4(col, 0.0f);
			}
		}
	}
}

// path.rgen:102-112, applied for the batch's frames in order
__global__ void __launch_bounds__(256) k_film(RenderParams rp, uint32_t n_batch_frames, int film_mode, const float4* __restrict__ colb,
											   float4* __restrict__ film, unsigned long long* stats) {
	uint32_t nan_count = 0;
	for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < rp.n_pix; pix += gridDim.x * blockDim.x) {
		const uint32_t gpix = (rp.row_first + (pix / rp.width) * rp.row_stride) * rp.width + pix % rp.width;  // position in the full image
		float4 acc = film[gpix];
		for (uint32_t fb = 0; fb < n_batch_frames; fb++) {
			const V3 col = xyz(colb[(size_t)fb * rp.n_pix + pix]);
			const float lum = luminance(col);
			if (lum != lum) {
				nan_count++;
				continue;
			}
			if (film_mode == LMB_FILM_RUNNING_MEAN) {
				const uint32_t frame = rp.first_frame + fb * rp.frame_stride;
				if (frame > 0) {
					const float w = 1.0f / float(frame + 1);
					const V3 m = mix(xyz(acc), col, w);
					acc = make_float4(m.x, m.y, m.z, 1.0f);
				} else {
					acc = make_float4(col.x, col.y, col.z, 1.0f);
				}
			} else {
				acc = make_float4(acc.x + col.x, acc.y + col.y, acc.z + col.z, acc.w + 1.0f);
			}
		}
		film[gpix] = acc;
	}
	flush_stats(stats, ST_NAN, nan_count);
}

__global__ void __launch_bounds__(256) k_resolve(uint32_t n_pix, float4* film) {
	for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < n_pix; pix += gridDim.x * blockDim.x) {
		const float4 a = film[pix];
		film[pix] = a.w > 0.0f ? make_float4(a.x / a.w, a.y / a.w, a.z / a.w, 1.0f) : make_float4(0.0f, 0.0f, 0.0f, 1.0f);
	}
}

// Ray source over a plain array of (o, tmin, d, tmax) rays: lmb_trace_closest / lmb_trace_any
struct ArraySource {
	const float4* __restrict__ rays;
	float4* __restrict__ hits;
	uint8_t* __restrict__ occ;
	bool any_hit;
	const uint32_t* __restrict__ order;  // optional: entry i = index of the i-th ray to trace (sort_rays)
	__device__ __forceinline__ bool is_any(uint32_t) const { return any_hit; }
	__device__ __forceinline__ void load(uint32_t i, V3& o, V3& d, float& tmin, float& tmax, uint32_t& tag) const {
		if (order) i = order[i];
		const float4 o4 = rays[2 * (size_t)i], d4 = rays[2 * (size_t)i + 1];
		o = xyz(o4), d = xyz(d4), tmin = o4.w, tmax = d4.w, tag = i;
	}
	__device__ __forceinline__ void store(uint32_t i, const Hit& h) const {
		if (any_hit)
			occ[i] = h.prim != 0xFFFFFFFFu;
		else
			hits[i] = make_float4(h.t, h.b1, h.b2, __uint_as_float(h.prim));
	}
};

template <bool PIN>
__global__ void __launch_bounds__(LMB_TRACE_THREADS, wide_blocks_per_sm(PIN)) k_trace_array(WideBvhView bvh, ArraySource src, uint32_t n, const uint32_t* __restrict__ n_dev,
																							 uint32_t* cursor, unsigned long long* stats, int count_rays) {
	if (n_dev) n = min(n, *n_dev);  // a list whose length lives on the device (BDPT's emitted connection rays)
	trace_wide_persistent<PIN>(bvh, src, n, cursor, stats, count_rays ? ST_CLOSEST : -1, count_rays ? ST_SHADOW : -1);
}
__global__ void __launch_bounds__(LMB_TRACE_THREADS) k_trace_array_bvh2(BvhView bvh, ArraySource src, uint32_t n, uint32_t* cursor, unsigned long long* stats, int count_rays) {
	trace_persistent(bvh, src, n, cursor, stats, count_rays ? ST_CLOSEST : -1, count_rays ? ST_SHADOW : -1);
}

BvhView view_of(const lmb_ctx* ctx) { return BvhView{ctx->bvh.nodes, ctx->bvh.tris, ctx->bvh.n}; }
static uint32_t env_u32(const char* name, uint32_t dflt) {
	const char* v = getenv(name);
	return (v && *v) ? (uint32_t)atoi(v) : dflt;
}
// Which instantiation of the wide walker to launch (trace_wide.cuh, PIN): the pinned one while nodes + triangles fit the L2 with room
// to spare (issue bound), the unpinned one beyond that (latency bound). LMB_TRACE_PIN=0|1 overrides for A/B runs.
bool trace_pinned(const lmb_ctx* ctx) {
	if (ctx->trace_pin >= 0) return ctx->trace_pin != 0;
	return (size_t)ctx->wide.n_nodes * 80 + (size_t)ctx->wide.n_tris * 48 <= (size_t)96 << 20;
}
WideBvhView wide_view_of(const lmb_ctx* ctx) {
	// scheduling thresholds of k_trace (results do not depend on them): tuning overrides for A/B runs
	static const uint32_t refill = env_u32("LMB_REFILL_LANES", LMB_WIDE_REFILL_LANES), round = env_u32("LMB_TRI_ROUND_LANES", 0);
	const uint32_t round_dflt = trace_pinned(ctx) ? LMB_TRI_ROUND_LANES : LMB_TRI_ROUND_LANES_UNPINNED;
	return WideBvhView{ctx->wide.nodes, ctx->wide.tris, ctx->wide.n_tris, 0x3F800000u, refill, round ? round : round_dflt};
}

}  // namespace

// rows of the image this context renders: row_first, row_first + row_stride, ... below height
uint32_t shard_rows(const lmb_ctx* ctx) {
	return ctx->row_first < ctx->height ? (ctx->height - ctx->row_first + ctx->row_stride - 1) / ctx->row_stride : 0u;
}

int wavefront_alloc(lmb_ctx* ctx, uint32_t frames_in_flight) {
	wavefront_free(ctx);
	Wavefront& wf = ctx->wf;
	const uint64_t n_pix = (uint64_t)ctx->width * shard_rows(ctx);
	if (frames_in_flight == 0) {
		// default: about 32 M path slots in flight, at most 64 frames (~11 GB of wavefront state out of 180 GB). Launch tails and the
		// thinning late bounces amortise over the batch: 1080p at 4 / 8 / 16 frames in flight = 2989 / 3038 / 3067 Mrays/s.
		frames_in_flight = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, (32ull << 20) / std::max<uint64_t>(n_pix, 1)));
	}
	const uint64_t n_slots = n_pix * frames_in_flight;
	if (n_slots == 0 || n_slots > 0x3FFFFFFFull) return set_error(ctx, LMB_ERR_INVALID, "width*height*frames_in_flight out of range");
	wf.n_slots = (uint32_t)n_slots;
	wf.frames_in_flight = frames_in_flight;
	auto alloc = [&](void** p, size_t bytes) { return check_cuda(ctx, cudaMalloc(p, bytes), "cudaMalloc(wavefront)"); };
	int rc;
	for (int p = 0; p < 2; p++) {
		if ((rc = alloc((void**)&wf.ray_o[p], n_slots * 16))) return rc;
		if ((rc = alloc((void**)&wf.ray_d[p], n_slots * 16))) return rc;
		if ((rc = alloc((void**)&wf.thr[p], n_slots * 16))) return rc;
		if ((rc = alloc((void**)&wf.col[p], n_slots * 16))) return rc;
		if ((rc = alloc((void**)&wf.pix[p], n_slots * 4))) return rc;
	}
	if ((rc = alloc((void**)&wf.hit, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.acc, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.nee, n_slots * 16 * NEE_PLANES))) return rc;
	if ((rc = alloc((void**)&wf.nee_path, n_slots * 4))) return rc;
	if ((rc = alloc((void**)&wf.miss_ray_o, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.miss_ray_d, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.miss_thr, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.miss_col, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.miss_pix, n_slots * 4))) return rc;
	if ((rc = alloc((void**)&wf.mat_queues, n_slots * 4 * N_MAT_QUEUES))) return rc;
	if ((rc = alloc((void**)&wf.trace_queue, n_slots * 4 * 3))) return rc;
	if ((rc = alloc((void**)&wf.probe_hit, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.shadow_occ, n_slots * 4))) return rc;
	if ((rc = alloc((void**)&wf.counters, CNT_COUNT * 4))) return rc;
	if ((rc = alloc((void**)&wf.miss_marks, (2 * MISS_RANGES + 2) * 4))) return rc;
	if ((rc = alloc((void**)&wf.stats, ST_COUNT * 8))) return rc;
	LMB_CUDA(ctx, cudaMemsetAsync(wf.stats, 0, ST_COUNT * 8, ctx->stream));
	return 0;
}

void wavefront_free(lmb_ctx* ctx) {
	bdpt_free(ctx);  // sized by the same image
	Wavefront& wf = ctx->wf;
	for (int p = 0; p < 2; p++) cudaFree(wf.ray_o[p]), cudaFree(wf.ray_d[p]), cudaFree(wf.thr[p]), cudaFree(wf.col[p]), cudaFree(wf.pix[p]);
	cudaFree(wf.hit), cudaFree(wf.acc), cudaFree(wf.nee), cudaFree(wf.nee_path);
	cudaFree(wf.miss_ray_o), cudaFree(wf.miss_ray_d), cudaFree(wf.miss_thr), cudaFree(wf.miss_col), cudaFree(wf.miss_pix);
	cudaFree(wf.mat_queues), cudaFree(wf.trace_queue), cudaFree(wf.probe_hit), cudaFree(wf.shadow_occ), cudaFree(wf.trace_cursor), cudaFree(wf.counters), cudaFree(wf.miss_marks), cudaFree(wf.stats);
	wf = Wavefront{};
}

int wavefront_render(lmb_ctx* ctx, const lmb_pc_path& pc, const lmb_scene_ubo& ubo, uint32_t first_frame, uint32_t n_frames, uint32_t stride,
					 int film_mode) {
	Wavefront& wf = ctx->wf;
	cudaStream_t st = ctx->stream;
	RenderParams rp;
	auto load = [](const float* p) {
		M4 m;
		for (int c = 0; c < 4; c++) m.c[c] = V4{p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]};
		return m;
	};
	rp.inv_view = load(ubo.inv_view);
	rp.inv_proj = load(ubo.inv_projection);
	rp.sky_col = V3{pc.sky_col[0], pc.sky_col[1], pc.sky_col[2]};
	rp.width = ctx->width, rp.height = ctx->height, rp.n_pix = ctx->width * shard_rows(ctx);
	rp.row_first = ctx->row_first, rp.row_stride = ctx->row_stride;
	rp.frame_stride = stride;
	rp.num_lights = pc.num_lights, rp.max_depth = pc.max_depth, rp.light_triangle_count = pc.light_triangle_count;
	rp.dir_light_idx = pc.dir_light_idx, rp.direct_lighting = pc.direct_lighting;
	const BvhView bvh = view_of(ctx);
	const WideBvhView wide = wide_view_of(ctx);
	const bool pinned = trace_pinned(ctx);
	const int grid_wide = ctx->sm_count * 16;
	// k_shade walks its queue with a grid-stride loop: the grid is a whole number of waves of the blocks one SM holds (A/B: LMB_SHADE_GRID_PER_SM)
	static const uint32_t shade_grid_per_sm = env_u32("LMB_SHADE_GRID_PER_SM", 16);
	const int grid_shade = ctx->sm_count * (int)shade_grid_per_sm;
	const int grid_256 = ctx->sm_count * 8;
	const int grid_trace = ctx->sm_count * 7;  // persistent: 7 blocks x 32 KB stack fit one SM's shared memory
	const int grid_trace_wide = ctx->sm_count * LMB_WIDE_BLOCKS_PER_SM;
	const PathPlanes planes[2] = {{wf.ray_o[0], wf.ray_d[0], wf.thr[0], wf.col[0], wf.pix[0]}, {wf.ray_o[1], wf.ray_d[1], wf.thr[1], wf.col[1], wf.pix[1]}};
	const MissPlanes miss{wf.miss_ray_o, wf.miss_ray_d, wf.miss_thr, wf.miss_col, wf.miss_pix};
	float ms;
	const bool prof = ctx->profile_stages;  // per-stage timing serialises the bounce loop; off by default
	// sky march beside the bounce loop (k_miss): LMB_MISS_SIDE_BLOCKS = blocks per SM of the side grid, 0 = march after the loop only
	static const uint32_t miss_side_blocks_env = env_u32("LMB_MISS_SIDE_BLOCKS", 0);
	const bool sky_march = pc.dir_light_idx != 0xFFFFFFFFu;
	const uint32_t miss_side_blocks = (sky_march && !prof) ? miss_side_blocks_env : 0u;
	bool side_launched = false;
	if (miss_side_blocks > 0 && !ctx->miss_stream) {
		int lo = 0, hi = 0;
		cudaDeviceGetStreamPriorityRange(&lo, &hi);
		LMB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->miss_stream, cudaStreamNonBlocking, hi));  // high priority: its few blocks get a seat as soon as one frees
		LMB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_miss_ready, cudaEventDisableTiming));
		LMB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_miss_done, cudaEventDisableTiming));
	}
	cudaEventRecord(ctx->ev[0], st);
	for (uint32_t done = 0; done < n_frames;) {
		const uint32_t nb = std::min(wf.frames_in_flight, n_frames - done);
		rp.first_frame = first_frame + done * stride;
		rp.n_active = nb * rp.n_pix;
		if (prof) cudaEventRecord(ctx->ev[1], st);
		k_begin_batch<<<1, 1, 0, st>>>(wf.counters, rp.n_active);
		k_raygen<<<grid_256, 256, 0, st>>>(rp, planes[0], wf.trace_queue, wf.stats);
		ctx->stats.kernel_launches += 2;
		if (prof) {
			cudaEventRecord(ctx->ev[2], st);
			cudaEventSynchronize(ctx->ev[2]);
			cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
			ctx->stats.ms_film += ms;
		}
		for (int depth = 0; depth < std::max(pc.max_depth, 1); depth++) {
			const int par = depth & 1;
			const WavefrontSource src{wf.trace_queue, wf.ray_o[par], wf.ray_d[par], wf.nee, wf.hit, wf.probe_hit, wf.shadow_occ, wf.n_slots};
			if (prof) cudaEventRecord(ctx->ev[1], st);
			if (ctx->use_bvh2)
				k_trace_bvh2<<<grid_trace, LMB_TRACE_THREADS, 0, st>>>(bvh, src, wf.counters, par, wf.stats);
			else if (pinned)
				k_trace<true><<<grid_trace_wide, LMB_TRACE_THREADS, 0, st>>>(wide, src, wf.counters, par, wf.stats);
			else
				k_trace<false><<<ctx->sm_count * wide_blocks_per_sm(false), LMB_TRACE_THREADS, 0, st>>>(wide, src, wf.counters, par, wf.stats);
			ctx->stats.kernel_launches += 1;
			if (prof) cudaEventRecord(ctx->ev[2], st);
			if (ctx->stats_per_launch) {
				// calibration aid (LMB_STATS_PER_LAUNCH=1, tools/calibrate_ktrace.py): the counters of THIS k_trace launch, in launch order,
				// to be joined with ncu's per-launch smsp__inst_executed.sum
				unsigned long long h[ST_COUNT];
				cudaMemcpyAsync(h, wf.stats, sizeof(h), cudaMemcpyDeviceToHost, st);
				cudaStreamSynchronize(st);
				static thread_local unsigned long long prev[ST_COUNT] = {};
				fprintf(stderr, "k_trace_launch {\"rays\": %llu, \"nodes\": %llu, \"tris\": %llu, \"iters\": %llu, \"node_trips\": %llu, \"rounds\": %llu, \"refills\": %llu}\n",
						(h[ST_CLOSEST] + h[ST_SHADOW] + h[ST_PROBE]) - (prev[ST_CLOSEST] + prev[ST_SHADOW] + prev[ST_PROBE]), h[ST_NODES] - prev[ST_NODES],
						h[ST_TRIS] - prev[ST_TRIS], h[ST_W_ITERS] - prev[ST_W_ITERS], h[ST_W_ITERS] - prev[ST_W_ITERS], h[ST_W_ROUNDS] - prev[ST_W_ROUNDS],
						h[ST_W_REFILLS] - prev[ST_W_REFILLS]);
				memcpy(prev, h, sizeof(h));
			}
			if (depth > 0) {
				k_connect<<<grid_wide, 128, 0, st>>>(rp, ctx->scene, wf.counters, par, wf.nee_path, wf.nee, wf.probe_hit, wf.shadow_occ, wf.col[par], wf.n_slots);
				ctx->stats.kernel_launches += 1;
			}
			if (prof) cudaEventRecord(ctx->ev[3], st);
#define LMB_SHADE_ARGS(mq, cidx) rp, ctx->scene, depth, wf.counters, par, cidx, mq, planes[par], planes[par ^ 1], wf.hit, wf.trace_queue, wf.nee, wf.nee_path, wf.acc, wf.n_slots, wf.stats
			const bool last = depth >= pc.max_depth - 1;  // the last bounce only collects emission (path.rgen:57-62)
			k_classify<<<grid_256, 256, 0, st>>>(rp, ctx->scene, depth, wf.counters, par, planes[par], wf.hit, wf.mat_queues, miss, wf.acc, wf.n_slots);
			ctx->stats.kernel_launches += 1;
			if (miss_side_blocks > 0 && depth < MISS_RANGES - 1) {  // the rays that escaped at this bounce: marched beside the bounces still to come
				k_miss_mark<<<1, 1, 0, st>>>(wf.counters, wf.miss_marks, depth, depth == 0);
				ctx->stats.kernel_launches += 1;
				if (miss_side_blocks > 0 && !last) {
					cudaEventRecord(ctx->ev_miss_ready, st);
					cudaStreamWaitEvent(ctx->miss_stream, ctx->ev_miss_ready, 0);
					k_miss<<<ctx->sm_count * miss_side_blocks, 128, 0, ctx->miss_stream>>>(rp, ctx->scene, wf.miss_marks, depth, depth, miss, wf.acc);
					ctx->stats.kernel_launches += 1;
					side_launched = true;
				}
			}
			if (last) {
				k_shade<0u, true><<<grid_shade, 128, 0, st>>>(LMB_SHADE_ARGS(wf.mat_queues, 0));
				ctx->stats.kernel_launches += 1;
			} else {
				for (int m = 0; m < N_MAT_QUEUES; m++) {
					if (!(ctx->mat_queue_mask & (1u << m))) continue;  // BSDF type absent from the scene (ENABLE_* macros, LumenScene.cpp:217-228)
					const uint32_t* mq = wf.mat_queues + (size_t)m * wf.n_slots;
					switch (m) {
						case 0: k_shade<LMB_BSDF_DIFFUSE, false><<<grid_shade * 128 / shade_threads(LMB_BSDF_DIFFUSE, false), shade_threads(LMB_BSDF_DIFFUSE, false), 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m * CNT_LINE)); break;
						case 1: k_shade<LMB_BSDF_MIRROR, false><<<grid_shade * 128 / shade_threads(LMB_BSDF_MIRROR, false), shade_threads(LMB_BSDF_MIRROR, false), 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m * CNT_LINE)); break;
						case 2: k_shade<LMB_BSDF_GLASS, false><<<grid_shade * 128 / shade_threads(LMB_BSDF_GLASS, false), shade_threads(LMB_BSDF_GLASS, false), 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m * CNT_LINE)); break;
						case 3: k_shade<LMB_BSDF_DIELECTRIC, false><<<grid_shade * 128 / shade_threads(LMB_BSDF_DIELECTRIC, false), shade_threads(LMB_BSDF_DIELECTRIC, false), 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m * CNT_LINE)); break;
						case 4: k_shade<LMB_BSDF_CONDUCTOR, false><<<grid_shade * 128 / shade_threads(LMB_BSDF_CONDUCTOR, false), shade_threads(LMB_BSDF_CONDUCTOR, false), 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m * CNT_LINE)); break;
						case 5: k_shade<LMB_BSDF_PRINCIPLED, false><<<grid_shade * 128 / shade_threads(LMB_BSDF_PRINCIPLED, false), shade_threads(LMB_BSDF_PRINCIPLED, false), 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m * CNT_LINE)); break;
						default: k_shade<0u, false><<<grid_shade * 128 / shade_threads(0u, false), shade_threads(0u, false), 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m * CNT_LINE)); break;
					}
					ctx->stats.kernel_launches += 1;
				}
			}
#undef LMB_SHADE_ARGS
			if (prof) {
				cudaEventRecord(ctx->ev[4], st);
				cudaEventSynchronize(ctx->ev[4]);
				cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
				ctx->stats.ms_extend += ms;
				cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
				ctx->stats.ms_connect += ms;
				cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
				ctx->stats.ms_shade += ms;
			}
		}
		if (prof) cudaEventRecord(ctx->ev[1], st);
		if (sky_march) {
			const int n_bounces = std::max(pc.max_depth, 1);
			int n_ranges = std::min(n_bounces, MISS_RANGES);
			if (miss_side_blocks == 0) {  // nothing marched beside the loop: ONE range over every record of the batch
				k_miss_mark<<<1, 1, 0, st>>>(wf.counters, wf.miss_marks, 0, true);
				ctx->stats.kernel_launches += 1;
				n_ranges = 1;
			} else if (n_bounces > MISS_RANGES - 1) {  // the bounces without a range of their own share the last one
				k_miss_mark<<<1, 1, 0, st>>>(wf.counters, wf.miss_marks, MISS_RANGES - 1, false);
				ctx->stats.kernel_launches += 1;
			}
			k_miss<<<grid_wide, 128, 0, st>>>(rp, ctx->scene, wf.miss_marks, 0, n_ranges - 1, miss, wf.acc);
			ctx->stats.kernel_launches += 1;
			if (side_launched) {
				cudaEventRecord(ctx->ev_miss_done, ctx->miss_stream);
				cudaStreamWaitEvent(st, ctx->ev_miss_done, 0);
				side_launched = false;
			}
		}
		k_film<<<grid_256, 256, 0, st>>>(rp, nb, film_mode, wf.acc, ctx->film, wf.stats);
		ctx->stats.kernel_launches += 1;
		if (prof) {
			cudaEventRecord(ctx->ev[2], st);
			cudaEventSynchronize(ctx->ev[2]);
			cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
			ctx->stats.ms_film += ms;
		}
		done += nb;
	}
	cudaEventRecord(ctx->ev[5], st);
	LMB_CUDA(ctx, cudaStreamSynchronize(st));
	LMB_CUDA(ctx, cudaGetLastError());
	cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[5]);
	ctx->stats.ms_render += ms;
	ctx->stats.frames += n_frames;
	return 0;
}

// count_rays = 0: the caller counts its own rays (BDPT's ray slots hold dead entries, which must not be counted)
static int launch_trace_array(lmb_ctx* ctx, const float4* d_rays, uint32_t n, float4* d_hits, uint8_t* d_occ, bool any, int count_rays = 1, const uint32_t* order = nullptr,
							  const uint32_t* n_dev = nullptr) {
	uint32_t* cursor = ctx->wf.trace_cursor;
	if (!cursor) {
		LMB_CUDA(ctx, cudaMalloc((void**)&ctx->wf.trace_cursor, 4));
		cursor = ctx->wf.trace_cursor;
	}
	LMB_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4, ctx->stream));
	const ArraySource src{d_rays, d_hits, d_occ, any, order};
	if (ctx->use_bvh2)
		k_trace_array_bvh2<<<ctx->sm_count * 7, LMB_TRACE_THREADS, 0, ctx->stream>>>(view_of(ctx), src, n, cursor, ctx->wf.stats, count_rays);
	else if (trace_pinned(ctx))
		k_trace_array<true><<<ctx->sm_count * LMB_WIDE_BLOCKS_PER_SM, LMB_TRACE_THREADS, 0, ctx->stream>>>(wide_view_of(ctx), src, n, n_dev, cursor, ctx->wf.stats, count_rays);
	else
		k_trace_array<false><<<ctx->sm_count * wide_blocks_per_sm(false), LMB_TRACE_THREADS, 0, ctx->stream>>>(wide_view_of(ctx), src, n, n_dev, cursor, ctx->wf.stats, count_rays);
	return check_cuda(ctx, cudaGetLastError(), "k_trace_array");
}
// Probe rays for the choice of the traversal tree (lbvh.cu): they leave a random point of a random triangle in a uniformly random
// direction, like the bounce rays of a path do. Pure integer hashing (pcg4d) + detmath: the same rays for both candidate trees.
__global__ void __launch_bounds__(256) k_probe_rays(uint32_t n_tris, const float4* __restrict__ tris, uint32_t n_rays, float4* __restrict__ rays) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_rays) return;
	Rng s{i, 0x9E3779B9u, 0x85EBCA6Bu, 0u};
	const V4 r = rand4(s);
	const float pick = rand1(s);
	const uint32_t t = min((uint32_t)(pick * (float)n_tris), n_tris - 1u);
	const float4 a = tris[3 * (size_t)t], b = tris[3 * (size_t)t + 1], c = tris[3 * (size_t)t + 2];
	float u = r.x, v = r.y;
	if (u + v > 1.0f) u = 1.0f - u, v = 1.0f - v;
	const V3 p = v3(a.x, a.y, a.z) + u * (v3(b.x, b.y, b.z) - v3(a.x, a.y, a.z)) + v * (v3(c.x, c.y, c.z) - v3(a.x, a.y, a.z));
	const float z = 1.0f - 2.0f * r.z, rad = sqrtf(fmaxf(0.0f, 1.0f - z * z));
	float sn, cs;
	lmb_sincosf(LMB_TWO_PI * r.w, &sn, &cs);
	rays[2 * (size_t)i] = make_float4(p.x, p.y, p.z, 1e-3f);
	rays[2 * (size_t)i + 1] = make_float4(rad * cs, rad * sn, z, 1e30f);
}

// Mean traversal steps (wide nodes visited + triangles tested) of n_rays probe rays through the CURRENT wide tree.
int probe_wide_tree(lmb_ctx* ctx, uint32_t n_rays, double* steps_per_ray) {
	*steps_per_ray = 0.0;
	if (ctx->bvh.n == 0 || n_rays == 0) return 0;
	float4 *rays = nullptr, *hits = nullptr;
	unsigned long long* st = nullptr;
	uint32_t* cursor = nullptr;
	LMB_CUDA(ctx, cudaMalloc((void**)&rays, (size_t)n_rays * 32));
	LMB_CUDA(ctx, cudaMalloc((void**)&hits, (size_t)n_rays * 16));
	LMB_CUDA(ctx, cudaMalloc((void**)&st, ST_COUNT * 8));
	LMB_CUDA(ctx, cudaMalloc((void**)&cursor, 4));
	LMB_CUDA(ctx, cudaMemsetAsync(st, 0, ST_COUNT * 8, ctx->stream));
	LMB_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4, ctx->stream));
	k_probe_rays<<<(n_rays + 255) / 256, 256, 0, ctx->stream>>>(ctx->bvh.n, ctx->bvh.tris, n_rays, rays);
	const ArraySource src{rays, hits, nullptr, false};
	if (trace_pinned(ctx))
		k_trace_array<true><<<ctx->sm_count * LMB_WIDE_BLOCKS_PER_SM, LMB_TRACE_THREADS, 0, ctx->stream>>>(wide_view_of(ctx), src, n_rays, nullptr, cursor, st, 1);
	else
		k_trace_array<false><<<ctx->sm_count * wide_blocks_per_sm(false), LMB_TRACE_THREADS, 0, ctx->stream>>>(wide_view_of(ctx), src, n_rays, nullptr, cursor, st, 1);
	unsigned long long h[ST_COUNT];
	LMB_CUDA(ctx, cudaMemcpyAsync(h, st, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	const int rc = check_cuda(ctx, cudaGetLastError(), "probe_wide_tree");
	cudaFree(rays), cudaFree(hits), cudaFree(st), cudaFree(cursor);
	*steps_per_ray = (double)(h[ST_NODES] + h[ST_TRIS]) / n_rays;
	return rc;
}

int launch_trace_closest(lmb_ctx* ctx, const float4* d_rays, uint32_t n, float4* d_hits, const uint32_t* order) {
	return launch_trace_array(ctx, d_rays, n, d_hits, nullptr, false, 1, order);
}
int launch_trace_any(lmb_ctx* ctx, const float4* d_rays, uint32_t n, uint8_t* d_occ) { return launch_trace_array(ctx, d_rays, n, nullptr, d_occ, true); }
int launch_trace_slots(lmb_ctx* ctx, const float4* d_rays, uint32_t n, float4* d_hits, uint8_t* d_occ, bool any) { return launch_trace_array(ctx, d_rays, n, d_hits, d_occ, any, 0); }
// the slots list[0 .. *n_dev) only (at most n_max of them); results land at the slots' own positions
int launch_trace_slot_list(lmb_ctx* ctx, const float4* d_rays, const uint32_t* list, const uint32_t* n_dev, uint32_t n_max, float4* d_hits, uint8_t* d_occ, bool any) {
	return launch_trace_array(ctx, d_rays, n_max, d_hits, d_occ, any, 0, list, n_dev);
}
int launch_resolve_on(lmb_ctx* ctx, float4* film, cudaStream_t stream) {
	k_resolve<<<ctx->sm_count * 8, 256, 0, stream>>>(ctx->width * ctx->height, film);
	return check_cuda(ctx, cudaGetLastError(), "k_resolve");
}
int launch_resolve(lmb_ctx* ctx) { return launch_resolve_on(ctx, ctx->film, ctx->stream); }

}  // namespace lmb
