// wavefront.cu -- the path integrator as wavefront kernels (replaces the single vkCmdTraceRaysKHR(W, H, 1) dispatch of
// src/RayTracer/Path.cpp:39-58 / path.rgen main()).
//
// One path slot per (pixel, frame-in-batch). Per bounce d, in stream order:
//   k_trace    ONE persistent traversal kernel for every ray in flight: continuation rays into bounce d (path.rgen:48),
//              and the shadow any-hit + MIS-probe closest-hit rays bounce d-1 generated (pt_commons.glsl:21-22, 32)
//   k_connect  MIS weights and radiance accumulation of bounce d-1's light samples (pt_commons.glsl:23-40, path.rgen:78)
//   k_surface  hit record, material (+texture), emission, depth cut, normal orientation; sorts paths by BSDF type
//              (path.rgen:49-74, ray.rchit, bsdf_commons.glsl:16-22)
//   k_nee<T>   per BSDF type: light sample, BSDF eval, MIS probe sample; writes the NEE record and its two rays
//              (path.rgen:75-80, pt_commons.glsl:3-20,28-30)
//   k_bsdf<T>  per BSDF type: continuation sample, throughput, Russian roulette; writes the next ray (path.rgen:81-100)
// then k_miss evaluates the sky for escaped rays (commons.glsl:156-168) and k_film applies the running-mean / sum film
// update in frame order (path.rgen:102-112).
// RNG state is a pure function of (x, y, frame, draw counter), so reordering paths across kernels cannot change any
// sample; every float expression keeps the order of the GLSL (see vec.cuh).
//
// HBM bytes per path-bounce (algorithmic): ray 32 rd + hit 16 wr (extend); ray 32 + hit 16 + state 32 rd, state 32 +
// ray 32 + NEE 128 wr, ~200 B scene gathers (shade); NEE 128 + col 16 rd, col 16 wr (connect). Film: 16 B per sample.
#include <cooperative_groups.h>
#include <stdio.h>

#include "context.h"
#include "scene_device.cuh"
#include "trace_persistent.cuh"
#include "trace_wide.cuh"

namespace cg = cooperative_groups;

namespace lmb {

namespace {

#ifndef LMB_SHADE_MIN_BLOCKS
#define LMB_SHADE_MIN_BLOCKS 4
#endif
constexpr float T_MIN = 0.001f;    // path.rgen:19
constexpr float T_MAX = 10000.0f;  // path.rgen:20

struct RenderParams {
	M4 inv_view, inv_proj;
	V3 sky_col;
	uint32_t width, height, n_pix;
	uint32_t first_frame, frame_stride;  // frame of batch slot fb = first_frame + fb * frame_stride
	uint32_t n_active;                    // live slots this batch = n_batch_frames * n_pix
	int32_t num_lights, max_depth, light_triangle_count;
	uint32_t dir_light_idx, direct_lighting;
};

// Work counters, double-buffered by bounce parity: launch d reads [d & 1] while k_shade(d) fills [(d + 1) & 1], which
// k_trace(d) zeroed when it started (together with the per-material counters). A shading warp appends to three lists per
// trip -- the typed ray queue, the light-sample (NEE) list and the next path list -- with two atomics issued back to back:
// one 32-bit add on the ray-queue size and one 64-bit add on the packed (path count << 32 | NEE count) pair.
enum Counter { CNT_TRACE = 0, CNT_CURSOR = 2, CNT_MISS = 4, CNT_PAIR = 6 /* 6,7 | 8,9: lo = NEE count, hi = path count */, CNT_MAT = 10, CNT_COUNT = 20 };
__device__ __forceinline__ uint32_t nee_count(const uint32_t* counters, int parity) { return counters[CNT_PAIR + 2 * parity]; }
__device__ __forceinline__ uint32_t path_count(const uint32_t* counters, int parity) { return counters[CNT_PAIR + 2 * parity + 1]; }

// material-sorted shade queues: index = log2(bsdf_type) for the six known types, 6 = unknown type (bsdf_type 0, quirk Q8)
constexpr int N_MAT_QUEUES = 7;

// trace queue entry = slot | type << 30
constexpr uint32_t RAY_CONTINUE = 0u, RAY_SHADOW = 1u, RAY_PROBE = 2u, SLOT_MASK = 0x3FFFFFFFu;

// NEE record: 8 float4 planes of n_slots each
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
	counters[CNT_TRACE] = n_active, counters[CNT_TRACE + 1] = 0;
	counters[CNT_CURSOR] = 0, counters[CNT_CURSOR + 1] = 0;
	counters[CNT_MISS] = 0;
	counters[CNT_PAIR] = 0, counters[CNT_PAIR + 1] = n_active;
	counters[CNT_PAIR + 2] = 0, counters[CNT_PAIR + 3] = 0;
}
// first thing k_trace(d) does: the lists k_classify(d) / k_shade(d) are about to fill start empty
__device__ __forceinline__ void reset_next_counters(uint32_t* counters, int parity) {
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		counters[CNT_TRACE + (parity ^ 1)] = 0;
		counters[CNT_CURSOR + (parity ^ 1)] = 0;
		counters[CNT_PAIR + 2 * (parity ^ 1)] = 0, counters[CNT_PAIR + 2 * (parity ^ 1) + 1] = 0;
		for (int m = 0; m < N_MAT_QUEUES; m++) counters[CNT_MAT + m] = 0;
	}
}

// path.rgen:23-45 + sample_camera (commons.glsl:30-33)
__global__ void __launch_bounds__(256) k_raygen(RenderParams rp, float4* __restrict__ ray_o, float4* __restrict__ ray_d, float4* __restrict__ thr,
												 float4* __restrict__ col, uint32_t* __restrict__ path_queue, uint32_t* __restrict__ trace_queue, unsigned long long* stats) {
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[ST_CLOSEST], (unsigned long long)rp.n_active);
	for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < rp.n_active; slot += gridDim.x * blockDim.x) {
		const uint32_t fb = slot / rp.n_pix, pix = slot - fb * rp.n_pix;
		const uint32_t px = pix % rp.width, py = pix / rp.width;
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
		ray_o[slot] = f4(origin, T_MIN);
		ray_d[slot] = f4(direction, T_MAX);
		thr[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(seed.w));
		col[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u));
		path_queue[slot] = slot;
		trace_queue[slot] = slot | (RAY_CONTINUE << 30);
	}
}

// Ray source of the wavefront: typed queue entries over the path-state and NEE planes.
struct WavefrontSource {
	const uint32_t* __restrict__ queue;
	const float4* __restrict__ ray_o;
	const float4* __restrict__ ray_d;
	const float4* __restrict__ nee;
	float4* __restrict__ hit;
	float4* __restrict__ probe_hit;
	uint32_t* __restrict__ shadow_occ;
	uint32_t n_slots;
	__device__ __forceinline__ void load(uint32_t i, V3& o, V3& d, float& tmin, float& tmax, bool& any) const {
		const uint32_t e = queue[i], type = e >> 30, slot = e & SLOT_MASK;
		if (type == RAY_CONTINUE) {
			const float4 o4 = ray_o[slot], d4 = ray_d[slot];
			o = xyz(o4), d = xyz(d4), tmin = o4.w, tmax = d4.w, any = false;
		} else if (type == RAY_SHADOW) {  // pt_commons.glsl:21-22: tmin 0, tmax = wi_len - EPS, terminate on first hit
			const float4 p4 = nee[NEE_P * (size_t)n_slots + slot];
			o = xyz(p4), d = xyz(nee[NEE_WI * (size_t)n_slots + slot]), tmin = 0.0f, tmax = p4.w, any = true;
		} else {  // pt_commons.glsl:32
			o = xyz(nee[NEE_P * (size_t)n_slots + slot]), d = xyz(nee[NEE_PROBE_WI * (size_t)n_slots + slot]), tmin = T_MIN, tmax = T_MAX, any = false;
		}
	}
	__device__ __forceinline__ void store(uint32_t i, const Hit& h, bool) const {
		const uint32_t e = queue[i], type = e >> 30, slot = e & SLOT_MASK;
		if (type == RAY_SHADOW)
			shadow_occ[slot] = h.prim != 0xFFFFFFFFu;
		else
			(type == RAY_CONTINUE ? hit : probe_hit)[slot] = make_float4(h.t, h.b1, h.b2, __uint_as_float(h.prim));
	}
};

__global__ void __launch_bounds__(LMB_TRACE_THREADS, LMB_WIDE_BLOCKS_PER_SM) k_trace(WideBvhView bvh, WavefrontSource src, uint32_t* counters, int parity, unsigned long long* stats) {
	reset_next_counters(counters, parity);
	trace_wide_persistent(bvh, src, counters[CNT_TRACE + parity], &counters[CNT_CURSOR + parity], stats, -1, -1);
}
// the same rays over the binary LBVH (LMB_TRAVERSAL=bvh2; A/B measurements and the canonical node counts)
__global__ void __launch_bounds__(LMB_TRACE_THREADS) k_trace_bvh2(BvhView bvh, WavefrontSource src, uint32_t* counters, int parity, unsigned long long* stats) {
	reset_next_counters(counters, parity);
	trace_persistent(bvh, src, counters[CNT_TRACE + parity], &counters[CNT_CURSOR + parity], stats, -1, -1);
}

// k_classify: sorts the live paths by the BSDF type of the surface they hit (one queue per type), so that k_shade<TYPE>
// runs a single lobe's code on full warps; retires escaped rays (path.rgen:49-55). Reads 4 B queue + 16 B hit + one byte of
// the L2-resident per-triangle queue table per path.
__global__ void __launch_bounds__(256) k_classify(RenderParams rp, DeviceScene sc, int depth, uint32_t* __restrict__ counters, int parity,
													   const uint32_t* __restrict__ queue, uint32_t* __restrict__ mat_queues, uint32_t* __restrict__ miss_queue,
													   const float4* __restrict__ hit, const float4* __restrict__ thr, float4* __restrict__ colb, uint32_t n_slots) {
	const uint32_t count = path_count(counters, parity);
	const int lane = threadIdx.x & 31;
	constexpr int U = 4;  // entries per lane and trip: four independent queue -> hit -> table chains in flight
	const uint32_t stride = gridDim.x * blockDim.x * U;
	for (uint32_t base = (blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) * U; base < count; base += stride) {  // warp-uniform trip count
		uint32_t slot[U], prim[U];
		int dest[U];  // 0..6 material queue, 7 miss queue, -1 retired
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t i = base + u * 32 + lane;
			slot[u] = i < count ? queue[i] : 0xFFFFFFFFu;
		}
#pragma unroll
		for (int u = 0; u < U; u++) prim[u] = slot[u] != 0xFFFFFFFFu ? __float_as_uint(hit[slot[u]].w) : 0xFFFFFFFFu;
#pragma unroll
		for (int u = 0; u < U; u++) {
			dest[u] = -1;
			if (slot[u] == 0xFFFFFFFFu) continue;
			if (prim[u] == 0xFFFFFFFFu) {  // path.rgen:49-55
				if (depth > 0 || rp.direct_lighting == 1) {
					if (rp.dir_light_idx == 0xFFFFFFFFu) {
						const float4 c4 = colb[slot[u]];
						const V3 col = xyz(c4) + xyz(thr[slot[u]]) * rp.sky_col;  // shade_atmosphere's constant-sky branch (commons.glsl:157-159)
						colb[slot[u]] = f4(col, c4.w);
					} else {
						dest[u] = 7;  // 64 x 8 step sky march: deferred to k_miss (ray_o / ray_d / thr / col of a dead path stay put)
					}
				}
			} else {
				dest[u] = sc.tri_matq[prim[u]];
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dest[u]);
			if (dest[u] >= 0) {
				const int leader = __ffs(peers) - 1;
				uint32_t at = 0;
				if (lane == leader) at = atomicAdd(dest[u] == 7 ? &counters[CNT_MISS] : &counters[CNT_MAT + dest[u]], (uint32_t)__popc(peers));
				at = __shfl_sync(peers, at, leader) + __popc(peers & ((1u << lane) - 1u));
				if (dest[u] == 7)
					miss_queue[at] = slot[u];
				else
					mat_queues[(size_t)dest[u] * n_slots + at] = slot[u];
			}
		}
	}
}

// k_shade<TYPE>: one whole bounce for the paths that hit a surface of BSDF type TYPE (path.rgen:56-100): hit record
// (ray.rchit), material (+texture), emission, depth cut, normal orientation; NEE up to its two traceRayEXT calls
// (pt_commons.glsl:3-20, 28-30); continuation sample, throughput update and Russian roulette. The path state stays in
// registers across the three stages: per path-bounce the kernel reads hit + ray direction + throughput + radiance (64 B,
// plus L2-resident scene data) and writes radiance, throughput, the next ray, the NEE record and three queue entries.
// LAST = true is the bounce at depth max_depth - 1, which only collects emission (path.rgen:57-62) for any BSDF type.
template <uint32_t TYPE, bool LAST>
__global__ void __launch_bounds__(128, LAST ? 8 : LMB_SHADE_MIN_BLOCKS) k_shade(RenderParams rp, DeviceScene sc, int depth, uint32_t* __restrict__ counters, int parity, int count_idx,
													const uint32_t* __restrict__ queue, uint32_t* __restrict__ path_queue, uint32_t* __restrict__ nee_queue,
													uint32_t* __restrict__ trace_queue, const float4* __restrict__ hit, float4* __restrict__ ray_o,
													float4* __restrict__ ray_d, float4* __restrict__ thr, float4* __restrict__ colb, float4* __restrict__ nee,
													uint32_t n_slots, unsigned long long* stats) {
	const uint32_t count = LAST ? path_count(counters, parity) : counters[count_idx];
	const int lane = threadIdx.x & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint32_t n_shadow = 0, n_probe = 0, n_cont = 0;
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < count; base += stride) {  // warp-uniform trip count
		const uint32_t i = base + lane;
		uint32_t slot = 0;
		bool do_shadow = false, do_probe = false, alive = false;
		do {
		if (i >= count) break;
		slot = queue[i];
		const float4 h4 = hit[slot];
		const float4 c4 = colb[slot];
		const float4 t4 = thr[slot];
		const float4 d4 = ray_d[slot];
		const uint32_t prim = __float_as_uint(h4.w);
		if (LAST && prim == 0xFFFFFFFFu) break;  // (escaped rays of the last bounce were retired by k_classify)
		V3 throughput = xyz(t4);
		V3 col = xyz(c4);
		const bool last_specular_in = (__float_as_uint(c4.w) & 1u) != 0;
		const HitPayload payload = build_hit(sc, prim, h4.y, h4.z);
		const lmb_material hit_mat = load_material(sc, payload.material_idx, payload.uv);
		if ((depth == 0 && rp.direct_lighting == 1) || last_specular_in) col += throughput * v3(hit_mat.emissive_factor);
		if (LAST || depth >= rp.max_depth - 1) {
			colb[slot] = f4(col, c4.w);
			break;
		}
		const V3 wo = -xyz(d4);
		V3 n_s = payload.n_s;
		bool side = true;
		V3 n_g = payload.n_g;
		if (dot(payload.n_g, wo) < 0.0f) n_g = -n_g;
		if (dot(n_g, payload.n_s) < 0) {
			n_s = -n_s;
			side = false;
		}
		const V3 origin = offset_ray(payload.pos, n_g);
		const bool last_specular = (hit_mat.bsdf_props & LMB_FLAG_SPECULAR) != 0;
		colb[slot] = f4(col, __uint_as_float(last_specular ? 1u : 0u));
		const V3 pos = payload.pos;
		const uint32_t fb = slot / rp.n_pix, pix = slot - fb * rp.n_pix;
		Rng seed{pix % rp.width, pix / rp.width, rp.first_frame + fb * rp.frame_stride, __float_as_uint(t4.w)};
		// ---- next-event estimation (path.rgen:75-80, pt_commons.glsl:3-20, 28-30)
		if (!last_specular && (depth > 0 || rp.direct_lighting == 1)) {
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
			do_shadow = true;
			n_shadow++;
			nee[NEE_P * (size_t)n_slots + slot] = f4(p, ls.wi_len - LMB_EPS);
			nee[NEE_LDIR * (size_t)n_slots + slot] = f4(ldir, ls.pdf_a);
			nee[NEE_T * (size_t)n_slots + slot] = f4(throughput, __uint_as_float(ls.instance_idx));
			if ((ls.flags & 0x7u) == LMB_LIGHT_AREA) {
				const V3 r3 = rand3(seed);
				const BsdfSample bs = sample_bsdf_t<TYPE>(n_s, wo, hit_mat, 1, side, r3);
				if (bs.pdf != 0) {
					flags |= NEE_FLAG_PROBE;
					// ray.rmiss writes only material_idx: when the probe misses, the reference still compares the ids of the
					// payload left by the previous trace, i.e. of the surface being shaded, and then uses its pos / n_s
					// (wi_len = |pos - pos| = 0). That can only match when this surface is the sampled light triangle.
					float g_stale = 0.0f;
					if (payload.instance_idx == ls.instance_idx && sc.tri_local[prim] == ls.triangle_idx) {
						flags |= NEE_FLAG_STALE_MATCH;
						const float wi_len = length(pos - pos);
						g_stale = fabsf(dot(payload.n_s, -bs.wi)) / (wi_len * wi_len);  // ray.rchit's un-flipped shading normal
					}
					nee[NEE_PROBE_WI * (size_t)n_slots + slot] = f4(bs.wi, bs.pdf);
					nee[NEE_F2 * (size_t)n_slots + slot] = f4(bs.f, fabsf(bs.cos_theta));
					nee[NEE_LE * (size_t)n_slots + slot] = f4(ls.Le, __uint_as_float(ls.triangle_idx));
					nee[NEE_POS * (size_t)n_slots + slot] = f4(pos, g_stale);
					do_probe = true;
					n_probe++;
				}
			}
			nee[NEE_WI * (size_t)n_slots + slot] = f4(ls.wi, __uint_as_float(flags));
		}
		// ---- continuation (path.rgen:81-100)
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
		if (alive) {
			thr[slot] = f4(throughput, __uint_as_float(seed.w));
			ray_o[slot] = f4(origin, T_MIN);
			ray_d[slot] = f4(bs.wi, T_MAX);
			n_cont++;
		}
		} while (0);
		if (!LAST) {
			// ---- the rays this warp generated go to the typed queue of the next launch: one atomic per warp and trip
			const uint32_t b_sh = __ballot_sync(0xFFFFFFFFu, do_shadow), b_pr = __ballot_sync(0xFFFFFFFFu, do_probe);
			const uint32_t b_ct = __ballot_sync(0xFFFFFFFFu, alive);
			const uint32_t n_sh = __popc(b_sh), n_pr = __popc(b_pr), n_ct = __popc(b_ct), total = n_sh + n_pr + n_ct;
			if (total) {
				uint32_t at = 0;
				unsigned long long at2 = 0;
				if (lane == 0) {
					at = atomicAdd(&counters[CNT_TRACE + (parity ^ 1)], total);
					at2 = atomicAdd(reinterpret_cast<unsigned long long*>(&counters[CNT_PAIR + 2 * (parity ^ 1)]), ((unsigned long long)n_ct << 32) | n_sh);
				}
				at = __shfl_sync(0xFFFFFFFFu, at, 0);
				const uint32_t at_nee = __shfl_sync(0xFFFFFFFFu, (uint32_t)at2, 0), at_path = __shfl_sync(0xFFFFFFFFu, (uint32_t)(at2 >> 32), 0);
				if (do_shadow) {
					const uint32_t r = __popc(b_sh & lt_mask);
					trace_queue[at + r] = slot | (RAY_SHADOW << 30);
					nee_queue[at_nee + r] = slot;
				}
				if (do_probe) trace_queue[at + n_sh + __popc(b_pr & lt_mask)] = slot | (RAY_PROBE << 30);
				if (alive) {
					const uint32_t r = __popc(b_ct & lt_mask);
					trace_queue[at + n_sh + n_pr + r] = slot | (RAY_CONTINUE << 30);
					path_queue[at_path + r] = slot;
				}
			}
		}
	}
	flush_stats(stats, ST_SHADOW, n_shadow);
	flush_stats(stats, ST_PROBE, n_probe);
	flush_stats(stats, ST_CLOSEST, n_cont);
}

// pt_commons.glsl:23-27, 33-39 and the accumulation of path.rgen:78, once both rays of the light sample are traced
__global__ void __launch_bounds__(128) k_connect(RenderParams rp, DeviceScene sc, const uint32_t* __restrict__ counters, int parity,
												  const uint32_t* __restrict__ nee_queue, const float4* __restrict__ nee, const float4* __restrict__ probe_hit,
												  const uint32_t* __restrict__ shadow_occ, float4* __restrict__ colb, uint32_t n_slots) {
	const uint32_t count = nee_count(counters, parity);
	const float light_pick_pdf = 1.0f / (float)rp.light_triangle_count;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		const uint32_t slot = nee_queue[i];
		const float4 wi4 = nee[NEE_WI * (size_t)n_slots + slot];
		const float4 l4 = nee[NEE_LDIR * (size_t)n_slots + slot];
		const float4 t4 = nee[NEE_T * (size_t)n_slots + slot];
		const uint32_t flags = __float_as_uint(wi4.w);
		V3 res = v3(0.0f);
		if (shadow_occ[slot] == 0u && (flags & NEE_FLAG_SHADOW_CONTRIB)) res += xyz(l4);
		if (flags & NEE_FLAG_PROBE) {
			const float4 pw4 = nee[NEE_PROBE_WI * (size_t)n_slots + slot];
			const float4 f4v = nee[NEE_F2 * (size_t)n_slots + slot];
			const float4 le4 = nee[NEE_LE * (size_t)n_slots + slot];
			const float4 pos4 = nee[NEE_POS * (size_t)n_slots + slot];
			const V3 wi = xyz(pw4);
			const float bsdf_pdf = pw4.w;
			const float4 ph = probe_hit[slot];
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
		const float4 c4 = colb[slot];
		const V3 col = xyz(c4) + xyz(t4) * res / light_pick_pdf;
		colb[slot] = f4(col, c4.w);
	}
}

// Escaped rays with a sun + sky light: col += throughput * shade_atmosphere(...) (path.rgen:50-53, commons.glsl:156-168).
// A path misses at most once and nothing is added to its radiance afterwards, so running this after the bounce loop keeps
// the order of the float additions.
__global__ void __launch_bounds__(128) k_miss(RenderParams rp, DeviceScene sc, const uint32_t* __restrict__ counters, const uint32_t* __restrict__ miss_queue,
											   const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, const float4* __restrict__ thr,
											   float4* __restrict__ colb) {
	const uint32_t count = counters[CNT_MISS];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
		const uint32_t slot = miss_queue[i];
		const float4 c4 = colb[slot];
		const V3 sky = shade_atmosphere(sc, rp.dir_light_idx, rp.sky_col, xyz(ray_o[slot]), xyz(ray_d[slot]), T_MAX);
		const V3 col = xyz(c4) + xyz(thr[slot]) * sky;
		colb[slot] = f4(col, c4.w);
	}
}

// path.rgen:102-112, applied for the batch's frames in order
__global__ void __launch_bounds__(256) k_film(RenderParams rp, uint32_t n_batch_frames, int film_mode, const float4* __restrict__ colb,
											   float4* __restrict__ film, unsigned long long* stats) {
	uint32_t nan_count = 0;
	for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < rp.n_pix; pix += gridDim.x * blockDim.x) {
		float4 acc = film[pix];
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
		film[pix] = acc;
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
	__device__ __forceinline__ void load(uint32_t i, V3& o, V3& d, float& tmin, float& tmax, bool& any) const {
		const float4 o4 = rays[2 * (size_t)i], d4 = rays[2 * (size_t)i + 1];
		o = xyz(o4), d = xyz(d4), tmin = o4.w, tmax = d4.w, any = any_hit;
	}
	__device__ __forceinline__ void store(uint32_t i, const Hit& h, bool) const {
		if (any_hit)
			occ[i] = h.prim != 0xFFFFFFFFu;
		else
			hits[i] = make_float4(h.t, h.b1, h.b2, __uint_as_float(h.prim));
	}
};

__global__ void __launch_bounds__(LMB_TRACE_THREADS, LMB_WIDE_BLOCKS_PER_SM) k_trace_array(WideBvhView bvh, ArraySource src, uint32_t n, uint32_t* cursor, unsigned long long* stats) {
	trace_wide_persistent(bvh, src, n, cursor, stats, ST_CLOSEST, ST_SHADOW);
}
__global__ void __launch_bounds__(LMB_TRACE_THREADS) k_trace_array_bvh2(BvhView bvh, ArraySource src, uint32_t n, uint32_t* cursor, unsigned long long* stats) {
	trace_persistent(bvh, src, n, cursor, stats, ST_CLOSEST, ST_SHADOW);
}

BvhView view_of(const lmb_ctx* ctx) { return BvhView{ctx->bvh.nodes, ctx->bvh.tris, ctx->bvh.n}; }
WideBvhView wide_view_of(const lmb_ctx* ctx) { return WideBvhView{ctx->wide.nodes, ctx->wide.tris, ctx->wide.n_tris, 0x3F800000u}; }

}  // namespace

int wavefront_alloc(lmb_ctx* ctx, uint32_t frames_in_flight) {
	wavefront_free(ctx);
	Wavefront& wf = ctx->wf;
	const uint64_t n_pix = (uint64_t)ctx->width * ctx->height;
	if (frames_in_flight == 0) {
		// default: about 8 M path slots in flight (enough to fill 148 SMs many times over), at most 64 frames
		frames_in_flight = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, (8ull << 20) / std::max<uint64_t>(n_pix, 1)));
	}
	const uint64_t n_slots = n_pix * frames_in_flight;
	if (n_slots == 0 || n_slots > 0x3FFFFFFFull) return set_error(ctx, LMB_ERR_INVALID, "width*height*frames_in_flight out of range");
	wf.n_slots = (uint32_t)n_slots;
	wf.frames_in_flight = frames_in_flight;
	auto alloc = [&](void** p, size_t bytes) { return check_cuda(ctx, cudaMalloc(p, bytes), "cudaMalloc(wavefront)"); };
	int rc;
	if ((rc = alloc((void**)&wf.ray_o, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.ray_d, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.hit, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.thr, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.col, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.nee, n_slots * 16 * NEE_PLANES))) return rc;
	if ((rc = alloc((void**)&wf.path_queue, n_slots * 4))) return rc;
	if ((rc = alloc((void**)&wf.nee_queue, n_slots * 4))) return rc;
	if ((rc = alloc((void**)&wf.miss_queue, n_slots * 4))) return rc;
	if ((rc = alloc((void**)&wf.mat_queues, n_slots * 4 * N_MAT_QUEUES))) return rc;
	if ((rc = alloc((void**)&wf.trace_queue, n_slots * 4 * 3))) return rc;
	if ((rc = alloc((void**)&wf.probe_hit, n_slots * 16))) return rc;
	if ((rc = alloc((void**)&wf.shadow_occ, n_slots * 4))) return rc;
	if ((rc = alloc((void**)&wf.counters, CNT_COUNT * 4))) return rc;
	if ((rc = alloc((void**)&wf.stats, ST_COUNT * 8))) return rc;
	LMB_CUDA(ctx, cudaMemsetAsync(wf.stats, 0, ST_COUNT * 8, ctx->stream));
	return 0;
}

void wavefront_free(lmb_ctx* ctx) {
	Wavefront& wf = ctx->wf;
	cudaFree(wf.ray_o), cudaFree(wf.ray_d), cudaFree(wf.hit), cudaFree(wf.thr), cudaFree(wf.col), cudaFree(wf.nee);
	cudaFree(wf.path_queue), cudaFree(wf.nee_queue), cudaFree(wf.miss_queue), cudaFree(wf.mat_queues), cudaFree(wf.trace_queue), cudaFree(wf.probe_hit), cudaFree(wf.shadow_occ), cudaFree(wf.trace_cursor), cudaFree(wf.counters), cudaFree(wf.stats);
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
	rp.width = ctx->width, rp.height = ctx->height, rp.n_pix = ctx->width * ctx->height;
	rp.frame_stride = stride;
	rp.num_lights = pc.num_lights, rp.max_depth = pc.max_depth, rp.light_triangle_count = pc.light_triangle_count;
	rp.dir_light_idx = pc.dir_light_idx, rp.direct_lighting = pc.direct_lighting;
	const BvhView bvh = view_of(ctx);
	const WideBvhView wide = wide_view_of(ctx);
	const int grid_wide = ctx->sm_count * 16;
	const int grid_256 = ctx->sm_count * 8;
	const int grid_trace = ctx->sm_count * 7;  // persistent: 7 blocks x 32 KB stack fit one SM's shared memory
	const int grid_trace_wide = ctx->sm_count * LMB_WIDE_BLOCKS_PER_SM;
	float ms;
	const bool prof = ctx->profile_stages;  // per-stage timing serialises the bounce loop; off by default
	cudaEventRecord(ctx->ev[0], st);
	for (uint32_t done = 0; done < n_frames;) {
		const uint32_t nb = std::min(wf.frames_in_flight, n_frames - done);
		rp.first_frame = first_frame + done * stride;
		rp.n_active = nb * rp.n_pix;
		if (prof) cudaEventRecord(ctx->ev[1], st);
		k_begin_batch<<<1, 1, 0, st>>>(wf.counters, rp.n_active);
		k_raygen<<<grid_256, 256, 0, st>>>(rp, wf.ray_o, wf.ray_d, wf.thr, wf.col, wf.path_queue, wf.trace_queue, wf.stats);
		ctx->stats.kernel_launches += 2;
		if (prof) {
			cudaEventRecord(ctx->ev[2], st);
			cudaEventSynchronize(ctx->ev[2]);
			cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
			ctx->stats.ms_film += ms;
		}
		const WavefrontSource src{wf.trace_queue, wf.ray_o, wf.ray_d, wf.nee, wf.hit, wf.probe_hit, wf.shadow_occ, wf.n_slots};
		for (int depth = 0; depth < std::max(pc.max_depth, 1); depth++) {
			const int par = depth & 1;
			if (prof) cudaEventRecord(ctx->ev[1], st);
			if (ctx->use_bvh2)
				k_trace_bvh2<<<grid_trace, LMB_TRACE_THREADS, 0, st>>>(bvh, src, wf.counters, par, wf.stats);
			else
				k_trace<<<grid_trace_wide, LMB_TRACE_THREADS, 0, st>>>(wide, src, wf.counters, par, wf.stats);
			ctx->stats.kernel_launches += 1;
			if (prof) cudaEventRecord(ctx->ev[2], st);
			if (depth > 0) {
				k_connect<<<grid_wide, 128, 0, st>>>(rp, ctx->scene, wf.counters, par, wf.nee_queue, wf.nee, wf.probe_hit, wf.shadow_occ, wf.col, wf.n_slots);
				ctx->stats.kernel_launches += 1;
			}
			if (prof) cudaEventRecord(ctx->ev[3], st);
#define LMB_SHADE_ARGS(mq, cidx) rp, ctx->scene, depth, wf.counters, par, cidx, mq, wf.path_queue, wf.nee_queue, wf.trace_queue, wf.hit, wf.ray_o, wf.ray_d, wf.thr, wf.col, wf.nee, wf.n_slots, wf.stats
			const bool last = depth >= pc.max_depth - 1;  // the last bounce only collects emission (path.rgen:57-62)
			if (!last || depth > 0 || pc.direct_lighting == 1) {
				k_classify<<<grid_256, 256, 0, st>>>(rp, ctx->scene, depth, wf.counters, par, wf.path_queue, wf.mat_queues, wf.miss_queue, wf.hit, wf.thr, wf.col,
													 wf.n_slots);
				ctx->stats.kernel_launches += 1;
			}
			if (last) {
				k_shade<0u, true><<<grid_wide, 128, 0, st>>>(LMB_SHADE_ARGS(wf.path_queue, 0));
				ctx->stats.kernel_launches += 1;
			} else {
				for (int m = 0; m < N_MAT_QUEUES; m++) {
					if (!(ctx->mat_queue_mask & (1u << m))) continue;  // BSDF type absent from the scene (ENABLE_* macros, LumenScene.cpp:217-228)
					const uint32_t* mq = wf.mat_queues + (size_t)m * wf.n_slots;
					switch (m) {
						case 0: k_shade<LMB_BSDF_DIFFUSE, false><<<grid_wide, 128, 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m)); break;
						case 1: k_shade<LMB_BSDF_MIRROR, false><<<grid_wide, 128, 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m)); break;
						case 2: k_shade<LMB_BSDF_GLASS, false><<<grid_wide, 128, 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m)); break;
						case 3: k_shade<LMB_BSDF_DIELECTRIC, false><<<grid_wide, 128, 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m)); break;
						case 4: k_shade<LMB_BSDF_CONDUCTOR, false><<<grid_wide, 128, 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m)); break;
						case 5: k_shade<LMB_BSDF_PRINCIPLED, false><<<grid_wide, 128, 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m)); break;
						default: k_shade<0u, false><<<grid_wide, 128, 0, st>>>(LMB_SHADE_ARGS(mq, CNT_MAT + m)); break;
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
		if (pc.dir_light_idx != 0xFFFFFFFFu) {
			k_miss<<<grid_wide, 128, 0, st>>>(rp, ctx->scene, wf.counters, wf.miss_queue, wf.ray_o, wf.ray_d, wf.thr, wf.col);
			ctx->stats.kernel_launches += 1;
		}
		k_film<<<grid_256, 256, 0, st>>>(rp, nb, film_mode, wf.col, ctx->film, wf.stats);
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

static int launch_trace_array(lmb_ctx* ctx, const float4* d_rays, uint32_t n, float4* d_hits, uint8_t* d_occ, bool any) {
	uint32_t* cursor = ctx->wf.trace_cursor;
	if (!cursor) {
		LMB_CUDA(ctx, cudaMalloc((void**)&ctx->wf.trace_cursor, 4));
		cursor = ctx->wf.trace_cursor;
	}
	LMB_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4, ctx->stream));
	const ArraySource src{d_rays, d_hits, d_occ, any};
	if (ctx->use_bvh2)
		k_trace_array_bvh2<<<ctx->sm_count * 7, LMB_TRACE_THREADS, 0, ctx->stream>>>(view_of(ctx), src, n, cursor, ctx->wf.stats);
	else
		k_trace_array<<<ctx->sm_count * LMB_WIDE_BLOCKS_PER_SM, LMB_TRACE_THREADS, 0, ctx->stream>>>(wide_view_of(ctx), src, n, cursor, ctx->wf.stats);
	return check_cuda(ctx, cudaGetLastError(), "k_trace_array");
}
int launch_trace_closest(lmb_ctx* ctx, const float4* d_rays, uint32_t n, float4* d_hits) { return launch_trace_array(ctx, d_rays, n, d_hits, nullptr, false); }
int launch_trace_any(lmb_ctx* ctx, const float4* d_rays, uint32_t n, uint8_t* d_occ) { return launch_trace_array(ctx, d_rays, n, nullptr, d_occ, true); }
int launch_resolve(lmb_ctx* ctx) {
	k_resolve<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->width * ctx->height, ctx->film);
	return check_cuda(ctx, cudaGetLastError(), "k_resolve");
}

}  // namespace lmb
