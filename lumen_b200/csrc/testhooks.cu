// testhooks.cu -- device-side known-answer probes (include/lumen_b200_testhooks.h). Each kernel calls the very device
// function the wavefront kernels use, so a bit-exact match here pins the product code, not a copy of it.
#include "lumen_b200_testhooks.h"

#include <vector>

#include "context.h"
#include "scene_device.cuh"

using namespace lmb;

namespace {

struct Buf {
	void* d = nullptr;
	size_t bytes = 0;
	~Buf() { cudaFree(d); }
};

int to_dev(lmb_ctx* ctx, Buf& b, const void* host, size_t bytes) {
	b.bytes = bytes;
	LMB_CUDA(ctx, cudaMalloc(&b.d, std::max<size_t>(bytes, 4)));
	if (host && bytes) LMB_CUDA(ctx, cudaMemcpyAsync(b.d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return 0;
}
int to_host(lmb_ctx* ctx, void* host, const Buf& b) {
	if (b.bytes) LMB_CUDA(ctx, cudaMemcpyAsync(host, b.d, b.bytes, cudaMemcpyDeviceToHost, ctx->stream));
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return check_cuda(ctx, cudaGetLastError(), "kat kernel");
}

#define GRID(n) ((n) + 127) / 128, 128, 0, ctx->stream

__global__ void k_pcg4d(const uint32_t* in, uint32_t n, uint32_t* out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t x = in[4 * i], y = in[4 * i + 1], z = in[4 * i + 2], w = in[4 * i + 3];
	pcg4d_full(x, y, z, w);
	out[4 * i] = x, out[4 * i + 1] = y, out[4 * i + 2] = z, out[4 * i + 3] = w;
}
__global__ void k_rand(const uint32_t* in, uint32_t n, uint32_t draws, float* out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Rng s{in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]};
	for (uint32_t k = 0; k < draws; k++) out[(size_t)i * draws + k] = rand1(s);
}
__global__ void k_detmath(const float* x, const float* y, uint32_t n, float* os, float* oc, float* oe, float* op) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float s, c;
	lmb_sincosf(x[i], &s, &c);
	os[i] = s, oc[i] = c;
	oe[i] = lmb_expf(x[i]);
	op[i] = lmb_powf(fabsf(x[i]), y[i]);
}
__global__ void k_offset_ray(const float* p, const float* nn, uint32_t n, float* a, float* b) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const V3 ra = offset_ray(v3(p + 3 * i), v3(nn + 3 * i));
	const V3 rb = offset_ray2(v3(p + 3 * i), v3(nn + 3 * i));
	a[3 * i] = ra.x, a[3 * i + 1] = ra.y, a[3 * i + 2] = ra.z;
	b[3 * i] = rb.x, b[3 * i + 1] = rb.y, b[3 * i + 2] = rb.z;
}
__global__ void k_sample_bsdf(const lmb_material* mat, const float* ns, const float* wo, const float* r, const uint8_t* side, uint32_t n, float* out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const lmb_material m = *mat;
	const BsdfSample s = sample_bsdf(v3(ns + 3 * i), v3(wo + 3 * i), m, 1, side[i] != 0, v3(r + 3 * i));
	float* o = out + 8 * (size_t)i;
	o[0] = s.f.x, o[1] = s.f.y, o[2] = s.f.z, o[3] = s.wi.x, o[4] = s.wi.y, o[5] = s.wi.z, o[6] = s.pdf, o[7] = s.cos_theta;
}
__global__ void k_eval_bsdf(const lmb_material* mat, const float* ns, const float* wo, const float* wi, const uint8_t* side, uint32_t n, float* out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const lmb_material m = *mat;
	float pdf;
	const V3 f = eval_bsdf(v3(ns + 3 * i), v3(wo + 3 * i), m, side[i] != 0, v3(wi + 3 * i), pdf);
	float* o = out + 4 * (size_t)i;
	o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = pdf;
}
__global__ void k_atmosphere(const float* org, const float* dir, const float* ld, const float* lL, uint32_t n, float* out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const V3 o = v3(org + 3 * i), d = v3(dir + 3 * i);
	float ray_length = 10000.0f;
	const V2 pi = atmo::planet_intersection(o, d);
	if (pi.x > 0) ray_length = gmin(ray_length, pi.x);
	const V3 r = atmo::integrate_scattering(o, d, ray_length, v3(ld), v3(lL));
	out[3 * i] = r.x, out[3 * i + 1] = r.y, out[3 * i + 2] = r.z;
}
__global__ void k_sample_light(DeviceScene sc, int num_lights, const float* r, const float* p, uint32_t n, float* out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const V4 rr = v4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
	const LightSample l = sample_light_Li(sc, rr, v3(p + 3 * i), num_lights);
	float* o = out + 16 * (size_t)i;
	o[0] = l.Le.x, o[1] = l.Le.y, o[2] = l.Le.z, o[3] = l.wi.x, o[4] = l.wi.y, o[5] = l.wi.z;
	o[6] = l.wi_len, o[7] = l.pdf_w, o[8] = l.pdf_a, o[9] = l.cos_from_light;
	o[10] = (float)(uint32_t)(rr.x * (float)num_lights), o[11] = (float)l.flags, o[12] = (float)l.triangle_idx, o[13] = (float)l.instance_idx;
	// bary as sampled (commons.glsl:133)
	const float sq = sqrtf(rr.z);
	const bool area = (l.flags & 7u) == LMB_LIGHT_AREA;
	o[14] = area ? 1 - sq : 0.0f, o[15] = area ? rr.w * sq : 0.0f;
}
__global__ void k_texture(DeviceScene sc, uint32_t tex, const float* uv, uint32_t n, float* out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const V3 c = sample_texture(sc, tex, v2(uv[2 * i], uv[2 * i + 1]));
	out[3 * i] = c.x, out[3 * i + 1] = c.y, out[3 * i + 2] = c.z;
}

}  // namespace

extern "C" {

int lmb_kat_pcg4d(lmb_ctx* ctx, const uint32_t* in4, uint32_t n, uint32_t* out4) {
	if (!ctx) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	Buf a, b;
	int rc;
	if ((rc = to_dev(ctx, a, in4, 16 * (size_t)n)) || (rc = to_dev(ctx, b, nullptr, 16 * (size_t)n))) return rc;
	k_pcg4d<<<GRID(n)>>>((const uint32_t*)a.d, n, (uint32_t*)b.d);
	return to_host(ctx, out4, b);
}
int lmb_kat_rand(lmb_ctx* ctx, const uint32_t* seed4, uint32_t n, uint32_t draws, float* out) {
	if (!ctx) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	Buf a, b;
	int rc;
	if ((rc = to_dev(ctx, a, seed4, 16 * (size_t)n)) || (rc = to_dev(ctx, b, nullptr, 4 * (size_t)n * draws))) return rc;
	k_rand<<<GRID(n)>>>((const uint32_t*)a.d, n, draws, (float*)b.d);
	return to_host(ctx, out, b);
}
int lmb_kat_detmath(lmb_ctx* ctx, const float* x, const float* y, uint32_t n, float* out_sin, float* out_cos, float* out_exp, float* out_pow) {
	if (!ctx) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	Buf a, b, s, c, e, p;
	int rc;
	const size_t nb = 4 * (size_t)n;
	if ((rc = to_dev(ctx, a, x, nb)) || (rc = to_dev(ctx, b, y, nb)) || (rc = to_dev(ctx, s, nullptr, nb)) || (rc = to_dev(ctx, c, nullptr, nb)) ||
		(rc = to_dev(ctx, e, nullptr, nb)) || (rc = to_dev(ctx, p, nullptr, nb)))
		return rc;
	k_detmath<<<GRID(n)>>>((const float*)a.d, (const float*)b.d, n, (float*)s.d, (float*)c.d, (float*)e.d, (float*)p.d);
	if ((rc = to_host(ctx, out_sin, s)) || (rc = to_host(ctx, out_cos, c)) || (rc = to_host(ctx, out_exp, e))) return rc;
	return to_host(ctx, out_pow, p);
}
int lmb_kat_offset_ray(lmb_ctx* ctx, const float* p3, const float* n3, uint32_t n, float* out3, float* out3_b) {
	if (!ctx) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	Buf a, b, c, d;
	int rc;
	const size_t nb = 12 * (size_t)n;
	if ((rc = to_dev(ctx, a, p3, nb)) || (rc = to_dev(ctx, b, n3, nb)) || (rc = to_dev(ctx, c, nullptr, nb)) || (rc = to_dev(ctx, d, nullptr, nb))) return rc;
	k_offset_ray<<<GRID(n)>>>((const float*)a.d, (const float*)b.d, n, (float*)c.d, (float*)d.d);
	if ((rc = to_host(ctx, out3, c))) return rc;
	return to_host(ctx, out3_b, d);
}
int lmb_kat_sample_bsdf(lmb_ctx* ctx, const lmb_material* mat, const float* n_s3, const float* wo3, const float* rands3, const uint8_t* side, uint32_t n,
						float* out8) {
	if (!ctx) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	Buf m, a, b, c, s, o;
	int rc;
	const size_t nb = 12 * (size_t)n;
	if ((rc = to_dev(ctx, m, mat, sizeof(lmb_material))) || (rc = to_dev(ctx, a, n_s3, nb)) || (rc = to_dev(ctx, b, wo3, nb)) ||
		(rc = to_dev(ctx, c, rands3, nb)) || (rc = to_dev(ctx, s, side, n)) || (rc = to_dev(ctx, o, nullptr, 32 * (size_t)n)))
		return rc;
	k_sample_bsdf<<<GRID(n)>>>((const lmb_material*)m.d, (const float*)a.d, (const float*)b.d, (const float*)c.d, (const uint8_t*)s.d, n, (float*)o.d);
	return to_host(ctx, out8, o);
}
int lmb_kat_eval_bsdf(lmb_ctx* ctx, const lmb_material* mat, const float* n_s3, const float* wo3, const float* wi3, const uint8_t* side, uint32_t n,
					  float* out4) {
	if (!ctx) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	Buf m, a, b, c, s, o;
	int rc;
	const size_t nb = 12 * (size_t)n;
	if ((rc = to_dev(ctx, m, mat, sizeof(lmb_material))) || (rc = to_dev(ctx, a, n_s3, nb)) || (rc = to_dev(ctx, b, wo3, nb)) ||
		(rc = to_dev(ctx, c, wi3, nb)) || (rc = to_dev(ctx, s, side, n)) || (rc = to_dev(ctx, o, nullptr, 16 * (size_t)n)))
		return rc;
	k_eval_bsdf<<<GRID(n)>>>((const lmb_material*)m.d, (const float*)a.d, (const float*)b.d, (const float*)c.d, (const uint8_t*)s.d, n, (float*)o.d);
	return to_host(ctx, out4, o);
}
int lmb_kat_atmosphere(lmb_ctx* ctx, const float* origin3, const float* dir3, const float* light_dir3, const float* light_L3, uint32_t n, float* out3) {
	if (!ctx) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	Buf a, b, c, d, o;
	int rc;
	const size_t nb = 12 * (size_t)n;
	if ((rc = to_dev(ctx, a, origin3, nb)) || (rc = to_dev(ctx, b, dir3, nb)) || (rc = to_dev(ctx, c, light_dir3, 12)) || (rc = to_dev(ctx, d, light_L3, 12)) ||
		(rc = to_dev(ctx, o, nullptr, nb)))
		return rc;
	k_atmosphere<<<GRID(n)>>>((const float*)a.d, (const float*)b.d, (const float*)c.d, (const float*)d.d, n, (float*)o.d);
	return to_host(ctx, out3, o);
}
int lmb_kat_sample_light(lmb_ctx* ctx, int32_t num_lights, const float* rands4, const float* p3, uint32_t n, float* out16) {
	if (!ctx || !ctx->scene_loaded) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	Buf a, b, o;
	int rc;
	if ((rc = to_dev(ctx, a, rands4, 16 * (size_t)n)) || (rc = to_dev(ctx, b, p3, 12 * (size_t)n)) || (rc = to_dev(ctx, o, nullptr, 64 * (size_t)n))) return rc;
	k_sample_light<<<GRID(n)>>>(ctx->scene, num_lights, (const float*)a.d, (const float*)b.d, n, (float*)o.d);
	return to_host(ctx, out16, o);
}
int lmb_kat_texture(lmb_ctx* ctx, uint32_t tex, const float* uv2, uint32_t n, float* out3) {
	if (!ctx || !ctx->scene_loaded || tex >= ctx->scene.n_textures) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	Buf a, o;
	int rc;
	if ((rc = to_dev(ctx, a, uv2, 8 * (size_t)n)) || (rc = to_dev(ctx, o, nullptr, 12 * (size_t)n))) return rc;
	k_texture<<<GRID(n)>>>(ctx->scene, tex, (const float*)a.d, n, (float*)o.d);
	return to_host(ctx, out3, o);
}

}  // extern "C"

// ---- structural check of the 8-wide traversal BVH (wide_bvh.cu), done on the host from a device read-back ----------
namespace {
struct WideCheck {
	const std::vector<float4>& nodes;
	const std::vector<float4>& tris;
	std::vector<uint8_t> seen;
	uint64_t errors = 0, dup = 0, max_depth = 0, internal = 0, leaves = 0, leaf_tris = 0, visited = 0;
	// returns the exact bounds of everything below `node`
	void walk(uint32_t node, uint64_t depth, double lo[3], double hi[3]) {
		visited++;
		max_depth = std::max(max_depth, depth);
		const float4* n = &nodes[5 * (size_t)node];
		const uint32_t ew = __builtin_bit_cast(uint32_t, n[0].w);
		const double p[3] = {n[0].x, n[0].y, n[0].z};
		double step[3];
		for (int a = 0; a < 3; a++) step[a] = ldexp(1.0, (int)((ew >> (8 * a)) & 0xFFu) - 15 - 127);
		const uint32_t child_base = __builtin_bit_cast(uint32_t, n[1].x), tri_base = __builtin_bit_cast(uint32_t, n[1].y);
		const uint32_t mw = __builtin_bit_cast(uint32_t, n[1].z), imask = mw >> 24, leaf24 = mw & 0xFFFFFFu;
		const uint32_t w[12] = {__builtin_bit_cast(uint32_t, n[2].x), __builtin_bit_cast(uint32_t, n[2].y), __builtin_bit_cast(uint32_t, n[2].z),
								__builtin_bit_cast(uint32_t, n[2].w), __builtin_bit_cast(uint32_t, n[3].x), __builtin_bit_cast(uint32_t, n[3].y),
								__builtin_bit_cast(uint32_t, n[3].z), __builtin_bit_cast(uint32_t, n[3].w), __builtin_bit_cast(uint32_t, n[4].x),
								__builtin_bit_cast(uint32_t, n[4].y), __builtin_bit_cast(uint32_t, n[4].z), __builtin_bit_cast(uint32_t, n[4].w)};
		auto q = [&](int plane /*0..5 = lo xyz, hi xyz*/, int slot) { return (double)((w[2 * plane + (slot >> 2)] >> (8 * (slot & 3))) & 0xFFu); };
		for (int a = 0; a < 3; a++) lo[a] = 1e300, hi[a] = -1e300;
		for (int s = 0; s < 8; s++) {
			const bool inner = (imask >> s) & 1u;
			const uint32_t bits = (leaf24 >> (3 * s)) & 7u;
			if (!inner && bits == 0) {  // empty slot: must carry an inverted box
				if (!(q(0, s) > q(3, s))) errors++;
				continue;
			}
			double blo[3], bhi[3], clo[3], chi[3];
			for (int a = 0; a < 3; a++) blo[a] = p[a] + q(a, s) * step[a], bhi[a] = p[a] + q(3 + a, s) * step[a];
			if (inner) {
				if (bits != 0) errors++;
				internal++;
				const uint32_t rel = (uint32_t)__builtin_popcount(imask & ((1u << s) - 1u));
				walk(child_base + rel, depth + 1, clo, chi);
			} else {
				leaves++;
				const uint32_t off = (uint32_t)__builtin_popcount(leaf24 & ((1u << (3 * s)) - 1u));
				const uint32_t cnt = bits == 1 ? 1 : (bits == 3 ? 2 : (bits == 7 ? 3 : 0));
				if (cnt == 0) errors++;
				for (int a = 0; a < 3; a++) clo[a] = 1e300, chi[a] = -1e300;
				for (uint32_t t = 0; t < cnt; t++) {
					const size_t ti = (size_t)tri_base + off + t;
					if (3 * ti + 2 >= tris.size()) {
						errors++;
						continue;
					}
					leaf_tris++;
					const uint32_t prim = __builtin_bit_cast(uint32_t, tris[3 * ti].w);
					if (prim >= seen.size() || seen[prim]) dup++;
					else seen[prim] = 1;
					for (int v = 0; v < 3; v++) {
						const float4& P = tris[3 * ti + v];
						const double c[3] = {P.x, P.y, P.z};
						for (int a = 0; a < 3; a++) clo[a] = std::min(clo[a], c[a]), chi[a] = std::max(chi[a], c[a]);
					}
				}
			}
			for (int a = 0; a < 3; a++) {
				if (!(blo[a] <= clo[a] && bhi[a] >= chi[a])) errors++;  // quantised child box must contain everything below it
				lo[a] = std::min(lo[a], clo[a]), hi[a] = std::max(hi[a], chi[a]);
			}
		}
	}
};
}  // namespace

extern "C" int lmb_kat_wide_bvh_check(lmb_ctx* ctx, uint64_t* out8) {
	if (!ctx || !out8 || !ctx->bvh.built) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	const DeviceWideBvh& wb = ctx->wide;
	for (int i = 0; i < 8; i++) out8[i] = 0;
	if (wb.n_tris == 0) return LMB_OK;
	std::vector<float4> nodes(5 * (size_t)wb.n_nodes), tris(3 * (size_t)wb.n_tris);
	LMB_CUDA(ctx, cudaMemcpy(nodes.data(), wb.nodes, nodes.size() * sizeof(float4), cudaMemcpyDeviceToHost));
	LMB_CUDA(ctx, cudaMemcpy(tris.data(), wb.tris, tris.size() * sizeof(float4), cudaMemcpyDeviceToHost));
	WideCheck wc{nodes, tris, std::vector<uint8_t>(wb.n_tris, 0)};
	double lo[3], hi[3];
	wc.walk(0, 1, lo, hi);
	uint64_t missing = 0;
	for (uint8_t f : wc.seen) missing += f ? 0 : 1;
	out8[0] = wb.n_nodes, out8[1] = wc.visited, out8[2] = wc.max_depth, out8[3] = wc.errors, out8[4] = wc.dup + missing, out8[5] = wc.internal,
	out8[6] = wc.leaves, out8[7] = wc.leaf_tris;
	return LMB_OK;
}
