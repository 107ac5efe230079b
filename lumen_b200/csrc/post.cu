// post.cu -- the steps either side of the integrator that SURVEY.md 8f ranks next, done on the device:
//   * scene ingest: the per-triangle tables the kernels index by global triangle id (mesh, local id, vertex record, shade
//     queue) are derived on the GPU from the uploaded PrimMeshInfo[] / indices[] instead of host loops over every triangle;
//   * EXR output: the RGBA32F film is converted to the three planar HALF channels B, G, R that ImageUtils::save_exr writes
//     (src/Framework/ImageUtils.cpp:22-89, tinyexr's float_to_half_full rounding) before it leaves the device: 6 bytes per
//     pixel cross PCIe instead of 16;
//   * RMSE against a ground-truth image: Lumen's own routine (src/shaders/rmse/calc_rmse.comp, reduce_rmse.comp,
//     output_rmse.comp, dispatched by RayTracer.cpp:215-241), quirks included, plus a true RMSE for convergence stops.
#include "context.h"

namespace lmb {

namespace {

// ------------------------------------------------------------------------------------------------ scene ingest
// One TLAS instance per prim mesh (Integrator.cpp:148-158) -> global triangle ids are the meshes' triangles concatenated.
// tri_first[m] = first global id of mesh m (n_meshes + 1 entries).
__global__ void __launch_bounds__(256) k_ingest_triangles(const lmb_prim_mesh_info* __restrict__ prim_infos, const uint32_t* __restrict__ indices,
															 const uint32_t* __restrict__ tri_first, const uint8_t* __restrict__ mat_q, uint32_t n_meshes,
															 uint32_t n_materials, uint32_t n_tris, uint32_t* __restrict__ tri_mesh, uint32_t* __restrict__ tri_local,
															 uint4* __restrict__ tri_rec, uint8_t* __restrict__ tri_matq, const lmb_vertex* __restrict__ vertices,
															 float4* __restrict__ tri_shade) {
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tris; t += gridDim.x * blockDim.x) {
		uint32_t lo = 0, hi = n_meshes;  // last m with tri_first[m] <= t
		while (hi - lo > 1) {
			const uint32_t mid = (lo + hi) >> 1;
			if (tri_first[mid] <= t) lo = mid;
			else hi = mid;
		}
		const uint32_t m = lo, local = t - tri_first[m];
		const lmb_prim_mesh_info pi = prim_infos[m];
		const uint32_t* ix = indices + pi.index_offset + 3 * (size_t)local;
		tri_mesh[t] = m;
		tri_local[t] = local;
		tri_rec[t] = make_uint4(ix[0] + pi.vertex_offset, ix[1] + pi.vertex_offset, ix[2] + pi.vertex_offset, m);
		tri_matq[t] = pi.material_index < n_materials ? mat_q[pi.material_index] : (uint8_t)6;
		// the shading record of build_hit (scene_device.cuh): positions, normals, uvs of the three corners, mesh, material
		const lmb_vertex a = vertices[ix[0] + pi.vertex_offset], b = vertices[ix[1] + pi.vertex_offset], c = vertices[ix[2] + pi.vertex_offset];
		float4* r = tri_shade + 8 * (size_t)t;
		r[0] = make_float4(a.pos[0], a.pos[1], a.pos[2], a.uv0[0]);
		r[1] = make_float4(b.pos[0], b.pos[1], b.pos[2], a.uv0[1]);
		r[2] = make_float4(c.pos[0], c.pos[1], c.pos[2], b.uv0[0]);
		r[3] = make_float4(a.normal[0], a.normal[1], a.normal[2], b.uv0[1]);
		r[4] = make_float4(b.normal[0], b.normal[1], b.normal[2], c.uv0[0]);
		r[5] = make_float4(c.normal[0], c.normal[1], c.normal[2], c.uv0[1]);
		r[6] = make_float4(__uint_as_float(m), __uint_as_float(pi.material_index), 0.0f, 0.0f);
		r[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
}

// ------------------------------------------------------------------------------------------------ EXR half planes
// tinyexr float_to_half_full (libs/tinyexr.h:898-934): truncate the mantissa, add one when the first dropped bit is set
// (round half UP in magnitude, not ties-to-even); NaN -> 0x200 quiet NaN; overflow -> infinity; fp32 subnormals -> signed 0.
__device__ __forceinline__ uint16_t float_to_half_tinyexr(float f) {
	const uint32_t u = __float_as_uint(f);
	const uint32_t sign = (u >> 16) & 0x8000u, exp8 = (u >> 23) & 0xFFu, man = u & 0x007FFFFFu;
	uint32_t o = 0;
	if (exp8 == 0u) {
		o = 0;
	} else if (exp8 == 255u) {
		o = 0x7C00u | (man ? 0x200u : 0u);
	} else {
		const int newexp = (int)exp8 - 127 + 15;
		if (newexp >= 31) {
			o = 0x7C00u;
		} else if (newexp <= 0) {
			if ((14 - newexp) <= 24) {
				const uint32_t mant = man | 0x800000u;
				o = mant >> (14 - newexp);
				if ((mant >> (13 - newexp)) & 1u) o++;
			}
		} else {
			o = ((uint32_t)newexp << 10) | (man >> 13);
			if (man & 0x1000u) o++;
		}
	}
	return (uint16_t)(sign | (o & 0x7FFFu));
}

// planes: B then G then R, n_pix halves each. Two pixels per thread so that the plane stores are 32-bit.
__global__ void __launch_bounds__(256) k_film_to_half_bgr(const float4* __restrict__ film, uint32_t n_pix, uint16_t* __restrict__ planes) {
	const uint32_t n_pairs = (n_pix + 1) / 2;
	for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += gridDim.x * blockDim.x) {
		const uint32_t i0 = 2 * p, i1 = 2 * p + 1;
		const float4 a = film[i0];
		const float4 b = i1 < n_pix ? film[i1] : make_float4(0, 0, 0, 0);
		const float ca[3] = {a.z, a.y, a.x}, cb[3] = {b.z, b.y, b.x};
#pragma unroll
		for (int c = 0; c < 3; c++) {
			const uint16_t h0 = float_to_half_tinyexr(ca[c]), h1 = float_to_half_tinyexr(cb[c]);
			uint16_t* dst = planes + (size_t)c * n_pix;
			if (i1 < n_pix && ((((size_t)c * n_pix + i0) & 1u) == 0u))
				*reinterpret_cast<uint32_t*>(dst + i0) = (uint32_t)h0 | ((uint32_t)h1 << 16);
			else {
				dst[i0] = h0;
				if (i1 < n_pix) dst[i1] = h1;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ RMSE
// calc_rmse.comp:21-42 with a 1024-thread workgroup and 32-wide subgroups. GLSL leaves the order of subgroupAdd and the
// content of shared slots of inactive subgroups undefined; the frozen choices are the oracle's (oracle.cpp orc_rmse_literal):
// sums run in lane order, inactive lanes add nothing, slots of wholly inactive subgroups read 0. The subgroupMin of
// calc_rmse.comp:37 (where a sum was meant) is kept. One double-precision sum of squares per block feeds the true RMSE.
__global__ void __launch_bounds__(1024) k_rmse_calc(const float4* __restrict__ gt, const float4* __restrict__ out, uint32_t n_pix, float* __restrict__ residual,
													   double* __restrict__ sq_sum) {
	__shared__ float data[32];
	__shared__ double dsum[32];
	const uint32_t idx = blockIdx.x * 1024u + threadIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	float val = 0.0f;
	double dval = 0.0;
	if (idx < n_pix) {
		const float4 a = gt[idx], b = out[idx];
		const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
		val = dx * dx + dy * dy + dz * dz;  // dot(diff, diff)
		const double ex = (double)a.x - (double)b.x, ey = (double)a.y - (double)b.y, ez = (double)a.z - (double)b.z;
		dval = ex * ex + ey * ey + ez * ez;
	}
	const uint32_t active = __ballot_sync(0xFFFFFFFFu, idx < n_pix);
	float sum = 0.0f;
	for (int l = 0; l < 32; l++) {
		const float v = __shfl_sync(0xFFFFFFFFu, val, l);
		if (active & (1u << l)) sum += v;
	}
	for (int o = 16; o > 0; o >>= 1) dval += __shfl_xor_sync(0xFFFFFFFFu, dval, o);
	if (lane == 0) data[warp] = active ? sum : 0.0f, dsum[warp] = dval;
	__syncthreads();
	if (threadIdx.x == 0) {
		float mn = data[0];
		double ds = dsum[0];
		for (int k = 1; k < 32; k++) {
			mn = data[k] < mn ? data[k] : mn;  // std::min(mn, data[k])
			ds += dsum[k];
		}
		residual[blockIdx.x] = mn;
		atomicAdd(sq_sum, ds);
	}
}
// reduce_rmse.comp:22-49: every workgroup of 1024 sums its slice of the residuals (in index order) into residual[group].
// The passes run in place exactly as the reference's dispatch loop does; one launch per pass, groups in ascending order are
// independent because group g only writes slot g <= its own first input.
__global__ void k_rmse_reduce(float* residual, uint32_t live, uint32_t group) {
	if (threadIdx.x != 0) return;
	const uint32_t first = group * 1024u, last = min(live, first + 1024u);
	float sum = 0.0f;
	for (uint32_t k = first; k < last; k++) sum += residual[k];
	residual[group] = sum;
}
// output_rmse.comp:21-24
__global__ void k_rmse_output(const float* residual, const double* sq_sum, uint32_t n_pix, float* rmse_literal, double* rmse_true) {
	*rmse_literal = sqrtf(n_pix ? residual[0] : 0.0f) / ((float)n_pix * 3.0f);
	*rmse_true = n_pix ? sqrt(*sq_sum / (3.0 * (double)n_pix)) : 0.0;
}

}  // namespace

int ingest_triangles(lmb_ctx* ctx, const uint32_t* d_tri_first, const uint8_t* d_mat_q, uint32_t n_meshes, uint32_t n_materials, uint32_t n_tris,
					 uint32_t* tri_mesh, uint32_t* tri_local, uint4* tri_rec, uint8_t* tri_matq, float4* tri_shade) {
	if (n_tris == 0) return 0;
	const int grid = std::min<uint32_t>((n_tris + 255) / 256, (uint32_t)ctx->sm_count * 8);
	k_ingest_triangles<<<grid, 256, 0, ctx->stream>>>(ctx->scene.prim_infos, ctx->scene.indices, d_tri_first, d_mat_q, n_meshes, n_materials, n_tris, tri_mesh,
														 tri_local, tri_rec, tri_matq, ctx->scene.vertices, tri_shade);
	return check_cuda(ctx, cudaGetLastError(), "k_ingest_triangles");
}

__global__ void __launch_bounds__(256) k_film_add(float4* __restrict__ film, const float4* __restrict__ other, uint32_t n_pix) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += gridDim.x * blockDim.x) {
		const float4 a = film[i], b = other[i];
		film[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
	}
}
int launch_film_add(lmb_ctx* ctx, const float4* d_other) {
	k_film_add<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->film, d_other, ctx->width * ctx->height);
	return check_cuda(ctx, cudaGetLastError(), "k_film_add");
}

int launch_film_to_half(lmb_ctx* ctx, uint16_t* d_planes) {
	const uint32_t n_pix = ctx->width * ctx->height;
	k_film_to_half_bgr<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->film, n_pix, d_planes);
	return check_cuda(ctx, cudaGetLastError(), "k_film_to_half_bgr");
}

// d_scratch: residual floats [(n_pix + 1023) / 1024] then, 8-byte aligned, one double (sum of squares), one double (true
// RMSE) and one float (literal RMSE); see rmse_scratch_bytes.
size_t rmse_scratch_bytes(uint32_t n_pix) { return (((size_t)(n_pix + 1023) / 1024 + 1) * 4 + 7) / 8 * 8 + 24; }
int launch_rmse(lmb_ctx* ctx, const float4* d_gt, void* d_scratch, float* h_literal, double* h_true) {
	const uint32_t n_pix = ctx->width * ctx->height;
	const uint32_t groups = (n_pix + 1023) / 1024;
	float* residual = (float*)d_scratch;
	double* dbl = (double*)((char*)d_scratch + (((size_t)groups + 1) * 4 + 7) / 8 * 8);
	float* lit = (float*)(dbl + 2);
	LMB_CUDA(ctx, cudaMemsetAsync(d_scratch, 0, rmse_scratch_bytes(n_pix), ctx->stream));  // .zero({residual_buffer, counter_buffer})
	if (groups) k_rmse_calc<<<groups, 1024, 0, ctx->stream>>>(d_gt, ctx->film, n_pix, residual, dbl);
	for (uint32_t live = groups; live > 1;) {
		const uint32_t g2 = (live + 1023) / 1024;
		for (uint32_t g = 0; g < g2; g++) k_rmse_reduce<<<1, 32, 0, ctx->stream>>>(residual, live, g);
		live = g2;
	}
	k_rmse_output<<<1, 1, 0, ctx->stream>>>(residual, dbl, n_pix, lit, dbl + 1);
	LMB_CUDA(ctx, cudaGetLastError());
	float hl = 0;
	double ht = 0;
	LMB_CUDA(ctx, cudaMemcpyAsync(&hl, lit, 4, cudaMemcpyDeviceToHost, ctx->stream));
	LMB_CUDA(ctx, cudaMemcpyAsync(&ht, dbl + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (h_literal) *h_literal = hl;
	if (h_true) *h_true = ht;
	return 0;
}

}  // namespace lmb
