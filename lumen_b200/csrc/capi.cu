// capi.cu -- implementation of the C ABI declared in include/lumen_b200.h (context lifetime, scene upload, film I/O,
// ray-query entry points). No CPU fallback: every entry point needs a CUDA device and says so when there is none.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "context.h"
#include "lumen_b200_testhooks.h"

static thread_local std::string g_create_error;

namespace lmb {

int set_error(lmb_ctx* ctx, int code, const std::string& msg) {
	if (ctx)
		ctx->err = msg;
	else
		g_create_error = msg;
	return code;
}

int check_cuda(lmb_ctx* ctx, cudaError_t e, const char* what) {
	if (e == cudaSuccess) return 0;
	const int code = (e == cudaErrorMemoryAllocation) ? LMB_ERR_OOM : LMB_ERR_CUDA;
	return set_error(ctx, code, std::string(what) + ": " + cudaGetErrorString(e));
}

namespace {
template <typename T>
int upload(lmb_ctx* ctx, const T* host, size_t count, const T** dev_out) {
	void* d = nullptr;
	LMB_CUDA(ctx, cudaMalloc(&d, std::max<size_t>(count, 1) * sizeof(T)));
	ctx->scene_allocs.push_back(d);
	if (count)
		LMB_CUDA(ctx, cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
	else
		LMB_CUDA(ctx, cudaMemsetAsync(d, 0, sizeof(T), ctx->stream));  // e.g. a scene without lights reads one all-zero Light
	*dev_out = (const T*)d;
	return 0;
}

// buffers of the post steps (half planes, ground-truth image, RMSE scratch) are sized by the film
void free_post(lmb_ctx* ctx) {
	if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
	ctx->copy_pending = false;
	cudaFree(ctx->half_planes), cudaFree(ctx->gt_img), cudaFree(ctx->rmse_scratch), cudaFree(ctx->film_snapshot);
	ctx->half_planes = nullptr, ctx->gt_img = nullptr, ctx->rmse_scratch = nullptr, ctx->film_snapshot = nullptr, ctx->has_gt = false;
}

void free_scene(lmb_ctx* ctx) {
	for (void* p : ctx->scene_allocs) cudaFree(p);
	ctx->scene_allocs.clear();
	ctx->scene = DeviceScene{};
	ctx->scene_loaded = false;
	free_bvh(ctx);
}
}  // namespace
}  // namespace lmb

using namespace lmb;

extern "C" {

const char* lmb_last_error(const lmb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int lmb_create(lmb_ctx** out, int device_id) {
	if (!out) return set_error(nullptr, LMB_ERR_INVALID, "lmb_create: out is NULL");
	int n = 0;
	const cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
		return set_error(nullptr, LMB_ERR_NO_DEVICE,
						 std::string("lmb_create: no CUDA device (") + cudaGetErrorString(e) + "); lumen_b200 has no CPU fallback");
	if (device_id < 0 || device_id >= n) return set_error(nullptr, LMB_ERR_INVALID, "lmb_create: device_id out of range");
	lmb_ctx* ctx = new lmb_ctx();
	ctx->device = device_id;
	if (cudaSetDevice(device_id) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
		const std::string msg = std::string("lmb_create: ") + cudaGetErrorString(cudaGetLastError());
		delete ctx;
		return set_error(nullptr, LMB_ERR_CUDA, msg);
	}
	cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_id);
	for (auto& ev : ctx->ev) cudaEventCreate(&ev);
	const char* trav = getenv("LMB_TRAVERSAL");
	ctx->use_bvh2 = trav && strcmp(trav, "bvh2") == 0;
	const char* tree = getenv("LMB_TREE");
	ctx->use_ploc = !(tree && strcmp(tree, "lbvh") == 0);
	ctx->tree_auto = !(tree && (strcmp(tree, "lbvh") == 0 || strcmp(tree, "ploc") == 0));
	const char* spl = getenv("LMB_STATS_PER_LAUNCH");
	ctx->stats_per_launch = spl && *spl && strcmp(spl, "0") != 0;
	const char* pin = getenv("LMB_TRACE_PIN");
	ctx->trace_pin = (pin && *pin) ? (atoi(pin) != 0) : -1;
	*out = ctx;
	return LMB_OK;
}

void lmb_destroy(lmb_ctx* ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	comm_free(ctx);
	cudaFree(ctx->ray_keys), cudaFree(ctx->ray_order), cudaFree(ctx->ray_hist);
	wavefront_free(ctx);
	cudaFree(ctx->film);
	free_post(ctx);
	free_scene(ctx);
	for (auto& ev : ctx->ev) cudaEventDestroy(ev);
	if (ctx->ev_snapshot) cudaEventDestroy(ctx->ev_snapshot);
	if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
	if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
	if (ctx->ev_miss_ready) cudaEventDestroy(ctx->ev_miss_ready);
	if (ctx->ev_miss_done) cudaEventDestroy(ctx->ev_miss_done);
	if (ctx->miss_stream) cudaStreamDestroy(ctx->miss_stream);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
}

// Every index the kernels will follow is checked here, once, on the host: the reference trusts its own loader, a C ABI cannot.
static int validate_scene(lmb_ctx* ctx, const lmb_scene_desc* sd) {
	auto bad = [&](const std::string& what) { return set_error(ctx, LMB_ERR_INVALID, "lmb_upload_scene: " + what); };
	if ((sd->n_vertices && !sd->vertices) || (sd->n_indices && !sd->indices) || (sd->n_materials && !sd->materials) || (sd->n_lights && !sd->lights) ||
		(sd->n_textures && !sd->textures) || (sd->n_prim_meshes && (!sd->prim_infos || !sd->prim_idx_counts || !sd->world_matrices || !sd->inv_world_matrices)))
		return bad("null array with a non-zero count");
	uint64_t total_tris = 0;
	for (uint32_t m = 0; m < sd->n_prim_meshes; m++) {
		const lmb_prim_mesh_info& pi = sd->prim_infos[m];
		const uint32_t cnt = sd->prim_idx_counts[m];
		if (cnt % 3) return bad("prim mesh " + std::to_string(m) + ": index count is not a multiple of 3");
		if ((uint64_t)pi.index_offset + cnt > sd->n_indices) return bad("prim mesh " + std::to_string(m) + ": index range exceeds n_indices");
		if (pi.material_index >= sd->n_materials) return bad("prim mesh " + std::to_string(m) + ": material_index out of range");
		for (uint32_t i = 0; i < cnt; i++)
			if ((uint64_t)sd->indices[pi.index_offset + i] + pi.vertex_offset >= sd->n_vertices)
				return bad("prim mesh " + std::to_string(m) + ": vertex index out of range");
		total_tris += cnt / 3;
	}
	if (total_tris > 0x07FFFFFFull) return bad("more than 2^27 - 1 triangles (the traversal packs triangle indices in 27 bits)");
	for (uint32_t i = 0; i < sd->n_materials; i++)
		if (sd->materials[i].texture_id >= (int32_t)sd->n_textures) return bad("material " + std::to_string(i) + ": texture_id out of range");
	for (uint32_t i = 0; i < sd->n_textures; i++)
		if (!sd->textures[i].rgba8 || sd->textures[i].width == 0 || sd->textures[i].height == 0) return bad("texture " + std::to_string(i) + ": empty");
	for (uint32_t i = 0; i < sd->n_lights; i++) {
		const lmb_light& l = sd->lights[i];
		if ((l.light_flags & 0x7u) != LMB_LIGHT_AREA) continue;
		if (l.prim_mesh_idx >= sd->n_prim_meshes) return bad("area light " + std::to_string(i) + ": prim_mesh_idx out of range");
		if (l.num_triangles == 0 || l.num_triangles > sd->prim_idx_counts[l.prim_mesh_idx] / 3) return bad("area light " + std::to_string(i) + ": num_triangles does not fit its mesh");
		// sample_triangle (commons.glsl:130) takes transpose(inverse(light.world_matrix)); the kernels read the host-inverted MESH
		// matrix instead (inv_world_matrices[prim_mesh_idx]), which is the same thing only while the two matrices are equal -- as
		// LumenScene.cpp:86-93 builds them
		if (memcmp(l.world_matrix, sd->world_matrices + 16 * (size_t)l.prim_mesh_idx, 64) != 0)
			return bad("area light " + std::to_string(i) + ": world_matrix differs from the world matrix of its mesh");
	}
	return 0;
}

int lmb_upload_scene(lmb_ctx* ctx, const lmb_scene_desc* sd) {
	if (!ctx || !sd) return LMB_ERR_INVALID;
	if (const int bad = validate_scene(ctx, sd)) return bad;
	cudaSetDevice(ctx->device);
	free_scene(ctx);
	DeviceScene& sc = ctx->scene;
	int rc;
	if ((rc = upload(ctx, sd->vertices, sd->n_vertices, &sc.vertices))) return rc;
	if ((rc = upload(ctx, sd->indices, sd->n_indices, &sc.indices))) return rc;
	if ((rc = upload(ctx, sd->materials, sd->n_materials, &sc.materials))) return rc;
	if ((rc = upload(ctx, sd->prim_infos, sd->n_prim_meshes, &sc.prim_infos))) return rc;
	if ((rc = upload(ctx, sd->world_matrices, 16 * (size_t)sd->n_prim_meshes, &sc.world_matrices))) return rc;
	if ((rc = upload(ctx, sd->inv_world_matrices, 16 * (size_t)sd->n_prim_meshes, &sc.inv_world_matrices))) return rc;
	if ((rc = upload(ctx, sd->lights, sd->n_lights, &sc.lights))) return rc;
	// global triangle numbering: prim meshes concatenated in order (one TLAS instance per mesh, Integrator.cpp:148-158).
	// The per-triangle tables are filled on the device (post.cu) from the arrays just uploaded.
	std::vector<uint32_t> tri_first(sd->n_prim_meshes + 1, 0u);
	for (uint32_t m = 0; m < sd->n_prim_meshes; m++) tri_first[m + 1] = tri_first[m] + sd->prim_idx_counts[m] / 3;
	const uint32_t n_tris = tri_first[sd->n_prim_meshes];
	// textures: RGBA8 texels + sRGB decode table (VK_FORMAT_R8G8B8A8_SRGB, LumenScene.cpp:204-213)
	std::vector<const uint8_t*> tex_ptrs;
	std::vector<uint2> tex_dims;
	for (uint32_t i = 0; i < sd->n_textures; i++) {
		const uint8_t* d = nullptr;
		if ((rc = upload(ctx, sd->textures[i].rgba8, 4 * (size_t)sd->textures[i].width * sd->textures[i].height, &d))) return rc;
		tex_ptrs.push_back(d);
		tex_dims.push_back(make_uint2(sd->textures[i].width, sd->textures[i].height));
	}
	const uint8_t* const* d_ptrs = nullptr;
	if ((rc = upload(ctx, tex_ptrs.data(), tex_ptrs.size(), (const uint8_t* const**)&d_ptrs))) return rc;
	sc.tex_data = d_ptrs;
	if ((rc = upload(ctx, tex_dims.data(), tex_dims.size(), &sc.tex_dims))) return rc;
	float lut[256];
	for (int i = 0; i < 256; i++) {
		const double c = i / 255.0;
		lut[i] = (float)(c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4));
	}
	if ((rc = upload(ctx, lut, 256, &sc.srgb_lut))) return rc;
	ctx->mat_queue_mask = 0;
	std::vector<uint8_t> mat_q(sd->n_materials);
	for (uint32_t i = 0; i < sd->n_materials; i++) {
		int q = 6;
		for (int m = 0; m < 6; m++)
			if (sd->materials[i].bsdf_type == (1u << m)) q = m;
		ctx->mat_queue_mask |= 1u << q;
		mat_q[i] = (uint8_t)q;
	}
	// per triangle: mesh, local id, vertex record and the shade queue its material's BSDF type maps to (k_classify reads one
	// byte per path)
	const uint32_t* d_tri_first = nullptr;
	const uint8_t* d_mat_q = nullptr;
	if ((rc = upload(ctx, tri_first.data(), tri_first.size(), &d_tri_first))) return rc;
	if ((rc = upload(ctx, mat_q.data(), mat_q.size(), &d_mat_q))) return rc;
	uint32_t *d_tri_mesh = nullptr, *d_tri_local = nullptr;
	uint4* d_tri_rec = nullptr;
	uint8_t* d_tri_matq = nullptr;
	auto dev_alloc = [&](void** p, size_t bytes) -> int {
		const int r = check_cuda(ctx, cudaMalloc(p, std::max<size_t>(bytes, 16)), "lmb_upload_scene: triangle tables");
		if (!r) ctx->scene_allocs.push_back(*p);
		return r;
	};
	if ((rc = dev_alloc((void**)&d_tri_mesh, (size_t)n_tris * 4))) return rc;
	if ((rc = dev_alloc((void**)&d_tri_local, (size_t)n_tris * 4))) return rc;
	if ((rc = dev_alloc((void**)&d_tri_rec, (size_t)n_tris * 16))) return rc;
	if ((rc = dev_alloc((void**)&d_tri_matq, (size_t)n_tris))) return rc;
	float4* d_tri_shade = nullptr;
	if ((rc = dev_alloc((void**)&d_tri_shade, (size_t)n_tris * 128))) return rc;
	if ((rc = ingest_triangles(ctx, d_tri_first, d_mat_q, sd->n_prim_meshes, sd->n_materials, n_tris, d_tri_mesh, d_tri_local, d_tri_rec, d_tri_matq, d_tri_shade))) return rc;
	sc.tri_mesh = d_tri_mesh, sc.tri_local = d_tri_local, sc.tri_rec = d_tri_rec, sc.tri_matq = d_tri_matq, sc.tri_shade = d_tri_shade;
	sc.n_tris = n_tris;
	sc.n_prim_meshes = sd->n_prim_meshes;
	sc.n_lights = sd->n_lights;
	sc.n_textures = sd->n_textures;
	ctx->h_light_flags.resize(sd->n_lights);
	for (uint32_t i = 0; i < sd->n_lights; i++) ctx->h_light_flags[i] = sd->lights[i].light_flags;
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host staging vectors go out of scope
	ctx->scene_loaded = true;
	return LMB_OK;
}

int lmb_build_accel(lmb_ctx* ctx) {
	if (!ctx) return LMB_ERR_INVALID;
	if (!ctx->scene_loaded) return set_error(ctx, LMB_ERR_INVALID, "lmb_build_accel: no scene uploaded");
	cudaSetDevice(ctx->device);
	return build_lbvh(ctx);
}

int lmb_init(lmb_ctx* ctx, uint32_t width, uint32_t height, uint32_t frames_in_flight) {
	if (!ctx || width == 0 || height == 0) return set_error(ctx, LMB_ERR_INVALID, "lmb_init: bad size");
	cudaSetDevice(ctx->device);
	cudaFree(ctx->film);
	ctx->film = nullptr;
	free_post(ctx);
	if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
	cudaFree(ctx->reduce_buf);  // sized by the film
	ctx->reduce_buf = nullptr, ctx->reduce_pending = false;
	ctx->width = width, ctx->height = height;
	if (shard_rows(ctx) == 0) return set_error(ctx, LMB_ERR_INVALID, "lmb_init: the pixel shard owns no row of this image");
	LMB_CUDA(ctx, cudaMalloc((void**)&ctx->film, (size_t)width * height * 16));
	LMB_CUDA(ctx, cudaMemsetAsync(ctx->film, 0, (size_t)width * height * 16, ctx->stream));
	const int rc = wavefront_alloc(ctx, frames_in_flight);
	if (rc) {  // leave no half-initialised context behind: lmb_render's guard is ctx->film
		wavefront_free(ctx);
		cudaFree(ctx->film);
		ctx->film = nullptr;
		return rc;
	}
	return lmb_reset_stats(ctx);
}

int lmb_set_pixel_shard(lmb_ctx* ctx, uint32_t row_first, uint32_t row_stride) {
	if (!ctx) return LMB_ERR_INVALID;
	if (row_stride == 0 || row_first >= row_stride) return set_error(ctx, LMB_ERR_INVALID, "lmb_set_pixel_shard: need row_first < row_stride, row_stride >= 1");
	if (ctx->film && (row_first != ctx->row_first || row_stride != ctx->row_stride)) {
		// the wavefront state is sized by the shard: drop it, the caller runs lmb_init again
		cudaSetDevice(ctx->device);
		cudaStreamSynchronize(ctx->stream);
		wavefront_free(ctx);
		cudaFree(ctx->film);
		ctx->film = nullptr;
		free_post(ctx);
	}
	ctx->row_first = row_first, ctx->row_stride = row_stride;
	return LMB_OK;
}

// The push-constant fields the kernels use as indices or divisors, against the uploaded scene: sample_light_Li reads
// lights[uint(rand * num_lights)], shade_atmosphere / k_miss read lights[dir_light_idx] (commons.glsl:227,160). The reference
// fills them from its own scene (Path.cpp:27-37); a C ABI has to check.
static int check_light_args(lmb_ctx* ctx, const char* who, int32_t num_lights, uint32_t dir_light_idx, int32_t light_triangle_count, int32_t max_depth,
							int32_t max_depth_limit) {
	const std::string w(who);
	const uint32_t n_lights = ctx->scene.n_lights;
	if (num_lights < 0 || (uint32_t)num_lights > n_lights)
		return set_error(ctx, LMB_ERR_INVALID, w + ": num_lights " + std::to_string(num_lights) + " is outside [0, " + std::to_string(n_lights) + "] (uploaded lights)");
	if (dir_light_idx != 0xFFFFFFFFu) {
		if (dir_light_idx >= n_lights) return set_error(ctx, LMB_ERR_INVALID, w + ": dir_light_idx is neither 0xFFFFFFFF nor an uploaded light");
		if ((ctx->h_light_flags[dir_light_idx] & 0x7u) != LMB_LIGHT_DIRECTIONAL)
			return set_error(ctx, LMB_ERR_INVALID, w + ": dir_light_idx does not name a directional light");
	}
	if (light_triangle_count < 0 || (light_triangle_count == 0 && num_lights > 0))
		return set_error(ctx, LMB_ERR_INVALID, w + ": light_triangle_count must be positive when there are lights (1 / light_triangle_count is the light pick pdf, path.rgen:76)");
	if (max_depth < 0 || max_depth > max_depth_limit)
		return set_error(ctx, LMB_ERR_INVALID, w + ": max_depth " + std::to_string(max_depth) + " is outside [0, " + std::to_string(max_depth_limit) + "]");
	return LMB_OK;
}

int lmb_render(lmb_ctx* ctx, const lmb_pc_path* pc, const lmb_scene_ubo* ubo, uint32_t first_frame, uint32_t n_frames, uint32_t frame_stride,
			   int film_mode) {
	if (!ctx || !pc || !ubo) return LMB_ERR_INVALID;
	if (!ctx->bvh.built) return set_error(ctx, LMB_ERR_INVALID, "lmb_render: call lmb_build_accel first");
	if (!ctx->film) return set_error(ctx, LMB_ERR_INVALID, "lmb_render: call lmb_init first");
	if (pc->size_x != ctx->width || pc->size_y != ctx->height) return set_error(ctx, LMB_ERR_INVALID, "lmb_render: PCPath size != lmb_init size");
	if (const int bad = check_light_args(ctx, "lmb_render", pc->num_lights, pc->dir_light_idx, pc->light_triangle_count, pc->max_depth, 4096)) return bad;
	if (frame_stride == 0) frame_stride = 1;
	if (film_mode == LMB_FILM_RUNNING_MEAN && frame_stride != 1)
		return set_error(ctx, LMB_ERR_INVALID, "lmb_render: running-mean film needs frame_stride 1");
	if (film_mode != LMB_FILM_RUNNING_MEAN && film_mode != LMB_FILM_SUM) return set_error(ctx, LMB_ERR_INVALID, "lmb_render: bad film_mode");
	cudaSetDevice(ctx->device);
	if (n_frames == 0) return LMB_OK;
	return wavefront_render(ctx, *pc, *ubo, first_frame, n_frames, frame_stride, film_mode);
}

static int check_bdpt_args(lmb_ctx* ctx, const lmb_pc_bdpt* pc) {
	if (!ctx->bvh.built) return set_error(ctx, LMB_ERR_INVALID, "lmb_render_bdpt: call lmb_build_accel first");
	if (!ctx->film) return set_error(ctx, LMB_ERR_INVALID, "lmb_render_bdpt: call lmb_init first");
	if (pc->size_x != ctx->width || pc->size_y != ctx->height) return set_error(ctx, LMB_ERR_INVALID, "lmb_render_bdpt: PCBDPT size != lmb_init size");
	// two sub-paths of max_depth + 1 vertices (92 B each) are kept per pixel: the depth bounds the allocation
	return check_light_args(ctx, "lmb_render_bdpt", pc->num_lights, pc->dir_light_idx, pc->light_triangle_count, pc->max_depth, 64);
}

int lmb_render_bdpt(lmb_ctx* ctx, const lmb_pc_bdpt* pc, const lmb_scene_ubo* ubo, uint32_t first_frame, uint32_t n_frames, uint32_t frame_stride,
					int film_mode) {
	if (!ctx || !pc || !ubo) return LMB_ERR_INVALID;
	const int rc = check_bdpt_args(ctx, pc);
	if (rc) return rc;
	if (frame_stride == 0) frame_stride = 1;
	if (film_mode == LMB_FILM_RUNNING_MEAN && frame_stride != 1)
		return set_error(ctx, LMB_ERR_INVALID, "lmb_render_bdpt: running-mean film needs frame_stride 1");
	if (film_mode != LMB_FILM_RUNNING_MEAN && film_mode != LMB_FILM_SUM) return set_error(ctx, LMB_ERR_INVALID, "lmb_render_bdpt: bad film_mode");
	cudaSetDevice(ctx->device);
	if (n_frames == 0) return LMB_OK;
	return bdpt_render(ctx, *pc, *ubo, first_frame, n_frames, frame_stride, film_mode, nullptr, nullptr);
}

int lmb_kat_bdpt_frame_raw(lmb_ctx* ctx, const lmb_pc_bdpt* pc, const lmb_scene_ubo* ubo, uint32_t frame, float* col_rgba, float* splat_rgb) {
	if (!ctx || !pc || !ubo || !col_rgba || !splat_rgb) return LMB_ERR_INVALID;
	const int rc = check_bdpt_args(ctx, pc);
	if (rc) return rc;
	cudaSetDevice(ctx->device);
	return bdpt_render(ctx, *pc, *ubo, frame, 1, 1, LMB_FILM_RUNNING_MEAN, col_rgba, splat_rgb);
}

int lmb_clear_film(lmb_ctx* ctx) {
	if (!ctx || !ctx->film) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	LMB_CUDA(ctx, cudaMemsetAsync(ctx->film, 0, (size_t)ctx->width * ctx->height * 16, ctx->stream));
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LMB_OK;
}

int lmb_resolve(lmb_ctx* ctx) {
	if (!ctx || !ctx->film) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	const int rc = launch_resolve(ctx);
	if (rc) return rc;
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LMB_OK;
}

int lmb_download(lmb_ctx* ctx, float* rgba) {
	if (!ctx || !ctx->film || !rgba) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	LMB_CUDA(ctx, cudaMemcpyAsync(rgba, ctx->film, (size_t)ctx->width * ctx->height * 16, cudaMemcpyDefault, ctx->stream));
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LMB_OK;
}

// The film is snapshotted on the render stream (a device-to-device copy, ~20 us at 1080p) and sent home by a second stream
// on the copy engine, so the PCIe transfer (33 MB at 1080p) overlaps whatever is rendered next instead of stalling it.
int lmb_download_async(lmb_ctx* ctx, float* rgba) {
	if (!ctx || !ctx->film || !rgba) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	const size_t bytes = (size_t)ctx->width * ctx->height * 16;
	if (!ctx->copy_stream) {
		LMB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
		LMB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_snapshot, cudaEventDisableTiming));
		LMB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
	}
	if (!ctx->film_snapshot) LMB_CUDA(ctx, cudaMalloc((void**)&ctx->film_snapshot, bytes));
	if (ctx->copy_pending) LMB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));  // the previous transfer still reads the snapshot
	LMB_CUDA(ctx, cudaMemcpyAsync(ctx->film_snapshot, ctx->film, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	LMB_CUDA(ctx, cudaEventRecord(ctx->ev_snapshot, ctx->stream));
	LMB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_snapshot, 0));
	LMB_CUDA(ctx, cudaMemcpyAsync(rgba, ctx->film_snapshot, bytes, cudaMemcpyDefault, ctx->copy_stream));
	LMB_CUDA(ctx, cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
	ctx->copy_pending = true;
	return LMB_OK;
}

int lmb_sync(lmb_ctx* ctx) {
	if (!ctx) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (ctx->copy_stream) LMB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
	if (ctx->comm_stream) LMB_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
	ctx->copy_pending = false, ctx->reduce_pending = false;
	return LMB_OK;
}

int lmb_download_half_bgr(lmb_ctx* ctx, uint16_t* planes) {
	if (!ctx || !ctx->film || !planes) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	const size_t bytes = (size_t)ctx->width * ctx->height * 3 * sizeof(uint16_t);
	if (!ctx->half_planes) LMB_CUDA(ctx, cudaMalloc((void**)&ctx->half_planes, bytes));
	const int rc = launch_film_to_half(ctx, ctx->half_planes);
	if (rc) return rc;
	LMB_CUDA(ctx, cudaMemcpyAsync(planes, ctx->half_planes, bytes, cudaMemcpyDefault, ctx->stream));
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LMB_OK;
}

int lmb_set_reference_image(lmb_ctx* ctx, const float* gt_rgba) {
	if (!ctx || !ctx->film || !gt_rgba) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	const size_t bytes = (size_t)ctx->width * ctx->height * 16;
	if (!ctx->gt_img) LMB_CUDA(ctx, cudaMalloc((void**)&ctx->gt_img, bytes));
	LMB_CUDA(ctx, cudaMemcpyAsync(ctx->gt_img, gt_rgba, bytes, cudaMemcpyDefault, ctx->stream));
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->has_gt = true;
	return LMB_OK;
}

int lmb_rmse(lmb_ctx* ctx, float* rmse_literal, double* rmse_true) {
	if (!ctx || !ctx->film) return LMB_ERR_INVALID;
	if (!ctx->has_gt) return set_error(ctx, LMB_ERR_INVALID, "lmb_rmse: call lmb_set_reference_image first (has_gt, RayTracer.cpp:215)");
	cudaSetDevice(ctx->device);
	if (!ctx->rmse_scratch) LMB_CUDA(ctx, cudaMalloc(&ctx->rmse_scratch, rmse_scratch_bytes(ctx->width * ctx->height)));
	return launch_rmse(ctx, ctx->gt_img, ctx->rmse_scratch, rmse_literal, rmse_true);
}

// The reduce of the sum films of a multi-GPU render without leaving the devices: src's film is copied device to device
// (cudaMemcpyPeerAsync: NVLink / NVSwitch where the peers are connected, through the host otherwise) into a staging buffer of dst
// and added there. fp32 adds in call order: the caller fixes the order, so the result is reproducible.
int lmb_film_add_from(lmb_ctx* dst, lmb_ctx* src) {
	if (!dst || !src || !dst->film || !src->film) return LMB_ERR_INVALID;
	if (dst == src) return set_error(dst, LMB_ERR_INVALID, "lmb_film_add_from: dst == src");
	if (dst->width != src->width || dst->height != src->height) return set_error(dst, LMB_ERR_INVALID, "lmb_film_add_from: image sizes differ");
	const size_t bytes = (size_t)dst->width * dst->height * 16;
	cudaSetDevice(src->device);
	LMB_CUDA(dst, cudaStreamSynchronize(src->stream));
	cudaSetDevice(dst->device);
	if (!dst->film_snapshot) LMB_CUDA(dst, cudaMalloc((void**)&dst->film_snapshot, bytes));
	if (dst->copy_pending) {  // an lmb_download_async transfer may still read the staging buffer
		LMB_CUDA(dst, cudaStreamSynchronize(dst->copy_stream));
		dst->copy_pending = false;
	}
	LMB_CUDA(dst, cudaMemcpyPeerAsync(dst->film_snapshot, dst->device, src->film, src->device, bytes, dst->stream));
	const int rc = launch_film_add(dst, dst->film_snapshot);
	if (rc) return rc;
	LMB_CUDA(dst, cudaStreamSynchronize(dst->stream));
	return LMB_OK;
}

int lmb_upload_film(lmb_ctx* ctx, const float* rgba) {
	if (!ctx || !ctx->film || !rgba) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	LMB_CUDA(ctx, cudaMemcpyAsync(ctx->film, rgba, (size_t)ctx->width * ctx->height * 16, cudaMemcpyDefault, ctx->stream));
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LMB_OK;
}

int lmb_film_device_ptr(lmb_ctx* ctx, void** dptr, uint64_t* n_floats) {
	if (!ctx || !ctx->film || !dptr) return LMB_ERR_INVALID;
	*dptr = ctx->film;
	if (n_floats) *n_floats = (uint64_t)ctx->width * ctx->height * 4;
	return LMB_OK;
}

int lmb_stream(lmb_ctx* ctx, void** cuda_stream) {
	if (!ctx || !cuda_stream) return LMB_ERR_INVALID;
	*cuda_stream = (void*)ctx->stream;
	return LMB_OK;
}

int lmb_set_profile_stages(lmb_ctx* ctx, int on) {
	if (!ctx) return LMB_ERR_INVALID;
	ctx->profile_stages = on != 0;
	return LMB_OK;
}

int lmb_get_stats(lmb_ctx* ctx, lmb_stats* out) {
	if (!ctx || !out) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	if (ctx->wf.stats) {
		unsigned long long h[ST_COUNT];
		LMB_CUDA(ctx, cudaMemcpyAsync(h, ctx->wf.stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
		LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		ctx->stats.rays_closest = h[ST_CLOSEST], ctx->stats.rays_shadow = h[ST_SHADOW], ctx->stats.rays_probe = h[ST_PROBE];
		ctx->stats.nodes_visited = h[ST_NODES], ctx->stats.tris_tested = h[ST_TRIS], ctx->stats.nan_samples = h[ST_NAN];
		ctx->stats.trace_warp_iters = h[ST_W_ITERS], ctx->stats.trace_node_trips = h[ST_W_ITERS];
		ctx->stats.trace_tri_rounds = h[ST_W_ROUNDS], ctx->stats.trace_refills = h[ST_W_REFILLS];
#ifdef LMB_TRACE_PROFILE
		fprintf(stderr, "k_trace profile: iters %llu node_trips %llu node_lanes %llu has_lanes %llu parked_lanes %llu rounds %llu pairs %llu refills %llu\n",
				h[ST_P_ITERS], h[ST_P_NODE_TRIPS], h[ST_P_NODE_LANES], h[ST_P_HAS_LANES], h[ST_P_PARKED_LANES], h[ST_P_ROUNDS], h[ST_P_PAIRS], h[ST_P_REFILLS]);
#endif
	}
	*out = ctx->stats;
	return LMB_OK;
}

int lmb_reset_stats(lmb_ctx* ctx) {
	if (!ctx) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	const lmb_stats keep = ctx->stats;
	ctx->stats = lmb_stats{};
	ctx->stats.ms_build_accel = keep.ms_build_accel, ctx->stats.ms_build_morton = keep.ms_build_morton;
	ctx->stats.ms_build_sort = keep.ms_build_sort, ctx->stats.ms_build_tree = keep.ms_build_tree;
	ctx->stats.ms_build_refit = keep.ms_build_refit, ctx->stats.ms_build_wide = keep.ms_build_wide;
	ctx->stats.wide_nodes = keep.wide_nodes, ctx->stats.wide_levels = keep.wide_levels;
	ctx->stats.ms_build_ploc = keep.ms_build_ploc, ctx->stats.ploc_iterations = keep.ploc_iterations;
	if (ctx->wf.stats) {
		LMB_CUDA(ctx, cudaMemsetAsync(ctx->wf.stats, 0, ST_COUNT * 8, ctx->stream));
		LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	}
	return LMB_OK;
}

static int ensure_stats(lmb_ctx* ctx) {
	if (ctx->wf.stats) return 0;
	LMB_CUDA(ctx, cudaMalloc((void**)&ctx->wf.stats, ST_COUNT * 8));
	LMB_CUDA(ctx, cudaMemsetAsync(ctx->wf.stats, 0, ST_COUNT * 8, ctx->stream));
	return 0;
}

int lmb_trace_closest(lmb_ctx* ctx, const float* rays, uint32_t n, lmb_hit* hits) {
	if (!ctx || !rays || !hits) return LMB_ERR_INVALID;
	if (!ctx->bvh.built) return set_error(ctx, LMB_ERR_INVALID, "lmb_trace_closest: call lmb_build_accel first");
	cudaSetDevice(ctx->device);
	if (n == 0) return LMB_OK;
	int rc = ensure_stats(ctx);
	if (rc) return rc;
	float4 *d_rays = nullptr, *d_hits = nullptr;
	rc = check_cuda(ctx, cudaMalloc((void**)&d_rays, (size_t)n * 32), "lmb_trace_closest: rays");
	if (!rc) rc = check_cuda(ctx, cudaMalloc((void**)&d_hits, (size_t)n * 16), "lmb_trace_closest: hits");
	if (!rc) rc = check_cuda(ctx, cudaMemcpyAsync(d_rays, rays, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream), "lmb_trace_closest: copy rays");
	if (!rc) rc = launch_trace_closest(ctx, d_rays, n, d_hits);
	if (!rc) rc = check_cuda(ctx, cudaMemcpyAsync(hits, d_hits, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream), "copy hits");
	if (!rc) rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "lmb_trace_closest");
	cudaFree(d_rays), cudaFree(d_hits);
	return rc;
}

int lmb_trace_any(lmb_ctx* ctx, const float* rays, uint32_t n, uint8_t* occluded) {
	if (!ctx || !rays || !occluded) return LMB_ERR_INVALID;
	if (!ctx->bvh.built) return set_error(ctx, LMB_ERR_INVALID, "lmb_trace_any: call lmb_build_accel first");
	cudaSetDevice(ctx->device);
	if (n == 0) return LMB_OK;
	int rc = ensure_stats(ctx);
	if (rc) return rc;
	float4* d_rays = nullptr;
	uint8_t* d_occ = nullptr;
	rc = check_cuda(ctx, cudaMalloc((void**)&d_rays, (size_t)n * 32), "lmb_trace_any: rays");
	if (!rc) rc = check_cuda(ctx, cudaMalloc((void**)&d_occ, n), "lmb_trace_any: occlusion bytes");
	if (!rc) rc = check_cuda(ctx, cudaMemcpyAsync(d_rays, rays, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream), "lmb_trace_any: copy rays");
	if (!rc) rc = launch_trace_any(ctx, d_rays, n, d_occ);
	if (!rc) rc = check_cuda(ctx, cudaMemcpyAsync(occluded, d_occ, n, cudaMemcpyDeviceToHost, ctx->stream), "copy occ");
	if (!rc) rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "lmb_trace_any");
	cudaFree(d_rays), cudaFree(d_occ);
	return rc;
}

int lmb_trace_closest_device(lmb_ctx* ctx, const void* d_rays, uint32_t n, void* d_hits, uint32_t repeat, float* ms_out) {
	return lmb_trace_closest_device_ex(ctx, d_rays, n, d_hits, repeat, 0, ms_out);
}

int lmb_trace_closest_device_ex(lmb_ctx* ctx, const void* d_rays, uint32_t n, void* d_hits, uint32_t repeat, int sort_rays_first, float* ms_out) {
	if (!ctx || !d_rays || !d_hits) return LMB_ERR_INVALID;
	if (!ctx->bvh.built) return set_error(ctx, LMB_ERR_INVALID, "lmb_trace_closest_device: call lmb_build_accel first");
	cudaSetDevice(ctx->device);
	int rc = ensure_stats(ctx);
	if (rc) return rc;
	if (repeat == 0) repeat = 1;
	if (sort_rays_first && n > 0) {  // scratch allocation outside the timed region
		const uint32_t* warm = nullptr;
		if ((rc = sort_rays(ctx, (const float4*)d_rays, n, &warm))) return rc;
		LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	}
	cudaEventRecord(ctx->ev[6], ctx->stream);
	for (uint32_t r = 0; r < repeat && !rc; r++) {
		const uint32_t* order = nullptr;
		if (sort_rays_first && n > 0) rc = sort_rays(ctx, (const float4*)d_rays, n, &order);  // the sort is part of every timed launch
		if (!rc) rc = launch_trace_closest(ctx, (const float4*)d_rays, n, (float4*)d_hits, order);
	}
	cudaEventRecord(ctx->ev[7], ctx->stream);
	if (rc) return rc;
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (ms_out) cudaEventElapsedTime(ms_out, ctx->ev[6], ctx->ev[7]);
	return LMB_OK;
}

int lmb_accel_num_tris(lmb_ctx* ctx, uint32_t* n) {
	if (!ctx || !n || !ctx->bvh.built) return LMB_ERR_INVALID;
	*n = ctx->bvh.n;
	return LMB_OK;
}

int lmb_accel_download(lmb_ctx* ctx, uint32_t* left, uint32_t* right, uint32_t* parent, uint32_t* leaf_prim, uint32_t* morton, uint64_t* keys,
					   float* aabb) {
	if (!ctx || !ctx->bvh.built) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	const DeviceBvh& b = ctx->bvh;
	const size_t n = b.n;
	if (n == 0) return LMB_OK;
	auto get = [&](void* dst, const void* src, size_t bytes) -> int {
		if (!dst || bytes == 0) return 0;
		return check_cuda(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream), "lmb_accel_download");
	};
	int rc;
	if ((rc = get(left, b.left, (n - 1) * 4))) return rc;
	if ((rc = get(right, b.right, (n - 1) * 4))) return rc;
	if ((rc = get(parent, b.parent, (2 * n - 1) * 4))) return rc;
	if ((rc = get(leaf_prim, b.leaf_prim, n * 4))) return rc;
	if ((rc = get(morton, b.morton, n * 4))) return rc;
	if ((rc = get(keys, b.keys, n * 8))) return rc;
	if ((rc = get(aabb, b.aabb, 6 * (2 * n - 1) * 4))) return rc;
	LMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LMB_OK;
}

}  // extern "C"
