// context.h -- internal state of an lmb_ctx (one GPU, one stream).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "lmb_types.h"
#include "lumen_b200.h"

namespace lmb {

// Scene arrays as the kernels see them (device pointers).
struct DeviceScene {
	const lmb_vertex* vertices;
	const uint32_t* indices;
	const lmb_material* materials;
	const lmb_prim_mesh_info* prim_infos;
	const float* world_matrices;
	const float* inv_world_matrices;
	const lmb_light* lights;
	const uint32_t* tri_mesh;   // global triangle id -> prim mesh index
	const uint32_t* tri_local;  // global triangle id -> mesh-local triangle number
	const uint4* tri_rec;       // global triangle id -> (vertex index of corner 0, 1, 2, prim mesh index)
	const float4* tri_shade;    // global triangle id -> 8 float4 (one 128-byte line): corner positions / normals / uvs, mesh, material (build_hit)
	const uint8_t* tri_matq;    // global triangle id -> shade queue of its material's BSDF type (0..5 = log2(bsdf_type), 6 = unknown)
	const uint8_t* const* tex_data;  // per texture: RGBA8 texels
	const uint2* tex_dims;
	const float* srgb_lut;  // 256 entries
	uint32_t n_tris, n_prim_meshes, n_lights, n_textures;
};

struct DeviceBvh {
	// canonical LBVH (parity surface)
	uint32_t* morton = nullptr;    // n, by global id
	uint64_t* keys = nullptr;      // n, sorted
	uint64_t* keys_tmp = nullptr;  // n, radix ping-pong
	uint32_t* leaf_prim = nullptr; // n
	uint32_t* left = nullptr;      // n-1
	uint32_t* right = nullptr;     // n-1
	uint32_t* parent = nullptr;    // 2n-1
	float* aabb = nullptr;         // 6*(2n-1)
	uint32_t* arrive = nullptr;    // n-1 refit counters
	uint32_t* bounds_enc = nullptr;  // 6 ordered-uint encoded scene bounds
	float4* tri_world = nullptr;   // 3n, by global id
	// traversal layout
	float4* nodes = nullptr;  // 4*(n-1)
	float4* tris = nullptr;   // 3n, leaf order
	uint32_t* radix_hist = nullptr;
	uint32_t* span_count = nullptr;  // n-1: leaves under each internal node (size of its Karras range)
	// quality tree for traversal (ploc.cu), same numbering as left/right/aabb; null when LMB_TREE=lbvh
	uint32_t* q_left = nullptr;
	uint32_t* q_right = nullptr;
	uint32_t* q_count = nullptr;  // 2n-1
	float* q_aabb = nullptr;      // 6*(2n-1)
	uint32_t n = 0;
	bool built = false;
};

// Compressed 8-wide BVH derived from the canonical LBVH (wide_bvh.cu); what k_trace walks.
struct DeviceWideBvh {
	float4* nodes = nullptr;  // 5 per node (80 B)
	float4* tris = nullptr;   // 3 per triangle, contiguous per node
	uint2* items[2] = {nullptr, nullptr};  // build scratch
	uint32_t* counters = nullptr;
	uint32_t n_nodes = 0, n_tris = 0, levels = 0;
};

// Wavefront state (wavefront.cu). Path planes are indexed by POSITION in the bounce's dense path list and double-buffered by
// bounce parity; acc / film are indexed by pixel slot (slot = frame_in_batch * W*H + y*W + x).
struct Wavefront {
	uint32_t n_slots = 0;
	float4* ray_o[2] = {nullptr, nullptr};  // origin.xyz, tmin
	float4* ray_d[2] = {nullptr, nullptr};  // direction.xyz, tmax
	float4* thr[2] = {nullptr, nullptr};    // throughput.xyz, rng counter (bits)
	float4* col[2] = {nullptr, nullptr};    // radiance.xyz, flags (bits): bit0 last_specular, bit1 ended (waits for its light sample)
	uint32_t* pix[2] = {nullptr, nullptr};  // pixel slot of the path
	float4* hit = nullptr;       // t, b1, b2, prim (bits) of the continuation ray, by path position
	float4* acc = nullptr;       // radiance of finished paths, by pixel slot (written exactly once per path)
	float4* nee = nullptr;       // 8 float4 planes, by position in the light-sample list
	uint32_t* nee_path = nullptr;  // light sample -> position of its path in the next list
	float4* probe_hit = nullptr;   // MIS-probe closest hit per light sample
	uint32_t* shadow_occ = nullptr;  // shadow-ray result per light sample
	float4* miss_ray_o = nullptr;  // escaped rays awaiting the sky march (k_miss): dense records
	float4* miss_ray_d = nullptr;
	float4* miss_thr = nullptr;
	float4* miss_col = nullptr;
	uint32_t* miss_pix = nullptr;
	uint32_t* mat_queues = nullptr;  // 7 x n_slots: path positions sorted by the BSDF type they hit (k_classify -> k_shade<TYPE>)
	uint32_t* trace_queue = nullptr; // typed ray entries for k_trace: index | type << 30 (up to 3 per path)
	uint32_t* trace_cursor = nullptr;  // work-fetch cursor of the array ray queries
	uint32_t* counters = nullptr;  // see enum Counter in wavefront.cu
	uint32_t* miss_marks = nullptr;  // sky march ranges (k_miss): [0..64] miss-record count after bounce d - 1's k_classify, [65..128] work cursors
	unsigned long long* stats = nullptr;  // device counters, see StatSlot
	uint32_t frames_in_flight = 0;
};

// BDPT state (bdpt.cu): the two sub-paths of every pixel (struct of arrays over pixels, 23 words per vertex), the radiance of the
// pixel's own strategies and the light-tracer splat image of the frame in flight. Allocated by the first lmb_render_bdpt.
struct BdptState {
	float* light_verts = nullptr;
	float* camera_verts = nullptr;
	float4* col = nullptr;
	float* splat = nullptr;
	// staged pipeline: walk state and per-pixel scalars (struct of arrays over pixels), ray slots and their results
	float* walk = nullptr;
	uint32_t* misc = nullptr;
	float4* rays = nullptr;   // 2 float4 per slot: walk rays use n_pix slots, connection rays n_conn_slots * n_pix
	float4* hits = nullptr;   // n_pix
	uint8_t* occ = nullptr;   // n_conn_slots * n_pix
	float4* contrib = nullptr;  // n_conn_slots * n_pix: weighted radiance of each pair (pair-parallel connections)
	uint8_t* pair_ts = nullptr; // (t, s) of each connection slot
	uint32_t* work_list = nullptr;   // n_conn_slots * n_pix: the (slot, pixel) entries that have work in the pass at hand (k_bdpt_worklist)
	uint32_t* work_count = nullptr;  // [0] emit pass, [1] resolve pass, [2] emitted connection rays
	uint32_t* alive_list[2] = {nullptr, nullptr};  // n_pix each: the pixels whose walk is in flight (ping-pong)
	uint32_t* alive_count = nullptr;               // 2 * max_depth + 2 list lengths of one frame
	uint32_t* emit_list = nullptr;   // the (slot, pixel) entries whose shadow ray the emit pass wrote: what the any-hit launch traces
	uint32_t n_pix = 0, n_verts = 0, n_conn_slots = 0;
};

enum StatSlot {
	ST_CLOSEST = 0, ST_SHADOW, ST_PROBE, ST_NODES, ST_TRIS, ST_NAN,
	// warp-level scheduling counters of the wide walker, always on (lmb_stats.trace_*)
	ST_W_ITERS, ST_W_NODE_TRIPS, ST_W_ROUNDS, ST_W_REFILLS,
	// k_trace scheduling counters, filled only by a -DLMB_TRACE_PROFILE build (tools/gpu_variants.sh): warp trips of the inner
	// loop, trips with a node step, lanes stepping, lanes owning a ray, lanes parked on triangles, rounds, pairs, refills
	ST_P_ITERS, ST_P_NODE_TRIPS, ST_P_NODE_LANES, ST_P_HAS_LANES, ST_P_PARKED_LANES, ST_P_ROUNDS, ST_P_PAIRS, ST_P_REFILLS,
	ST_COUNT
};

}  // namespace lmb

struct lmb_ctx {
	int device = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	std::string err;
	// scene
	bool scene_loaded = false;
	uint32_t mat_queue_mask = 0;  // bit m: some material maps to shade queue m (0..5 = log2(bsdf_type), 6 = unknown type)
	lmb::DeviceScene scene{};
	std::vector<void*> scene_allocs;
	std::vector<lmb_prim_mesh_info> h_prim_infos;
	std::vector<uint32_t> h_idx_counts;
	std::vector<uint32_t> h_light_flags;  // host copy of Light.light_flags: lmb_render checks PCPath's light indices against it
	lmb::DeviceBvh bvh;
	lmb::DeviceWideBvh wide;
	bool use_ploc = true;   // LMB_TREE=lbvh: collapse the canonical Karras tree instead of the PLOC tree
	int trace_pin = -1;     // LMB_TRACE_PIN=0|1 forces the unpinned / pinned instantiation of the wide walker (-1: by BVH footprint)
	bool tree_auto = true;  // LMB_TREE unset: build both binary trees, walk the one with the lower surface-area cost
	bool use_bvh2 = false;  // LMB_TRAVERSAL=bvh2: walk the binary LBVH instead of the 8-wide BVH (A/B measurements)
	// film / wavefront
	uint32_t width = 0, height = 0;
	uint32_t row_first = 0, row_stride = 1;  // pixel shard: this context renders image rows row_first + k * row_stride
	float4* film = nullptr;
	lmb::Wavefront wf;
	lmb::BdptState bdpt;
	// post steps (post.cu)
	float4* film_snapshot = nullptr;  // lmb_download_async: copy of the film the copy stream sends home while rendering goes on
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_snapshot = nullptr, ev_copied = nullptr;
	bool copy_pending = false;
	uint16_t* half_planes = nullptr;  // 3 x W*H halves: B, G, R planes of the EXR writer
	float4* gt_img = nullptr;         // ground-truth image of the RMSE routine ("gt_img_addr")
	void* rmse_scratch = nullptr;
	bool has_gt = false;
	// ray ordering scratch of the array queries (lbvh.cu sort_rays)
	uint32_t *ray_keys = nullptr, *ray_order = nullptr, *ray_hist = nullptr;
	uint32_t ray_sort_cap = 0;
	// multi-GPU exchange (comm.cu): NCCL communicator (ncclComm_t) of this rank, its own high-priority stream, the film snapshot the
	// overlapped reduce works on
	void* comm = nullptr;
	int comm_rank = 0, comm_size = 1;
	cudaStream_t comm_stream = nullptr;
	cudaEvent_t ev_comm_ready = nullptr, ev_comm_done = nullptr;
	float4* reduce_buf = nullptr;
	bool reduce_pending = false;
	// stats
	lmb_stats stats{};
	bool profile_stages = false;
	bool stats_per_launch = false;  // LMB_STATS_PER_LAUNCH=1: print k_trace's counters after every launch (calibration runs only)
	cudaEvent_t ev[8]{};
	// the sky march of escaped rays runs beside the bounce loop on its own stream (wavefront.cu k_miss)
	cudaStream_t miss_stream = nullptr;
	cudaEvent_t ev_miss_ready = nullptr, ev_miss_done = nullptr;
};

namespace lmb {
int set_error(lmb_ctx* ctx, int code, const std::string& msg);
int check_cuda(lmb_ctx* ctx, cudaError_t e, const char* what);
int build_lbvh(lmb_ctx* ctx);
void free_bvh(lmb_ctx* ctx);
int build_wide_bvh(lmb_ctx* ctx);
void free_wide_bvh(lmb_ctx* ctx);
int build_ploc(lmb_ctx* ctx);
void free_ploc(lmb_ctx* ctx);
int probe_wide_tree(lmb_ctx* ctx, uint32_t n_rays, double* steps_per_ray);
int wavefront_alloc(lmb_ctx* ctx, uint32_t frames_in_flight);
uint32_t shard_rows(const lmb_ctx* ctx);
void wavefront_free(lmb_ctx* ctx);
int wavefront_render(lmb_ctx* ctx, const lmb_pc_path& pc, const lmb_scene_ubo& ubo, uint32_t first_frame, uint32_t n_frames, uint32_t stride,
					 int film_mode);
int bdpt_render(lmb_ctx* ctx, const lmb_pc_bdpt& pc, const lmb_scene_ubo& ubo, uint32_t first_frame, uint32_t n_frames, uint32_t frame_stride, int film_mode,
				float* raw_col, float* raw_splat);
void bdpt_free(lmb_ctx* ctx);
int launch_trace_closest(lmb_ctx* ctx, const float4* d_rays, uint32_t n, float4* d_hits, const uint32_t* order = nullptr);
int sort_rays(lmb_ctx* ctx, const float4* d_rays, uint32_t n, const uint32_t** order_out);
int launch_trace_any(lmb_ctx* ctx, const float4* d_rays, uint32_t n, uint8_t* d_occ);
// ray slots with dead entries (NaN origin = hits nothing): closest hits into d_hits or occlusion bytes into d_occ; rays are NOT counted
int launch_trace_slots(lmb_ctx* ctx, const float4* d_rays, uint32_t n, float4* d_hits, uint8_t* d_occ, bool any);
int launch_trace_slot_list(lmb_ctx* ctx, const float4* d_rays, const uint32_t* list, const uint32_t* n_dev, uint32_t n_max, float4* d_hits, uint8_t* d_occ, bool any);
int launch_resolve(lmb_ctx* ctx);
int launch_resolve_on(lmb_ctx* ctx, float4* film, cudaStream_t stream);  // k_resolve on any RGBA32F sum image of the film's size
void comm_free(lmb_ctx* ctx);
int ingest_triangles(lmb_ctx* ctx, const uint32_t* d_tri_first, const uint8_t* d_mat_q, uint32_t n_meshes, uint32_t n_materials, uint32_t n_tris,
					 uint32_t* tri_mesh, uint32_t* tri_local, uint4* tri_rec, uint8_t* tri_matq, float4* tri_shade);
int launch_film_to_half(lmb_ctx* ctx, uint16_t* d_planes);
int launch_film_add(lmb_ctx* ctx, const float4* d_other);
size_t rmse_scratch_bytes(uint32_t n_pix);
int launch_rmse(lmb_ctx* ctx, const float4* d_gt, void* d_scratch, float* h_literal, double* h_true);
}  // namespace lmb

#define LMB_CUDA(ctx, call)                                            \
	do {                                                               \
		const int _rc = lmb::check_cuda((ctx), (call), #call);         \
		if (_rc != 0) return _rc;                                      \
	} while (0)
