/*
 * lumen_b200.h -- C ABI of the B200-native Path integrator (liblumen_b200.so, CUDA sm_100a).
 *
 * This is the drop-in boundary for Lumen's `Path` integrator. Reference interface being replaced:
 *   class Integrator { init(); render(); update(); destroy(); create_accel(); output_tex; frame_num; }
 *       src/RayTracer/Integrator.h:14-33, src/RayTracer/Path.h:4-23, src/RayTracer/Path.cpp:4-70
 *   scene upload                     src/RayTracer/LumenScene.cpp:134-216
 *   acceleration-structure build     src/RayTracer/Integrator.cpp:137-160, src/Framework/AccelerationStructure.cpp:171-315
 *   per-frame dispatch               src/RayTracer/Path.cpp:27-59 -> vkCmdTraceRaysKHR(W, H, 1)
 * Mapping (see INTEGRATION.md for the C++ shim `class PathB200 final : public Integrator`):
 *   RayTracer::init                  -> lmb_create, lmb_upload_scene
 *   Integrator::create_accel         -> lmb_build_accel
 *   Path::init / Integrator::init    -> lmb_init(width, height)
 *   Path::render (one frame)         -> lmb_render(pc, ubo, frame_num, 1)   (batched: n_frames > 1)
 *   output_tex readback (F10 -> EXR) -> lmb_download
 *   Path::destroy / cleanup          -> lmb_destroy
 * Plain pointers and sizes only; no C++ or torch types cross this boundary. Every call returns 0 on success or a
 * negative lmb_status; nothing throws. Calls on one context must be serialised by the caller; one context drives one
 * GPU (one process per GPU; see DESIGN.md "Multi-GPU").
 */
#ifndef LUMEN_B200_H
#define LUMEN_B200_H

#include <stdint.h>
#include "lmb_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lmb_ctx lmb_ctx;

enum lmb_status {
	LMB_OK = 0,
	LMB_ERR_INVALID = -1,   /* bad argument or call order */
	LMB_ERR_CUDA = -2,      /* CUDA runtime error, see lmb_last_error */
	LMB_ERR_NO_DEVICE = -3, /* no usable CUDA device: there is no CPU fallback */
	LMB_ERR_OOM = -4
};

/* Film update rule for lmb_render. */
enum lmb_film_mode {
	/* path.rgen:102-112: running mean in frame order, NaN samples skipped, alpha = 1. Requires frame_stride == 1. */
	LMB_FILM_RUNNING_MEAN = 0,
	/* Sharded rendering: rgb += sample, alpha += 1 per non-NaN sample. After the cross-GPU sum, lmb_resolve divides.  */
	LMB_FILM_SUM = 1
};

typedef struct lmb_stats {
	uint64_t rays_closest; /* continuation rays (path.rgen:48) */
	uint64_t rays_shadow;  /* any-hit rays (pt_commons.glsl:21) */
	uint64_t rays_probe;   /* MIS BSDF-probe rays (pt_commons.glsl:32) */
	uint64_t nodes_visited; /* traversal steps: 80-byte 8-wide nodes (LMB_TRAVERSAL=bvh2: 64-byte binary nodes) */
	uint64_t tris_tested;
	uint64_t nan_samples;  /* samples dropped by the NaN guard */
	uint64_t frames;       /* frames rendered since lmb_init / lmb_reset_stats */
	uint64_t kernel_launches; /* CUDA kernels launched by lmb_render since lmb_init / lmb_reset_stats */
	float ms_render;       /* device time of all lmb_render calls (CUDA events on the context's stream) */
	float ms_extend;       /* k_trace: BVH traversal of continuation + shadow + MIS-probe rays (profile_stages only) */
	float ms_shade;        /* k_shade: hit record, NEE generation, BSDF sampling */
	float ms_connect;      /* k_connect: MIS weights + radiance accumulation */
	float ms_film;         /* ray generation + sky (k_miss) + film kernels */
	float ms_build_accel;  /* last lmb_build_accel: total */
	float ms_build_morton; /* flatten + bounds + Morton codes */
	float ms_build_sort;   /* radix sort */
	float ms_build_tree;   /* Karras hierarchy */
	float ms_build_refit;  /* bottom-up AABB refit + node packing */
	float ms_build_wide;   /* collapse of the LBVH into the compressed 8-wide traversal BVH (included in ms_build_accel) */
	uint32_t wide_nodes;   /* nodes of the 8-wide BVH (80 B each) */
	uint32_t wide_levels;  /* depth of the 8-wide BVH */
	uint32_t ploc_iterations; /* clustering iterations of the traversal tree (0: the Karras tree is walked) */
	float ms_build_ploc;   /* PLOC clustering over the sorted leaves (included in ms_build_accel) */
	float tree_cost_ratio; /* probe-ray traversal steps of the clustered 8-wide tree / of the Karras one (the cheaper is walked; 0: not compared) */
	/* k_trace's own scheduling counters (warp level, always on, one integer add each): trips of the traversal loop (every trip steps a
	 * node in some lane: measured), warp-cooperative triangle rounds, 32-ray refills. With the per-counter instruction costs fitted on
	 * the committed ncu launch list (profiles/ktrace_calibration.json) they give the warp instructions the kernel issued -- the
	 * numerator of bench.py's issue roofline. trace_node_trips is kept for layout stability and equals trace_warp_iters. */
	uint64_t trace_warp_iters, trace_node_trips, trace_tri_rounds, trace_refills;
} lmb_stats;

typedef struct lmb_hit {
	float t, b1, b2;
	uint32_t prim; /* global triangle id (prim meshes concatenated in order); 0xFFFFFFFF = miss */
} lmb_hit;

/* Creates a context on CUDA device `device_id`. Fails with LMB_ERR_NO_DEVICE when there is none. */
int lmb_create(lmb_ctx** out, int device_id);
void lmb_destroy(lmb_ctx* ctx);
const char* lmb_last_error(const lmb_ctx* ctx); /* ctx may be NULL: error of the last failed lmb_create */

/* Copies the scene arrays to the device (LumenScene.cpp:134-216). Invalidates any previous accel. */
int lmb_upload_scene(lmb_ctx* ctx, const lmb_scene_desc* scene);
/* Builds the world-space LBVH on the GPU: Morton codes -> radix sort -> Karras hierarchy -> bottom-up refit. */
int lmb_build_accel(lmb_ctx* ctx);
/* Allocates the RGBA32F film ("output_tex", Integrator.cpp:33-40) and the wavefront state for width x height.
 * `frames_in_flight` = how many frames (samples per pixel) one wavefront batch carries; 0 picks a default. */
int lmb_init(lmb_ctx* ctx, uint32_t width, uint32_t height, uint32_t frames_in_flight);

/* Pixel sharding (secondary axis; BASELINE config 4 "tile + sample sharding"): this context renders only the
 * image rows row_first, row_first + row_stride, ... (full-width tiles one row high, interleaved over the shards, so every
 * shard sees the same mix of sky, walls and floor). Call BEFORE lmb_init: the wavefront state is sized for the shard's
 * pixels (calling it afterwards with different values drops film + wavefront; run lmb_init again). The film keeps the full
 * width x height; rows of other shards are never touched, so an LMB_FILM_SUM film of the shards adds up to the whole image
 * with one all-reduce. RNG seeds stay (x, y, frame, 0) of the FULL image (path.rgen:23): a sharded render is bit-identical
 * to the same rows of an unsharded one. Default (0, 1) = every row. Combine with frame_stride for tile x sample sharding. */
int lmb_set_pixel_shard(lmb_ctx* ctx, uint32_t row_first, uint32_t row_stride);

/* Renders frames first_frame, first_frame + frame_stride, ... (n_frames of them) and updates the film.
 * pc->frame_num is ignored; RNG seed of a sample is (x, y, frame, 0) exactly as path.rgen:23. Synchronous. */
int lmb_render(lmb_ctx* ctx, const lmb_pc_path* pc, const lmb_scene_ubo* ubo, uint32_t first_frame, uint32_t n_frames,
			   uint32_t frame_stride, int film_mode);
/* Lumen's BDPT integrator on the same scene, accel and film. Replaces BDPT::render
 * (src/RayTracer/BDPT.cpp:55-95) = one dispatch of src/shaders/integrators/bdpt/bdpt.rgen per frame: renders frames
 * first_frame .. first_frame + n_frames - 1 into the running-mean film (bdpt.rgen:79-89). pc->frame_num is ignored; the RNG
 * seed of a sample is (x, y, frame ^ pc->time, 0) as bdpt.rgen:36-37 -- `time` is the caller's (BDPT.cpp:57 draws rand() per
 * frame; pass a fresh value per call to do the same). Light-tracer splats of a frame land in that frame. Vertex storage
 * (2 x (max_depth + 1) x 92 B per pixel) is allocated on first use. frame_stride / film_mode as lmb_render: LMB_FILM_SUM with
 * frames first_frame + k * frame_stride is the sample-index shard of one GPU; the caller adds the films and calls
 * lmb_resolve. Not available with pixel shards (splats cross rows). Synchronous. */
int lmb_render_bdpt(lmb_ctx* ctx, const lmb_pc_bdpt* pc, const lmb_scene_ubo* ubo, uint32_t first_frame, uint32_t n_frames,
					uint32_t frame_stride, int film_mode);
/* Zeroes the film (Path::update resets frame_num to 0 on camera change; sum mode needs an explicit clear). */
int lmb_clear_film(lmb_ctx* ctx);
/* LMB_FILM_SUM epilogue: rgb /= alpha (pixels with alpha 0 stay 0), alpha = 1. */
int lmb_resolve(lmb_ctx* ctx);
/* Copies the film to `rgba` (host or device pointer, resolved by UVA): width*height*4 floats, row-major, pixel (x, y) at
 * 4*(y*width + x). */
int lmb_download(lmb_ctx* ctx, float* rgba);
/* lmb_download without the wait: the film is snapshotted in stream order and sent to `rgba` by a second stream, so the
 * transfer overlaps the next lmb_render (pass pinned host memory for a truly asynchronous transfer). `rgba` holds the
 * film as of this call once lmb_sync returns; lmb_sync waits for everything queued so far, transfers included. */
int lmb_download_async(lmb_ctx* ctx, float* rgba);
int lmb_sync(lmb_ctx* ctx);
/* EXR payload made on the device: the film's R, G, B converted to HALF with tinyexr's float_to_half_full rounding
 * (libs/tinyexr.h:898-934) and laid out as the three planes ImageUtils::save_exr writes (src/Framework/ImageUtils.cpp:22-89):
 * planes[0 .. n) = B, [n .. 2n) = G, [2n .. 3n) = R, n = width*height. 6 bytes per pixel leave the device instead of 16. */
int lmb_download_half_bgr(lmb_ctx* ctx, uint16_t* planes);
/* Ground-truth image of the RMSE routine (RayTracer.cpp:117-126 load_reference / has_gt): width*height*4 floats, host or
 * device pointer. Dropped by lmb_init. */
int lmb_set_reference_image(lmb_ctx* ctx, const float* gt_rgba);
/* RMSE of the film against the ground-truth image, on the device. *rmse_literal = Lumen's own routine
 * (src/shaders/rmse/calc_rmse.comp + reduce_rmse.comp + output_rmse.comp as dispatched by RayTracer.cpp:215-241, with its
 * subgroupMin and sqrt(S)/(3N) quirks); *rmse_true = sqrt(mean over RGB of (film - gt)^2) in fp64. Either may be NULL. */
int lmb_rmse(lmb_ctx* ctx, float* rmse_literal, double* rmse_true);
/* Copies an image (host or device pointer) into the film: resume an accumulation, or write back a reduced sum. */
int lmb_upload_film(lmb_ctx* ctx, const float* rgba);
/* dst.film += src.film, element-wise: the reduce of the LMB_FILM_SUM films of a multi-GPU render (same image size; the two
 * contexts may sit on different devices -- the copy goes device to device, over NVLink where the peers are connected).
 * Synchronous on return. No reference equivalent (Lumen drives one GPU). */
int lmb_film_add_from(lmb_ctx* dst, lmb_ctx* src);
/* ---- the exchange step of a multi-GPU render (SURVEY.md 8e / 2.3 "Collectives": no reference equivalent, Lumen drives one GPU).
 * One rank = one context = one GPU. The only message is the fp32 sum of the LMB_FILM_SUM films (RGB sums + valid-sample count in
 * alpha), reduced by ncclAllReduce over NVLink / NVSwitch with the "/ count" epilogue of lmb_resolve queued behind it on the same
 * stream. NCCL is bound at run time (libnccl.so.2); without it these calls fail with a message and everything else works. */
#define LMB_COMM_ID_BYTES 128
/* rank 0 makes the id (ncclGetUniqueId) and hands the 128 bytes to the other ranks by whatever means the host has (MPI, a file,
 * torch.distributed ...); every rank then calls lmb_comm_init -- collectively, it returns once all n_ranks have joined. */
int lmb_comm_get_unique_id(uint8_t* id128);
int lmb_comm_init(lmb_ctx* ctx, const uint8_t* id128, int rank, int n_ranks);
/* the same for n contexts of ONE process (one per device): rank i = ctxs[i] */
int lmb_comm_init_all(lmb_ctx** ctxs, int n);
int lmb_comm_destroy(lmb_ctx* ctx);
int lmb_comm_info(lmb_ctx* ctx, int* rank, int* n_ranks, int* nccl_version); /* n_ranks = 0: no communicator */
/* film -> sum over the ranks -> rgb / alpha, alpha = 1 (every rank gets the whole image). Collective; nothing waits on the host
 * (lmb_sync does). out_rgba == NULL: in place on the render stream, the film IS the resolved image afterwards. out_rgba != NULL
 * (host, pinned for a truly asynchronous copy, or device; width*height*4 floats): the film is snapshotted in stream order and, with
 * clear_film != 0, zeroed for the next batch; reduce, resolve and the copy to out_rgba run on the context's own high-priority
 * stream while lmb_render goes on with the next frames. */
int lmb_film_allreduce(lmb_ctx* ctx, float* out_rgba, int clear_film);
/* The same with ONE receiver (ncclReduce): rank `root` ends with the resolved image of all ranks (in its film, or in out_rgba with
 * the snapshot form); the other ranks only contribute -- out_rgba is ignored there and may be NULL, clear_film alone selects the
 * snapshot form. A progressive render shows one image: N - 1 ranks need neither the sum nor the copy home. */
int lmb_film_reduce(lmb_ctx* ctx, int root, float* out_rgba, int clear_film);

/* Device pointer of the film and the CUDA stream (cudaStream_t) the context works on, for zero-copy consumers
 * (e.g. an NCCL all-reduce issued by the caller). */
int lmb_film_device_ptr(lmb_ctx* ctx, void** dptr, uint64_t* n_floats);
int lmb_stream(lmb_ctx* ctx, void** cuda_stream);

/* on != 0: time extend / shade / connect / film separately (adds a host sync per bounce; off by default). */
int lmb_set_profile_stages(lmb_ctx* ctx, int on);
int lmb_get_stats(lmb_ctx* ctx, lmb_stats* out);
int lmb_reset_stats(lmb_ctx* ctx);

/* Raw ray queries against the built accel; rays = n x 8 floats (ox, oy, oz, tmin, dx, dy, dz, tmax), HOST pointers. */
int lmb_trace_closest(lmb_ctx* ctx, const float* rays, uint32_t n, lmb_hit* hits);
int lmb_trace_any(lmb_ctx* ctx, const float* rays, uint32_t n, uint8_t* occluded);
/* Same on DEVICE pointers, timed with CUDA events (ms_out may be NULL); `repeat` launches back to back. */
int lmb_trace_closest_device(lmb_ctx* ctx, const void* d_rays, uint32_t n, void* d_hits, uint32_t repeat, float* ms_out);
/* ... with sort_rays != 0 every launch first orders the rays by (origin cell, direction octant) -- 15-bit keys, a counting sort on the
 * device, timed with the launch -- so that the persistent walker's 32-ray fetches are coherent; hits land at the rays' own indices and
 * do not depend on the order. Pays when the BVH does not fit the L2 (BASELINE config 5). */
int lmb_trace_closest_device_ex(lmb_ctx* ctx, const void* d_rays, uint32_t n, void* d_hits, uint32_t repeat, int sort_rays, float* ms_out);

/* LBVH read-back for the bit-exact topology check. n = triangle count.
 * left/right: n-1, parent: 2n-1, leaf_prim / morton: n, keys: n (uint64), aabb: 6*(2n-1). NULL pointers are skipped. */
int lmb_accel_num_tris(lmb_ctx* ctx, uint32_t* n);
int lmb_accel_download(lmb_ctx* ctx, uint32_t* left, uint32_t* right, uint32_t* parent, uint32_t* leaf_prim, uint32_t* morton,
					   uint64_t* keys, float* aabb);

#ifdef __cplusplus
}
#endif
#endif /* LUMEN_B200_H */
