// comm.cu -- the one exchange step of a multi-GPU render: the fp32 sum-reduce of the LMB_FILM_SUM films over
// NCCL (NVLink 5 / NVSwitch), with the "divide by the valid-sample count" epilogue (k_resolve) queued right behind it on the same
// stream. One rank = one lmb_ctx = one GPU; ranks may be processes (lmb_comm_init with a shared unique id) or contexts of one
// process (lmb_comm_init_all). No reference equivalent: Lumen drives one GPU.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy already in the process if the host application loaded one): a
// single-GPU drop-in needs nothing but the CUDA runtime, and lmb_comm_* says so when the library is absent.
#include <dlfcn.h>
#include <nccl.h>  // types only; no symbol of libnccl is linked

#include <mutex>

#include "context.h"

namespace lmb {
namespace {
struct NcclApi {
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	ncclResult_t (*GetVersion)(int*) = nullptr;
	std::string error;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

const NcclApi& nccl() {
	std::call_once(g_nccl_once, [] {
		NcclApi& a = g_nccl;
		for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
			a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (a.handle) break;
		}
		if (!a.handle) {
			a.error = std::string("libnccl.so.2 not found (") + dlerror() + ")";
			return;
		}
		auto sym = [&](const char* n) {
			void* p = dlsym(a.handle, n);
			if (!p && a.error.empty()) a.error = std::string("libnccl lacks ") + n;
			return p;
		};
		a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
		a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
		a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
		a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
		a.Reduce = (decltype(a.Reduce))sym("ncclReduce");
		a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
		a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
		a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
		a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
	});
	return g_nccl;
}

int nccl_check(lmb_ctx* ctx, ncclResult_t r, const char* what) {
	if (r == ncclSuccess) return 0;
	return set_error(ctx, LMB_ERR_CUDA, std::string(what) + ": " + nccl().GetErrorString(r));
}
#define LMB_NCCL(ctx, call)                                   \
	do {                                                      \
		const int _rc = nccl_check((ctx), (call), #call);     \
		if (_rc != 0) return _rc;                             \
	} while (0)

int need_nccl(lmb_ctx* ctx, const char* who) {
	if (!nccl().error.empty()) return set_error(ctx, LMB_ERR_INVALID, std::string(who) + ": " + nccl().error);
	return 0;
}

int comm_streams(lmb_ctx* ctx) {
	if (ctx->comm_stream) return 0;
	int lo = 0, hi = 0;
	cudaDeviceGetStreamPriorityRange(&lo, &hi);
	// highest priority: the reduce is a few hundred microseconds of a few CTAs and should not queue behind the persistent
	// traversal blocks of the batch that renders meanwhile
	LMB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
	LMB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comm_ready, cudaEventDisableTiming));
	LMB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comm_done, cudaEventDisableTiming));
	return 0;
}
}  // namespace

void comm_free(lmb_ctx* ctx) {
	if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
	if (ctx->comm) nccl().CommDestroy((ncclComm_t)ctx->comm);
	ctx->comm = nullptr, ctx->comm_rank = 0, ctx->comm_size = 1;
	if (ctx->ev_comm_ready) cudaEventDestroy(ctx->ev_comm_ready);
	if (ctx->ev_comm_done) cudaEventDestroy(ctx->ev_comm_done);
	if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
	ctx->ev_comm_ready = ctx->ev_comm_done = nullptr, ctx->comm_stream = nullptr;
	cudaFree(ctx->reduce_buf);
	ctx->reduce_buf = nullptr;
	ctx->reduce_pending = false;
}
}  // namespace lmb

using namespace lmb;

extern "C" {

int lmb_comm_get_unique_id(uint8_t* id128) {
	if (!id128) return LMB_ERR_INVALID;
	if (const int bad = need_nccl(nullptr, "lmb_comm_get_unique_id")) return bad;
	static_assert(sizeof(ncclUniqueId) == LMB_COMM_ID_BYTES, "ncclUniqueId size");
	ncclUniqueId id;
	if (const int bad = nccl_check(nullptr, nccl().GetUniqueId(&id), "ncclGetUniqueId")) return bad;
	memcpy(id128, &id, sizeof(id));
	return LMB_OK;
}

int lmb_comm_init(lmb_ctx* ctx, const uint8_t* id128, int rank, int n_ranks) {
	if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return set_error(ctx, LMB_ERR_INVALID, "lmb_comm_init: bad arguments");
	if (const int bad = need_nccl(ctx, "lmb_comm_init")) return bad;
	if (ctx->comm) return set_error(ctx, LMB_ERR_INVALID, "lmb_comm_init: the context already has a communicator");
	cudaSetDevice(ctx->device);
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	ncclComm_t comm = nullptr;
	LMB_NCCL(ctx, nccl().CommInitRank(&comm, n_ranks, id, rank));
	ctx->comm = comm, ctx->comm_rank = rank, ctx->comm_size = n_ranks;
	return comm_streams(ctx);
}

int lmb_comm_init_all(lmb_ctx** ctxs, int n) {
	if (!ctxs || n < 1) return LMB_ERR_INVALID;
	for (int i = 0; i < n; i++) {
		if (!ctxs[i]) return LMB_ERR_INVALID;
		if (ctxs[i]->comm) return set_error(ctxs[i], LMB_ERR_INVALID, "lmb_comm_init_all: the context already has a communicator");
		for (int k = 0; k < i; k++)
			if (ctxs[k]->device == ctxs[i]->device) return set_error(ctxs[i], LMB_ERR_INVALID, "lmb_comm_init_all: two contexts on one device (NCCL wants one rank per GPU)");
	}
	if (const int bad = need_nccl(ctxs[0], "lmb_comm_init_all")) return bad;
	ncclUniqueId id;
	LMB_NCCL(ctxs[0], nccl().GetUniqueId(&id));
	std::vector<ncclComm_t> comms((size_t)n, nullptr);
	LMB_NCCL(ctxs[0], nccl().GroupStart());
	for (int i = 0; i < n; i++) {
		cudaSetDevice(ctxs[i]->device);
		const ncclResult_t r = nccl().CommInitRank(&comms[i], n, id, i);
		if (r != ncclSuccess) {
			nccl().GroupEnd();
			return nccl_check(ctxs[i], r, "ncclCommInitRank");
		}
	}
	LMB_NCCL(ctxs[0], nccl().GroupEnd());
	for (int i = 0; i < n; i++) {
		ctxs[i]->comm = comms[i], ctxs[i]->comm_rank = i, ctxs[i]->comm_size = n;
		cudaSetDevice(ctxs[i]->device);
		if (const int rc = comm_streams(ctxs[i])) return rc;
	}
	return LMB_OK;
}

int lmb_comm_destroy(lmb_ctx* ctx) {
	if (!ctx) return LMB_ERR_INVALID;
	cudaSetDevice(ctx->device);
	comm_free(ctx);
	return LMB_OK;
}

int lmb_comm_info(lmb_ctx* ctx, int* rank, int* n_ranks, int* nccl_version) {
	if (!ctx) return LMB_ERR_INVALID;
	if (rank) *rank = ctx->comm_rank;
	if (n_ranks) *n_ranks = ctx->comm ? ctx->comm_size : 0;
	if (nccl_version) {
		*nccl_version = 0;
		if (nccl().error.empty()) nccl().GetVersion(nccl_version);
	}
	return LMB_OK;
}

// film (LMB_FILM_SUM, this rank's samples) -> sum over the ranks -> rgb / alpha, on every rank (root < 0: ncclAllReduce) or on one
// (root >= 0: ncclReduce; the other ranks only contribute -- a progressive renderer shows ONE image, so N - 1 ranks need neither the
// sum nor the trip home of 133 MB at 4K).
//   out_rgba == NULL: in place, in stream order on the render stream; the film holds the resolved image of ALL ranks afterwards
//     (on the root only with a root; the other ranks' films are left as they were).
//   out_rgba != NULL: the film is snapshotted in stream order (device-to-device copy) and, when clear_film is set, zeroed for the next
//     batch; the snapshot is reduced, resolved and copied to out_rgba (host -- pinned for a truly asynchronous copy -- or device) on
//     the context's high-priority comm stream while the render stream goes on with the next lmb_render. With a root, out_rgba is
//     written on the root only and may be NULL elsewhere (the snapshot form is then chosen by clear_film).
// Nothing waits on the host in either form; lmb_sync does.
static int film_reduce(lmb_ctx* ctx, int root, float* out_rgba, int clear_film, const char* who) {
	if (!ctx || !ctx->film) return set_error(ctx, LMB_ERR_INVALID, std::string(who) + ": call lmb_init first");
	if (!ctx->comm) return set_error(ctx, LMB_ERR_INVALID, std::string(who) + ": call lmb_comm_init first");
	if (root >= ctx->comm_size) return set_error(ctx, LMB_ERR_INVALID, std::string(who) + ": root is not a rank of the communicator");
	cudaSetDevice(ctx->device);
	const size_t n_pix = (size_t)ctx->width * ctx->height, bytes = n_pix * 16;
	ncclComm_t comm = (ncclComm_t)ctx->comm;
	const bool mine = root < 0 || root == ctx->comm_rank;  // this rank receives the sum
	auto reduce = [&](float4* buf, cudaStream_t st) {
		return root < 0 ? nccl().AllReduce(buf, buf, n_pix * 4, ncclFloat, ncclSum, comm, st) : nccl().Reduce(buf, buf, n_pix * 4, ncclFloat, ncclSum, root, comm, st);
	};
	if (!out_rgba && !(root >= 0 && !mine && clear_film)) {
		if (clear_film) return set_error(ctx, LMB_ERR_INVALID, std::string(who) + ": clear_film needs out_rgba (the in-place form leaves the result in the film)");
		LMB_NCCL(ctx, reduce(ctx->film, ctx->stream));
		return mine ? launch_resolve_on(ctx, ctx->film, ctx->stream) : LMB_OK;
	}
	if (!ctx->reduce_buf) LMB_CUDA(ctx, cudaMalloc((void**)&ctx->reduce_buf, bytes));
	if (ctx->reduce_pending) LMB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm_done, 0));  // the previous reduce still owns the buffer
	LMB_CUDA(ctx, cudaMemcpyAsync(ctx->reduce_buf, ctx->film, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	if (clear_film) LMB_CUDA(ctx, cudaMemsetAsync(ctx->film, 0, bytes, ctx->stream));
	LMB_CUDA(ctx, cudaEventRecord(ctx->ev_comm_ready, ctx->stream));
	LMB_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_comm_ready, 0));
	LMB_NCCL(ctx, reduce(ctx->reduce_buf, ctx->comm_stream));
	if (mine) {
		if (const int rc = launch_resolve_on(ctx, ctx->reduce_buf, ctx->comm_stream)) return rc;
		LMB_CUDA(ctx, cudaMemcpyAsync(out_rgba, ctx->reduce_buf, bytes, cudaMemcpyDefault, ctx->comm_stream));
	}
	LMB_CUDA(ctx, cudaEventRecord(ctx->ev_comm_done, ctx->comm_stream));
	ctx->reduce_pending = true;
	return LMB_OK;
}

int lmb_film_allreduce(lmb_ctx* ctx, float* out_rgba, int clear_film) { return film_reduce(ctx, -1, out_rgba, clear_film, "lmb_film_allreduce"); }

int lmb_film_reduce(lmb_ctx* ctx, int root, float* out_rgba, int clear_film) {
	if (root < 0) return set_error(ctx, LMB_ERR_INVALID, "lmb_film_reduce: root must be a rank (lmb_film_allreduce gives every rank the image)");
	if (ctx && ctx->comm && root == ctx->comm_rank && clear_film && !out_rgba)
		return set_error(ctx, LMB_ERR_INVALID, "lmb_film_reduce: clear_film needs out_rgba on the root (the in-place form leaves the result in the film)");
	return film_reduce(ctx, root, out_rgba, clear_film, "lmb_film_reduce");
}

}  // extern "C"
