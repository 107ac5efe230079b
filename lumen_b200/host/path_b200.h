// PathB200: drop-in for Lumen's `class Path final : public Integrator` (reference: src/RayTracer/Path.h:4-23,
// src/RayTracer/Path.cpp:4-77) that drives the CUDA library through the C ABI of include/lumen_b200.h.
//   Path::init      -> lmb_create + lmb_upload_scene + lmb_init          (Integrator::init allocates output_tex + UBO)
//   create_accel    -> lmb_build_accel                                   (Integrator.cpp:137-160)
//   Path::render    -> lmb_render(pc, ubo, frame_num, frames_per_call)   (Path.cpp:27-59; frames_per_call = 1 in Lumen)
//   Path::update    -> frame_num += frames_per_call; camera change => frame_num = 0 (Path.cpp:61-68)
//   Path::destroy   -> lmb_destroy
#pragma once
#include <stdexcept>
#include <string>

#include "integrator.h"
#include "lumen_b200.h"

class PathB200 final : public Integrator {
  public:
	PathB200(lmh::Scene* scene, int device = 0, uint32_t frames_per_call = 1)
		: Integrator(scene), device(device), frames_per_call(frames_per_call), path_length((uint32_t)scene->config.path_length) {}
	~PathB200() override { destroy(); }

	void init() override {
		check(lmb_create(&ctx, device), "lmb_create");
		const lmb_scene_desc desc = lumen_scene->desc();
		check(lmb_upload_scene(ctx, &desc), "lmb_upload_scene");
		check(lmb_init(ctx, lumen_scene->width, lumen_scene->height, frames_per_call), "lmb_init");
		scene_ubo = lumen_scene->make_ubo();
		frame_num = 0;
	}
	void create_accel() override { check(lmb_build_accel(ctx), "lmb_build_accel"); }
	void render() override {
		pc_ray = lumen_scene->make_pc((int)path_length, direct_lighting);
		pc_ray.frame_num = frame_num;
		check(lmb_render(ctx, &pc_ray, &scene_ubo, frame_num, frames_per_call, 1, LMB_FILM_RUNNING_MEAN), "lmb_render");
	}
	bool update() override {
		frame_num += frames_per_call;
		if (updated) {  // camera moved: restart accumulation (Path.cpp:63-66)
			scene_ubo = lumen_scene->make_ubo();
			frame_num = 0;
		}
		const bool r = updated;
		updated = false;
		return r;
	}
	void destroy() override {
		if (ctx) lmb_destroy(ctx);
		ctx = nullptr;
	}
	const std::vector<float>& read_output() override {
		film.resize((size_t)lumen_scene->width * lumen_scene->height * 4);
		check(lmb_download(ctx, film.data()), "lmb_download");
		return film;
	}
	// F10 -> save_exr("out.exr") (RayTracer.cpp:447-452): half conversion and B, G, R planes are made on the device.
	void save_exr(const char* path) {
		const size_t n = (size_t)lumen_scene->width * lumen_scene->height;
		half_planes.resize(3 * n);
		check(lmb_download_half_bgr(ctx, half_planes.data()), "lmb_download_half_bgr");
		std::string err;
		if (!lmh::save_exr_half_bgr(half_planes.data(), (int)lumen_scene->width, (int)lumen_scene->height, path, &err)) throw std::runtime_error("save_exr: " + err);
	}
	// load_reference / has_gt + calc_rmse (RayTracer.cpp:117-126, 215-241, 456-461)
	void set_reference(const float* gt_rgba) { check(lmb_set_reference_image(ctx, gt_rgba), "lmb_set_reference_image"); }
	void rmse(float* literal, double* true_rmse) { check(lmb_rmse(ctx, literal, true_rmse), "lmb_rmse"); }
	// Checkpoint / resume of a progressive render: the running-mean film plus frame_num continue the accumulation exactly
	// (path.rgen:106-109 reads the old value when frame_num > 0).
	void save_checkpoint(const char* path) {
		std::string err;
		if (!lmh::save_checkpoint(path, read_output().data(), lumen_scene->width, lumen_scene->height, frame_num, path_length, &err))
			throw std::runtime_error("save_checkpoint: " + err);
	}
	void load_checkpoint(const char* path) {
		uint32_t w = 0, h = 0, frames = 0, depth = 0;
		std::string err;
		if (!lmh::load_checkpoint(path, film, w, h, frames, depth, &err)) throw std::runtime_error("load_checkpoint: " + err);
		if (w != lumen_scene->width || h != lumen_scene->height || depth != path_length)
			throw std::runtime_error("load_checkpoint: checkpoint is for another image size or path length");
		check(lmb_upload_film(ctx, film.data()), "lmb_upload_film");
		frame_num = frames;
	}
	lmb_stats stats() {
		lmb_stats s{};
		check(lmb_get_stats(ctx, &s), "lmb_get_stats");
		return s;
	}
	uint32_t path_length;         // ImGui "Path length" slider, Path.cpp:74
	bool direct_lighting = true;  // ImGui "Direct lighting", Path.cpp:75

  private:
	void check(int rc, const char* what) {
		if (rc != 0) throw std::runtime_error(std::string(what) + ": " + lmb_last_error(ctx));
	}
	int device;
	uint32_t frames_per_call;
	lmb_ctx* ctx = nullptr;
	lmb_pc_path pc_ray{};
	lmb_scene_ubo scene_ubo{};
	std::vector<float> film;
	std::vector<uint16_t> half_planes;
};

// PathB200Multi: the same integrator over several GPUs of one box. Every device holds a full scene + BVH replica
// and renders the frames f with f mod N == its rank over the whole image (sample-index sharding: perfect balance, no seams)
// into an un-normalised sum film with the per-pixel valid-sample count in alpha (LMB_FILM_SUM); read_output() reduces the films
// with ONE NCCL reduce to device 0 over NVLink / NVSwitch followed by the "/ count" epilogue there (lmb_comm_init_all +
// lmb_film_reduce; the contexts must sit on distinct GPUs). Where the same GPU is listed twice (one-GPU test boxes) NCCL cannot
// be used (one rank per device) and the films are added on device 0 by device-to-device copies instead (lmb_film_add_from).
// One host thread per device, because lmb_render is synchronous on return. render() advances frame_num by N * frames_per_call.
#include <thread>

// ShardedMulti holds everything but the render call; PathB200Multi and BDPTB200Multi (bdpt_b200.h) supply it.
class ShardedMulti : public Integrator {
  public:
	ShardedMulti(lmh::Scene* scene, std::vector<int> devices, uint32_t frames_per_call = 1)
		: Integrator(scene), path_length((uint32_t)scene->config.path_length), devices(std::move(devices)), frames_per_call(frames_per_call) {
		if (this->devices.empty()) throw std::runtime_error("ShardedMulti: no device");
	}
	~ShardedMulti() override { destroy(); }

	void init() override {
		ctx.assign(devices.size(), nullptr);
		const lmb_scene_desc desc = lumen_scene->desc();
		each([&](size_t r) {
			check(r, lmb_create(&ctx[r], devices[r]), "lmb_create");
			check(r, lmb_upload_scene(ctx[r], &desc), "lmb_upload_scene");
			check(r, lmb_init(ctx[r], lumen_scene->width, lumen_scene->height, frames_per_call), "lmb_init");
		});
		scene_ubo = lumen_scene->make_ubo();
		frame_num = 0;
		reduced = false;
		bool distinct = ctx.size() > 1;
		for (size_t a = 0; a < devices.size(); a++)
			for (size_t b = 0; b < a; b++) distinct = distinct && devices[a] != devices[b];
		use_nccl = false;
		if (distinct) {
			check(0, lmb_comm_init_all(ctx.data(), (int)ctx.size()), "lmb_comm_init_all");
			use_nccl = true;
		}
	}
	bool reduces_with_nccl() const { return use_nccl; }
	void create_accel() override {
		each([&](size_t r) { check(r, lmb_build_accel(ctx[r]), "lmb_build_accel"); });
	}
	void render() override {
		if (reduced) throw std::runtime_error("ShardedMulti: render() after read_output() needs init() (the films hold the reduced image)");
		prepare_frame();
		const uint32_t n = (uint32_t)ctx.size();
		each([&](size_t r) { check(r, render_shard(ctx[r], frame_num + (uint32_t)r, frames_per_call, n), what()); });
	}
	bool update() override {
		frame_num += frames_per_call * (uint32_t)ctx.size();
		return false;
	}
	void destroy() override {
		for (auto& c : ctx)
			if (c) lmb_destroy(c);
		ctx.clear();
	}
	const std::vector<float>& read_output() override {
		reduce();
		film.resize((size_t)lumen_scene->width * lumen_scene->height * 4);
		check(0, lmb_download(ctx[0], film.data()), "lmb_download");
		return film;
	}
	void save_exr(const char* path) {
		reduce();
		const size_t n = (size_t)lumen_scene->width * lumen_scene->height;
		half_planes.resize(3 * n);
		check(0, lmb_download_half_bgr(ctx[0], half_planes.data()), "lmb_download_half_bgr");
		std::string err;
		if (!lmh::save_exr_half_bgr(half_planes.data(), (int)lumen_scene->width, (int)lumen_scene->height, path, &err)) throw std::runtime_error("save_exr: " + err);
	}
	// summed over the devices (ms_render: the slowest device)
	lmb_stats stats() {
		lmb_stats total{};
		for (size_t r = 0; r < ctx.size(); r++) {
			lmb_stats s{};
			check(r, lmb_get_stats(ctx[r], &s), "lmb_get_stats");
			if (r == 0) total = s;
			else {
				total.rays_closest += s.rays_closest, total.rays_shadow += s.rays_shadow, total.rays_probe += s.rays_probe;
				total.nodes_visited += s.nodes_visited, total.tris_tested += s.tris_tested, total.nan_samples += s.nan_samples;
				total.frames += s.frames, total.kernel_launches += s.kernel_launches;
				total.ms_render = std::max(total.ms_render, s.ms_render);
			}
		}
		return total;
	}
	uint32_t path_length;
	bool direct_lighting = true;

  protected:
	virtual void prepare_frame() = 0;  // push constants of this round
	virtual int render_shard(lmb_ctx* c, uint32_t first_frame, uint32_t n_frames, uint32_t frame_stride) = 0;  // into the LMB_FILM_SUM film
	virtual const char* what() const = 0;
	lmb_scene_ubo scene_ubo{};

  private:
	// sum films -> device 0, in rank order, then rgb /= valid-sample count
	void reduce() {
		if (reduced) return;
		if (use_nccl) {
			// collective: every rank enqueues its part of the reduce to device 0 (the one read_output / save_exr read; + resolve there),
			// then waits for its own stream
			each([&](size_t r) {
				check(r, lmb_film_reduce(ctx[r], 0, nullptr, 0), "lmb_film_reduce");
				check(r, lmb_sync(ctx[r]), "lmb_sync");
			});
		} else {
			for (size_t r = 1; r < ctx.size(); r++) check(0, lmb_film_add_from(ctx[0], ctx[r]), "lmb_film_add_from");
			check(0, lmb_resolve(ctx[0]), "lmb_resolve");
		}
		reduced = true;
	}
	template <typename F>
	void each(F f) {
		std::vector<std::thread> th;
		std::vector<std::string> errs(devices.size());
		for (size_t r = 0; r < devices.size(); r++)
			th.emplace_back([&, r] {
				try {
					f(r);
				} catch (const std::exception& e) {
					errs[r] = e.what();
				}
			});
		for (auto& t : th) t.join();
		for (size_t r = 0; r < errs.size(); r++)
			if (!errs[r].empty()) throw std::runtime_error("device " + std::to_string(devices[r]) + ": " + errs[r]);
	}
	void check(size_t r, int rc, const char* what) {
		if (rc != 0) throw std::runtime_error(std::string(what) + ": " + lmb_last_error(r < ctx.size() ? ctx[r] : nullptr));
	}
	std::vector<int> devices;
	uint32_t frames_per_call;
	std::vector<lmb_ctx*> ctx;
	bool reduced = false;
	bool use_nccl = false;
	std::vector<float> film;
	std::vector<uint16_t> half_planes;
};

class PathB200Multi final : public ShardedMulti {
  public:
	using ShardedMulti::ShardedMulti;

  protected:
	void prepare_frame() override {
		pc_ray = lumen_scene->make_pc((int)path_length, direct_lighting);
		pc_ray.frame_num = frame_num;
	}
	int render_shard(lmb_ctx* c, uint32_t first_frame, uint32_t n_frames, uint32_t frame_stride) override {
		return lmb_render(c, &pc_ray, &scene_ubo, first_frame, n_frames, frame_stride, LMB_FILM_SUM);
	}
	const char* what() const override { return "lmb_render"; }

  private:
	lmb_pc_path pc_ray{};
};
