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
