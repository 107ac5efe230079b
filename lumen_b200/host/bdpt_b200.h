// BDPTB200: drop-in for Lumen's `class BDPT final : public Integrator` (reference: src/RayTracer/BDPT.h, src/RayTracer/BDPT.cpp:4-108)
// over the C ABI of include/lumen_b200.h.
//   BDPT::init      -> lmb_create + lmb_upload_scene + lmb_init   (the light / camera path buffers and the colour storage of
//                      BDPT.cpp:7-24 are device memory of the context, allocated by the first lmb_render_bdpt)
//   create_accel    -> lmb_build_accel
//   BDPT::render    -> lmb_render_bdpt(pc, ubo, frame_num, frames_per_call)   (BDPT.cpp:55-95: PCBDPT filled from the scene;
//                      pc.time = rand() % UINT_MAX as BDPT.cpp:57 unless fixed_time is set -- reproducible renders)
//   BDPT::update    -> frame_num += frames_per_call; camera change => frame_num = 0 (BDPT.cpp:97-104)
//   BDPT::destroy   -> lmb_destroy
#pragma once
#include <climits>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "integrator.h"
#include "lumen_b200.h"
#include "path_b200.h"  // ShardedMulti

class BDPTB200 final : public Integrator {
  public:
	BDPTB200(lmh::Scene* scene, int device = 0, uint32_t frames_per_call = 1)
		: Integrator(scene), path_length((uint32_t)scene->config.path_length), device(device), frames_per_call(frames_per_call) {}
	~BDPTB200() override { destroy(); }

	void init() override {
		check(lmb_create(&ctx, device), "lmb_create");
		const lmb_scene_desc desc = lumen_scene->desc();
		check(lmb_upload_scene(ctx, &desc), "lmb_upload_scene");
		check(lmb_init(ctx, lumen_scene->width, lumen_scene->height, 1), "lmb_init");  // 1: no Path wavefront batch is needed
		scene_ubo = lumen_scene->make_ubo();
		frame_num = 0;
	}
	void create_accel() override { check(lmb_build_accel(ctx), "lmb_build_accel"); }
	void render() override {
		const lmb_pc_path p = lumen_scene->make_pc((int)path_length, true);  // the same scene-derived fields as Path.cpp:27-38
		static_assert(sizeof(lmb_pc_bdpt) + sizeof(uint32_t) == sizeof(lmb_pc_path), "PCBDPT is PCPath without direct_lighting");
		std::memcpy(&pc_ray, &p, sizeof(pc_ray));
		pc_ray.frame_num = frame_num;
		pc_ray.time = has_fixed_time ? fixed_time : (uint32_t)(rand() % UINT_MAX);
		check(lmb_render_bdpt(ctx, &pc_ray, &scene_ubo, frame_num, frames_per_call, 1, LMB_FILM_RUNNING_MEAN), "lmb_render_bdpt");
	}
	bool update() override {
		frame_num += frames_per_call;
		if (updated) {
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
	void save_exr(const char* path) {
		const size_t n = (size_t)lumen_scene->width * lumen_scene->height;
		half_planes.resize(3 * n);
		check(lmb_download_half_bgr(ctx, half_planes.data()), "lmb_download_half_bgr");
		std::string err;
		if (!lmh::save_exr_half_bgr(half_planes.data(), (int)lumen_scene->width, (int)lumen_scene->height, path, &err)) throw std::runtime_error("save_exr: " + err);
	}
	lmb_stats stats() {
		lmb_stats s{};
		check(lmb_get_stats(ctx, &s), "lmb_get_stats");
		return s;
	}
	void set_time(uint32_t t) { has_fixed_time = true, fixed_time = t; }
	uint32_t path_length;

  private:
	void check(int rc, const char* what) {
		if (rc != 0) throw std::runtime_error(std::string(what) + ": " + lmb_last_error(ctx));
	}
	int device;
	uint32_t frames_per_call;
	bool has_fixed_time = false;
	uint32_t fixed_time = 0;
	lmb_ctx* ctx = nullptr;
	lmb_pc_bdpt pc_ray{};
	lmb_scene_ubo scene_ubo{};
	std::vector<float> film;
	std::vector<uint16_t> half_planes;
};

// BDPTB200Multi: BDPT over several GPUs of one box (SURVEY.md 8e on the 8f-3 row; the reference's BDPT drives one GPU,
// src/RayTracer/BDPT.cpp:55-95). Same sharding and the same reduce as PathB200Multi: device r renders the frames f with f mod N == r
// through lmb_render_bdpt into its LMB_FILM_SUM film (col + the frame's own light-tracer splats), one NCCL all-reduce + resolve.
// pc.time is one value per render() round, the same on every device (bdpt.rgen:36-37 seeds with frame ^ time).
class BDPTB200Multi final : public ShardedMulti {
  public:
	using ShardedMulti::ShardedMulti;
	void set_time(uint32_t t) { has_fixed_time = true, fixed_time = t; }

  protected:
	void prepare_frame() override {
		const lmb_pc_path p = lumen_scene->make_pc((int)path_length, true);
		std::memcpy(&pc_ray, &p, sizeof(pc_ray));
		pc_ray.frame_num = frame_num;
		pc_ray.time = has_fixed_time ? fixed_time : (uint32_t)(rand() % UINT_MAX);
	}
	int render_shard(lmb_ctx* c, uint32_t first_frame, uint32_t n_frames, uint32_t frame_stride) override {
		return lmb_render_bdpt(c, &pc_ray, &scene_ubo, first_frame, n_frames, frame_stride, LMB_FILM_SUM);
	}
	const char* what() const override { return "lmb_render_bdpt"; }

  private:
	bool has_fixed_time = false;
	uint32_t fixed_time = 0;
	lmb_pc_bdpt pc_ray{};
};
