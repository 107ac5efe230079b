// lumen_headless -- headless equivalent of Lumen's application shell for the Path integrator.
// Reference lifecycle: main (src/main.cpp:7-31) -> RayTracer::init (src/RayTracer/RayTracer.cpp:19-69: load_scene,
// create_integrator, init, create_accel) -> per frame render()/update() (:167-197) -> F10: save_exr("out.exr") (:447-452)
// -> cleanup (:491-499). The window size (hard-coded 1920x1080 in main.cpp:14-18) is a CLI option here.
//
//   lumen_headless scene.(json|xml) [--width W] [--height H] [--spp N] [--depth D] [--out out.exr] [--device i] [--batch F]
//                  [--ref gt.exr [--target-rmse X]] [--checkpoint file [--checkpoint-every N] [--resume]]
//                  [--integrator path|bdpt|scene] [--time T]   BDPT (BDPTB200); --time fixes PCBDPT.time
//                  [--devices 0,1,...]   several GPUs of one box: sample-index sharding, sum films reduced by one NCCL all-reduce
//                                        (PathB200Multi; a device may be listed twice; not combined with --ref / --checkpoint)
// Progressive service: with --ref the RMSE against the ground-truth image is computed on the device
// after every batch (Lumen does it every 5 s, RayTracer.cpp:453-461, and prints rmse * 1e6) and rendering stops early once
// the true RMSE falls to --target-rmse; --checkpoint writes film + frame count every N frames (and at the end), --resume
// continues from it with results bit-identical to an uninterrupted run.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <regex>
#include <string>
#include <vector>
#include <stdexcept>

#include "bdpt_b200.h"
#include "path_b200.h"

int main(int argc, char** argv) {
	std::string scene_name = "scenes/caustics.json";  // RayTracer.cpp:14-17 default
	uint32_t width = 1920, height = 1080, spp = 64, batch = 8;
	int depth = 0, device = 0;
	std::string out = "out.exr", ref_path, ckpt_path;
	double target_rmse = -1.0;
	uint32_t ckpt_every = 0;
	bool resume = false;
	std::string integrator_name = "path";
	long bdpt_time = -1;  // --time: fixed PCBDPT.time (default: rand() per frame as BDPT.cpp:57)
	std::vector<int> devices;
	const std::regex fn("(.*).(.json|.xml)");  // RayTracer::parse_args, RayTracer.cpp:468-476
	for (int i = 1; i < argc; i++) {
		const std::string a = argv[i];
		auto next = [&]() { return (i + 1 < argc) ? argv[++i] : ""; };
		if (a == "--width") width = (uint32_t)atoi(next());
		else if (a == "--height") height = (uint32_t)atoi(next());
		else if (a == "--spp") spp = (uint32_t)atoi(next());
		else if (a == "--depth") depth = atoi(next());
		else if (a == "--out") out = next();
		else if (a == "--device") device = atoi(next());
		else if (a == "--batch") batch = (uint32_t)atoi(next());
		else if (a == "--ref") ref_path = next();
		else if (a == "--target-rmse") target_rmse = atof(next());
		else if (a == "--checkpoint") ckpt_path = next();
		else if (a == "--checkpoint-every") ckpt_every = (uint32_t)atoi(next());
		else if (a == "--resume") resume = true;
		else if (a == "--integrator") integrator_name = next();
		else if (a == "--time") bdpt_time = atol(next());
		else if (a == "--devices") {
			const std::string list = next();
			for (size_t b = 0; b < list.size();) {
				const size_t e = list.find(',', b);
				devices.push_back(atoi(list.substr(b, e == std::string::npos ? std::string::npos : e - b).c_str()));
				if (e == std::string::npos) break;
				b = e + 1;
			}
		}
		else if (std::regex_match(a, fn)) scene_name = a;
	}
	try {
		lmh::Scene scene;
		scene.load(scene_name, width, height);
		// create_integrator (RayTracer.cpp:244-282) picks by scene.config; here --integrator path | bdpt | scene (default path)
		if (integrator_name == "scene") integrator_name = scene.config.integrator_name;
		if (integrator_name != "path" && integrator_name != "bdpt") {
			fprintf(stderr, "note: integrator '%s' is not provided (path, bdpt); using path\n", integrator_name.c_str());
			integrator_name = "path";
		} else if (scene.config.integrator_name != integrator_name) {
			fprintf(stderr, "note: scene asks for integrator '%s'; rendering with '%s'\n", scene.config.integrator_name.c_str(), integrator_name.c_str());
		}
		batch = std::max(1u, std::min(batch, spp));
		if (integrator_name == "bdpt" && devices.size() > 1) {
			if (!ref_path.empty() || !ckpt_path.empty()) throw std::runtime_error("--devices cannot be combined with --ref / --checkpoint");
			const uint32_t n = (uint32_t)devices.size();
			BDPTB200Multi multi(&scene, devices, batch);
			if (depth > 0) multi.path_length = (uint32_t)depth;
			if (bdpt_time >= 0) multi.set_time((uint32_t)bdpt_time);
			multi.init();
			multi.create_accel();
			while (multi.frame_num + batch * n <= spp) {
				multi.render();
				multi.update();
			}
			const lmb_stats st = multi.stats();
			const double rays = (double)(st.rays_closest + st.rays_shadow);
			printf("%u x %u, %llu frames on %u devices, BDPT depth %u: %.1f ms on the slowest device, %.1f Mrays/s, %.2f spp/s\n", width, height,
				   (unsigned long long)st.frames, n, multi.path_length, st.ms_render, rays / st.ms_render / 1e3, st.frames / (st.ms_render * 1e-3));
			multi.save_exr(out.c_str());
			printf("wrote %s (films reduced by %s)\n", out.c_str(), multi.reduces_with_nccl() ? "one NCCL all-reduce" : "device-to-device adds on device 0");
			multi.destroy();
			return 0;
		}
		if (integrator_name == "bdpt") {
			if (!ref_path.empty() || !ckpt_path.empty()) throw std::runtime_error("--integrator bdpt cannot be combined with --ref / --checkpoint");
			BDPTB200 bdpt(&scene, device, batch);
			if (depth > 0) bdpt.path_length = (uint32_t)depth;
			if (bdpt_time >= 0) bdpt.set_time((uint32_t)bdpt_time);
			bdpt.init();
			bdpt.create_accel();
			while (bdpt.frame_num + batch <= spp) {
				bdpt.render();
				bdpt.update();
			}
			const lmb_stats st = bdpt.stats();
			const double rays = (double)(st.rays_closest + st.rays_shadow);
			printf("%u x %u, %llu frames, BDPT depth %u: %.1f ms on device, %.1f Mrays/s, %.2f spp/s\n", width, height, (unsigned long long)st.frames,
				   bdpt.path_length, st.ms_render, rays / st.ms_render / 1e3, st.frames / (st.ms_render * 1e-3));
			bdpt.save_exr(out.c_str());
			printf("wrote %s\n", out.c_str());
			bdpt.destroy();
			return 0;
		}
		if (devices.size() > 1) {
			if (!ref_path.empty() || !ckpt_path.empty()) throw std::runtime_error("--devices cannot be combined with --ref / --checkpoint");
			const uint32_t n = (uint32_t)devices.size();
			PathB200Multi multi(&scene, devices, batch);
			if (depth > 0) multi.path_length = (uint32_t)depth;
			multi.init();
			multi.create_accel();
			while (multi.frame_num + batch * n <= spp) {  // whole rounds of N x batch frames
				multi.render();
				multi.update();
			}
			const lmb_stats st = multi.stats();
			const double rays = (double)(st.rays_closest + st.rays_shadow + st.rays_probe);
			printf("%u x %u, %llu frames on %u devices, depth %u: %.1f ms on the slowest device, %.1f Mrays/s, %.2f spp/s\n", width, height,
				   (unsigned long long)st.frames, n, multi.path_length, st.ms_render, rays / st.ms_render / 1e3, st.frames / (st.ms_render * 1e-3));
			multi.save_exr(out.c_str());
			printf("wrote %s (films reduced by %s)\n", out.c_str(), multi.reduces_with_nccl() ? "one NCCL all-reduce" : "device-to-device adds on device 0");
			multi.destroy();
			return 0;
		}
		PathB200 integrator(&scene, device, batch);
		if (depth > 0) integrator.path_length = (uint32_t)depth;
		integrator.init();
		integrator.create_accel();
		if (!ref_path.empty()) {
			std::vector<float> gt;
			int gw = 0, gh = 0;
			std::string err;
			if (!lmh::load_exr(ref_path.c_str(), gt, gw, gh, &err)) throw std::runtime_error("--ref: " + err);
			if ((uint32_t)gw != width || (uint32_t)gh != height) throw std::runtime_error("--ref: image size differs from --width/--height");
			integrator.set_reference(gt.data());
		}
		if (resume && !ckpt_path.empty()) {
			integrator.load_checkpoint(ckpt_path.c_str());
			printf("resumed %s at frame %u\n", ckpt_path.c_str(), integrator.frame_num);
		}
		uint32_t last_ckpt = integrator.frame_num;
		while (integrator.frame_num + batch <= spp) {
			integrator.render();
			integrator.update();
			if (!ref_path.empty()) {
				float lit = 0;
				double tru = 0;
				integrator.rmse(&lit, &tru);
				printf("frame %u: RMSE %g (Lumen's routine x 1e6), true RMSE %g\n", integrator.frame_num, (double)lit * 1e6, tru);
				if (target_rmse >= 0 && tru <= target_rmse) break;
			}
			if (!ckpt_path.empty() && ckpt_every && integrator.frame_num - last_ckpt >= ckpt_every) {
				integrator.save_checkpoint(ckpt_path.c_str());
				last_ckpt = integrator.frame_num;
			}
		}
		if (!ckpt_path.empty()) integrator.save_checkpoint(ckpt_path.c_str());
		const lmb_stats st = integrator.stats();
		const double rays = (double)(st.rays_closest + st.rays_shadow + st.rays_probe);
		printf("%u x %u, %llu frames, depth %u: %.1f ms on device, %.1f Mrays/s, %.2f spp/s, LBVH build %.2f ms\n", width, height,
			   (unsigned long long)st.frames, integrator.path_length, st.ms_render, rays / st.ms_render / 1e3, st.frames / (st.ms_render * 1e-3),
			   st.ms_build_accel);
		integrator.save_exr(out.c_str());
		printf("wrote %s\n", out.c_str());
		integrator.destroy();
	} catch (const std::exception& e) {
		fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
