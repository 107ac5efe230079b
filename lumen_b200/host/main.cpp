// lumen_headless -- headless equivalent of Lumen's application shell for the Path integrator.
// Reference lifecycle: main (src/main.cpp:7-31) -> RayTracer::init (src/RayTracer/RayTracer.cpp:19-69: load_scene,
// create_integrator, init, create_accel) -> per frame render()/update() (:167-197) -> F10: save_exr("out.exr") (:447-452)
// -> cleanup (:491-499). The window size (hard-coded 1920x1080 in main.cpp:14-18) is a CLI option here.
//
//   lumen_headless scene.(json|xml) [--width W] [--height H] [--spp N] [--depth D] [--out out.exr] [--device i] [--batch F]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <regex>
#include <string>

#include "path_b200.h"

int main(int argc, char** argv) {
	std::string scene_name = "scenes/caustics.json";  // RayTracer.cpp:14-17 default
	uint32_t width = 1920, height = 1080, spp = 64, batch = 8;
	int depth = 0, device = 0;
	std::string out = "out.exr";
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
		else if (std::regex_match(a, fn)) scene_name = a;
	}
	try {
		lmh::Scene scene;
		scene.load(scene_name, width, height);
		if (scene.config.integrator_name != "path")
			fprintf(stderr, "note: scene asks for integrator '%s'; this build provides the Path integrator and uses it\n", scene.config.integrator_name.c_str());
		batch = std::max(1u, std::min(batch, spp));
		PathB200 integrator(&scene, device, batch);
		if (depth > 0) integrator.path_length = (uint32_t)depth;
		integrator.init();
		integrator.create_accel();
		while (integrator.frame_num + batch <= spp) {
			integrator.render();
			integrator.update();
		}
		const lmb_stats st = integrator.stats();
		const double rays = (double)(st.rays_closest + st.rays_shadow + st.rays_probe);
		printf("%u x %u, %llu frames, depth %u: %.1f ms on device, %.1f Mrays/s, %.2f spp/s, LBVH build %.2f ms\n", width, height,
			   (unsigned long long)st.frames, integrator.path_length, st.ms_render, rays / st.ms_render / 1e3, st.frames / (st.ms_render * 1e-3),
			   st.ms_build_accel);
		std::string err;
		if (!lmh::save_exr(integrator.read_output().data(), (int)width, (int)height, out.c_str(), &err)) {
			fprintf(stderr, "save_exr: %s\n", err.c_str());
			return 1;
		}
		printf("wrote %s\n", out.c_str());
		integrator.destroy();
	} catch (const std::exception& e) {
		fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
