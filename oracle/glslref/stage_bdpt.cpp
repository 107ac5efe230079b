// ORACLE -- TEST INFRASTRUCTURE ONLY. src/shaders/integrators/bdpt/bdpt.rgen (+ commons.glsl, bsdf_commons.glsl, bsdf/*.glsl,
// integrators/bdpt_commons.glsl), translated by glsl2cpp.py, and what BDPT::render (src/RayTracer/BDPT.cpp:55-95) sets up around it:
// the light / camera path buffers and the colour storage behind SceneDesc addresses, zeroed per frame (BDPT.cpp:79-80).
//
// The GLSL splats light-tracer samples with a NON-ATOMIC `tmp_col.d[idx] += splat` into a buffer other invocations read and clear in
// the same dispatch (bdpt.rgen:69-78): what a pixel reads there depends on the GPU's scheduling. To have a function of the inputs,
// every invocation here gets a private, zeroed colour storage: main() then stores (own strategies + the splats the pixel sends to
// itself) into the image, and the splats it sends to OTHER pixels are harvested from the private storage into a separate image.
// Their per-pixel sum is what any scheduling of the reference adds up to over a frame (oracle/bdpt.h point B2).
#include <omp.h>
#include "harness.h"
#include "integrators/bdpt/bdpt_commons.h"

static_assert(sizeof(::PCBDPT) == sizeof(lmb_pc_bdpt), "PCBDPT");

namespace glslref {
struct BdptRgen : Stage {
	using Stage::Stage;
#include "gen/integrators/bdpt/bdpt.rgen.inc"
};
}  // namespace glslref

using namespace glslref;

extern "C" int ref_render_bdpt_frame(ref_scene* s, const lmb_pc_bdpt* pc_in, const lmb_scene_ubo* ubo, uint32_t frame, float* image_rgba,
									 float* splat_rgb, uint64_t* rays3, int n_threads) {
	const int W = (int)pc_in->size_x, H = (int)pc_in->size_y;
	const size_t n_pix = (size_t)W * H, n_vtx = n_pix * (size_t)(pc_in->max_depth + 1);
	::PCBDPT pc;
	std::memcpy(&pc, pc_in, sizeof(pc));
	// the seed is (x, y, frame_num ^ time, 0) (bdpt.rgen:36-37); frame_num = 0 makes main() store the frame's own value
	// (bdpt.rgen:79-89) instead of mixing it into the running mean
	pc.time = frame ^ pc_in->time;
	pc.frame_num = 0;
	// BDPT.cpp:7-24 + :79-80: both path buffers, zeroed per frame. bdpt_connect_cam / calc_mis_weight form light_vtx(s - 2) with s = 1,
	// the slot before the pixel's own (a dead read: the value is never used); for the first pixel that is element 0xFFFFFFFF in the
	// shader's uint arithmetic, which glslref::BufArray answers with zeros instead of a fault (prelude.h)
	std::vector<::PathVertex> light(n_vtx), camera(n_vtx);
	std::memset(light.data(), 0, light.size() * sizeof(::PathVertex));
	std::memset(camera.data(), 0, camera.size() * sizeof(::PathVertex));
	BufferRegistry::get().add(light.data(), light.size() * sizeof(::PathVertex));
	BufferRegistry::get().add(camera.data(), camera.size() * sizeof(::PathVertex));
	const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
	std::vector<std::vector<float>> splats((size_t)nt, std::vector<float>(n_pix * 3, 0.0f));
	uint64_t r0 = 0, r1 = 0, r2 = 0;
#pragma omp parallel num_threads(nt) reduction(+ : r0, r1, r2)
	{
		std::vector<vec3> tmp_col(n_pix, vec3(0.0f));
		// private SceneDesc, SceneUBO and Env: the bindings of the handle, with this thread's own colour storage
		::SceneDesc desc = s->scene_desc;
		::SceneUBO ubo_local;
		std::memcpy(&ubo_local, ubo, sizeof(ubo_local));
		desc.light_path_addr = (uint64_t)(uintptr_t)light.data();
		desc.camera_path_addr = (uint64_t)(uintptr_t)camera.data();
		desc.color_storage_addr = (uint64_t)(uintptr_t)tmp_col.data();
		Env env = s->env;
		env.sets[0][1] = &ubo_local;
		env.sets[0][2] = &desc;
		env.push_constants = &pc;
		env.images[0] = image2D{image_rgba, W, H};
		float* my_splat = splats[(size_t)omp_get_thread_num()].data();
		t_rays[0] = t_rays[1] = t_rays[2] = 0;
#pragma omp for schedule(dynamic, 1)
		for (int y = 0; y < H; y++) {
			for (int x = 0; x < W; x++) {
				Stage::Inputs in;
				in.env = &env;
				in.launch_id = uvec3(x, y, 0);
				in.launch_size = uvec3(W, H, 1);
				BdptRgen inv(in);
				inv.main();
				// harvest what this pixel sent to other pixels (its own entry was read and cleared by main())
				for (size_t i = 0; i < n_pix; i++) {
					vec3& c = tmp_col[i];
					if (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f || c.x != c.x || c.y != c.y || c.z != c.z) {
						my_splat[3 * i] += c.x, my_splat[3 * i + 1] += c.y, my_splat[3 * i + 2] += c.z;
						c = vec3(0.0f);
					}
				}
			}
		}
		r0 += t_rays[0], r1 += t_rays[1], r2 += t_rays[2];
	}
	// the shader numbers pixels x * H + y (bdpt.rgen:35,68); the image is row-major
	for (int y = 0; y < H; y++)
		for (int x = 0; x < W; x++)
			for (int c = 0; c < 3; c++) {
				float sum = 0.0f;
				for (int t = 0; t < nt; t++) sum += splats[(size_t)t][3 * ((size_t)x * H + y) + c];
				splat_rgb[3 * ((size_t)y * W + x) + c] = sum;
			}
	BufferRegistry::get().remove(light.data()), BufferRegistry::get().remove(camera.data());
	if (rays3) rays3[0] += r0, rays3[1] += r1, rays3[2] += r2;
	return 0;
}
