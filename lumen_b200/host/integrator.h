// Headless counterpart of Lumen's abstract integrator (reference: src/RayTracer/Integrator.h:14-33). Same lifecycle and
// public state; the Vulkan members (vk::Texture* output_tex, vk::BVH) become a host-readable film and the C-ABI context.
#pragma once
#include <cstdint>
#include <vector>

#include "lumen_scene.h"

class Integrator {
  public:
	explicit Integrator(lmh::Scene* lumen_scene) : lumen_scene(lumen_scene) {}
	virtual ~Integrator() = default;
	virtual void init() = 0;
	virtual void render() = 0;
	virtual bool gui() { return false; }  // no GUI on a headless B200 box
	virtual bool update() = 0;
	virtual void destroy() = 0;
	virtual void create_accel() = 0;  // Integrator::create_accel(vk::BVH&, std::vector<vk::BVH>&) without the Vulkan handles
	// output_tex equivalent: RGBA32F, row-major, filled by read_output()
	virtual const std::vector<float>& read_output() = 0;
	bool updated = false;
	uint32_t frame_num = 0;

  protected:
	lmh::Scene* lumen_scene = nullptr;
};
