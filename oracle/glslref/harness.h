// ORACLE -- TEST INFRASTRUCTURE ONLY. What the stage translation units share: the scene handle with the descriptor bindings of
// Path::init / Path::render (Path.cpp:6-11, 49-57) and the per-thread ray counters of the shader-binding-table emulation (stage_path.cpp).
#pragma once
#include <vector>
#include "glslref.h"
#include "stage_common.h"

struct ref_scene {
	lmb_scene_desc sd;
	const void* user;
	ref_trace1_fn trace1;
	ref_texture_fn texture;
	::SceneDesc scene_desc;  // commons.h:237-312; Path.cpp:6-11 fills four addresses
	::SceneUBO ubo;
	::PCPath pc;
	std::vector<glslref::sampler2D> samplers;
	glslref::Env env;
};


namespace glslref {
extern thread_local uint64_t t_rays[3];  // traceRayEXT calls of this thread: closest (cull mask 0xFF), any-hit, closest (cull mask 0x1)
}
