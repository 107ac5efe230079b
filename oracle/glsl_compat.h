// ORACLE -- TEST INFRASTRUCTURE ONLY. Never linked into, imported by or executed from the product path
// (lumen_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// glsl_compat.h: the GLSL built-ins the reference shaders rely on, mapped onto glm + include/lmb_detmath.h.
// GLSL float literals are single precision; every constant below carries an `f` suffix for that reason.
#pragma once
#include <cstdint>
#include <cmath>
#include <glm/glm.hpp>
#include "lmb_detmath.h"
#include "lmb_types.h"

namespace orc {
using glm::vec2;
using glm::vec3;
using glm::vec4;
using glm::mat4;
using glm::uvec4;
using glm::ivec3;
using uint = uint32_t;

// utils.glsl:4-9
constexpr float PI = 3.14159265359f;
constexpr float TWO_PI = 6.28318530718f;
constexpr float INV_PI = 1.0f / PI;
constexpr float EPS = 0.001f;

inline float g_sin(float x) { return lmb_sinf(x); }
inline float g_cos(float x) { return lmb_cosf(x); }
inline float g_exp(float x) { return lmb_expf(x); }
inline float g_pow(float x, float y) { return lmb_powf(x, y); }
inline vec3 g_exp(const vec3& v) { return vec3(lmb_expf(v.x), lmb_expf(v.y), lmb_expf(v.z)); }
inline vec3 g_sqrt(const vec3& v) { return vec3(std::sqrt(v.x), std::sqrt(v.y), std::sqrt(v.z)); }
inline float g_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }  // GLSL sign(0) == 0
inline bool g_isinf(float x) { return std::isinf(x); }
inline bool g_isnan(float x) { return x != x; }

inline vec3 v3(const float* p) { return vec3(p[0], p[1], p[2]); }
inline mat4 m4(const float* p) {
	mat4 m;
	for (int c = 0; c < 4; c++)
		for (int r = 0; r < 4; r++) m[c][r] = p[4 * c + r];
	return m;
}
}  // namespace orc
