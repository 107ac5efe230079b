// ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or executed from lumen_b200/).
//
// prelude.h: the GLSL execution environment the mechanically translated reference shaders (glsl2cpp.py) compile
// against. A translated stage is the body of `struct Prog : glslref::Stage { ... }`: GLSL globals are members (one
// struct instance = one shader invocation), GLSL built-in functions are static members of Stage (class-scope lookup
// finds them before anything in <cmath>), GLSL built-in variables (gl_LaunchIDEXT, ...) are members of Stage, and
// the descriptor bindings / push constants / ray tracing calls go through `Stage::env`.
//
// What is DEFINED here rather than taken from the reference (GLSL leaves it to the implementation, so some
// definition has to be chosen; the same choices as oracle/glsl_compat.h so that results can be compared bit for bit):
//   * sin / cos / exp / pow / log / exp2 / log2 -> include/lmb_detmath.h   (GLSL: precision implementation-defined)
//   * normalize, length, distance, reflect, mix, clamp, inverse, transpose, cross, dot -> glm's definitions
//     (v * (1/sqrt(dot)), sqrt(dot), I - N*dot(N,I)*2, x*(1-a) + y*a, min(max(x,lo),hi), cofactor inverse)
//   * min(x,y) = (y < x) ? y : x and max(x,y) = (x < y) ? y : x   (the GLSL spec's own formulae)
//   * no floating-point contraction (-ffp-contract=off); GLSL allows it unless `precise`
//   * ray / triangle intersection, instance transforms and texture filtering are outside the shaders (Vulkan driver
//     and hardware): they are callbacks of the environment (Env::intersect, Env::texture), supplied by the test
//     harness from oracle/liboracle.so
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#define GLM_FORCE_XYZW_ONLY
#include <glm/glm.hpp>
#include "lmb_detmath.h"

namespace glslref {
using glm::ivec2;
using glm::ivec3;
using glm::ivec4;
using glm::mat3;
using glm::mat4;
using glm::mat4x3;
using glm::uvec2;
using glm::uvec3;
using glm::uvec4;
using glm::vec2;
using glm::vec3;
using glm::vec4;
using uint = unsigned int;
using mat4x4 = glm::mat4;
using mat3x3 = glm::mat3;

// ------------------------------------------------------------------------------------------------ swizzles
template <int A, int B, class V>
inline glm::vec<2, typename V::value_type> swz(const V& v) {
	return glm::vec<2, typename V::value_type>(v[A], v[B]);
}
template <int A, int B, int C, class V>
inline glm::vec<3, typename V::value_type> swz(const V& v) {
	return glm::vec<3, typename V::value_type>(v[A], v[B], v[C]);
}
template <int A, int B, int C, int D, class V>
inline glm::vec<4, typename V::value_type> swz(const V& v) {
	return glm::vec<4, typename V::value_type>(v[A], v[B], v[C], v[D]);
}
template <int A, int B, class V, class R>
inline void swz_set(V& v, const R& r) {
	const glm::vec<2, typename V::value_type> t(r);
	v[A] = t[0], v[B] = t[1];
}
template <int A, int B, int C, class V, class R>
inline void swz_set(V& v, const R& r) {
	const glm::vec<3, typename V::value_type> t(r);
	v[A] = t[0], v[B] = t[1], v[C] = t[2];
}

// ------------------------------------------------------------------------------------------------ mixed int / float operators
// GLSL converts an int operand implicitly when the other one is a float vector; glm's templates do not deduce that.
#define GLSLREF_MIXED_OPS(V)                                                             \
	inline V operator*(int s, const V& v) { return (float)s * v; }                        \
	inline V operator*(const V& v, int s) { return v * (float)s; }                        \
	inline V operator/(const V& v, int s) { return v / (float)s; }                        \
	inline V operator/(int s, const V& v) { return (float)s / v; }                        \
	inline V operator+(int s, const V& v) { return (float)s + v; }                        \
	inline V operator+(const V& v, int s) { return v + (float)s; }                        \
	inline V operator-(int s, const V& v) { return (float)s - v; }                        \
	inline V operator-(const V& v, int s) { return v - (float)s; }                        \
	inline V operator*(uint s, const V& v) { return (float)s * v; }                       \
	inline V operator*(const V& v, uint s) { return v * (float)s; }                       \
	inline V operator/(const V& v, uint s) { return v / (float)s; }
GLSLREF_MIXED_OPS(vec2)
GLSLREF_MIXED_OPS(vec3)
GLSLREF_MIXED_OPS(vec4)
#undef GLSLREF_MIXED_OPS
// GLSL's vector == / != give one bool (glm's too); ivec2 == ivec2 is used by the (disabled) logging macros only.

// ------------------------------------------------------------------------------------------------ buffer references
// `layout(buffer_reference) buffer B { T d[]; }` is a raw 64-bit address on the GPU: indexing past the allocation is undefined there
// and, in practice, reads whatever is mapped. The reference does it on purpose-free dead paths (bdpt_commons.glsl:332 forms
// light_vtx(s - 2) with s = 1: `bdpt_path_idx + s - 2` in uint arithmetic, element 0xFFFFFFFF for the first pixel; the value is never
// used). On the CPU such a read must not fault: the harness registers every buffer it binds, a BufArray resolves its element count
// once per invocation, and an out-of-range element is a zeroed dummy (reads give 0, writes are dropped).
struct BufferRegistry {
	struct Range {
		uintptr_t base;
		size_t bytes;
	};
	static constexpr int MAX = 64;
	Range ranges[MAX];
	int count = 0;
	static BufferRegistry& get() {
		static BufferRegistry r;
		return r;
	}
	void add(const void* p, size_t bytes) {  // not thread-safe: called while no stage runs
		for (int i = 0; i < count; i++)
			if (ranges[i].base == (uintptr_t)p) {
				ranges[i].bytes = bytes;
				return;
			}
		if (count < MAX) ranges[count++] = Range{(uintptr_t)p, bytes};
	}
	void remove(const void* p) {
		for (int i = 0; i < count; i++)
			if (ranges[i].base == (uintptr_t)p) ranges[i] = ranges[--count];
	}
	uint64_t elems_from(uint64_t addr, size_t elem) const {
		for (int i = 0; i < count; i++)
			if (addr >= ranges[i].base && addr < ranges[i].base + ranges[i].bytes) return (ranges[i].base + ranges[i].bytes - addr) / elem;
		return addr ? ~0ull : 0ull;  // unregistered: unchecked; null: empty
	}
};
template <class T>
struct BufArray {
	T* p = nullptr;
	uint64_t n = 0;
	BufArray() = default;
	explicit BufArray(uint64_t addr) : p(reinterpret_cast<T*>(addr)), n(BufferRegistry::get().elems_from(addr, sizeof(T))) {}
	T& operator[](uint64_t i) const {
		if (i < n) return p[i];
		static thread_local std::remove_const_t<T> dummy;
		std::memset((void*)&dummy, 0, sizeof(dummy));
		return dummy;
	}
};

// ------------------------------------------------------------------------------------------------ opaque resources
struct image2D {
	float* rgba;  // RGBA32F, row-major
	int width, height;
};
struct sampler2D {
	const void* user;  // harness cookie (texture id of the scene)
	uint32_t id;
};
struct accelerationStructureEXT {
	const void* user;
};

struct Intersection {
	float t, b1, b2;
	uint32_t instance_custom_index;  // gl_InstanceCustomIndexEXT = prim-mesh index (Integrator.cpp:151)
	uint32_t primitive_id;           // gl_PrimitiveID = triangle index inside that mesh
	uint32_t hit;                    // 0 = miss
};

struct Env {
	const void* push_constants = nullptr;
	void* sets[2][8] = {};
	image2D images[8] = {};
	const sampler2D* sampler_arrays[8] = {};
	accelerationStructureEXT tlas = {};
	void* user = nullptr;
	// the parts of the pipeline that live in the Vulkan driver / hardware
	void (*intersect)(const Env*, const float ray8[8], int terminate_on_first_hit, Intersection* out) = nullptr;
	void (*texture)(const Env*, const sampler2D*, const float uv[2], float rgba[4]) = nullptr;
	// the shader binding table: runs closest-hit / miss stages on a payload (programs.cpp)
	void (*trace_ray)(const Env*, uint ray_flags, uint cull_mask, uint sbt_offset, uint sbt_stride, uint miss_index, const vec3& origin,
					  float tmin, const vec3& dir, float tmax, void* payload) = nullptr;

	void* buffer(int set, int binding) const { return sets[set][binding]; }
	image2D image(int, int binding) const { return images[binding]; }
	const sampler2D* samplers(int, int binding) const { return sampler_arrays[binding]; }
	accelerationStructureEXT accel(int, int) const { return tlas; }
};

// ------------------------------------------------------------------------------------------------ scalar promotion
template <class A, class B>
using prom_t = std::conditional_t<std::is_floating_point_v<A> || std::is_floating_point_v<B>, float,
								  std::conditional_t<std::is_unsigned_v<A> || std::is_unsigned_v<B>, uint, int>>;
template <class X>
constexpr bool is_sc = std::is_arithmetic_v<X>;
template <class X>
struct is_glm_vec : std::false_type {};
template <int N, class X, glm::qualifier Q>
struct is_glm_vec<glm::vec<N, X, Q>> : std::true_type {};
template <class X>
constexpr bool is_vec = is_glm_vec<X>::value;

struct Stage {
	const Env* env = nullptr;
	// built-in variables (ray tracing stages)
	uvec3 gl_LaunchIDEXT{0}, gl_LaunchSizeEXT{1};
	int gl_PrimitiveID = 0, gl_InstanceCustomIndexEXT = 0;
	mat4x3 gl_ObjectToWorldEXT{1.0f}, gl_WorldToObjectEXT{1.0f};
	float gl_RayTminEXT = 0, gl_HitTEXT = 0;
	uint gl_HitKindEXT = 0xFEu;  // gl_HitKindFrontFacingTriangleEXT
	vec2 glsl_hit_attribs{0};
	void* glsl_incoming_payload = nullptr;
	static constexpr uint gl_RayFlagsNoneEXT = 0u, gl_RayFlagsOpaqueEXT = 1u, gl_RayFlagsNoOpaqueEXT = 2u,
						  gl_RayFlagsTerminateOnFirstHitEXT = 4u, gl_RayFlagsSkipClosestHitShaderEXT = 8u;

	// everything the pipeline hands to one invocation; set before the stage's own globals are initialised
	struct Inputs {
		const Env* env = nullptr;
		uvec3 launch_id{0}, launch_size{1};
		const Intersection* hit = nullptr;  // closest-hit stages
		float ray_tmin = 0;
		mat4 object_to_world{1.0f}, world_to_object{1.0f};
		void* incoming_payload = nullptr;
	};
	explicit Stage(const Inputs& in)
		: env(in.env), gl_LaunchIDEXT(in.launch_id), gl_LaunchSizeEXT(in.launch_size), gl_ObjectToWorldEXT(in.object_to_world),
		  gl_WorldToObjectEXT(in.world_to_object), gl_RayTminEXT(in.ray_tmin), glsl_incoming_payload(in.incoming_payload) {
		if (in.hit) {
			gl_PrimitiveID = (int)in.hit->primitive_id;
			gl_InstanceCustomIndexEXT = (int)in.hit->instance_custom_index;
			gl_HitTEXT = in.hit->t;
			glsl_hit_attribs = vec2(in.hit->b1, in.hit->b2);
		}
	}

	// ---- scalar built-ins (GLSL genType with implicit int -> float conversion)
	template <class A, class B, std::enable_if_t<is_sc<A> && is_sc<B>, int> = 0>
	static prom_t<A, B> min(A a, B b) {
		const prom_t<A, B> x = (prom_t<A, B>)a, y = (prom_t<A, B>)b;
		return (y < x) ? y : x;
	}
	template <class A, class B, std::enable_if_t<is_sc<A> && is_sc<B>, int> = 0>
	static prom_t<A, B> max(A a, B b) {
		const prom_t<A, B> x = (prom_t<A, B>)a, y = (prom_t<A, B>)b;
		return (x < y) ? y : x;
	}
	template <class A, class B, class C, std::enable_if_t<is_sc<A> && is_sc<B> && is_sc<C>, int> = 0>
	static prom_t<A, prom_t<B, C>> clamp(A x, B lo, C hi) {
		using P = prom_t<A, prom_t<B, C>>;
		return min(max((P)x, (P)lo), (P)hi);
	}
	template <class A, class B, class C, std::enable_if_t<is_sc<A> && is_sc<B> && is_sc<C> && !std::is_same_v<C, bool>, int> = 0>
	static float mix(A x, B y, C a) {
		return (float)x * (1.0f - (float)a) + (float)y * (float)a;
	}
	static float abs(float x) { return std::fabs(x); }
	static int abs(int x) { return x < 0 ? -x : x; }
	static float sqrt(float x) { return std::sqrt(x); }
	static float sqrt(int x) { return std::sqrt((float)x); }
	static float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
	static float sin(float x) { return lmb_sinf(x); }
	static float cos(float x) { return lmb_cosf(x); }
	static float tan(float x) { return lmb_sinf(x) / lmb_cosf(x); }
	static float exp(float x) { return lmb_expf(x); }
	static float exp2(float x) { return lmb_exp2f(x); }
	static float log2(float x) { return lmb_log2f(x); }
	static float log(float x) { return lmb_log2f(x) * 0.693147180559945f; }
	template <class A, class B, std::enable_if_t<is_sc<A> && is_sc<B>, int> = 0>
	static float pow(A x, B y) {
		return lmb_powf((float)x, (float)y);
	}
	static float acos(float x) { return std::acos(x); }
	static float asin(float x) { return std::asin(x); }
	static float atan(float x) { return std::atan(x); }
	static float atan(float y, float x) { return std::atan2(y, x); }
	static float floor(float x) { return std::floor(x); }
	static float ceil(float x) { return std::ceil(x); }
	static float fract(float x) { return x - std::floor(x); }
	static float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
	static float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
	static float radians(float d) { return d * 0.01745329251994329577f; }
	static bool isnan(float x) { return x != x; }
	static bool isinf(float x) { return std::isinf(x); }
	static float length(float x) { return std::fabs(x); }
	static int floatBitsToInt(float f) { return (int)lmb_f2bits(f); }
	static uint floatBitsToUint(float f) { return lmb_f2bits(f); }
	static float intBitsToFloat(int i) { return lmb_bits2f((uint32_t)i); }
	static float uintBitsToFloat(uint u) { return lmb_bits2f(u); }

	// ---- vector built-ins (glm's definitions, component-wise through the scalar ones above where GLSL leaves precision open)
	template <int N, class X>
	static glm::vec<N, X> min(const glm::vec<N, X>& a, const glm::vec<N, X>& b) {
		glm::vec<N, X> r;
		for (int i = 0; i < N; i++) r[i] = min(a[i], b[i]);
		return r;
	}
	template <int N, class X>
	static glm::vec<N, X> max(const glm::vec<N, X>& a, const glm::vec<N, X>& b) {
		glm::vec<N, X> r;
		for (int i = 0; i < N; i++) r[i] = max(a[i], b[i]);
		return r;
	}
	template <int N, class X, class S, std::enable_if_t<is_sc<S>, int> = 0>
	static glm::vec<N, X> min(const glm::vec<N, X>& a, S b) {
		return min(a, glm::vec<N, X>((X)b));
	}
	template <int N, class X, class S, std::enable_if_t<is_sc<S>, int> = 0>
	static glm::vec<N, X> max(const glm::vec<N, X>& a, S b) {
		return max(a, glm::vec<N, X>((X)b));
	}
	template <int N, class X, class S, class U, std::enable_if_t<is_sc<S> && is_sc<U>, int> = 0>
	static glm::vec<N, X> clamp(const glm::vec<N, X>& a, S lo, U hi) {
		return min(max(a, glm::vec<N, X>((X)lo)), glm::vec<N, X>((X)hi));
	}
	template <int N, class X>
	static glm::vec<N, X> clamp(const glm::vec<N, X>& a, const glm::vec<N, X>& lo, const glm::vec<N, X>& hi) {
		return min(max(a, lo), hi);
	}
	template <int N, class S, std::enable_if_t<is_sc<S>, int> = 0>
	static glm::vec<N, float> mix(const glm::vec<N, float>& x, const glm::vec<N, float>& y, S a) {
		return x * (1.0f - (float)a) + y * (float)a;
	}
	template <int N>
	static glm::vec<N, float> mix(const glm::vec<N, float>& x, const glm::vec<N, float>& y, const glm::vec<N, float>& a) {
		return x * (glm::vec<N, float>(1.0f) - a) + y * a;
	}
#define GLSLREF_CW1(name)                                                  \
	template <int N>                                                        \
	static glm::vec<N, float> name(const glm::vec<N, float>& v) {           \
		glm::vec<N, float> r;                                               \
		for (int i = 0; i < N; i++) r[i] = name(v[i]);                      \
		return r;                                                           \
	}
	GLSLREF_CW1(abs)
	GLSLREF_CW1(sqrt)
	GLSLREF_CW1(sin)
	GLSLREF_CW1(cos)
	GLSLREF_CW1(exp)
	GLSLREF_CW1(exp2)
	GLSLREF_CW1(log)
	GLSLREF_CW1(floor)
	GLSLREF_CW1(fract)
	GLSLREF_CW1(sign)
#undef GLSLREF_CW1
	template <int N>
	static glm::vec<N, float> pow(const glm::vec<N, float>& a, const glm::vec<N, float>& b) {
		glm::vec<N, float> r;
		for (int i = 0; i < N; i++) r[i] = pow(a[i], b[i]);
		return r;
	}
	template <int N>
	static glm::vec<N, bool> isnan(const glm::vec<N, float>& v) {
		glm::vec<N, bool> r;
		for (int i = 0; i < N; i++) r[i] = isnan(v[i]);
		return r;
	}
	template <int N>
	static bool any(const glm::vec<N, bool>& v) {
		bool r = false;
		for (int i = 0; i < N; i++) r = r || v[i];
		return r;
	}
	template <int N>
	static bool all(const glm::vec<N, bool>& v) {
		bool r = true;
		for (int i = 0; i < N; i++) r = r && v[i];
		return r;
	}
	template <int N>
	static float dot(const glm::vec<N, float>& a, const glm::vec<N, float>& b) {
		return glm::dot(a, b);
	}
	static vec3 cross(const vec3& a, const vec3& b) { return glm::cross(a, b); }
	template <int N>
	static float length(const glm::vec<N, float>& v) {
		return glm::length(v);
	}
	template <int N>
	static float distance(const glm::vec<N, float>& a, const glm::vec<N, float>& b) {
		return glm::distance(a, b);
	}
	template <int N>
	static glm::vec<N, float> normalize(const glm::vec<N, float>& v) {
		return glm::normalize(v);
	}
	template <int N>
	static glm::vec<N, float> reflect(const glm::vec<N, float>& I, const glm::vec<N, float>& Nn) {
		return glm::reflect(I, Nn);
	}
	static mat4 inverse(const mat4& m) { return glm::inverse(m); }
	static mat4 transpose(const mat4& m) { return glm::transpose(m); }
	static mat3 inverse(const mat3& m) { return glm::inverse(m); }
	static mat3 transpose(const mat3& m) { return glm::transpose(m); }

	// ---- images and textures
	static vec4 imageLoad(const image2D& img, const ivec2& p) {
		const float* q = img.rgba + 4 * ((size_t)p.y * img.width + p.x);
		return vec4(q[0], q[1], q[2], q[3]);
	}
	static void imageStore(const image2D& img, const ivec2& p, const vec4& v) {
		float* q = img.rgba + 4 * ((size_t)p.y * img.width + p.x);
		q[0] = v.x, q[1] = v.y, q[2] = v.z, q[3] = v.w;
	}
	vec4 texture(const sampler2D& s, const vec2& uv) const {
		float rgba[4];
		const float t[2] = {uv.x, uv.y};
		env->texture(env, &s, t, rgba);
		return vec4(rgba[0], rgba[1], rgba[2], rgba[3]);
	}
};

}  // namespace glslref

// traceRayEXT(topLevel, rayFlags, cullMask, sbtRecordOffset, sbtRecordStride, missIndex, origin, Tmin, direction, Tmax, payload location)
#define traceRayEXT(tlas_, flags_, mask_, sbt_off_, sbt_stride_, miss_, o_, tmin_, d_, tmax_, loc_) \
	env->trace_ray(env, (flags_), (mask_), (sbt_off_), (sbt_stride_), (miss_), (o_), (tmin_), (d_), (tmax_), payload_at(loc_))

// commons.h:65-158 defines these for GLSL only (debugPrintfEXT); the C++ side of that header leaves them undefined.
// They have no effect on results (DISABLE_LOGGING gives the same empty expansions, commons.h:139-156).
#define LOG_CLICKED0(str)
#define LOG_CLICKED(str, args)
#define LOG_CLICKED2(str, args1, args2)
#define LOG_CLICKED3(str, args1, args2, args3)
#define LOG_CLICKED4(str, args1, args2, args3, args4)
#define LOG_VAL(str, args, coord)
#define LOG(str, coord)
#define LOG0(str)
#define LOG1(str, val)
#define LOG2(str, val, val2)
#define LOG3(str, val, val2, val3)
#define ASSERT_CLICKED_STR(cond, expected, str, val)
#define ASSERT_CLICKED(cond, expected)
#define ASSERT(cond)
#define ASSERT0(cond, str)
#define ASSERT1(cond, str, val1)
