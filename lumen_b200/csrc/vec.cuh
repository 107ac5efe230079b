// vec.cuh -- minimal fp32 vector algebra for the device code.
//
// Operation ORDER is part of the contract: the CUDA path must reproduce the CPU transliteration of Lumen's GLSL bit for
// bit, so every helper spells out the same expression tree GLSL/glm use (dot = (x*x' + y*y') + z*z',
// normalize = v * (1/sqrt(dot)), reflect = I - N*dot(N,I)*2, mix = x*(1-a) + y*a, min/max as ternaries with GLSL/glm
// NaN behaviour). The TU is compiled with -fmad=false; FMAs appear only where written explicitly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LMB_D __device__ __forceinline__

namespace lmb {

struct V2 {
	float x, y;
};
struct V3 {
	float x, y, z;
};
struct V4 {
	float x, y, z, w;
};
struct M4 {  // column-major, c[col]
	V4 c[4];
};

LMB_D V2 v2(float x, float y) { return V2{x, y}; }
LMB_D V3 v3(float x, float y, float z) { return V3{x, y, z}; }
LMB_D V3 v3(float s) { return V3{s, s, s}; }
LMB_D V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
LMB_D V4 v4(float x, float y, float z, float w) { return V4{x, y, z, w}; }
LMB_D V4 v4(const V3& a, float w) { return V4{a.x, a.y, a.z, w}; }
LMB_D V3 xyz(const V4& a) { return V3{a.x, a.y, a.z}; }

LMB_D V3 operator+(const V3& a, const V3& b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
LMB_D V3 operator-(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
LMB_D V3 operator*(const V3& a, const V3& b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
LMB_D V3 operator/(const V3& a, const V3& b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
LMB_D V3 operator*(const V3& a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
LMB_D V3 operator*(float s, const V3& a) { return V3{s * a.x, s * a.y, s * a.z}; }
LMB_D V3 operator/(const V3& a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
LMB_D V3 operator+(const V3& a, float s) { return V3{a.x + s, a.y + s, a.z + s}; }
LMB_D V3 operator-(const V3& a, float s) { return V3{a.x - s, a.y - s, a.z - s}; }
LMB_D V3 operator-(float s, const V3& a) { return V3{s - a.x, s - a.y, s - a.z}; }
LMB_D V3 operator-(const V3& a) { return V3{-a.x, -a.y, -a.z}; }
LMB_D V3& operator+=(V3& a, const V3& b) {
	a = a + b;
	return a;
}
LMB_D V3& operator-=(V3& a, const V3& b) {
	a = a - b;
	return a;
}
LMB_D V3& operator*=(V3& a, const V3& b) {
	a = a * b;
	return a;
}
LMB_D V3& operator*=(V3& a, float s) {
	a = a * s;
	return a;
}
LMB_D V3& operator/=(V3& a, float s) {
	a = a / s;
	return a;
}

LMB_D V2 operator+(const V2& a, const V2& b) { return V2{a.x + b.x, a.y + b.y}; }
LMB_D V2 operator-(const V2& a, const V2& b) { return V2{a.x - b.x, a.y - b.y}; }
LMB_D V2 operator*(const V2& a, const V2& b) { return V2{a.x * b.x, a.y * b.y}; }
LMB_D V2 operator/(const V2& a, const V2& b) { return V2{a.x / b.x, a.y / b.y}; }
LMB_D V2 operator*(const V2& a, float s) { return V2{a.x * s, a.y * s}; }
LMB_D V2 operator*(float s, const V2& a) { return V2{s * a.x, s * a.y}; }
LMB_D V2 operator/(const V2& a, float s) { return V2{a.x / s, a.y / s}; }
LMB_D V2 operator+(const V2& a, float s) { return V2{a.x + s, a.y + s}; }
LMB_D V2 operator-(const V2& a, float s) { return V2{a.x - s, a.y - s}; }

LMB_D V4 operator+(const V4& a, const V4& b) { return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
LMB_D V4 operator*(const V4& a, float s) { return V4{a.x * s, a.y * s, a.z * s, a.w * s}; }

// glm::min / glm::max / glm::clamp (func_common.inl:17-35, 649-653)
LMB_D float gmin(float x, float y) { return (y < x) ? y : x; }
LMB_D float gmax(float x, float y) { return (x < y) ? y : x; }
LMB_D float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
LMB_D V2 gclamp(const V2& v, float lo, float hi) { return V2{gclamp(v.x, lo, hi), gclamp(v.y, lo, hi)}; }
LMB_D V3 gmax(const V3& a, float s) { return V3{gmax(a.x, s), gmax(a.y, s), gmax(a.z, s)}; }

LMB_D float dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
LMB_D float dot(const V2& a, const V2& b) { return a.x * b.x + a.y * b.y; }
LMB_D V3 cross(const V3& x, const V3& y) { return V3{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
LMB_D float length(const V3& a) { return sqrtf(dot(a, a)); }
LMB_D V3 normalize(const V3& a) { return a * (1.0f / sqrtf(dot(a, a))); }
LMB_D V3 reflect(const V3& I, const V3& N) { return I - N * dot(N, I) * 2.0f; }
LMB_D float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
LMB_D V3 mix(const V3& x, const V3& y, float a) { return x * (1.0f - a) + y * a; }
LMB_D V3 vsqrt(const V3& a) { return V3{sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)}; }
LMB_D float gsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// glm mat4 * vec4 (type_mat4x4.inl:562-573): (c0*x + c1*y) + (c2*z + c3*w)
LMB_D V4 mul(const M4& m, const V4& v) { return (m.c[0] * v.x + m.c[1] * v.y) + (m.c[2] * v.z + m.c[3] * v.w); }
LMB_D M4 load_m4(const float* p) {
	M4 m;
#pragma unroll
	for (int c = 0; c < 4; c++) m.c[c] = V4{p[4 * c + 0], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]};
	return m;
}
LMB_D M4 transpose(const M4& a) {
	M4 t;
	t.c[0] = V4{a.c[0].x, a.c[1].x, a.c[2].x, a.c[3].x};
	t.c[1] = V4{a.c[0].y, a.c[1].y, a.c[2].y, a.c[3].y};
	t.c[2] = V4{a.c[0].z, a.c[1].z, a.c[2].z, a.c[3].z};
	t.c[3] = V4{a.c[0].w, a.c[1].w, a.c[2].w, a.c[3].w};
	return t;
}
// vec3 * mat4x3 (row vector): component i = dot(v, column i)
LMB_D V3 mul_row(const V3& v, const M4& m) { return V3{dot(v, xyz(m.c[0])), dot(v, xyz(m.c[1])), dot(v, xyz(m.c[2]))}; }

LMB_D float comp(const V3& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

}  // namespace lmb
