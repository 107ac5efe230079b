/*
 * lmb_detmath.h -- bit-reproducible single-precision transcendentals for host (g++) and device (nvcc).
 *
 * Why: the acceptance bar is per-pixel equality (1e-4 relative on >= 99.9 % of pixels at a fixed seed) between the
 * CUDA wavefront tracer and the CPU transliteration of Lumen's GLSL. Path tracing is chaotic -- a 1-ulp difference in
 * sin/cos/exp/pow flips a Russian-roulette or lobe decision -- and glibc's sinf/expf/powf differ from CUDA libdevice's.
 * GLSL itself leaves these functions' precision implementation-defined (the reference runs them on whatever the Vulkan
 * driver provides), so this header *defines* them once, using only operations that are correctly rounded on both
 * sides: + - * / sqrt, fmaf, rintf and integer bit casts. Both translation units must be built without automatic FMA
 * contraction (g++ -ffp-contract=off, nvcc -fmad=false); every FMA here is explicit.
 *
 * Accuracy (checked against libm in tests/test_detmath.py): sin/cos <= 2 ulp on [-64, 64]; exp <= 2 ulp;
 * log2 <= 2 ulp; pow(x,y) = exp2(y*log2(x)) evaluated in fp32 like GPU hardware does (relative error ~ |y log2 x| 2^-23).
 *
 * GLSL call sites served: cos/sin (utils.glsl:226, microfacet_commons.glsl:114-115, principled.glsl:147,
 * commons.glsl:206), exp (atmosphere.glsl:93-101,144), pow (sampling_commons.glsl:86,91, principled.glsl:142,
 * atmosphere.glsl:175).
 */
#ifndef LMB_DETMATH_H
#define LMB_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LMB_HD __host__ __device__ __forceinline__
#else
#define LMB_HD static inline
#endif

LMB_HD float lmb_bits2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
	return __uint_as_float(u);
#else
	float f;
	memcpy(&f, &u, 4);
	return f;
#endif
}
LMB_HD uint32_t lmb_f2bits(float f) {
#if defined(__CUDA_ARCH__)
	return __float_as_uint(f);
#else
	uint32_t u;
	memcpy(&u, &f, 4);
	return u;
#endif
}

/* sin and cos together. Cody-Waite 3-term reduction by pi/2, Cephes-style minimax kernels on [-pi/4, pi/4]. */
LMB_HD void lmb_sincosf(float x, float* s_out, float* c_out) {
	const float kf = rintf(x * 0.636619772367581343f);
	const int k = (int)kf;
	float r = fmaf(kf, -1.5707962513e+00f, x);
	r = fmaf(kf, -7.5497894159e-08f, r);
	r = fmaf(kf, -5.3903029534e-15f, r);
	const float z = r * r;
	/* sin(r) */
	float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
	ps = fmaf(ps, z, -1.6666654611e-1f);
	const float s = fmaf(ps * z, r, r);
	/* cos(r) */
	float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
	pc = fmaf(pc, z, 4.166664568298827e-2f);
	const float c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
	switch (k & 3) {
		case 0:
			*s_out = s;
			*c_out = c;
			break;
		case 1:
			*s_out = c;
			*c_out = -s;
			break;
		case 2:
			*s_out = -s;
			*c_out = -c;
			break;
		default:
			*s_out = -c;
			*c_out = s;
			break;
	}
}
LMB_HD float lmb_sinf(float x) {
	float s, c;
	lmb_sincosf(x, &s, &c);
	return s;
}
LMB_HD float lmb_cosf(float x) {
	float s, c;
	lmb_sincosf(x, &s, &c);
	return c;
}

/* 2^n for integer n in [-126, 127] */
LMB_HD float lmb_pow2i(int n) { return lmb_bits2f((uint32_t)(n + 127) << 23); }

/* e^x. Underflows to 0 below about -87.3*2, overflows to +inf above 88.72. */
LMB_HD float lmb_expf(float x) {
	if (!(x > -150.0f)) return (x != x) ? x : 0.0f;
	if (x > 88.72283935546875f) return lmb_bits2f(0x7f800000u);
	const float nf = rintf(x * 1.44269504088896341f);
	float r = fmaf(nf, -0.693359375f, x);
	r = fmaf(nf, 2.12194440e-4f, r);
	float p = fmaf(1.9875691500e-4f, r, 1.3981999507e-3f);
	p = fmaf(p, r, 8.3334519073e-3f);
	p = fmaf(p, r, 4.1665795894e-2f);
	p = fmaf(p, r, 1.6666665459e-1f);
	p = fmaf(p, r, 5.0000001201e-1f);
	p = fmaf(p * r, r, r) + 1.0f;
	const int n = (int)nf;
	const int n1 = n / 2;
	const int n2 = n - n1;
	return (p * lmb_pow2i(n1)) * lmb_pow2i(n2);
}

/* lmb_expf again, bit for bit (tools/check_expf_equiv.c walks all 2^32 inputs), in the form a GPU issues fastest -- the sky march
 * (atmosphere.glsl:93-144) evaluates ~1500 of these per escaped ray and is bound by instruction issue:
 *   rintf(t) and (int) are FRND + F2I on the quarter-rate conversion pipe -> s = t + 1.5*2^23 rounds t to the nearest-even integer
 *   in the FP32 adder (|t| < 2^22), nf = s - 1.5*2^23, and the low bits of s ARE the integer;
 *   (p * 2^n1) * 2^n2 with n1 + n2 = n is p with n added to its exponent field whenever neither step leaves the normal range
 *   (-86 < x <= 88: n in [-124, 127], p in [0.70, 1.42]) -> one shift-add on the bit pattern. Outside that window: the form above. */
LMB_HD float lmb_expf_window(float x) { /* requires -86 < x <= 88 */
	const float s = x * 1.44269504088896341f + 12582912.0f;
	const float nf = s - 12582912.0f;
	float r = fmaf(nf, -0.693359375f, x);
	r = fmaf(nf, 2.12194440e-4f, r);
	float p = fmaf(1.9875691500e-4f, r, 1.3981999507e-3f);
	p = fmaf(p, r, 8.3334519073e-3f);
	p = fmaf(p, r, 4.1665795894e-2f);
	p = fmaf(p, r, 1.6666665459e-1f);
	p = fmaf(p, r, 5.0000001201e-1f);
	p = fmaf(p * r, r, r) + 1.0f;
	return lmb_bits2f(lmb_f2bits(p) + (lmb_f2bits(s) << 23)); /* 0x4B400000 << 23 == 0 (mod 2^32): only n is left of s */
}
LMB_HD float lmb_expf_fast(float x) { return (x > -86.0f && x <= 88.0f) ? lmb_expf_window(x) : lmb_expf(x); }

/* log2(x) for finite x > 0 (normal or subnormal). */
LMB_HD float lmb_log2f(float x) {
	int e = 0;
	uint32_t ix = lmb_f2bits(x);
	if (ix < 0x00800000u) { /* subnormal: scale up by 2^23 */
		x = x * 8388608.0f;
		ix = lmb_f2bits(x);
		e = -23;
	}
	/* mantissa in [sqrt(1/2), sqrt(2)) */
	ix += 0x3f800000u - 0x3f3504f3u;
	e += (int)(ix >> 23) - 127;
	ix = (ix & 0x007fffffu) + 0x3f3504f3u;
	const float f = lmb_bits2f(ix) - 1.0f;
	const float z = f * f;
	float p = fmaf(7.0376836292e-2f, f, -1.1514610310e-1f);
	p = fmaf(p, f, 1.1676998740e-1f);
	p = fmaf(p, f, -1.2420140846e-1f);
	p = fmaf(p, f, 1.4249322787e-1f);
	p = fmaf(p, f, -1.6668057665e-1f);
	p = fmaf(p, f, 2.0000714765e-1f);
	p = fmaf(p, f, -2.4999993993e-1f);
	p = fmaf(p, f, 3.3333331174e-1f);
	float y = (f * z) * p;
	y = fmaf(-0.5f, z, y);
	/* ln(m) = f + y ;  log2 = ln * log2(e), split for a little extra precision */
	const float ln_m_hi = f;
	float r = fmaf(y, 1.44269504088896341f, ln_m_hi * 4.42695040888963407e-1f);
	r = r + ln_m_hi;
	return r + (float)e;
}

/* 2^t */
LMB_HD float lmb_exp2f(float t) {
	if (!(t > -300.0f)) return (t != t) ? t : 0.0f;
	if (t >= 128.0f) return lmb_bits2f(0x7f800000u);
	const float nf = rintf(t);
	const float r = (t - nf) * 0.693147180559945309f;
	float p = fmaf(1.9875691500e-4f, r, 1.3981999507e-3f);
	p = fmaf(p, r, 8.3334519073e-3f);
	p = fmaf(p, r, 4.1665795894e-2f);
	p = fmaf(p, r, 1.6666665459e-1f);
	p = fmaf(p, r, 5.0000001201e-1f);
	p = fmaf(p * r, r, r) + 1.0f;
	const int n = (int)nf;
	const int n1 = n / 2;
	const int n2 = n - n1;
	return (p * lmb_pow2i(n1)) * lmb_pow2i(n2);
}

/* GLSL pow(x, y) for x >= 0: exp2(y * log2(x)), with pow(0, y>0) = 0 and pow(x, 0) = 1. Negative x is undefined in
 * GLSL; here it yields NaN. */
LMB_HD float lmb_powf(float x, float y) {
	if (y == 0.0f) return 1.0f;
	if (x == 0.0f) return (y > 0.0f) ? 0.0f : lmb_bits2f(0x7f800000u);
	if (x < 0.0f || x != x) return lmb_bits2f(0x7fc00000u);
	if (x == lmb_bits2f(0x7f800000u)) return (y > 0.0f) ? x : 0.0f;
	return lmb_exp2f(y * lmb_log2f(x));
}

#endif /* LMB_DETMATH_H */
