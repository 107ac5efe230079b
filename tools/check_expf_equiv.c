// Exhaustive check (all 2^32 bit patterns) that lmb_expf_fast == lmb_expf bit for bit (include/lmb_detmath.h).
//   gcc -O2 -fopenmp -ffp-contract=off -mfma -I../include check_expf_equiv.c -lm && ./a.out
// NaN inputs compare by bit pattern too (both forms return the input). -DSTRIDE=n walks every n-th bit pattern (tests/test_detmath.py).
#include <stdio.h>
#include <omp.h>
#include "lmb_detmath.h"
#ifndef STRIDE
#define STRIDE 1
#endif
int main() {
	unsigned long long bad = 0, fast = 0;
#pragma omp parallel for reduction(+ : bad, fast) schedule(static)
	for (long long u = 0; u < (1ll << 32); u += STRIDE) {
		const float x = lmb_bits2f((uint32_t)u);
		const uint32_t a = lmb_f2bits(lmb_expf(x)), b = lmb_f2bits(lmb_expf_fast(x));
		if (x > -86.0f && x <= 88.0f) fast++;
		if (a != b) {
			bad++;
			if (bad < 8) printf("x = %.9g (0x%08x): %08x vs %08x\n", x, (uint32_t)u, a, b);
		}
	}
	printf("mismatches = %llu of %lld inputs (%llu take the short form)\n", bad, ((1ll << 32) + STRIDE - 1) / STRIDE, fast);
	return bad != 0;
}
