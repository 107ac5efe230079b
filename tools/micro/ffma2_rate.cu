// Issue / pipe rate of FFMA2 (fma.rn.f32x2, sm_100) against FFMA on a B200: same number of fused multiply-adds, half the warp instructions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_rate tools/micro/ffma2_rate.cu && /tmp/ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
template <bool PACKED>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
	float2 x[8];
	for (int j = 0; j < 8; j++) x[j] = make_float2(threadIdx.x * 1e-3f + j, blockIdx.x * 1e-3f - j);
	const float2 A = make_float2(a, a), B = make_float2(b, b);
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int j = 0; j < 8; j++) {
			if (PACKED) x[j] = __ffma2_rn(x[j], A, B);
			else x[j].x = fmaf(x[j].x, a, b), x[j].y = fmaf(x[j].y, a, b);
		}
	}
	float s = 0;
	for (int j = 0; j < 8; j++) s += x[j].x + x[j].y;
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
	float* out;
	const int blocks = 148 * 8, iters = 1 << 14;
	cudaMalloc(&out, blocks * 256 * 4);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0), cudaEventCreate(&e1);
	for (int packed = 0; packed < 2; packed++) {
		for (int rep = 0; rep < 2; rep++) {
			cudaEventRecord(e0);
			if (packed) k<true><<<blocks, 256>>>(out, 0.999f, 1e-3f, iters);
			else k<false><<<blocks, 256>>>(out, 0.999f, 1e-3f, iters);
			cudaEventRecord(e1);
			cudaEventSynchronize(e1);
			float ms;
			cudaEventElapsedTime(&ms, e0, e1);
			const double fma = (double)blocks * 256 * iters * 16;
			if (rep) printf("%s: %.3f ms, %.1f T fused multiply-adds / s\n", packed ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9);
		}
	}
	return 0;
}
