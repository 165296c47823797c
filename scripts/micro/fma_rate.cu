// Micro-benchmark: issue rate of FFMA vs packed FFMA2 (fma.rn.f32x2) vs DFMA on sm_100a (development aid).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float x[16]; double y[8]; float2 z[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { y[i] = x[i]; z[i] = make_float2(x[2 * i], x[2 * i + 1]); }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = __ffma2_rn(z[i], a2, b2);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fma(y[i], (double)a, (double)b);
        } else {   // mixed: 8 FFMA2 + 8 integer ops per iteration: does FFMA2 leave issue slots free?
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = __ffma2_rn(z[i], a2, b2);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __int_as_float(__float_as_int(x[i]) * 3 + it);
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += x[i];
    for (int i = 0; i < 8; ++i) s += (float)y[i] + z[i].x + z[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    const char* names[] = {"FFMA x16", "FFMA2 x8", "DFMA x8", "FFMA2 x8 + IMAD x8"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
            if (mode == 1) k<1><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
            if (mode == 2) k<2><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
            if (mode == 3) k<3><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fmas = (double)148 * 8 * 256 * iters * (mode == 2 ? 8 : 16);
        printf("%-22s: %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", names[mode], ms, fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.9e9);
    }
    return 0;
}
