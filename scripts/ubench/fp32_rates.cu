// Micro-benchmark: issue rates of un-fused FMUL+FADD, FFMA and packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rates fp32_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = (threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { x[i] = __fmul_rn(x[i], a); x[i] = __fadd_rn(x[i], b); }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { x[i] = __fmaf_rn(x[i], a, b); x[i] = __fmaf_rn(x[i], a, b); }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v, aa, bb;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
                asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
        }
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}

template <int MODE> double run(float* d, int lane_ops_per_iter) {
    const int iters = 1 << 13, blocks = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, 0.999999f, 1e-7f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    return (double)blocks * 256 * iters * lane_ops_per_iter / (best * 1e-3) / 1e12;
}

int main() {
    float* d; cudaMalloc(&d, 64);
    printf("un-fused FMUL+FADD : %.2f T lane-instr/s\n", run<0>(d, 32));
    printf("FFMA               : %.2f T lane-instr/s (x2 flop)\n", run<1>(d, 32));
    printf("FFMA2 (f32x2)      : %.2f T lane-fma/s   (packed: 2 fma per lane per instr)\n", run<2>(d, 64));
    return 0;
}
