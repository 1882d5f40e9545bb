// Issue rate of scalar FFMA against packed FFMA2 (fma.rn.f32x2, sm_100a): does one FFMA2 cost one issue slot
// or two?  Register-only loops, 16 independent chains per thread, every SM filled.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(256) scalar_kernel(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = (float)(threadIdx.x + k) * 1e-3f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], a, b);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) packed_kernel(float* out, int iters, float a, float b) {
    unsigned long long x[16], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        float v = (float)(threadIdx.x + k) * 1e-3f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x[k]) : "f"(v), "f"(v + 1.0f));
    }
#pragma unroll 1
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 16; ++k) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(aa), "l"(bb));
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[k])); s += lo + hi; }
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float* out;
    cudaMalloc(&out, 1 << 24);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) scalar_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
            else packed_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double instr = (double)blocks * threads * iters * 64.0;          // thread-level instructions
        const double flops = instr * (which ? 4.0 : 2.0);
        printf("%s: %.3f ms  %.1f G thread-instr/s  %.1f TFLOP/s\n", which ? "FFMA2 (packed)" : "FFMA  (scalar)", best, instr / best / 1e6, flops / best / 1e9);
    }
    return 0;
}
