// Microbenchmark (round 2): does the packed fp32x2 arithmetic of sm_100a (FMUL2 / FADD2 / FFMA2) raise the
// per-SM fp32 rate, or only free issue slots?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
// Each kernel runs `iters` rounds of 8 independent dependency chains per thread; rates are per-lane scalar flops.
#include <cstdio>
#include <cuda_runtime.h>

#define PACK(lo, hi, out) asm("mov.b64 %0, {%1, %2};" : "=l"(out) : "f"(lo), "f"(hi))
#define UNPACK(in, lo, hi) asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(in))

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    // MODE 0: scalar FMUL+FADD pairs; 1: FMUL2+FADD2; 2: scalar FFMA; 3: FFMA2;
    // MODE 4: scalar FMUL+FADD + one integer op per pair; 5: FMUL2+FADD2 + one integer op per pair of pairs
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    unsigned long long X[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) PACK(x[2 * i], x[2 * i + 1], X[i]);
    unsigned long long A, B;
    PACK(a, a, A); PACK(b, b, B);
    unsigned int z = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { x[i] = x[i] * a; x[i] = x[i] + b; if (MODE == 4) z = (z ^ (z << 1)) + i; }
        } else if (MODE == 1 || MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(X[i]) : "l"(X[i]), "l"(A));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(X[i]) : "l"(X[i]), "l"(B));
                if (MODE == 5) { z = (z ^ (z << 1)) + i; z = (z ^ (z << 1)) + i + 1; }
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(X[i]) : "l"(X[i]), "l"(A), "l"(B));
        }
    }
    float s = 0.f;
    if (MODE == 1 || MODE == 3 || MODE == 5) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float lo, hi; UNPACK(X[i], lo, hi); s += lo + hi; }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) s += x[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)z;
}

template <int MODE>
void run(const char* name, int flops_per_elem_iter) {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    float* out; cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    k<MODE><<<blocks, threads>>>(out, 100, 1.0001f, 0.0001f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * 16.0 * flops_per_elem_iter;
    printf("%-44s %8.3f ms  %7.2f T scalar-ops/s  (%.1f ops/clk/SM at %d MHz)\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000);
    cudaFree(out);
}

int main() {
    run<0>("scalar FMUL + FADD", 2);
    run<1>("packed FMUL2 + FADD2", 2);
    run<2>("scalar FFMA (counted as 1 op)", 1);
    run<3>("packed FFMA2 (counted as 1 op per half)", 1);
    run<4>("scalar FMUL + FADD + 2 int ops per element", 2);
    run<5>("packed FMUL2 + FADD2 + 2 int ops per element", 2);
    return 0;
}
