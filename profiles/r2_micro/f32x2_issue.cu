// Microbenchmark (round 2): an FFMA2 keeps the fp32 pipe busy for two cycles -- can the scheduler issue other
// instructions (integer ALU, shuffles, predicate ops) in the second cycle?  Per iteration every thread issues
// 8 independent FFMA2 (or 16 scalar FFMA) plus M integer LOP3/IADD ops, M = 0, 8, 16, 24, 32.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define PACK(lo, hi, out) asm("mov.b64 %0, {%1, %2};" : "=l"(out) : "f"(lo), "f"(hi))
#define UNPACK(in, lo, hi) asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(in))

template <int PACKED, int M>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    unsigned long long X[8], A, B;
#pragma unroll
    for (int i = 0; i < 8; ++i) PACK(x[2 * i], x[2 * i + 1], X[i]);
    PACK(a, a, A); PACK(b, b, B);
    unsigned int z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(X[i]) : "l"(X[i]), "l"(A), "l"(B));
            else { x[2 * i] = __fmaf_rn(x[2 * i], a, b); x[2 * i + 1] = __fmaf_rn(x[2 * i + 1], a, b); }
#pragma unroll
            for (int m = 0; m < M / 8; ++m) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[(i + m) & 7]) : "r"(z[(i + m + 3) & 7]), "r"(it));
        }
    }
    float s = 0.f;
    if (PACKED) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float lo, hi; UNPACK(X[i], lo, hi); s += lo + hi; }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) s += x[i];
    }
    unsigned int zz = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) zz ^= z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)zz;
}

template <int PACKED, int M>
void run() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    float* out; cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    k<PACKED, M><<<blocks, threads>>>(out, 100, 1.0001f, 0.0001f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<PACKED, M><<<blocks, threads>>>(out, iters, 1.0001f, 0.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // cycles per iteration per SM sub-partition: warps per SMSP = 8 blocks * 8 warps / 4
    const double cyc = ms * 1e-3 * p.clockRate * 1e3 / iters / (8.0 * 8 / 4);
    printf("%s fma x16 lanes-ops + %2d int ops / iteration: %7.3f ms, %5.1f issue cycles per warp-iteration (fp pipe needs 16)\n",
           PACKED ? "8 FFMA2 " : "16 FFMA ", M, ms, cyc);
    cudaFree(out);
}

int main() {
    run<0, 0>(); run<0, 8>(); run<0, 16>(); run<0, 32>();
    run<1, 0>(); run<1, 8>(); run<1, 16>(); run<1, 24>(); run<1, 32>();
    return 0;
}
