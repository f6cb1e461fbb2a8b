// PCIe copy bandwidth, both directions at once, pinned host memory: contiguous 1-D copies against the 2-D copies the
// streamer issues (host rows of ny + 2 = 8194 floats, device pitch 8224 floats: every host row starts 8 bytes further
// off a 64-byte boundary).  nvcc -O2 -o pcie_2d pcie_2d.cu
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    const size_t w = 8194 * 4, dp = 8224 * 4, rows = 8194;
    float *hin, *hout, *din, *dout;
    cudaHostAlloc(&hin, w * rows, cudaHostAllocDefault); cudaHostAlloc(&hout, w * rows, cudaHostAllocDefault);
    cudaMalloc(&din, dp * rows); cudaMalloc(&dout, dp * rows);
    cudaMemset(din, 0, dp * rows); cudaMemset(dout, 0, dp * rows);
    for (size_t i = 0; i < w * rows / 4; ++i) { hin[i] = 1.0f; hout[i] = 0.0f; }
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int slabs = 16, reps = 4;
    for (int mode = 0; mode < 4; ++mode) {      // 0: 1-D both ways, 1: 2-D both ways, 2: 2-D up only, 3: 2-D down only
        cudaDeviceSynchronize();
        cudaEventRecord(e0, 0);
        for (int r = 0; r < reps; ++r)
            for (int s = 0; s < slabs; ++s) {
                const size_t r0 = rows * s / slabs, nr = rows * (s + 1) / slabs - r0;
                if (mode == 0) {
                    cudaMemcpyAsync((char*)din + r0 * w, (char*)hin + r0 * w, nr * w, cudaMemcpyHostToDevice, s1);
                    cudaMemcpyAsync((char*)hout + r0 * w, (char*)dout + r0 * w, nr * w, cudaMemcpyDeviceToHost, s2);
                } else {
                    if (mode != 3) cudaMemcpy2DAsync((char*)din + r0 * dp, dp, (char*)hin + r0 * w, w, w, nr, cudaMemcpyHostToDevice, s1);
                    if (mode != 2) cudaMemcpy2DAsync((char*)hout + r0 * w, w, (char*)dout + r0 * dp, dp, w, nr, cudaMemcpyDeviceToHost, s2);
                }
            }
        cudaEventRecord(e1, 0);      // legacy default stream: waits for s1 and s2
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const char* names[] = {"1-D both ways", "2-D both ways", "2-D up only", "2-D down only"};
        printf("%s: %.1f GB/s per direction (%.2f ms per 268.5 MB field)\n", names[mode], reps * w * rows / ms / 1e6, ms / reps);
    }
    return 0;
}
