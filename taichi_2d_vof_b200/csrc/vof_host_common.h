// Host-side helpers shared by the 2-D and 3-D translation units of libvof (errors, launch checks, constants).
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "vof2d_jacobi_tb.cuh"
#include "vof_common.cuh"

namespace vofhost {

inline thread_local char g_err[512] = "";
inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int launch_ok(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return VOF_OK;
}
inline vof::ConstDiv make_const_div(float b) {
    vof::ConstDiv d;
    d.b = b; d.bd = (double)b;
    d.rd = 1.0 / d.bd;            // RN64(1/b)
    d.r = (float)d.rd;            // RN32(1/b) up to double rounding; the exhaustive device check is the proof
    return d;
}
inline void node_coords(std::vector<float>& x, int n, double L) {
    // np.hstack((0.0, np.linspace(0, L, n + 1), L)).astype(float32), 2dvof.py:43-46
    x.assign((size_t)n + 3, 0.0f);
    const double step = L / n;
    for (int k = 0; k <= n; ++k) x[(size_t)k + 1] = (float)(k == n ? L : k * step);
    x[0] = 0.0f;
    x[(size_t)n + 2] = (float)L;
}


}  // namespace vofhost

#define CU(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return vofhost::fail((int)e_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CHECK_CTX(c) \
    do { if (!(c)) return vofhost::fail(VOF_EINVAL, "null context"); } while (0)
#define TRY(x) do { int rc_ = (x); if (rc_ != VOF_OK) return rc_; } while (0)

// Non-stalling device -> host read of one field (the -s / VTK output path, 2dvof.py:565, 3dvof.py:627): the field is
// snapshotted device-to-device on the compute stream (268 MB at 8192^2: ~0.1 ms), and the snapshot travels to the
// (pinned) host buffer on a side stream while the time loop goes on.  One read in flight per context.
namespace vofhost {
struct AsyncGet {
    float* snap = nullptr;
    size_t snap_bytes = 0;
    cudaStream_t side = nullptr;
    cudaEvent_t ready = nullptr, done = nullptr;
    bool pending = false;
};
// src_row0: start of the pitched field (first row, pad columns included); the host array is dense rows of width_bytes
inline int async_get_begin(AsyncGet& a, cudaStream_t compute, const float* src_row0, size_t field_bytes, size_t pitch_bytes,
                           size_t width_bytes, size_t rows, size_t col_off_bytes, float* host_dst) {
    if (!a.side) {
        CU(cudaStreamCreateWithFlags(&a.side, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&a.ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming));
    }
    if (a.snap_bytes < field_bytes) {
        if (a.pending) CU(cudaEventSynchronize(a.done));
        if (a.snap) cudaFree(a.snap);
        a.snap = nullptr; a.snap_bytes = 0;
        if (cudaMalloc((void**)&a.snap, field_bytes) != cudaSuccess)
            return fail(VOF_ENOMEM, "cudaMalloc(%zu bytes) for the output snapshot failed", field_bytes);
        a.snap_bytes = field_bytes;
    }
    if (a.pending) CU(cudaStreamWaitEvent(compute, a.done, 0));     // the previous read still owns the snapshot
    CU(cudaMemcpyAsync(a.snap, src_row0, field_bytes, cudaMemcpyDeviceToDevice, compute));
    CU(cudaEventRecord(a.ready, compute));
    CU(cudaStreamWaitEvent(a.side, a.ready, 0));
    CU(cudaMemcpy2DAsync(host_dst, width_bytes, (const char*)a.snap + col_off_bytes, pitch_bytes, width_bytes, rows,
                         cudaMemcpyDeviceToHost, a.side));
    CU(cudaEventRecord(a.done, a.side));
    a.pending = true;
    return VOF_OK;
}
inline int async_get_wait(AsyncGet& a) {
    if (a.pending) { CU(cudaEventSynchronize(a.done)); a.pending = false; }
    return VOF_OK;
}
inline void async_get_free(AsyncGet& a) {
    if (a.pending) cudaEventSynchronize(a.done);
    if (a.snap) cudaFree(a.snap);
    if (a.ready) cudaEventDestroy(a.ready);
    if (a.done) cudaEventDestroy(a.done);
    if (a.side) cudaStreamDestroy(a.side);
    a = AsyncGet();
}
}  // namespace vofhost
