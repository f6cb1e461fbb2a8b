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
