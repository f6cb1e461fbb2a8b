// Shared definitions for the libvof CUDA translation units (sm_100a only).
//
// Arithmetic contract: every kernel evaluates the reference's expressions
// (/root/reference/2dvof.py) in IEEE fp32, left to right, with NO FMA contraction
// (this directory is compiled with -fmad=false) and correctly rounded / and sqrt
// (-prec-div=true -prec-sqrt=true, the nvcc defaults).  Where an FMA is wanted for
// speed AND proven value-identical it is written explicitly with __fmaf_rn.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vof.h"

namespace vof {

// Column offset of logical j = 0 inside a pitched row: logical j = 1 (first interior
// column) sits on a 128-byte boundary so interior float4 / warp accesses are aligned.
constexpr int kColOff = 31;
constexpr int kPitchAlign = 32;   // floats

// fp32 constants exactly as the reference's kernels see them (double-folded Python
// scalars rounded once).
struct Consts {
    float dt, dx, dy, dxi, dyi, dxi2, dyi2, dxdy, dtdy, dtdx;
    float m1_2dx, m1_2dy;    // -1/(2*dx), -1/(2*dy)          2dvof.py:287-294
    float i_dx_2, i_dy_2;    // 1/dx/2, 1/dy/2                2dvof.py:308-309
    float neg_sigma;         // -sigma[None]                  2dvof.py:213
    float rho_l, rho_g, nu_l, nu_g, gx, gy;
    float cflx, cfly;        // 0.25*dx, 0.25*dy              2dvof.py:274, 279
    float one, neg_zero;     // 1.0f and -0.0f as RUN-TIME values: operands of the packed fp32x2 forms that ptxas must not
                             // recognise as a plain add / multiply (it would contract the pair into one FFMA2, see Pk2)
};

// Geometry of one context's local arrays.  Local row l holds global row gi0 + l.
struct Grid {
    int nx, ny;      // global interior size
    int gi0;         // global i of local row 0 (0 for a full-domain context)
    int nrows;       // local rows allocated
    int pitch;       // floats per row
};

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- order-exact scalar helpers -------------------------------------------------
// 2dvof.py:192-195  var(a,b,c) = a+b+c - max(a,b,c) - min(a,b,c)
__device__ __forceinline__ float var3(float a, float b, float c) {
    float s = (a + b) + c;
    float mx = fmaxf(fmaxf(a, b), c);
    float mn = fminf(fminf(a, b), c);
    return (s - mx) - mn;
}
// var(0, 1, x) and var(x, 0, 1) are the same function of x: (1+x) - max(1,x) - min(0,x)
__device__ __forceinline__ float var01(float x) {
    return ((1.0f + x) - fmaxf(1.0f, x)) - fminf(0.0f, x);
}
// 2dvof.py:201-203
__device__ __forceinline__ float rho_of(float F, const Consts& c) {
    float f = var01(F);
    return c.rho_g * (1.0f - f) + c.rho_l * f;
}
__device__ __forceinline__ float nu_of(float F, const Consts& c) {
    float f = var01(F);
    return c.nu_l * f + c.nu_g * (1.0f - f);
}

// IEEE quotient with the zero numerator taken out: nvcc's division sends t = 0 (ubiquitous while the pressure front has
// not reached a cell) through its out-of-line slow path (~35 instructions); 0 / b is a zero with the product's sign.
__device__ __forceinline__ float div_nz(float t, float b) {
    float r = t * (b < 0.0f ? -1.0f : 1.0f);
    if (t != 0.0f) { asm volatile("" ::: "memory"); r = t / b; }
    return r;
}

// ---- Blackwell packed fp32 (PTX fma.rn.f32x2, SASS FFMA2: one issue slot, two independent fp32 operations) ----------
// Only the fma form is used, with operands that make it a single separately rounded operation of the reference:
//   a * b = fma(a, b, -0)   the exact product plus -0 rounds once, and x + (-0) = x for every x including +-0
//   a + b = fma(a, 1, b)    a * 1 is exact
//   a - b = fma(b, -1, a)   b * (-1) is exact; signs of zero as in a - b
// (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under -fmad=false, so those forms are not used.)
typedef unsigned long long f32x2;      // two fp32 in an aligned register pair: low half = the lower column

__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 pk2(float both) { return pk2(both, both); }
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
constexpr f32x2 kNegZero2 = 0x8000000080000000ull, kOne2 = 0x3f8000003f800000ull;
// literal forms: ptxas may fuse them with their neighbours (tolerance-mode code only)
__device__ __forceinline__ f32x2 loose_mul2(f32x2 a, f32x2 b) { return fma2(a, b, kNegZero2); }
__device__ __forceinline__ f32x2 loose_add2(f32x2 a, f32x2 b) { return fma2(a, kOne2, b); }
// With LITERAL 1 and -0 ptxas rewrites the two forms above to FADD2 / FMUL2 and then contracts an adjacent FMUL2 + FADD2
// pair into one FFMA2 -- even under -fmad=false (seen in SASS; three cells of a 300 x 700 step differed).  The operations
// of a kernel whose products feed sums directly therefore take 1 and -0 from kernel parameters (Consts::one, neg_zero):
// every operation stays a genuine three-operand FFMA2, and two fmas cannot be fused.
struct Pk2 {
    f32x2 one, nz, m1;
    __device__ __forceinline__ f32x2 mul(f32x2 a, f32x2 b) const { return fma2(a, b, nz); }
    __device__ __forceinline__ f32x2 add(f32x2 a, f32x2 b) const { return fma2(a, one, b); }
    __device__ __forceinline__ f32x2 sub(f32x2 a, f32x2 b) const { return fma2(b, m1, a); }
};
__device__ __forceinline__ Pk2 pk2_ops(const Consts& k) {
    Pk2 o;
    o.one = pk2(k.one); o.nz = pk2(k.neg_zero); o.m1 = pk2(-k.one);
    return o;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace vof
