// SURVEY.md section 8(f) rank 4 -- the two opt-in pieces next to the hot path.  Neither is on the default step.
//
//  * k_fct_forward<DIR>: the stand-alone FCT variant of the reference's test/forward_fct.py:267-351 (the version its
//    differentiable experiments are built on).  It differs from the production sweeps of 2dvof.py:321-448 in five
//    places: the limiter ratios divide by (p + eps); there is no var() clamp, neither on the transported-diffused
//    value nor on the result; the divergence correction dx dy / dv multiplies the flux term only; both sweeps scale q
//    by dx; and each of the five loops writes a zero-initialised array of its own time level, so a never-written
//    entry is 0 (the production arrays keep last step's values there).  One thread per cell recomputes what it needs
//    (5 transported-diffused values, 4 antidiffusive fluxes, 3 ratio pairs, 2 limiters): 20-odd redundant flops per
//    cell instead of four intermediate arrays; the expressions and their order are the reference's, so the result is
//    bit-identical to the reference run under the taichi stand-in (tests/golden/ref_fct_*.npz).
//  * k_jacobi_cheb: one sweep of the Chebyshev semi-iterative acceleration of the reference's Jacobi iteration
//    (2dvof.py:236-266): x(k+1) = x(k-1) + w(k+1) (J x(k) - x(k-1)) [Golub & Varga 1961], w from the spectral radius
//    of the Jacobi matrix on the Neumann grid.  x(k-1) is exactly what the other ping-pong buffer holds, so the
//    three-term recurrence needs no extra field.  It changes the pressure (that is its point): outside parity mode.
#pragma once
#include "vof_common.cuh"

namespace vof {

struct FwdC {
    float dt, dx, dy, dxdy, dtd;     // dtd = dt * dy (x sweep, forward_fct.py:269) or dt * dx (y sweep, :312), folded in double
    float eps;
};

// F, the advecting velocity and the geometry of one sweep direction: the stencil runs along `s` (element stride),
// cell index k = 1 .. n along it (0 and n + 1 are ghosts)
struct FwdLine {
    const float* F;
    const float* w;      // u (x sweep) or v (y sweep)
    int s, n;
};

__device__ __forceinline__ float fw_F(const FwdLine& L, const float* base, int k) { return base[(long long)k * L.s]; }

// low-order flux through the lower face of cell k (forward_fct.py:270 / :276): (w dt) F_upwind, >= picks the lower cell
__device__ __forceinline__ float fw_flux_L(const FwdLine& L, const float* Fb, const float* wb, int k, const FwdC& c) {
    const float w = wb[(long long)k * L.s];
    return w >= 0.0f ? w * c.dt * Fb[(long long)(k - 1) * L.s] : w * c.dt * Fb[(long long)k * L.s];
}
__device__ __forceinline__ float fw_flux_H(const FwdLine& L, const float* Fb, const float* wb, int k, const FwdC& c) {
    const float w = wb[(long long)k * L.s];
    return w <= 0.0f ? w * c.dt * Fb[(long long)(k - 1) * L.s] : w * c.dt * Fb[(long long)k * L.s];
}
__device__ __forceinline__ float fw_dv(const FwdLine& L, const float* wb, int k, const FwdC& c) {
    return c.dxdy - c.dtd * (wb[(long long)(k + 1) * L.s] - wb[(long long)k * L.s]);
}
// transported-diffused value (forward_fct.py:268-272): F + (((fl - fr) dy / (dx dy)) dx dy) / dv; 0 outside 1 .. n
// (the y sweep writes fb - ft with the same scaling, :311-315)
__device__ __forceinline__ float fw_Ftd(const FwdLine& L, const float* Fb, const float* wb, int k, const FwdC& c) {
    if (k < 1 || k > L.n) return 0.0f;
    const float lo = fw_flux_L(L, Fb, wb, k, c), hi = fw_flux_L(L, Fb, wb, k + 1, c);
    float t = (lo - hi) * c.dy;
    t = t / c.dxdy;
    t = t * c.dx;
    t = t * c.dy;
    t = t / fw_dv(L, wb, k, c);
    return Fb[(long long)k * L.s] + t;
}
// antidiffusive flux of face k (:274-277), written for k = 1 .. n + 1
__device__ __forceinline__ float fw_a(const FwdLine& L, const float* Fb, const float* wb, int k, const FwdC& c) {
    if (k < 1 || k > L.n + 1) return 0.0f;
    return fw_flux_H(L, Fb, wb, k, c) - fw_flux_L(L, Fb, wb, k, c);
}

template <int DIR>      // 0: fct_x_sweep (stencil along i), 1: fct_y_sweep (along j)
__global__ void __launch_bounds__(128)
k_fct_forward(Grid g, FwdC c, const float* __restrict__ F, const float* __restrict__ w, float* __restrict__ Fn) {
    const int j = 1 + blockIdx.x * 128 + threadIdx.x, i = 1 + blockIdx.y;
    if (j > g.ny || i > g.nx) return;
    FwdLine L;
    L.F = F; L.w = w;
    L.s = DIR == 0 ? g.pitch : 1;
    L.n = DIR == 0 ? g.nx : g.ny;
    const int k = DIR == 0 ? i : j;
    // base pointers of the line through this cell: element m of the line is base[m * s]
    const long long o0 = DIR == 0 ? (long long)j : (long long)i * g.pitch;
    const float* Fb = F + o0;
    const float* wb = w + o0;
    float td[5], a[4];
#pragma unroll
    for (int d = 0; d < 5; ++d) td[d] = fw_Ftd(L, Fb, wb, k - 2 + d, c);       // cells k-2 .. k+2
#pragma unroll
    for (int d = 0; d < 4; ++d) a[d] = fw_a(L, Fb, wb, k - 1 + d, c);          // faces k-1 .. k+2
    // limiter ratios of cells k-1, k, k+1 (:279-297); a never-written entry (cell 0 or n + 1) is 0
    float rp[3], rm[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int m = k - 1 + d;
        rp[d] = rm[d] = 0.0f;
        if (m >= 1 && m <= L.n) {
            const float t0 = td[d + 1], tm = td[d], tp = td[d + 2];
            const float fmax = fmaxf(fmaxf(t0, tm), tp), fmin = fminf(fminf(t0, tm), tp);
            const float alo = a[d], ahi = a[d + 1];                          // faces m and m + 1
            const float pp = fmaxf(0.0f, alo) - fminf(0.0f, ahi);
            const float qp = (fmax - t0) * c.dx;
            if (pp > 0.0f) rp[d] = fminf(1.0f, qp / (pp + c.eps));
            const float pm = fmaxf(0.0f, ahi) - fminf(0.0f, alo);
            const float qm = (t0 - fmin) * c.dx;
            if (pm > 0.0f) rm[d] = fminf(1.0f, qm / (pm + c.eps));
        }
    }
    // face limiters (:299-303): face m + 1 is written from cell m = 1 .. n, so face 1 keeps its 0
    float clo = 0.0f, chi;
    if (k >= 2) clo = a[1] >= 0.0f ? fminf(rp[1], rm[0]) : fminf(rp[0], rm[1]);          // face k from cell k - 1
    chi = a[2] >= 0.0f ? fminf(rp[2], rm[1]) : fminf(rp[1], rm[2]);                        // face k + 1 from cell k
    // corrective update (:305-308)
    float t = a[2] * chi - a[1] * clo;
    t = t / c.dy;
    t = t * c.dx;
    t = t * c.dy;
    t = t / fw_dv(L, wb, k, c);
    Fn[(long long)i * g.pitch + j] = td[2] - t;
}

// ---- Chebyshev-accelerated Jacobi ------------------------------------------------------------------------------------
// one sweep from the hoisted rhs: pn (holding x(k-1) on entry) <- x(k-1) + omega (J p - x(k-1)); ghosts copied from p
__global__ void __launch_bounds__(kBlockJ)
k_jacobi_cheb(Grid g, Consts c, const float* __restrict__ p, float* __restrict__ pn, const float* __restrict__ rhs,
              float omega, int r0, int r1, int rows_per_block) {
    const int j = blockIdx.x * kBlockJ + threadIdx.x;
    if (j > g.ny + 1) return;
    const int ia = r0 + blockIdx.y * rows_per_block;
    const int ib = min(r1, ia + rows_per_block - 1);
    const int P = g.pitch;
    const bool jin = (j >= 1 && j <= g.ny);
    const float an = (j != g.ny) ? c.dyi2 : 0.0f;
    const float as = (j != 1) ? c.dyi2 : 0.0f;
    for (int i = ia; i <= ib; ++i) {
        const size_t o = (size_t)i * P + j;
        const int gi = g.gi0 + i;
        float out = p[o];
        if (jin && gi >= 1 && gi <= g.nx) {
            const float ae = (gi != g.nx) ? c.dxi2 : 0.0f;
            const float aw = (gi != 1) ? c.dxi2 : 0.0f;
            const float ap = -1.0f * (((ae + aw) + an) + as);
            float t = rhs[o] - ae * p[o + P];
            t = t - aw * p[o - P];
            t = t - an * p[o + 1];
            t = t - as * p[o - 1];
            const float jac = t / ap, prev = pn[o];
            out = prev + omega * (jac - prev);
        }
        pn[o] = out;
    }
}

}  // namespace vof
