// sm_100a kernels of the 3-D VOF step (/root/reference/3dvof.py:141-302, 351-547).
// Layout: (nx+2, ny+2, nz+2) fp32, k contiguous (pitched so that k = 1 sits on a 128-byte boundary), j stride
// = pk, i stride = pj.  Threads map to k (coalesced), blocks tile j, each thread marches along i with the
// i-neighbours rolling through registers; j/k-neighbours are read through L1.  Same arithmetic contract as the
// 2-D kernels (order-exact fp32, no FMA contraction).  kappa is never computed by 3dvof.py (line 607), so the
// CSF terms are exactly +-0 and are not evaluated.
#pragma once
#include "vof2d_fct.cuh"
#include "vof2d_jacobi_tb.cuh"
#include "vof_common.cuh"

namespace vof {

struct Grid3 {
    int nx, ny, nz;
    int gi0, nrows;      // local plane l holds global plane gi0 + l
    int pk;              // floats per k-row
    long long pj;        // floats per i-plane = (ny+2) * pk
};

struct Consts3 {
    float dt, dx, dy, dz, dxi, dyi, dzi, dxi2, dyi2, dzi2;
    float vol, dxdy, dt_yz, dt_xz, dt_xy;          // dx*dy*dz, dx*dy, dt*dy*dz, dt*dx*dz, dt*dx*dy (double-folded)
    float rho_l, rho_g, nu_l, nu_g, gx, gy, gz, cflx, cfly;
    float ap[2][2][2];                             // Poisson diagonal by [i-wall][j-wall][k-wall] (3dvof.py:268-274)
};

constexpr int kB3 = 128;   // threads along k

__device__ __forceinline__ float rho3(float F, const Consts3& c) { const float f = var01(F); return c.rho_g * (1.0f - f) + c.rho_l * f; }
__device__ __forceinline__ float nu3(float F, const Consts3& c) { const float f = var01(F); return c.nu_l * f + c.nu_g * (1.0f - f); }

// ---- cal_nu_rho (3dvof.py:199-204): all cells --------------------------------------------------------------
__global__ void __launch_bounds__(kB3)
k3_cal_nu_rho(Grid3 g, Consts3 c, const float* __restrict__ F, float* __restrict__ rho, float* __restrict__ nu, int r0, int r1) {
    const int k = blockIdx.x * kB3 + threadIdx.x, j = blockIdx.y;
    if (k > g.nz + 1) return;
    for (int i = r0 + blockIdx.z; i <= r1; i += gridDim.z) {
        const size_t o = (size_t)i * g.pj + (size_t)j * g.pk + k;
        const float f = F[o];
        rho[o] = rho3(f, c);
        nu[o] = nu3(f, c);
    }
}

// ---- advect_upwind (3dvof.py:207-258): u*, v*, w* in one pass -----------------------------------------------
template <bool INLINE_PROPS>
__global__ void __launch_bounds__(kB3)
k3_advect(Grid3 g, Consts3 c, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w,
          const float* __restrict__ F, const float* __restrict__ rho, const float* __restrict__ nu,
          float* __restrict__ us, float* __restrict__ vs, float* __restrict__ ws, int r0, int r1, int rows_per_block) {
    const int k = 1 + blockIdx.x * kB3 + threadIdx.x, j = 1 + blockIdx.y;
    if (k > g.nz) return;
    const int ia = r0 + blockIdx.z * rows_per_block, ib = min(r1, ia + rows_per_block - 1);
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    for (int i = ia; i <= ib; ++i) {
        const int gi = g.gi0 + i;
        const size_t o = (size_t)i * si + (size_t)j * sj + k;
        const float nu_c = INLINE_PROPS ? nu3(F[o], c) : nu[o];
        const float uc = u[o], vc = v[o], wc = w[o];
        if (gi >= 2 && gi <= g.nx) {   // 3dvof.py:210-225
            const float v_here = 0.25f * (((v[o - si] + v[o - si + sj]) + vc) + v[o + sj]);
            const float w_here = 0.25f * (((w[o - si] + w[o - si + 1]) + wc) + w[o + 1]);
            const float um = u[o - si], up = u[o + si], ujm = u[o - sj], ujp = u[o + sj], ukm = u[o - 1], ukp = u[o + 1];
            const float dudx = uc > 0.0f ? (uc - um) * c.dxi : (up - uc) * c.dxi;
            const float dudy = v_here > 0.0f ? (uc - ujm) * c.dyi : (ujp - uc) * c.dyi;
            const float dudz = w_here > 0.0f ? (uc - ukm) * c.dzi : (ukp - uc) * c.dzi;
            float acc = (nu_c * ((um - 2.0f * uc) + up)) * c.dxi2;
            acc = acc + (nu_c * ((ujm - 2.0f * uc) + ujp)) * c.dyi2;
            acc = acc + (nu_c * ((ukm - 2.0f * uc) + ukp)) * c.dzi2;
            acc = acc - uc * dudx; acc = acc - v_here * dudy; acc = acc - w_here * dudz;
            acc = acc + c.gx;                                // + fx_kappa*2/(rho+rho) is +-0: kappa == 0
            us[o] = uc + c.dt * acc;
        }
        if (j >= 2 && gi >= 1 && gi <= g.nx) {   // 3dvof.py:226-241
            const float u_here = 0.25f * (((u[o - sj] + uc) + u[o + si - sj]) + u[o + si]);
            const float w_here = 0.25f * (((w[o - sj + 1] + w[o - sj]) + wc) + w[o + 1]);
            const float vm = v[o - si], vp = v[o + si], vjm = v[o - sj], vjp = v[o + sj], vkm = v[o - 1], vkp = v[o + 1];
            const float dvdx = u_here > 0.0f ? (vc - vm) * c.dxi : (vp - vc) * c.dxi;
            const float dvdy = vc > 0.0f ? (vc - vjm) * c.dyi : (vjp - vc) * c.dyi;
            const float dvdz = w_here > 0.0f ? (vc - vkm) * c.dzi : (vkp - vc) * c.dzi;
            float acc = (nu_c * ((vm - 2.0f * vc) + vp)) * c.dxi2;
            acc = acc + (nu_c * ((vjm - 2.0f * vc) + vjp)) * c.dyi2;
            acc = acc + (nu_c * ((vkm - 2.0f * vc) + vkp)) * c.dzi2;
            acc = acc - u_here * dvdx; acc = acc - vc * dvdy; acc = acc - w_here * dvdz;
            acc = acc + c.gy;
            vs[o] = vc + c.dt * acc;
        }
        if (k >= 2 && gi >= 1 && gi <= g.nx) {   // 3dvof.py:242-258
            const float u_here = 0.25f * (((u[o + si - 1] + u[o - 1]) + u[o + si]) + uc);
            const float v_here = 0.25f * (((v[o + sj - 1] + v[o - 1]) + vc) + v[o + sj]);
            const float wm = w[o - si], wp = w[o + si], wjm = w[o - sj], wjp = w[o + sj], wkm = w[o - 1], wkp = w[o + 1];
            const float dwdx = u_here > 0.0f ? (wc - wm) * c.dxi : (wp - wc) * c.dxi;
            const float dwdy = v_here > 0.0f ? (wc - wjm) * c.dyi : (wjp - wc) * c.dyi;
            const float dwdz = wc > 0.0f ? (wc - wkm) * c.dzi : (wkp - wc) * c.dzi;
            float acc = (nu_c * ((wm - 2.0f * wc) + wp)) * c.dxi2;
            acc = acc + (nu_c * ((wjm - 2.0f * wc) + wjp)) * c.dyi2;
            acc = acc + (nu_c * ((wkm - 2.0f * wc) + wkp)) * c.dzi2;
            acc = acc - u_here * dwdx; acc = acc - v_here * dwdy; acc = acc - wc * dwdz;
            acc = acc + c.gz;
            ws[o] = wc + c.dt * acc;
        }
    }
    (void)rho;
}

// ---- advect_upwind, second generation: float4 lanes along k, a warp per j-line, 4 lines per block, marching i.  The
// first-generation kernel issues ~28 scalar loads per cell (302 instructions per cell, ncu); here a lane loads ten float4
// per plane step for its 4 cells, the i-neighbours roll through registers and the k-neighbours come from the adjacent
// lanes.  Properties from F (the fused step); expressions and their order are those of k3_advect / 3dvof.py:207-258.
__global__ void __launch_bounds__(128)
k3_advect5(Grid3 g, Consts3 c, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w,
           const float* __restrict__ F, float* __restrict__ us, float* __restrict__ vs, float* __restrict__ ws,
           int r0, int r1, int rows_per_block) {
    const int lane = threadIdx.x & 31;
    const int j = 1 + blockIdx.y * 4 + (threadIdx.x >> 5);
    if (j > g.ny) return;                                                 // warp-uniform
    const int kl = 1 + (blockIdx.x * 32 + lane) * 4;
    const bool active = kl <= g.nz;
    const int ia = r0 + blockIdx.z * rows_per_block, ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    size_t o = (size_t)ia * si + (size_t)j * sj + kl;
    struct R4 { float x[4]; };
    auto ld = [&](const float* b, size_t q) { R4 r; const float4 t = active ? *reinterpret_cast<const float4*>(b + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                                              r.x[0] = t.x; r.x[1] = t.y; r.x[2] = t.z; r.x[3] = t.w; return r; };
    const bool ledge = active && lane == 0;                               // reads column kl-1 itself
    const bool redge = active && (lane == 31 || kl + 4 == g.nz + 1);      // reads column kl+4 itself
    auto left = [&](const R4& r, const float* b, size_t q) { float x = __shfl_up_sync(0xffffffffu, r.x[3], 1); if (ledge) x = b[q - 1]; return x; };
    auto right = [&](const R4& r, const float* b, size_t q) { float x = __shfl_down_sync(0xffffffffu, r.x[0], 1); if (redge) x = b[q + 4]; return x; };
    const bool full = kl + 3 <= g.nz;
    bool colin[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) colin[q] = kl + q <= g.nz;
    // rolled along i: planes i-1 and i of the line j, u of the line j-1 (plane i), v of the line j+1 (plane i-1)
    R4 u_m = ld(u, o - si), u_c = ld(u, o), u_cjm = ld(u, o - sj);
    R4 v_m = ld(v, o - si), v_c = ld(v, o), v_mjp = ld(v, o - si + sj);
    R4 w_m = ld(w, o - si), w_c = ld(w, o);
    for (int i = ia; i <= ib; ++i, o += si) {
        const int gi = g.gi0 + i;
        const R4 u_p = ld(u, o + si), u_pjm = ld(u, o + si - sj), v_p = ld(v, o + si), w_p = ld(w, o + si);
        const R4 u_jp = ld(u, o + sj), v_jm = ld(v, o - sj), v_jp = ld(v, o + sj), w_jm = ld(w, o - sj), w_jp = ld(w, o + sj);
        const R4 f_c = ld(F, o);
        // k-neighbours (column kl-1 / kl+4) of the lines that need them
        const float u_c_l = left(u_c, u, o), u_c_r = right(u_c, u, o), u_p_l = left(u_p, u, o + si);
        const float v_c_l = left(v_c, v, o), v_c_r = right(v_c, v, o), v_jp_l = left(v_jp, v, o + sj);
        const float w_c_l = left(w_c, w, o), w_c_r = right(w_c, w, o), w_m_r = right(w_m, w, o - si), w_jm_r = right(w_jm, w, o - sj);
        if (active) {
            float ou[4], ov[4], ow[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float nu_c = nu3(f_c.x[q], c);
                const float uc = u_c.x[q], vc = v_c.x[q], wc = w_c.x[q];
                const float ukm = q ? u_c.x[q - 1] : u_c_l, ukp = q < 3 ? u_c.x[q + 1] : u_c_r;
                const float vkm = q ? v_c.x[q - 1] : v_c_l, vkp = q < 3 ? v_c.x[q + 1] : v_c_r;
                const float wkm = q ? w_c.x[q - 1] : w_c_l, wkp = q < 3 ? w_c.x[q + 1] : w_c_r;
                {   // 3dvof.py:210-225
                    const float v_here = 0.25f * (((v_m.x[q] + v_mjp.x[q]) + vc) + v_jp.x[q]);
                    const float w_here = 0.25f * (((w_m.x[q] + (q < 3 ? w_m.x[q + 1] : w_m_r)) + wc) + wkp);
                    const float um = u_m.x[q], up = u_p.x[q], ujm = u_cjm.x[q], ujp = u_jp.x[q];
                    const float dudx = uc > 0.0f ? (uc - um) * c.dxi : (up - uc) * c.dxi;
                    const float dudy = v_here > 0.0f ? (uc - ujm) * c.dyi : (ujp - uc) * c.dyi;
                    const float dudz = w_here > 0.0f ? (uc - ukm) * c.dzi : (ukp - uc) * c.dzi;
                    float acc = (nu_c * ((um - 2.0f * uc) + up)) * c.dxi2;
                    acc = acc + (nu_c * ((ujm - 2.0f * uc) + ujp)) * c.dyi2;
                    acc = acc + (nu_c * ((ukm - 2.0f * uc) + ukp)) * c.dzi2;
                    acc = acc - uc * dudx; acc = acc - v_here * dudy; acc = acc - w_here * dudz;
                    acc = acc + c.gx;                                // + fx_kappa*2/(rho+rho) is +-0: kappa == 0
                    ou[q] = uc + c.dt * acc;
                }
                {   // 3dvof.py:226-241
                    const float u_here = 0.25f * (((u_cjm.x[q] + uc) + u_pjm.x[q]) + u_p.x[q]);
                    const float w_here = 0.25f * ((((q < 3 ? w_jm.x[q + 1] : w_jm_r) + w_jm.x[q]) + wc) + wkp);
                    const float vm = v_m.x[q], vp = v_p.x[q], vjm = v_jm.x[q], vjp = v_jp.x[q];
                    const float dvdx = u_here > 0.0f ? (vc - vm) * c.dxi : (vp - vc) * c.dxi;
                    const float dvdy = vc > 0.0f ? (vc - vjm) * c.dyi : (vjp - vc) * c.dyi;
                    const float dvdz = w_here > 0.0f ? (vc - vkm) * c.dzi : (vkp - vc) * c.dzi;
                    float acc = (nu_c * ((vm - 2.0f * vc) + vp)) * c.dxi2;
                    acc = acc + (nu_c * ((vjm - 2.0f * vc) + vjp)) * c.dyi2;
                    acc = acc + (nu_c * ((vkm - 2.0f * vc) + vkp)) * c.dzi2;
                    acc = acc - u_here * dvdx; acc = acc - vc * dvdy; acc = acc - w_here * dvdz;
                    acc = acc + c.gy;
                    ov[q] = vc + c.dt * acc;
                }
                {   // 3dvof.py:242-258
                    const float u_here = 0.25f * ((((q ? u_p.x[q - 1] : u_p_l) + ukm) + u_p.x[q]) + uc);
                    const float v_here = 0.25f * ((((q ? v_jp.x[q - 1] : v_jp_l) + vkm) + vc) + v_jp.x[q]);
                    const float wm = w_m.x[q], wp = w_p.x[q], wjm = w_jm.x[q], wjp = w_jp.x[q];
                    const float dwdx = u_here > 0.0f ? (wc - wm) * c.dxi : (wp - wc) * c.dxi;
                    const float dwdy = v_here > 0.0f ? (wc - wjm) * c.dyi : (wjp - wc) * c.dyi;
                    const float dwdz = wc > 0.0f ? (wc - wkm) * c.dzi : (wkp - wc) * c.dzi;
                    float acc = (nu_c * ((wm - 2.0f * wc) + wp)) * c.dxi2;
                    acc = acc + (nu_c * ((wjm - 2.0f * wc) + wjp)) * c.dyi2;
                    acc = acc + (nu_c * ((wkm - 2.0f * wc) + wkp)) * c.dzi2;
                    acc = acc - u_here * dwdx; acc = acc - v_here * dwdy; acc = acc - wc * dwdz;
                    acc = acc + c.gz;
                    ow[q] = wc + c.dt * acc;
                }
            }
            const bool irow = gi >= 1 && gi <= g.nx;
            if (gi >= 2 && gi <= g.nx) {
                if (full) *reinterpret_cast<float4*>(us + o) = make_float4(ou[0], ou[1], ou[2], ou[3]);
                else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (colin[q]) us[o + q] = ou[q];
                }
            }
            if (irow && j >= 2) {
                if (full) *reinterpret_cast<float4*>(vs + o) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (colin[q]) vs[o + q] = ov[q];
                }
            }
            if (irow) {
                if (full && kl >= 2) *reinterpret_cast<float4*>(ws + o) = make_float4(ow[0], ow[1], ow[2], ow[3]);
                else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (colin[q] && kl + q >= 2) ws[o + q] = ow[q];
                }
            }
        }
        u_m = u_c; u_c = u_p; u_cjm = u_pjm;
        v_m = v_c; v_c = v_p; v_mjp = v_jp;
        w_m = w_c; w_c = w_p;
    }
}

// ---- set_BC (3dvof.py:141-190): three loops, launched in the reference's order ---------------------------------
// face 0: j-faces over (i, k); 1: i-faces over (j, k); 2: k-faces over (i, j).  mask: 1 u, 2 v, 4 w, 8 F, 16 p, 32 rho
__global__ void __launch_bounds__(kB3)
k3_set_bc(Grid3 g, float* __restrict__ u, float* __restrict__ v, float* __restrict__ w, float* __restrict__ F,
          float* __restrict__ p, float* __restrict__ rho, int face, int ra, int rb, int has_lo, int has_hi, unsigned mask) {
    float* X[3] = {F, p, rho};
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    if (face == 0) {          // over (i, k): planes ra..rb
        const int k = blockIdx.x * kB3 + threadIdx.x, i = ra + blockIdx.y;
        if (k > g.nz + 1 || i > rb) return;
        const size_t b = (size_t)i * si + k, lo0 = b, lo1 = b + sj, hi1 = b + (size_t)(g.ny + 1) * sj, hi0 = b + (size_t)g.ny * sj;
        if (mask & 1u) { u[lo0] = u[lo1]; u[hi1] = u[hi0]; }
        if (mask & 2u) { v[lo1] = 0.0f; v[hi1] = 0.0f; }
        if (mask & 4u) { w[lo0] = w[lo1]; w[hi1] = w[hi0]; }
#pragma unroll
        for (int q = 0; q < 3; ++q) if (mask & (8u << q)) { X[q][lo0] = X[q][lo1]; X[q][hi1] = X[q][hi0]; }
    } else if (face == 1) {   // over (j, k): only contexts that hold a physical i-wall
        const int k = blockIdx.x * kB3 + threadIdx.x, j = blockIdx.y;
        if (k > g.nz + 1) return;
        const size_t b = (size_t)j * sj + k;
        if (has_lo) {
            const size_t g0 = (size_t)(0 - g.gi0) * si + b, g1 = g0 + si;
            if (mask & 1u) u[g1] = 0.0f;
            if (mask & 2u) v[g0] = v[g1];
            if (mask & 4u) w[g0] = w[g1];
#pragma unroll
            for (int q = 0; q < 3; ++q) if (mask & (8u << q)) X[q][g0] = X[q][g1];
        }
        if (has_hi) {
            const size_t gn = (size_t)(g.nx - g.gi0) * si + b, gn1 = gn + si;
            if (mask & 1u) u[gn1] = 0.0f;
            if (mask & 2u) v[gn1] = v[gn];
            if (mask & 4u) w[gn1] = w[gn];
#pragma unroll
            for (int q = 0; q < 3; ++q) if (mask & (8u << q)) X[q][gn1] = X[q][gn];
        }
    } else {                  // over (i, j): planes ra..rb; the k-neighbours are 1 float apart
        const int j = blockIdx.x * kB3 + threadIdx.x, i = ra + blockIdx.y;
        if (j > g.ny + 1 || i > rb) return;
        const size_t b = (size_t)i * si + (size_t)j * sj;
        const size_t lo0 = b, lo1 = b + 1, hi0 = b + g.nz, hi1 = b + g.nz + 1;
        if (mask & 1u) { u[lo0] = u[lo1]; u[hi1] = u[hi0]; }
        if (mask & 2u) { v[lo0] = v[lo1]; v[hi1] = v[hi0]; }
        if (mask & 4u) { w[lo1] = 0.0f; w[hi1] = 0.0f; }
#pragma unroll
        for (int q = 0; q < 3; ++q) if (mask & (8u << q)) { X[q][lo0] = X[q][lo1]; X[q][hi1] = X[q][hi0]; }
    }
}

// ---- Poisson rhs (3dvof.py:264-267), hoisted out of the sweeps ----------------------------------------------------
template <bool INLINE_PROPS>
__global__ void __launch_bounds__(kB3)
k3_rhs(Grid3 g, Consts3 c, const float* __restrict__ rhoF, const float* __restrict__ us, const float* __restrict__ vs,
       const float* __restrict__ ws, float* __restrict__ rhs, int r0, int r1, int rows_per_block) {
    const int k = 1 + blockIdx.x * kB3 + threadIdx.x, j = 1 + blockIdx.y;
    if (k > g.nz) return;
    const int ia = r0 + blockIdx.z * rows_per_block, ib = min(r1, ia + rows_per_block - 1);
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    size_t o = (size_t)ia * si + (size_t)j * sj + k;
    float us_c = us[o];
    for (int i = ia; i <= ib; ++i, o += si) {
        const float us_p = us[o + si];
        const float r = INLINE_PROPS ? rho3(rhoF[o], c) : rhoF[o];
        rhs[o] = (r / c.dt) * (((us_p - us_c) * c.dxi + (vs[o + sj] - vs[o]) * c.dyi) + (ws[o + 1] - ws[o]) * c.dzi);
        us_c = us_p;
    }
}

// Poisson rhs, second generation: float4 lanes along k, a warp per j-line, u* rolled along i, w*[k+1] from the adjacent
// lane, everything of plane i+1 requested while plane i is computed.  rho from F (the fused step).
__global__ void __launch_bounds__(128)
k3_rhs5(Grid3 g, Consts3 c, const float* __restrict__ F, const float* __restrict__ us, const float* __restrict__ vs,
        const float* __restrict__ ws, float* __restrict__ rhs, int r0, int r1, int rows_per_block) {
    const int lane = threadIdx.x & 31;
    const int j = 1 + blockIdx.y * 4 + (threadIdx.x >> 5);
    if (j > g.ny) return;                                                 // warp-uniform
    const int kl = 1 + (blockIdx.x * 32 + lane) * 4;
    const bool active = kl <= g.nz;
    const int ia = r0 + blockIdx.z * rows_per_block, ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    size_t o = (size_t)ia * si + (size_t)j * sj + kl;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ld4 = [&](const float* b, size_t q) { return active ? *reinterpret_cast<const float4*>(b + q) : z4; };
    const bool redge = active && (lane == 31 || kl + 4 == g.nz + 1);      // reads column kl+4 itself
    const bool full = kl + 3 <= g.nz;
    float4 us_c = ld4(us, o);
    float4 us_n = ld4(us, o + si), vs_n = ld4(vs, o), vj_n = ld4(vs, o + sj), ws_n = ld4(ws, o), f_n = ld4(F, o);
    float we_n = redge ? ws[o + 4] : 0.0f;
    for (int i = ia; i <= ib; ++i, o += si) {
        const float4 us_p = us_n, vs4 = vs_n, vj4 = vj_n, ws4 = ws_n, f4 = f_n;
        const float we = we_n;
        if (i < ib) {
            const size_t on = o + si;
            us_n = ld4(us, on + si); vs_n = ld4(vs, on); vj_n = ld4(vs, on + sj); ws_n = ld4(ws, on); f_n = ld4(F, on);
            if (redge) we_n = ws[on + 4];
        }
        float w_r = __shfl_down_sync(0xffffffffu, ws4.x, 1);
        if (redge) w_r = we;
        if (active) {
            const float uc[4] = {us_c.x, us_c.y, us_c.z, us_c.w}, up[4] = {us_p.x, us_p.y, us_p.z, us_p.w};
            const float vc[4] = {vs4.x, vs4.y, vs4.z, vs4.w}, vp[4] = {vj4.x, vj4.y, vj4.z, vj4.w};
            const float wc[4] = {ws4.x, ws4.y, ws4.z, ws4.w}, ff[4] = {f4.x, f4.y, f4.z, f4.w};
            float out[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float r = rho3(ff[q], c);
                const float wp = q < 3 ? wc[q + 1] : w_r;
                out[q] = (r / c.dt) * (((up[q] - uc[q]) * c.dxi + (vp[q] - vc[q]) * c.dyi) + (wp - wc[q]) * c.dzi);
            }
            if (full) *reinterpret_cast<float4*>(rhs + o) = make_float4(out[0], out[1], out[2], out[3]);
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (kl + q <= g.nz) rhs[o + q] = out[q];
            }
        }
        us_c = us_p;
    }
}

// ---- solve_p_jacobi (3dvof.py:261-283): one sweep p -> pn; non-interior cells are copied -----------------------
// RHS_MODE 0: hoisted rhs; 1: recomputed from the rho array (the reference's structure)
template <int RHS_MODE>
__global__ void __launch_bounds__(kB3)
k3_jacobi(Grid3 g, Consts3 c, const float* __restrict__ p, float* __restrict__ pn, const float* __restrict__ rhs,
          const float* __restrict__ rho, const float* __restrict__ us, const float* __restrict__ vs,
          const float* __restrict__ ws, int r0, int r1, int rows_per_block) {
    const int k = blockIdx.x * kB3 + threadIdx.x, j = blockIdx.y;
    if (k > g.nz + 1) return;
    const int ia = r0 + blockIdx.z * rows_per_block, ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    size_t o = (size_t)ia * si + (size_t)j * sj + k;
    const bool jkin = j >= 1 && j <= g.ny && k >= 1 && k <= g.nz;
    const float an = (j != g.ny) ? c.dyi2 : 0.0f, as = (j != 1) ? c.dyi2 : 0.0f;
    const float af = (k != g.nz) ? c.dzi2 : 0.0f, ab = (k != 1) ? c.dzi2 : 0.0f;
    const int jw = (j == 1 || j == g.ny) ? 1 : 0, kw = (k == 1 || k == g.nz) ? 1 : 0;
    float p_m = ia > 0 ? p[o - si] : 0.0f, p_c = p[o];
    for (int i = ia; i <= ib; ++i, o += si) {
        const int gi = g.gi0 + i;
        const float p_p = (i + 1 < g.nrows) ? p[o + si] : 0.0f;
        float out = p_c;
        if (jkin && gi >= 1 && gi <= g.nx) {
            float b;
            if (RHS_MODE == 0) b = rhs[o];
            else b = (rho[o] / c.dt) * (((us[o + si] - us[o]) * c.dxi + (vs[o + sj] - vs[o]) * c.dyi) + (ws[o + 1] - ws[o]) * c.dzi);
            const float ae = (gi != g.nx) ? c.dxi2 : 0.0f, aw = (gi != 1) ? c.dxi2 : 0.0f;
            const float ap = c.ap[(gi == 1 || gi == g.nx) ? 1 : 0][jw][kw];
            float t = b - ae * p_p;
            t = t - aw * p_m; t = t - an * p[o + sj]; t = t - as * p[o - sj];
            t = t - af * p[o + 1]; t = t - ab * p[o - 1];
            out = t / ap;
        }
        pn[o] = out;
        p_m = p_c; p_c = p_p;
    }
}

// One sweep, 4 consecutive k per lane (LDG.128 / STG.128), hoisted rhs.  Cells away from every wall use literal
// coefficients and the exact reciprocal division (ConstDiv, verified at context creation); the rest (wall
// rows / columns, ghosts) take the selected-coefficient IEEE path.  Same arithmetic, same order.
struct Jac3C { float cx, cy, cz; ConstDiv dv; int fast_div_ok; int bare_div_ok; };   // bare: no sub-normal fix-up needed (proven, see k_check_div_by_const)

__global__ void __launch_bounds__(128)
k3_jacobi4(Grid3 g, Consts3 c, Jac3C jc, const float* __restrict__ p, float* __restrict__ pn, const float* __restrict__ rhs,
           int r0, int r1, int rows_per_block) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int kl = 1 + (blockIdx.x * 32 + lane) * 4;                      // == 1 (mod 4): 16-byte aligned
    if (j > g.ny + 1) return;
    const bool active = kl <= g.nz + 1;
    const int ia = r0 + blockIdx.z * rows_per_block, ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    size_t o = (size_t)ia * si + (size_t)j * sj + kl;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ld = [&](const float* b, size_t q) { return active ? *reinterpret_cast<const float4*>(b + q) : z4; };
    const bool jin = j >= 1 && j <= g.ny;
    const bool lane_interior = j >= 2 && j <= g.ny - 1 && kl >= 2 && kl + 3 <= g.nz - 1;   // no j- or k-wall coefficient is zero
    float4 p_m = ia > 0 ? ld(p, o - si) : z4, p_c = ld(p, o);
    for (int i = ia; i <= ib; ++i, o += si) {
        const int gi = g.gi0 + i;
        const float4 p_p = (i + 1 < g.nrows) ? ld(p, o + si) : z4;
        // k-neighbours of the 4 cells: own values, the two outer ones from the adjacent lanes (or memory at the strip edge)
        float left = __shfl_up_sync(0xffffffffu, p_c.w, 1), right = __shfl_down_sync(0xffffffffu, p_c.x, 1);
        if (active) {
            if (lane == 0) left = p[o - 1];
            if (lane == 31 && kl + 4 <= g.nz + 1) right = p[o + 4];
        }
        float4 out = p_c;
        if (active && jin && gi >= 1 && gi <= g.nx) {
            const float4 pjp = ld(p, o + sj), pjm = ld(p, o - sj), b4 = ld(rhs, o);
            const float pc[4] = {p_c.x, p_c.y, p_c.z, p_c.w}, pp[4] = {p_p.x, p_p.y, p_p.z, p_p.w}, pm[4] = {p_m.x, p_m.y, p_m.z, p_m.w};
            const float jp[4] = {pjp.x, pjp.y, pjp.z, pjp.w}, jm[4] = {pjm.x, pjm.y, pjm.z, pjm.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
            float r[4];
            if (lane_interior && gi >= 2 && gi <= g.nx - 1) {
                float t[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float kp = q < 3 ? pc[q + 1] : right, km = q > 0 ? pc[q - 1] : left;
                    t[q] = bb[q] - jc.cx * pp[q];
                    t[q] = t[q] - jc.cx * pm[q];
                    t[q] = t[q] - jc.cy * jp[q];
                    t[q] = t[q] - jc.cy * jm[q];
                    t[q] = t[q] - jc.cz * kp;
                    t[q] = t[q] - jc.cz * km;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) r[q] = jc.fast_div_ok ? div_by_const(t[q], jc.dv) : t[q] / jc.dv.b;
            } else {
                const float ae = (gi != g.nx) ? c.dxi2 : 0.0f, aw = (gi != 1) ? c.dxi2 : 0.0f;
                const float an = (j != g.ny) ? c.dyi2 : 0.0f, as = (j != 1) ? c.dyi2 : 0.0f;
                const int iw = (gi == 1 || gi == g.nx) ? 1 : 0, jw = (j == 1 || j == g.ny) ? 1 : 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = kl + q;
                    const float kp = q < 3 ? pc[q + 1] : right, km = q > 0 ? pc[q - 1] : left;
                    const float af = (k != g.nz) ? c.dzi2 : 0.0f, ab = (k != 1) ? c.dzi2 : 0.0f;
                    const float ap = c.ap[iw][jw][(k == 1 || k == g.nz) ? 1 : 0];
                    float t = bb[q] - ae * pp[q];
                    t = t - aw * pm[q]; t = t - an * jp[q]; t = t - as * jm[q];
                    t = t - af * kp; t = t - ab * km;
                    r[q] = (k >= 1 && k <= g.nz) ? t / ap : pc[q];        // ghost / pad columns pass through
                }
            }
            out = make_float4(r[0], r[1], r[2], r[3]);
        }
        if (active) *reinterpret_cast<float4*>(pn + o) = out;
        if (kl == 1 && active) pn[o - 1] = p[o - 1];                      // ghost column k = 0 passes through
        p_m = p_c; p_c = p_p;
    }
}

// ---- 7-point sweep, second generation ---------------------------------------------------------------------------
// k3_jacobi4 ran 116 instructions per cell (ncu): a duplicated wall path that every warp with one wall lane executed
// in full, four separately diverging sub-normal fix-ups, an almost empty fifth strip at nz = 512, and every load
// issued in the iteration that consumes it.  Here: one expression with the coefficients selected per plane (ae, aw),
// per warp (an, as) and per lane (af, ab); the constant-divisor quotient for every cell plus an IEEE division only on
// the cells whose diagonal differs; one warp-level fix-up for sub-normal quotients; everything row i+1 needs is
// requested while row i is computed; strips cover interior columns only (the lane next to a ghost column copies it).
template <bool BARE>
__global__ void __launch_bounds__(128, BARE ? 8 : 7)      // 28 (32) warps per SM: 512^3 step 7.14 -> 6.60 ms (6 blocks: 6.85, 8 blocks with the sub-normal path spill: 7.21)
k3_jacobi5(Grid3 g, Consts3 c, Jac3C jc, const float* __restrict__ p, float* __restrict__ pn, const float* __restrict__ rhs,
           int r0, int r1, int rows_per_block) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int kl = 1 + (blockIdx.x * 32 + lane) * 4;                      // == 1 (mod 4): 16-byte aligned
    if (j > g.ny + 1) return;                                             // warp-uniform
    const bool active = kl <= g.nz;
    const int ia = r0 + blockIdx.z * rows_per_block, ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    size_t o = (size_t)ia * si + (size_t)j * sj + kl;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int last = g.nrows - 1;
    const bool jin = j >= 1 && j <= g.ny;
    const bool jw = j == 1 || j == g.ny;
    const float an = (j != g.ny) ? c.dyi2 : 0.0f, as = (j != 1) ? c.dyi2 : 0.0f;
    bool kin[4], kw[4];
    float af[4], ab[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = kl + q;
        kin[q] = k <= g.nz; kw[q] = k == 1 || k == g.nz;
        af[q] = (k != g.nz) ? c.dzi2 : 0.0f; ab[q] = (k != 1) ? c.dzi2 : 0.0f;
    }
    const bool strip_kwall = __any_sync(0xffffffffu, active && (kw[0] || kw[1] || kw[2] || kw[3]));
    const bool left_edge = active && lane == 0;                           // reads column kl-1 itself
    const bool left_ghost = active && kl == 1;                            // ... and copies it (ghost column 0)
    const bool right_edge = active && (lane == 31 || kl + 4 == g.nz + 1); // reads column kl+4 itself
    const bool right_ghost = active && kl + 4 == g.nz + 1;                // ... and copies it (ghost column nz+1)
    auto ldc = [&](const float* b, size_t q, bool ok) { return ok ? *reinterpret_cast<const float4*>(b + q) : z4; };
    // row i: centre planes i-1, i, i+1, the j-neighbours and rhs of plane i, the two k-edge values
    float4 p_m = (active && ia > 0) ? ldc(p, o - si, true) : z4;
    float4 p_c = ldc(p, o, active);
    float4 p_p = ldc(p, o + si, active && ia + 1 <= last);
    float4 jp_c = ldc(p, o + sj, active && jin), jm_c = ldc(p, o - sj, active && jin), b_c = ldc(rhs, o, active && jin);
    float le_c = left_edge ? p[o - 1] : 0.0f, re_c = right_edge ? p[o + 4] : 0.0f;
    for (int i = ia; i <= ib; ++i, o += si) {
        const int gi = g.gi0 + i;
        // requests for row i+1
        const bool more = i + 1 <= ib;
        const size_t on = o + si;
        const float4 p_pp = ldc(p, on + si, active && more && i + 2 <= last);
        const float4 jp_n = ldc(p, on + sj, active && more && jin), jm_n = ldc(p, on - sj, active && more && jin);
        const float4 b_n = ldc(rhs, on, active && more && jin);
        const float le_n = (left_edge && more) ? p[on - 1] : 0.0f, re_n = (right_edge && more) ? p[on + 4] : 0.0f;

        float left = __shfl_up_sync(0xffffffffu, p_c.w, 1), right = __shfl_down_sync(0xffffffffu, p_c.x, 1);
        if (left_edge) left = le_c;
        if (right_edge) right = re_c;
        float4 out = p_c;
        const bool rowin = gi >= 1 && gi <= g.nx;
        if (rowin && jin) {                                               // warp-uniform
            const bool iw = gi == 1 || gi == g.nx;
            const float ae = (gi != g.nx) ? c.dxi2 : 0.0f, aw = (gi != 1) ? c.dxi2 : 0.0f;
            const float pc[4] = {p_c.x, p_c.y, p_c.z, p_c.w}, pp[4] = {p_p.x, p_p.y, p_p.z, p_p.w}, pm[4] = {p_m.x, p_m.y, p_m.z, p_m.w};
            const float jp[4] = {jp_c.x, jp_c.y, jp_c.z, jp_c.w}, jm[4] = {jm_c.x, jm_c.y, jm_c.z, jm_c.w}, bb[4] = {b_c.x, b_c.y, b_c.z, b_c.w};
            float t[4], r[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float kp = q < 3 ? pc[q + 1] : right, km = q > 0 ? pc[q - 1] : left;
                t[q] = bb[q] - ae * pp[q];
                t[q] = t[q] - aw * pm[q];
                t[q] = t[q] - an * jp[q];
                t[q] = t[q] - as * jm[q];
                t[q] = t[q] - af[q] * kp;
                t[q] = t[q] - ab[q] * km;
            }
            if (jc.fast_div_ok && !iw && !jw) {
                bool slow = false;
#pragma unroll
                for (int q = 0; q < 4; ++q) { r[q] = div_by_const_core(t[q], jc.dv.b, jc.dv.r); slow = slow || (!BARE && div_needs_ieee(t[q])); }
                if (!BARE && __any_sync(0xffffffffu, slow)) {             // sub-normal quotients: the fp64 scheme (not needed where the
                                                                          // bare form is proven exact for every numerator: BARE)
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (div_needs_ieee(t[q])) r[q] = div_slow(t[q], jc.dv);
                }
                if (strip_kwall) {                                        // warp-uniform: first and last strip only
                    const float ap1 = c.ap[0][0][1];
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (kw[q]) r[q] = div_nz(t[q], ap1);   // k-wall cells: another diagonal
                }
            } else {
                const float ap0 = c.ap[iw ? 1 : 0][jw ? 1 : 0][0], ap1 = c.ap[iw ? 1 : 0][jw ? 1 : 0][1];
#pragma unroll
                for (int q = 0; q < 4; ++q) r[q] = div_nz(t[q], kw[q] ? ap1 : ap0);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) r[q] = kin[q] ? r[q] : pc[q];      // ghost / pad columns pass through
            out = make_float4(r[0], r[1], r[2], r[3]);
        }
        if (active) *reinterpret_cast<float4*>(pn + o) = out;
        if (left_ghost) pn[o - 1] = le_c;                                 // ghost column k = 0 passes through
        if (right_ghost) pn[o + 4] = re_c;                                // ghost column k = nz+1 (when it starts a float4)
        p_m = p_c; p_c = p_p; p_p = p_pp; jp_c = jp_n; jm_c = jm_n; b_c = b_n; le_c = le_n; re_c = re_n;
    }
}

// ---- update_uv (3dvof.py:286-302) -----------------------------------------------------------------------------------
template <bool INLINE_PROPS>
__global__ void __launch_bounds__(kB3)
k3_project(Grid3 g, Consts3 c, const float* __restrict__ rhoF, const float* __restrict__ p, const float* __restrict__ us,
           const float* __restrict__ vs, const float* __restrict__ ws, float* __restrict__ u, float* __restrict__ v,
           float* __restrict__ w, unsigned long long* __restrict__ courant_count, int r0, int r1, int rows_per_block,
           int own_a, int own_b) {
    const int k = 1 + blockIdx.x * kB3 + threadIdx.x, j = 1 + blockIdx.y;
    if (k > g.nz) return;
    const int ia = r0 + blockIdx.z * rows_per_block, ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    size_t o = (size_t)ia * si + (size_t)j * sj + k;
    auto R = [&](size_t q) { return INLINE_PROPS ? rho3(rhoF[q], c) : rhoF[q]; };
    float p_m = p[o - si], rho_m = R(o - si);
    unsigned flags = 0;
    for (int i = ia; i <= ib; ++i, o += si) {
        const int gi = g.gi0 + i;
        const bool own = i >= own_a && i <= own_b;
        const float p_c = p[o], rho_c = R(o);
        if (gi >= 2 && gi <= g.nx) {
            const float r = (rho_c + rho_m) * 0.5f;
            const float un = us[o] - ((c.dt / r) * (p_c - p_m)) * c.dxi;
            u[o] = un; flags += own && (un * c.dt > c.cflx);
        }
        if (gi >= 1 && gi <= g.nx) {
            if (j >= 2) {
                const float r = (rho_c + R(o - sj)) * 0.5f;
                const float vn = vs[o] - ((c.dt / r) * (p_c - p[o - sj])) * c.dyi;
                v[o] = vn; flags += own && (vn * c.dt > c.cfly);
            }
            if (k >= 2) {
                const float r = (rho_c + R(o - 1)) * 0.5f;
                const float wn = ws[o] - ((c.dt / r) * (p_c - p[o - 1])) * c.dzi;
                w[o] = wn; flags += own && (wn * c.dt > c.cflx);   // 0.25*dx, 3dvof.py:301
            }
        }
        p_m = p_c; rho_m = rho_c;
    }
    if (flags) atomicAdd(courant_count, (unsigned long long)flags);
}

// ---- update_uv, second generation: float4 lanes along k, a warp per j-line, 4 lines per block, marching i.  Everything
// plane i+1 needs is requested while plane i is computed (the first-generation kernel was latency bound: every load
// was issued where it was consumed); rho is evaluated once per cell and line instead of four times per cell; the
// k-1 neighbours come from the adjacent lane.  Properties from F (the fused step); same expressions, same order.
__global__ void __launch_bounds__(128)
k3_project5(Grid3 g, Consts3 c, const float* __restrict__ F, const float* __restrict__ p, const float* __restrict__ us,
            const float* __restrict__ vs, const float* __restrict__ ws, float* __restrict__ u, float* __restrict__ v,
            float* __restrict__ w, unsigned long long* __restrict__ courant_count, int r0, int r1, int rows_per_block,
            int own_a, int own_b) {
    const int lane = threadIdx.x & 31;
    const int j = 1 + blockIdx.y * 4 + (threadIdx.x >> 5);
    if (j > g.ny) return;                                                 // warp-uniform
    const int kl = 1 + (blockIdx.x * 32 + lane) * 4;
    const bool active = kl <= g.nz;
    const int ia = r0 + blockIdx.z * rows_per_block, ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    size_t o = (size_t)ia * si + (size_t)j * sj + kl;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ld4 = [&](const float* b, size_t q, bool ok) { return ok ? *reinterpret_cast<const float4*>(b + q) : z4; };
    const bool full = kl + 3 <= g.nz;                                     // all four columns interior
    const bool edge = active && lane == 0;                                // reads column kl-1 itself
    bool colin[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) colin[q] = kl + q <= g.nz;
    auto rho4 = [&](float4 f, float (&r)[4]) { r[0] = rho3(f.x, c); r[1] = rho3(f.y, c); r[2] = rho3(f.z, c); r[3] = rho3(f.w, c); };
    // plane ia-1: p and rho only
    float pm[4], rm[4];
    { const float4 t = ld4(p, o - si, active); pm[0] = t.x; pm[1] = t.y; pm[2] = t.z; pm[3] = t.w; rho4(ld4(F, o - si, active), rm); }
    float4 p_n = ld4(p, o, active), F_n = ld4(F, o, active), us_n = ld4(us, o, active), vs_n = ld4(vs, o, active), ws_n = ld4(ws, o, active);
    float4 pj_n = ld4(p, o - sj, active), Fj_n = ld4(F, o - sj, active);
    float pe_n = edge ? p[o - 1] : 0.0f, Fe_n = edge ? F[o - 1] : 0.0f;
    unsigned flags = 0;
    for (int i = ia; i <= ib; ++i, o += si) {
        const float4 p4 = p_n, F4 = F_n, us4 = us_n, vs4 = vs_n, ws4 = ws_n, pj4 = pj_n, Fj4 = Fj_n;
        const float pe = pe_n, Fe = Fe_n;
        if (i < ib) {
            const size_t on = o + si;
            p_n = ld4(p, on, active); F_n = ld4(F, on, active); us_n = ld4(us, on, active); vs_n = ld4(vs, on, active); ws_n = ld4(ws, on, active);
            pj_n = ld4(p, on - sj, active); Fj_n = ld4(F, on - sj, active);
            if (edge) { pe_n = p[on - 1]; Fe_n = F[on - 1]; }
        }
        const int gi = g.gi0 + i;
        const bool own = i >= own_a && i <= own_b;
        const float pc[4] = {p4.x, p4.y, p4.z, p4.w}, usv[4] = {us4.x, us4.y, us4.z, us4.w}, vsv[4] = {vs4.x, vs4.y, vs4.z, vs4.w};
        const float wsv[4] = {ws4.x, ws4.y, ws4.z, ws4.w}, pj[4] = {pj4.x, pj4.y, pj4.z, pj4.w};
        float rc[4], rj[4];
        rho4(F4, rc);
        rho4(Fj4, rj);
        float p_l = __shfl_up_sync(0xffffffffu, pc[3], 1), r_l = __shfl_up_sync(0xffffffffu, rc[3], 1);
        if (edge) { p_l = pe; r_l = rho3(Fe, c); }
        if (active) {
            float ou[4], ov[4], ow[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float ru = (rc[q] + rm[q]) * 0.5f;
                ou[q] = usv[q] - ((c.dt / ru) * (pc[q] - pm[q])) * c.dxi;
                const float rv = (rc[q] + rj[q]) * 0.5f;
                ov[q] = vsv[q] - ((c.dt / rv) * (pc[q] - pj[q])) * c.dyi;
                const float rw = (rc[q] + (q ? rc[q - 1] : r_l)) * 0.5f;
                ow[q] = wsv[q] - ((c.dt / rw) * (pc[q] - (q ? pc[q - 1] : p_l))) * c.dzi;
            }
            const bool urow = gi >= 2 && gi <= g.nx, vrow = gi >= 1 && gi <= g.nx;
            if (urow) {
                if (full) *reinterpret_cast<float4*>(u + o) = make_float4(ou[0], ou[1], ou[2], ou[3]);
                else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (colin[q]) u[o + q] = ou[q];
                }
                if (own) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) flags += colin[q] && (ou[q] * c.dt > c.cflx);
                }
            }
            if (vrow) {
                if (j >= 2) {
                    if (full) *reinterpret_cast<float4*>(v + o) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                    else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) if (colin[q]) v[o + q] = ov[q];
                    }
                    if (own) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) flags += colin[q] && (ov[q] * c.dt > c.cfly);
                    }
                }
                if (full && kl >= 2) *reinterpret_cast<float4*>(w + o) = make_float4(ow[0], ow[1], ow[2], ow[3]);
                else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (colin[q] && kl + q >= 2) w[o + q] = ow[q];
                }
                if (own) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) flags += colin[q] && kl + q >= 2 && (ow[q] * c.dt > c.cflx);   // 0.25*dx, 3dvof.py:301
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { pm[q] = pc[q]; rm[q] = rc[q]; }
    }
    if (flags) atomicAdd(courant_count, (unsigned long long)flags);
}

// ---- FCT sweeps (3dvof.py:366-541) -----------------------------------------------------------------------------------
struct Fct3C {
    float dt, dx, dy, dz, vol, dtd;   // dtd: dt*dy*dz | dt*dx*dz | dt*dx*dy by axis
    float m1, m2;                     // factors applied to (lo - hi) before the division (m2 = 1 means "absent")
    int has_m2;
    float d1;                         // divisor of the flux term: vol (x, z) or dx*dy (y)
    float qs;                         // dx (x, y) or dz (z) in the limiter ratios
    float d2;                         // dy (x, y) or dz (z) in the corrective update
};

__device__ __forceinline__ float f3_lo(float vel, float Fm, float Fc, float dt) { const float vd = vel * dt; return vel >= 0.0f ? vd * Fm : vd * Fc; }
__device__ __forceinline__ float f3_hi(float vel, float Fm, float Fc, float dt) { const float vd = vel * dt; return vel <= 0.0f ? vd * Fm : vd * Fc; }
// Exact short-cuts as in the 2-D sweeps (vof2d_fct.cuh): x + (+-0 / d) = x for the F >= +0 this code produces,
// 0 * dx * dy * dz / dv = 0 (div_nz keeps the sign), and where both faces of a cell carry no antidiffusive flux the
// limiter ratios are 0 and the corrective term is +-0.
__device__ __forceinline__ float f3_ftd(float Fc, float lo, float hi, float dv, const Fct3C& c) {
    const float d = lo - hi;
    float T = Fc;
    if (d != 0.0f) {
        float s = d * c.m1;
        if (c.has_m2) s = s * c.m2;
        T = Fc + div_nz(s, c.d1);
    }
    float t = div_nz(((T * c.dx) * c.dy) * c.dz, dv);
    if (t > 1.0f || t < 0.0f) t = var01(t);
    return t;
}
__device__ __forceinline__ void f3_ratios(float tm, float tc, float tp, float a_c, float a_p, const Fct3C& c, float& rp, float& rm) {
    rp = 0.0f; rm = 0.0f;
    if (a_c != 0.0f || a_p != 0.0f) {                 // otherwise pp = pm = 0 (or NaN never: a is finite or NaN != 0)
        const float pp = fmaxf(0.0f, a_c) - fminf(0.0f, a_p), pm = fmaxf(0.0f, a_p) - fminf(0.0f, a_c);
        if (pp > 0.0f) {
            const float fmax = fmaxf(fmaxf(tc, tm), tp);
            const float qp = (fmax - tc) * c.qs;
            rp = fminf(1.0f, div_nz(qp, pp));
        }
        if (pm > 0.0f) {
            const float fmin = fminf(fminf(tc, tm), tp);
            const float qm = (tc - fmin) * c.qs;
            rm = fminf(1.0f, div_nz(qm, pm));
        }
    }
}
__device__ __forceinline__ float f3_cface(float a_f, float rp_m, float rm_m, float rp_c, float rm_c) {
    return a_f >= 0.0f ? fminf(rp_c, rm_m) : fminf(rp_m, rm_c);
}
template <bool POST>
__device__ __forceinline__ float f3_update(float td, float a_c, float c_c, float a_p, float c_p, float dv, const Fct3C& c) {
    float fn = td;
    if (a_c != 0.0f || a_p != 0.0f) {
        const float t = a_p * c_p - a_c * c_c;
        fn = td - div_nz((((div_nz(t, c.d2)) * c.dx) * c.dy) * c.dz, dv);
    }
    float f = var01(fn);
    if (POST) f = var01(f);
    return f;
}

// The 3-D expressions as the policy of the second-generation kernels of vof2d_fct.cuh (the z-sweep runs on k_fct_y5:
// stencil along the contiguous axis by shuffles, rows = the (i, j) lines of the field).
struct FctOps3 {                                   // 3dvof.py:366-541
    using C = Fct3C;
    static __device__ __forceinline__ void face(float vel, float F_m, float F_c, const C& c, float& lo, float& a) {
        lo = f3_lo(vel, F_m, F_c, c.dt);
        a = f3_hi(vel, F_m, F_c, c.dt) - lo;
    }
    static __device__ __forceinline__ float lo_uniform(float vel, float c0, const C& c) { return (vel * c.dt) * c0; }
    static __device__ __forceinline__ float dv(const C& c, float dvel) { return c.vol - c.dtd * dvel; }
    static __device__ __forceinline__ float td_one(const C& c) {      // a full cell whose faces move alike: a constant
        float t = div_nz(((1.0f * c.dx) * c.dy) * c.dz, c.vol);
        if (t > 1.0f || t < 0.0f) t = var01(t);
        return t;
    }
    static __device__ __forceinline__ float ftd(float Fc, float lo, float hi, float dv, bool interior, const C& c, float td_one) {
        const float d = lo - hi;
        float T = Fc;
        if (d != 0.0f) {
            float s = d * c.m1;
            if (c.has_m2) s = s * c.m2;
            T = Fc + div_nz(s, c.d1);
        }
        float t;
        if (T == 1.0f && dv == c.vol) t = td_one;
        else {
            t = div_nz(((T * c.dx) * c.dy) * c.dz, dv);
            if (t > 1.0f || t < 0.0f) t = var01(t);
        }
        return interior ? t : 0.0f;
    }
    static __device__ __forceinline__ void ratios(float td_m, float td_c, float td_p, float a_c, float a_p, bool interior, const C& c,
                                                  float& rp, float& rm) {
        rp = 0.0f; rm = 0.0f;
        if (interior) f3_ratios(td_m, td_c, td_p, a_c, a_p, c, rp, rm);
    }
    static __device__ __forceinline__ float cface(float a_f, float rp_m, float rm_m, float rp_c, float rm_c, bool valid) {
        return valid ? f3_cface(a_f, rp_m, rm_m, rp_c, rm_c) : 0.0f;
    }
    template <bool POST>
    static __device__ __forceinline__ float update(float td_c, float a_c, float c_c, float a_p, float c_p, float dv, const C& c) {
        return f3_update<POST>(td_c, a_c, c_c, a_p, c_p, dv, c);
    }
};

// Sweep along a strided axis (x: stride pj, y: stride pk): one thread per line position, the chain of radius 3
// rolls through registers exactly as in the 2-D x-sweep.  `n` = interior cells along the axis, `goff` = global
// index of local index 0 along the axis (gi0 for x, 0 for y), `la..lb` = local interior range to produce.
template <bool POST>
__device__ __forceinline__ void f3_line(const float* __restrict__ Fin, const float* __restrict__ vel, float* __restrict__ Fout,
                                        size_t base, size_t st, int la, int lb, int lmax, int n, int goff, const Fct3C& c) {
    auto ldF = [&](int l) { return (l >= 0 && l <= lmax) ? Fin[base + (size_t)l * st] : 0.0f; };
    auto ldV = [&](int l) { return (l >= 0 && l <= lmax) ? vel[base + (size_t)l * st] : 0.0f; };
    auto interior = [&](int l) { const int gl = goff + l; return gl >= 1 && gl <= n; };
    float F1 = ldF(la - 3), u1 = 0.0f, lo1 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    float td1 = 0.0f, td2 = 0.0f, td3 = 0.0f, dv1 = 1.0f, dv2 = 1.0f, dv3 = 1.0f;
    float rp2 = 0.0f, rm2 = 0.0f, c2 = 0.0f;
    float Fk = ldF(la - 2), uk = ldV(la - 2);
    for (int k = la - 2; k <= lb + 3; ++k) {
        const float Fn = ldF(k + 1), un = ldV(k + 1);
        const float lo0 = f3_lo(uk, F1, Fk, c.dt);
        const float a0 = f3_hi(uk, F1, Fk, c.dt) - lo0;
        const float dv_n = c.vol - c.dtd * (uk - u1);
        const float td_n = interior(k - 1) ? f3_ftd(F1, lo1, lo0, dv_n, c) : 0.0f;
        td3 = td2; td2 = td1; td1 = td_n;
        dv3 = dv2; dv2 = dv1; dv1 = dv_n;
        float rp_n = 0.0f, rm_n = 0.0f;
        if (interior(k - 2)) f3_ratios(td3, td2, td1, a2, a1, c, rp_n, rm_n);
        const int gf = goff + k - 2;
        const float c_n = (gf >= 2 && gf <= n + 1) ? f3_cface(a2, rp2, rm2, rp_n, rm_n) : 0.0f;
        const int io = k - 3;
        if (io >= la && io <= lb && interior(io)) Fout[base + (size_t)io * st] = f3_update<POST>(td3, a3, c2, a2, c_n, dv3, c);
        rp2 = rp_n; rm2 = rm_n; c2 = c_n;
        a3 = a2; a2 = a1; a1 = a0; lo1 = lo0;
        F1 = Fk; Fk = Fn; u1 = uk; uk = un;
    }
}

// AXIS 0: threads over (j, k), march i.  AXIS 1: threads over (i, k), march j.  Cells the sweep never writes
// (ghosts) are copied so that the out-of-place result is a complete field.
template <int AXIS, bool POST>
__global__ void __launch_bounds__(kB3)
k3_fct_strided(Grid3 g, Fct3C c, const float* __restrict__ Fin, const float* __restrict__ vel, float* __restrict__ Fout,
               int ra, int rb, int rows_per_block) {
    const int k = blockIdx.x * kB3 + threadIdx.x;
    if (k > g.nz + 1) return;
    const size_t si = (size_t)g.pj, sj = (size_t)g.pk;
    if (AXIS == 0) {
        const int j = blockIdx.y;
        const int ia = ra + blockIdx.z * rows_per_block, ib = min(rb, ia + rows_per_block - 1);
        if (ia > ib) return;
        const size_t base = (size_t)j * sj + k;
        const bool lo_wall = g.gi0 + ia == 1, hi_wall = g.gi0 + ib == g.nx;
        if (lo_wall) Fout[base + (size_t)(ia - 1) * si] = Fin[base + (size_t)(ia - 1) * si];
        if (hi_wall) Fout[base + (size_t)(ib + 1) * si] = Fin[base + (size_t)(ib + 1) * si];
        if (j == 0 || j == g.ny + 1 || k == 0 || k == g.nz + 1) {
            for (int i = ia; i <= ib; ++i) Fout[base + (size_t)i * si] = Fin[base + (size_t)i * si];
            return;
        }
        f3_line<POST>(Fin, vel, Fout, base, si, ia, ib, g.nrows - 1, g.nx, g.gi0, c);
    } else {
        const int i = ra + blockIdx.y;     // every local plane, ghost planes included (copied)
        if (i > rb) return;
        const size_t base = (size_t)i * si + k;
        const int gi = g.gi0 + i;
        const bool line_in = gi >= 1 && gi <= g.nx && k >= 1 && k <= g.nz;
        Fout[base] = Fin[base];
        Fout[base + (size_t)(g.ny + 1) * sj] = Fin[base + (size_t)(g.ny + 1) * sj];
        if (!line_in) {
            for (int j = 1; j <= g.ny; ++j) Fout[base + (size_t)j * sj] = Fin[base + (size_t)j * sj];
            return;
        }
        f3_line<POST>(Fin, vel, Fout, base, sj, 1, g.ny, g.ny + 1, g.ny, 0, c);
    }
}

// z-sweep: the chain runs along the contiguous axis.  One thread per cell, a block stages a k-row segment (+3 halo)
// of F and w for a few (i, j) rows in shared memory and runs the chain in three barrier-separated phases.
template <bool POST, int TR, int TK>
__global__ void __launch_bounds__(256)
k3_fct_z(Grid3 g, Fct3C c, const float* __restrict__ Fin, const float* __restrict__ w, float* __restrict__ Fout,
         int ra, int rb) {
    constexpr int W = TK + 8;
    __shared__ float sF[TR][W], sV[TR][W], sTd[TR][W], sA[TR][W], sRp[TR][W], sRm[TR][W];
    const int nrows_total = (rb - ra + 1) * (g.ny + 2);        // (i, j) rows, ghost rows included (copied)
    const int row0 = blockIdx.y * TR;
    const int k0 = 1 + blockIdx.x * TK, kb = k0 - 4;
    const int tid = threadIdx.x;
    auto row_off = [&](int r, int& gi, int& j) -> size_t {
        const int rr = row0 + r;
        const int i = ra + rr / (g.ny + 2); j = rr - (rr / (g.ny + 2)) * (g.ny + 2);
        gi = g.gi0 + i;
        return (size_t)i * g.pj + (size_t)j * g.pk;
    };
    for (int q = tid; q < TR * W; q += 256) {
        const int r = q / W, s = q - r * W, k = kb + s;
        float f = 0.0f, vv = 0.0f;
        if (row0 + r < nrows_total && k >= 0 && k <= g.nz + 1) {
            int gi, j; const size_t o = row_off(r, gi, j) + k;
            f = Fin[o]; vv = w[o];
        }
        sF[r][s] = f; sV[r][s] = vv;
    }
    __syncthreads();
    for (int q = tid; q < TR * (TK + 4); q += 256) {
        const int r = q / (TK + 4), s = 2 + (q - r * (TK + 4)), k = kb + s;
        const float Fm = sF[r][s - 1], Fc = sF[r][s], Fp = sF[r][s + 1], vc = sV[r][s], vp = sV[r][s + 1];
        const float lo = f3_lo(vc, Fm, Fc, c.dt), hi = f3_lo(vp, Fc, Fp, c.dt);
        sA[r][s] = f3_hi(vc, Fm, Fc, c.dt) - lo;
        const float dv = c.vol - c.dtd * (vp - vc);
        sTd[r][s] = (k >= 1 && k <= g.nz) ? f3_ftd(Fc, lo, hi, dv, c) : 0.0f;
    }
    __syncthreads();
    for (int q = tid; q < TR * (TK + 2); q += 256) {
        const int r = q / (TK + 2), s = 3 + (q - r * (TK + 2)), k = kb + s;
        float rp = 0.0f, rm = 0.0f;
        if (k >= 1 && k <= g.nz) f3_ratios(sTd[r][s - 1], sTd[r][s], sTd[r][s + 1], sA[r][s], sA[r][s + 1], c, rp, rm);
        sRp[r][s] = rp; sRm[r][s] = rm;
    }
    __syncthreads();
    for (int q = tid; q < TR * TK; q += 256) {
        const int r = q / TK, s = 4 + (q - r * TK), k = kb + s;
        if (row0 + r >= nrows_total || k > g.nz) continue;
        int gi, j; const size_t o = row_off(r, gi, j) + k;
        if (gi < 1 || gi > g.nx || j < 1 || j > g.ny) { Fout[o] = sF[r][s]; continue; }
        const float a_c = sA[r][s], a_p = sA[r][s + 1];
        const float c_c = (k >= 2) ? f3_cface(a_c, sRp[r][s - 1], sRm[r][s - 1], sRp[r][s], sRm[r][s]) : 0.0f;
        const float c_p = f3_cface(a_p, sRp[r][s], sRm[r][s], sRp[r][s + 1], sRm[r][s + 1]);
        const float dv = c.vol - c.dtd * (sV[r][s + 1] - sV[r][s]);
        Fout[o] = f3_update<POST>(sTd[r][s], a_c, c_c, a_p, c_p, dv, c);
    }
    if (blockIdx.x == 0 && tid < TR && row0 + tid < nrows_total) {   // ghost columns k = 0, nz+1
        int gi, j; const size_t o = row_off(tid, gi, j);
        Fout[o] = Fin[o];
        Fout[o + g.nz + 1] = Fin[o + g.nz + 1];
    }
}

// ---- post_process_f (3dvof.py:544-547) and set_init_F (126-138) -------------------------------------------------------
__global__ void __launch_bounds__(kB3)
k3_post_process_f(Grid3 g, float* __restrict__ F, int r0, int r1) {
    const int k = blockIdx.x * kB3 + threadIdx.x, j = blockIdx.y;
    if (k > g.nz + 1) return;
    for (int i = r0 + blockIdx.z; i <= r1; i += gridDim.z) {
        const size_t o = (size_t)i * g.pj + (size_t)j * g.pk + k;
        F[o] = var3(F[o], 0.0f, 1.0f);
    }
}

__global__ void __launch_bounds__(kB3)
k3_set_init_F(Grid3 g, float x2, float y2, float z2, const float* __restrict__ xs, const float* __restrict__ ys,
              const float* __restrict__ zs, float* __restrict__ F, int r0, int r1) {
    const int k = blockIdx.x * kB3 + threadIdx.x, j = blockIdx.y;
    if (k > g.nz + 1) return;
    const float yj = ys[j], zk = zs[k];
    for (int i = r0 + blockIdx.z; i <= r1; i += gridDim.z) {
        const int gi = g.gi0 + i;
        if (gi < 0 || gi > g.nx + 1) continue;
        const float xi = xs[gi];
        if (xi >= 0.0f && xi <= x2 && yj >= 0.0f && yj <= y2 && zk >= 0.0f && zk <= z2) F[(size_t)i * g.pj + (size_t)j * g.pk + k] = 1.0f;
    }
}

// ---- diagnostics ----------------------------------------------------------------------------------------------------------
struct Diag3 {
    double mass;
    unsigned int max_cfl_bits;
    unsigned int pad;
    unsigned long long courant_count;
    unsigned int wq[2];          // WorkQueue counters of the queue-scheduled kernels (vof2d_stream.cuh), zero between launches
};

__global__ void __launch_bounds__(256)
k3_diag(Grid3 g, Consts3 c, const float* __restrict__ F, const float* __restrict__ u, const float* __restrict__ v,
        const float* __restrict__ w, Diag3* __restrict__ d, int r0, int r1) {
    __shared__ double s_mass[8];
    __shared__ float s_cfl[8];
    double mass = 0.0; float cfl = 0.0f;
    for (int i = r0 + blockIdx.z; i <= r1; i += gridDim.z) {
        const int gi = g.gi0 + i;
        if (gi < 1 || gi > g.nx) continue;
        for (int j = 1 + blockIdx.y; j <= g.ny; j += gridDim.y)
            for (int k = 1 + threadIdx.x; k <= g.nz; k += 256) {
                const size_t o = (size_t)i * g.pj + (size_t)j * g.pk + k;
                mass += (double)F[o];
                cfl = fmaxf(cfl, fmaxf(fmaxf(fabsf(u[o]) * c.dt * c.dxi, fabsf(v[o]) * c.dt * c.dyi), fabsf(w[o]) * c.dt * c.dzi));
            }
    }
    mass = warp_sum(mass); cfl = warp_max(cfl);
    const int wdx = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_mass[wdx] = mass; s_cfl[wdx] = cfl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) { mass += s_mass[q]; cfl = fmaxf(cfl, s_cfl[q]); }
        atomicAdd(&d->mass, mass);
        atomicMax(&d->max_cfl_bits, __float_as_uint(cfl));
    }
}

}  // namespace vof
