// sm_100a kernels of the 2-D VOF step.  One kernel (group) per reference kernel of
// /root/reference/2dvof.py; each cites the lines it replaces.  All are HBM-bound stencils:
// threads map to the contiguous axis j (coalesced 128-byte warp accesses, interior column 1
// on a 128-byte boundary), blocks tile i, values that are reused along i stay in registers
// (row marching) and values reused along j go through shared memory.
#pragma once
#include "vof_common.cuh"

namespace vof {

constexpr int kBlockJ = 128;   // threads along j for the marching kernels

// ======================================================================================
// cal_nu_rho  (2dvof.py:198-203) -- all cells including ghosts.  Sequence mode only: the
// fused step recomputes rho/nu from F in registers and never stores them.
// ======================================================================================
__global__ void __launch_bounds__(kBlockJ)
k_cal_nu_rho(Grid g, Consts c, const float* __restrict__ F, float* __restrict__ rho,
             float* __restrict__ nu, int r0, int r1, int rows_per_block) {
    const int j = blockIdx.x * kBlockJ + threadIdx.x;
    if (j > g.ny + 1) return;
    const int ia = r0 + blockIdx.y * rows_per_block;
    const int ib = min(r1, ia + rows_per_block - 1);
    for (int i = ia; i <= ib; ++i) {
        const size_t o = (size_t)i * g.pitch + j;
        const float f = F[o];
        rho[o] = rho_of(f, c);
        nu[o] = nu_of(f, c);
    }
}

// ======================================================================================
// get_normal_young  (2dvof.py:283-309): Youngs normal -> unit normal -> curvature.
// The reference stores 13 scratch arrays; here a tile of F goes to shared memory once, the
// four corner gradients of a cell are recognised as the SAME expression evaluated at the four
// surrounding corners (mx1(i,j) == mx3(i+1,j+1) == mx2(i,j+1) == mx4(i+1,j) term for term), so
// each corner is computed once, then the normals on tile+1, then kappa.  R F, W kappa = 8 B/cell.
// mx/my of cells outside the global interior are 0 (never written in the reference).
// ======================================================================================
template <int TI, int TJ>
__global__ void __launch_bounds__(256)
k_kappa(Grid g, Consts c, const float* __restrict__ F, float* __restrict__ kappa, int r0, int r1) {
    constexpr int FW = TJ + 4, FH = TI + 4;     // F tile (halo 2)
    constexpr int CW = TJ + 3, CH = TI + 3;     // corners
    constexpr int MW = TJ + 2, MH = TI + 2;     // normals (halo 1)
    __shared__ float sF[FH][FW + 1];
    __shared__ float sG[CH][CW + 1], sH[CH][CW + 1];
    __shared__ float sMx[MH][MW + 1], sMy[MH][MW + 1];
    const int ti0 = r0 + blockIdx.y * TI;       // first output row (local)
    const int tj0 = 1 + blockIdx.x * TJ;        // first output column
    const int tid = threadIdx.x;

    for (int k = tid; k < FH * FW; k += 256) {
        const int li = k / FW, lj = k - li * FW;
        const int i = ti0 - 2 + li, j = tj0 - 2 + lj;
        float f = 0.0f;
        if (i >= 0 && i < g.nrows && j >= 0 && j <= g.ny + 1) f = F[(size_t)i * g.pitch + j];
        sF[li][lj] = f;
    }
    __syncthreads();
    // corner (li, lj) sits between cells (li, lj), (li+1, lj), (li, lj+1), (li+1, lj+1) of the F tile
    for (int k = tid; k < CH * CW; k += 256) {
        const int li = k / CW, lj = k - li * CW;
        const float f00 = sF[li][lj], f01 = sF[li][lj + 1], f10 = sF[li + 1][lj], f11 = sF[li + 1][lj + 1];
        sG[li][lj] = c.m1_2dx * (((f11 + f10) - f01) - f00);   // 2dvof.py:287 (mx1) and its three aliases
        sH[li][lj] = c.m1_2dy * (((f11 - f10) + f01) - f00);   // 2dvof.py:288 (my1)
    }
    __syncthreads();
    for (int k = tid; k < MH * MW; k += 256) {
        const int li = k / MW, lj = k - li * MW;
        const int gi = g.gi0 + ti0 - 1 + li, j = tj0 - 1 + lj;
        float mx = 0.0f, my = 0.0f;
        if (gi >= 1 && gi <= g.nx && j >= 1 && j <= g.ny) {
            // mx1 = G(i,j), mx2 = G(i,j-1), mx3 = G(i-1,j-1), mx4 = G(i-1,j); 2dvof.py:296-297
            const float mxs = (((sG[li + 1][lj + 1] + sG[li + 1][lj]) + sG[li][lj]) + sG[li][lj + 1]) / 4.0f;
            const float mys = (((sH[li + 1][lj + 1] + sH[li + 1][lj]) + sH[li][lj]) + sH[li][lj + 1]) / 4.0f;
            if (fabsf(mxs) < 1e-10f && fabsf(mys) < 1e-10f) {   // 2dvof.py:300-302
                mx = mxs; my = mys;
            } else {                                            // 2dvof.py:304-306
                const float mag = sqrtf(mxs * mxs + mys * mys);
                mx = mxs / mag; my = mys / mag;
            }
        }
        sMx[li][lj] = mx; sMy[li][lj] = my;
    }
    __syncthreads();
    for (int k = tid; k < TI * TJ; k += 256) {
        const int li = k / TJ, lj = k - li * TJ;
        const int i = ti0 + li, j = tj0 + lj;
        if (i <= r1 && j <= g.ny) {
            // 2dvof.py:308-309
            kappa[(size_t)i * g.pitch + j] =
                -(c.i_dx_2 * (sMx[li + 2][lj + 1] - sMx[li][lj + 1]) + c.i_dy_2 * (sMy[li + 1][lj + 2] - sMy[li + 1][lj]));
        }
    }
}

// ======================================================================================
// advect_upwind  (2dvof.py:206-233): u*, v* in one pass (the reference makes two).
// INLINE_PROPS: rho/nu recomputed from F (fused step) instead of read from their arrays.
// Thread = one column j, marching over a chunk of rows; i-neighbours roll through registers.
// ======================================================================================
template <bool INLINE_PROPS>
__global__ void __launch_bounds__(kBlockJ)
k_advect(Grid g, Consts c, const float* __restrict__ u, const float* __restrict__ v,
         const float* __restrict__ F, const float* __restrict__ kappa, const float* __restrict__ rho,
         const float* __restrict__ nu, float* __restrict__ us, float* __restrict__ vs, int r0, int r1,
         int rows_per_block) {
    const int j = 1 + blockIdx.x * kBlockJ + threadIdx.x;
    if (j > g.ny) return;
    const int ia = r0 + blockIdx.y * rows_per_block;
    const int ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const int P = g.pitch;
    size_t o = (size_t)ia * P + j;
    // rolling state along i: rows i-1 (m), i (c); row i+1 (p) is loaded each iteration
    float u_m = u[o - P], u_c = u[o], u_c_jm = u[o - 1];
    float v_m = v[o - P], v_c = v[o], v_m_jp = v[o - P + 1], v_c_jp = v[o + 1];
    float F_m = F[o - P], F_c = F[o];
    float k_m = kappa[o - P], k_c = kappa[o];
    for (int i = ia; i <= ib; ++i, o += P) {
        const int gi = g.gi0 + i;
        const float u_p = u[o + P], v_p = v[o + P];
        const float u_c_jp = u[o + 1];
        const float u_p_jm = u[o + P - 1];
        const float v_c_jm = v[o - 1];
        const float F_jm = F[o - 1], k_jm = kappa[o - 1];
        float rho_c, rho_m, rho_jm, nu_c;
        if (INLINE_PROPS) {
            rho_c = rho_of(F_c, c); rho_m = rho_of(F_m, c); rho_jm = rho_of(F_jm, c); nu_c = nu_of(F_c, c);
        } else {
            rho_c = rho[o]; rho_m = rho[o - P]; rho_jm = rho[o - 1]; nu_c = nu[o];
        }
        if (gi >= 2 && gi <= g.nx) {   // 2dvof.py:208-220
            const float v_here = 0.25f * (((v_m + v_m_jp) + v_c) + v_c_jp);
            const float dudx = u_c > 0.0f ? (u_c - u_m) * c.dxi : (u_p - u_c) * c.dxi;
            const float dudy = v_here > 0.0f ? (u_c - u_c_jm) * c.dyi : (u_c_jp - u_c) * c.dyi;
            const float kappa_ave = (k_c + k_m) / 2.0f;
            const float fx_kappa = ((c.neg_sigma * (F_c - F_m)) * kappa_ave) / c.dx;
            float acc = (nu_c * ((u_m - 2.0f * u_c) + u_p)) * c.dxi2;
            acc = acc + (nu_c * ((u_c_jm - 2.0f * u_c) + u_c_jp)) * c.dyi2;
            acc = acc - u_c * dudx;
            acc = acc - v_here * dudy;
            acc = acc + c.gx;
            acc = acc + (fx_kappa * 2.0f) / (rho_c + rho_m);
            us[o] = u_c + c.dt * acc;
        }
        if (j >= 2 && gi >= 1 && gi <= g.nx) {   // 2dvof.py:221-233
            const float u_p_j = u_p;
            const float u_here = 0.25f * (((u_c_jm + u_c) + u_p_jm) + u_p_j);
            const float dvdx = u_here > 0.0f ? (v_c - v_m) * c.dxi : (v_p - v_c) * c.dxi;
            const float dvdy = v_c > 0.0f ? (v_c - v_c_jm) * c.dyi : (v_c_jp - v_c) * c.dyi;
            const float kappa_ave = (k_c + k_jm) / 2.0f;
            const float fy_kappa = ((c.neg_sigma * (F_c - F_jm)) * kappa_ave) / c.dy;
            float acc = (nu_c * ((v_m - 2.0f * v_c) + v_p)) * c.dxi2;
            acc = acc + (nu_c * ((v_c_jm - 2.0f * v_c) + v_c_jp)) * c.dyi2;
            acc = acc - u_here * dvdx;
            acc = acc - v_c * dvdy;
            acc = acc + c.gy;
            acc = acc + (fy_kappa * 2.0f) / (rho_c + rho_jm);
            vs[o] = v_c + c.dt * acc;
        }
        // roll
        u_m = u_c; u_c = u_p; u_c_jm = u_p_jm;
        v_m = v_c; v_c = v_p; v_m_jp = v_c_jp; v_c_jp = v[o + P + 1];
        F_m = F_c; F_c = F[o + P];
        k_m = k_c; k_c = kappa[o + P];
    }
}

// ======================================================================================
// set_BC  (2dvof.py:162-189): row loop then column loop.  One race-free launch: every output
// is expressed through PRE-call interior values (corner = loop B applied to loop A's result).
// Threads [0, nA) handle loop A (y-walls) for local rows ra0..ra1; then nB threads per x-wall.
// mask bit: 1 u, 2 v, 4 F, 8 p, 16 rho.
// ======================================================================================
__global__ void __launch_bounds__(128)
k_set_bc(Grid g, float* __restrict__ u, float* __restrict__ v, float* __restrict__ F,
         float* __restrict__ p, float* __restrict__ rho, int ra0, int ra1, int has_lo, int has_hi,
         unsigned mask) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    const int nA = ra1 - ra0 + 1;
    const int nB = g.ny + 2;
    const int P = g.pitch, ny = g.ny;
    float* X[3] = {F, p, rho};
    if (t < nA) {
        const int i = ra0 + t, gi = g.gi0 + i;
        const size_t o = (size_t)i * P;
        const bool ghost_row = (gi == 0 || gi == g.nx + 1);
        const bool wall_face_row = (gi == 1 || gi == g.nx + 1);   // u = 0 there (loop B wins)
        if ((mask & 1u) && !wall_face_row) {
            u[o + 0] = u[o + 1];
            u[o + ny + 1] = u[o + ny];
        }
        if ((mask & 2u) && !ghost_row) {
            v[o + 1] = 0.0f;
            v[o + ny + 1] = 0.0f;
        }
        if (!ghost_row) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (mask & (4u << k)) {
                    X[k][o + 0] = X[k][o + 1];
                    X[k][o + ny + 1] = X[k][o + ny];
                }
        }
        return;
    }
    int tb = t - nA;
    int side = -1;
    if (has_lo) { if (tb < nB) side = 0; else tb -= nB; }
    if (side < 0 && has_hi) { if (tb < nB) side = 1; }
    if (side < 0) return;
    const int j = tb;
    const int jc = min(max(j, 1), ny);                       // column whose loop-A value lands in (·, j)
    const int gi_in = side == 0 ? 1 : g.nx;                  // interior row copied from
    const int gi_gh = side == 0 ? 0 : g.nx + 1;              // ghost row written
    const int gi_wall = side == 0 ? 1 : g.nx + 1;            // u-face row set to 0
    const size_t o_in = (size_t)(gi_in - g.gi0) * P, o_gh = (size_t)(gi_gh - g.gi0) * P;
    if (mask & 1u) u[(size_t)(gi_wall - g.gi0) * P + j] = 0.0f;
    if (mask & 2u) v[o_gh + j] = (j == 1 || j == ny + 1) ? 0.0f : v[o_in + j];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (mask & (4u << k)) X[k][o_gh + j] = X[k][o_in + jc];
}

// ======================================================================================
// Poisson right-hand side, hoisted out of the sweeps (2dvof.py:239-241; value-identical in
// all sweeps of a step).
// ======================================================================================
__device__ __forceinline__ float poisson_rhs(float rho_c, float us_c, float us_p, float vs_c, float vs_jp,
                                             const Consts& c) {
    return (rho_c / c.dt) * ((us_p - us_c) * c.dxi + (vs_jp - vs_c) * c.dyi);
}

template <bool INLINE_PROPS>
__global__ void __launch_bounds__(kBlockJ)
k_rhs(Grid g, Consts c, const float* __restrict__ rhoF, const float* __restrict__ us,
      const float* __restrict__ vs, float* __restrict__ rhs, int r0, int r1, int rows_per_block) {
    const int j = 1 + blockIdx.x * kBlockJ + threadIdx.x;
    if (j > g.ny) return;
    const int ia = r0 + blockIdx.y * rows_per_block;
    const int ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const int P = g.pitch;
    size_t o = (size_t)ia * P + j;
    float us_c = us[o];
    for (int i = ia; i <= ib; ++i, o += P) {
        const float us_p = us[o + P];
        const float r = INLINE_PROPS ? rho_of(rhoF[o], c) : rhoF[o];
        rhs[o] = poisson_rhs(r, us_c, us_p, vs[o], vs[o + 1], c);
        us_c = us_p;
    }
}

// ======================================================================================
// solve_p_jacobi  (2dvof.py:236-266) -- ONE sweep p -> pn (ping-pong instead of the
// reference's pt->p copy loop).  Walls are zeroed coefficients, not ghosts (258-262).
// Non-interior cells of this context's rows are copied so pn is a complete field.
// RHS_MODE 0: read hoisted rhs; 1: recompute from rho array (reference structure);
// 2: recompute with rho inline from F.
// ======================================================================================
template <int RHS_MODE>
__global__ void __launch_bounds__(kBlockJ)
k_jacobi(Grid g, Consts c, const float* __restrict__ p, float* __restrict__ pn,
         const float* __restrict__ rhs, const float* __restrict__ rhoF, const float* __restrict__ us,
         const float* __restrict__ vs, int r0, int r1, int rows_per_block) {
    const int j = blockIdx.x * kBlockJ + threadIdx.x;
    if (j > g.ny + 1) return;
    const int ia = r0 + blockIdx.y * rows_per_block;
    const int ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const int P = g.pitch;
    size_t o = (size_t)ia * P + j;
    const bool jin = (j >= 1 && j <= g.ny);
    const float an = (j != g.ny) ? c.dyi2 : 0.0f;
    const float as = (j != 1) ? c.dyi2 : 0.0f;
    float p_m = ia > 0 ? p[o - P] : 0.0f, p_c = p[o];
    float us_c = (RHS_MODE != 0 && jin) ? us[o] : 0.0f;
    for (int i = ia; i <= ib; ++i, o += P) {
        const int gi = g.gi0 + i;
        const float p_p = (i + 1 < g.nrows) ? p[o + P] : 0.0f;
        float out = p_c;
        float us_p = 0.0f;
        if (RHS_MODE != 0 && jin && i + 1 < g.nrows) us_p = us[o + P];
        if (jin && gi >= 1 && gi <= g.nx) {
            float b;
            if (RHS_MODE == 0) b = rhs[o];
            else {
                const float r = RHS_MODE == 2 ? rho_of(rhoF[o], c) : rhoF[o];
                b = poisson_rhs(r, us_c, us_p, vs[o], vs[o + 1], c);
            }
            const float ae = (gi != g.nx) ? c.dxi2 : 0.0f;
            const float aw = (gi != 1) ? c.dxi2 : 0.0f;
            const float ap = -1.0f * (((ae + aw) + an) + as);
            float t = b - ae * p_p;
            t = t - aw * p_m;
            t = t - an * p[o + 1];
            t = t - as * p[o - 1];
            out = t / ap;
        }
        pn[o] = out;
        p_m = p_c; p_c = p_p; us_c = us_p;
    }
}

// ======================================================================================
// update_uv  (2dvof.py:269-280): projection.  The reference prints per offending face when
// u*dt > 0.25*dx; here offenders are counted (one atomic per offending thread, normally none).
// ======================================================================================
template <bool INLINE_PROPS>
__global__ void __launch_bounds__(kBlockJ)
k_project(Grid g, Consts c, const float* __restrict__ rhoF, const float* __restrict__ p,
          const float* __restrict__ us, const float* __restrict__ vs, float* __restrict__ u,
          float* __restrict__ v, unsigned long long* __restrict__ courant_count, int r0, int r1,
          int rows_per_block, int own_a, int own_b) {
    const int j = 1 + blockIdx.x * kBlockJ + threadIdx.x;
    if (j > g.ny) return;
    const int ia = r0 + blockIdx.y * rows_per_block;
    const int ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const int P = g.pitch;
    size_t o = (size_t)ia * P + j;
    float p_m = p[o - P];
    float rho_m = INLINE_PROPS ? rho_of(rhoF[o - P], c) : rhoF[o - P];
    unsigned flags = 0;
    for (int i = ia; i <= ib; ++i, o += P) {
        const int gi = g.gi0 + i;
        const float p_c = p[o];
        const float rho_c = INLINE_PROPS ? rho_of(rhoF[o], c) : rhoF[o];
        if (gi >= 2 && gi <= g.nx) {
            const float r = (rho_c + rho_m) * 0.5f;
            const float un = us[o] - ((c.dt / r) * (p_c - p_m)) * c.dxi;
            u[o] = un;
            flags += (un * c.dt > c.cflx) && i >= own_a && i <= own_b;
        }
        if (j >= 2 && gi >= 1 && gi <= g.nx) {
            const float rho_jm = INLINE_PROPS ? rho_of(rhoF[o - 1], c) : rhoF[o - 1];
            const float r = (rho_c + rho_jm) * 0.5f;
            const float vn = vs[o] - ((c.dt / r) * (p_c - p[o - 1])) * c.dyi;
            v[o] = vn;
            flags += (vn * c.dt > c.cfly) && i >= own_a && i <= own_b;
        }
        p_m = p_c; rho_m = rho_c;
    }
    if (flags) atomicAdd(courant_count, (unsigned long long)flags);
}

// ======================================================================================
// FCT building blocks shared by both sweeps (2dvof.py:321-448).  `m`/`c`/`p` = minus /
// centre / plus along the sweep direction.
// ======================================================================================
// low-order (donor-cell) face flux, 2dvof.py:325-326 / 391-392
__device__ __forceinline__ float flux_lo(float vel, float F_m, float F_c, float dt) {
    const float vd = vel * dt;
    return vel >= 0.0f ? vd * F_m : vd * F_c;
}
// high-order = DOWNWIND face flux, 2dvof.py:342-343 / 408-409
__device__ __forceinline__ float flux_hi(float vel, float F_m, float F_c, float dt) {
    const float vd = vel * dt;
    return vel <= 0.0f ? vd * F_m : vd * F_c;
}
// transported-diffused value, 2dvof.py:329-331 / 393-395 (`lo` = flux through the minus face,
// `hi` = through the plus face; the other direction's pair is the literal 0 in the source)
__device__ __forceinline__ float fct_ftd(float F_c, float lo, float hi, float dv, const Consts& c) {
    const float sum = lo - hi;
    float t = ((((F_c + (sum * c.dy) / c.dxdy) * c.dx) * c.dy)) / dv;
    if (t > 1.0f || t < 0.0f) t = var01(t);
    return t;
}
// Zalesak limiter ratios of one cell, 2dvof.py:334-335, 352-363 / 398-399, 416-427
__device__ __forceinline__ void fct_ratios(float td_m, float td_c, float td_p, float a_c, float a_p,
                                           const Consts& c, float& rp, float& rm) {
    const float fmax = fmaxf(fmaxf(td_c, td_m), td_p);
    const float fmin = fminf(fminf(td_c, td_m), td_p);
    const float pp = fmaxf(0.0f, a_c) - fminf(0.0f, a_p);
    const float qp = (fmax - td_c) * c.dx;
    rp = pp > 0.0f ? fminf(1.0f, qp / pp) : 0.0f;
    const float pm = fmaxf(0.0f, a_p) - fminf(0.0f, a_c);
    const float qm = (td_c - fmin) * c.dx;
    rm = pm > 0.0f ? fminf(1.0f, qm / pm) : 0.0f;
}
// face limiter, 2dvof.py:366-369 / 435-438 (rp/rm of the minus cell = *_m, of the plus cell = *_c)
__device__ __forceinline__ float fct_cface(float a_f, float rp_m, float rm_m, float rp_c, float rm_c) {
    return a_f >= 0.0f ? fminf(rp_c, rm_m) : fminf(rp_m, rm_c);
}
// final update, 2dvof.py:377-382 / 442-448 (+ post_process_f 452-455 when POST)
template <bool POST>
__device__ __forceinline__ float fct_update(float td_c, float a_c, float c_c, float a_p, float c_p, float dv,
                                            const Consts& c) {
    const float t = a_p * c_p - a_c * c_c;
    const float fn = td_c - (((t / c.dy) * c.dx) * c.dy) / dv;
    float f = var01(fn);
    if (POST) f = var01(f);
    return f;
}

// ======================================================================================
// fct_x_sweep  (2dvof.py:321-382): the four loops fused into one pass.  The sweep couples
// cells only along i, so each thread owns a column j and marches up a chunk of rows with the
// whole dependency chain (F[i-3..i+3], u[i-2..i+3]) rolling through registers: R F, R u, W F =
// 12 B/cell, no scratch arrays, no shared memory.  Out of place (Fin -> Fout) because a chunk's
// warm-up rows are another chunk's outputs.  Ghost columns/rows are copied through.
// Ghost semantics of the reference's never-written scratch: Ftd, rp, rm = 0 outside the global
// interior; cx = 0 on face 1 (and outside [2, nx+1]).
// ======================================================================================
template <bool POST>
__global__ void __launch_bounds__(kBlockJ)
k_fct_x(Grid g, Consts c, const float* __restrict__ Fin, const float* __restrict__ u,
        float* __restrict__ Fout, int r0, int r1, int rows_per_block) {
    const int j = blockIdx.x * kBlockJ + threadIdx.x;
    if (j > g.ny + 1) return;
    const int ia = r0 + blockIdx.y * rows_per_block;
    const int ib = min(r1, ia + rows_per_block - 1);
    if (ia > ib) return;
    const int P = g.pitch, nrows = g.nrows;
    const float* Fc = Fin + j;
    const float* uc = u + j;
    float* Fo = Fout + j;
    if (j == 0 || j == g.ny + 1) {   // ghost columns: unchanged by the sweep
        int a = ia, b = ib;
        if (g.gi0 + ia == 1) a = ia - 1;
        if (g.gi0 + ib == g.nx) b = ib + 1;
        for (int i = a; i <= b; ++i) Fo[(size_t)i * P] = Fc[(size_t)i * P];
        return;
    }
    if (g.gi0 + ia == 1) Fo[(size_t)(ia - 1) * P] = Fc[(size_t)(ia - 1) * P];       // ghost row 0
    if (g.gi0 + ib == g.nx) Fo[(size_t)(ib + 1) * P] = Fc[(size_t)(ib + 1) * P];    // ghost row nx+1

    auto ldF = [&](int i) { return (i >= 0 && i < nrows) ? Fc[(size_t)i * P] : 0.0f; };
    auto ldU = [&](int i) { return (i >= 0 && i < nrows) ? uc[(size_t)i * P] : 0.0f; };
    auto interior = [&](int i) { const int gi = g.gi0 + i; return gi >= 1 && gi <= g.nx; };

    // rolling registers; suffix = age in rows behind the row k being loaded
    float F1 = ldF(ia - 3);              // F[k-1]
    float u1 = 0.0f;                      // u[k-1]
    float lo1 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;   // face k-1.. : low flux, antidiffusive flux
    float td1 = 0.0f, td2 = 0.0f, td3 = 0.0f;            // Ftd[k-1], [k-2], [k-3]
    float dv1 = 1.0f, dv2 = 1.0f, dv3 = 1.0f;            // dv[k-1] ..
    float rp2 = 0.0f, rm2 = 0.0f, rp3 = 0.0f, rm3 = 0.0f;// ratios of cells k-2, k-3
    float c2 = 0.0f, c3 = 0.0f;                          // face limiter of faces k-2, k-3
    float Fk = ldF(ia - 2), uk = ldU(ia - 2);
    for (int k = ia - 2; k <= ib + 3; ++k) {
        const float Fn = ldF(k + 1), un = ldU(k + 1);    // prefetch next row
        // face k (between cells k-1 and k)
        const float lo0 = flux_lo(uk, F1, Fk, c.dt);
        const float a0 = flux_hi(uk, F1, Fk, c.dt) - lo0;
        // cell k-1: Ftd, dv                                      (loop 1)
        const float dv_n = c.dxdy - c.dtdy * (uk - u1);
        const float td_n = interior(k - 1) ? fct_ftd(F1, lo1, lo0, dv_n, c) : 0.0f;
        // roll cell-centred pipelines: now td1 = Ftd[k-1]
        td3 = td2; td2 = td1; td1 = td_n;
        dv3 = dv2; dv2 = dv1; dv1 = dv_n;
        // cell k-2: ratios need Ftd[k-3..k-1], a[k-2] (= a2 before roll... see below)
        // antidiffusive fluxes: before rolling, a1 = a[k-1], a2 = a[k-2], a3 = a[k-3]
        float rp_n = 0.0f, rm_n = 0.0f;
        if (interior(k - 2)) fct_ratios(td3, td2, td1, a2, a1, c, rp_n, rm_n);   // (loop 2)
        // face k-2 limiter needs ratios of cells k-3 (rp2/rm2 before roll) and k-2 (new)
        const int gf = g.gi0 + k - 2;
        const float c_n = (gf >= 2 && gf <= g.nx + 1) ? fct_cface(a2, rp2, rm2, rp_n, rm_n) : 0.0f;   // (loop 3)
        // cell k-3 update needs Ftd[k-3] (= td3... careful: after roll td3 = Ftd[k-3]), a[k-3] = a3,
        // c[k-3] = c2 (before roll), a[k-2] = a2, c[k-2] = c_n, dv[k-3] = dv3
        const int io = k - 3;
        if (io >= ia && io <= ib && interior(io))
            Fo[(size_t)io * P] = fct_update<POST>(td3, a3, c2, a2, c_n, dv3, c);  // (loop 4)
        // roll face / ratio pipelines
        rp3 = rp2; rm3 = rm2; rp2 = rp_n; rm2 = rm_n;
        c3 = c2; c2 = c_n;
        a3 = a2; a2 = a1; a1 = a0; lo1 = lo0;
        F1 = Fk; Fk = Fn; u1 = uk; uk = un;
    }
    (void)rp3; (void)rm3; (void)c3;
}

// ======================================================================================
// fct_y_sweep  (2dvof.py:385-448): same chain along the contiguous axis j.  A block stages
// TR rows x (TJ + halo) columns of F and v in shared memory; the chain runs in three phases
// (Ftd + antidiffusive flux | limiter ratios | face limiter + update) with a barrier between
// them; phases read neighbours along j from shared memory.  R F, R v, W F = 12 B/cell.
// ======================================================================================
template <bool POST, int TR, int TJ>
__global__ void __launch_bounds__(256)
k_fct_y(Grid g, Consts c, const float* __restrict__ Fin, const float* __restrict__ v,
        float* __restrict__ Fout, int r0, int r1) {
    constexpr int W = TJ + 8;          // staged columns: j0-4 .. j0+TJ+3 (one spare keeps 16-byte alignment)
    __shared__ float sF[TR][W], sV[TR][W], sTd[TR][W], sA[TR][W], sRp[TR][W], sRm[TR][W];
    const int i0 = r0 + blockIdx.y * TR;
    const int j0 = 1 + blockIdx.x * TJ;          // first output column
    const int jb = j0 - 4;                       // column of staged index 0
    const int tid = threadIdx.x;
    const int P = g.pitch, ny = g.ny;

    for (int k = tid; k < TR * W; k += 256) {
        const int r = k / W, s = k - r * W;
        const int i = i0 + r, j = jb + s;
        float f = 0.0f, vv = 0.0f;
        if (i <= r1 && j >= 0 && j <= ny + 1) {
            const size_t o = (size_t)i * P + j;
            f = Fin[o]; vv = v[o];
        }
        sF[r][s] = f; sV[r][s] = vv;
    }
    __syncthreads();
    // phase A: cells s in [2, TJ+6): Ftd[j], a[j] (face j = between j-1 and j)      (loops 1 and 2a)
    for (int k = tid; k < TR * (TJ + 4); k += 256) {
        const int r = k / (TJ + 4), s = 2 + (k - r * (TJ + 4));
        const int j = jb + s;
        const float F_m = sF[r][s - 1], F_c = sF[r][s], F_p = sF[r][s + 1];
        const float v_c = sV[r][s], v_p = sV[r][s + 1];
        const float lo = flux_lo(v_c, F_m, F_c, c.dt);
        const float hi = flux_lo(v_p, F_c, F_p, c.dt);
        sA[r][s] = flux_hi(v_c, F_m, F_c, c.dt) - lo;
        const float dv = c.dxdy - c.dtdx * (v_p - v_c);
        sTd[r][s] = (j >= 1 && j <= ny) ? fct_ftd(F_c, lo, hi, dv, c) : 0.0f;
    }
    __syncthreads();
    // phase B: cells s in [3, TJ+5): limiter ratios                                  (loop 2b)
    for (int k = tid; k < TR * (TJ + 2); k += 256) {
        const int r = k / (TJ + 2), s = 3 + (k - r * (TJ + 2));
        const int j = jb + s;
        float rp = 0.0f, rm = 0.0f;
        if (j >= 1 && j <= ny) fct_ratios(sTd[r][s - 1], sTd[r][s], sTd[r][s + 1], sA[r][s], sA[r][s + 1], c, rp, rm);
        sRp[r][s] = rp; sRm[r][s] = rm;
    }
    __syncthreads();
    // phase C: output cells s in [4, TJ+4)                                           (loops 3 and 4)
    for (int k = tid; k < TR * TJ; k += 256) {
        const int r = k / TJ, s = 4 + (k - r * TJ);
        const int i = i0 + r, j = jb + s;
        if (i > r1 || j > ny) continue;
        const int gi = g.gi0 + i;
        if (gi < 1 || gi > g.nx) {   // ghost rows are not touched by the sweep
            Fout[(size_t)i * P + j] = sF[r][s];
            continue;
        }
        const float a_c = sA[r][s], a_p = sA[r][s + 1];
        // cy on face j is 0 for j = 1 (never written), 2dvof.py:435-438 writes faces 2..ny+1
        const float c_c = (j >= 2) ? fct_cface(a_c, sRp[r][s - 1], sRm[r][s - 1], sRp[r][s], sRm[r][s]) : 0.0f;
        const float c_p = fct_cface(a_p, sRp[r][s], sRm[r][s], sRp[r][s + 1], sRm[r][s + 1]);
        const float dv = c.dxdy - c.dtdx * (sV[r][s + 1] - sV[r][s]);
        Fout[(size_t)i * P + j] = fct_update<POST>(sTd[r][s], a_c, c_c, a_p, c_p, dv, c);
    }
    // ghost columns of these rows pass through unchanged
    if (blockIdx.x == 0 && tid < TR && i0 + tid <= r1) {
        const size_t o = (size_t)(i0 + tid) * P;
        Fout[o] = Fin[o];
        Fout[o + ny + 1] = Fin[o + ny + 1];
    }
}

// ======================================================================================
// post_process_f  (2dvof.py:452-455) -- all cells.  Sequence mode only (fused into the second
// FCT sweep otherwise; on ghost cells the following set_BC overwrites it).
// ======================================================================================
__global__ void __launch_bounds__(kBlockJ)
k_post_process_f(Grid g, float* __restrict__ F, int r0, int r1, int rows_per_block) {
    const int j = blockIdx.x * kBlockJ + threadIdx.x;
    if (j > g.ny + 1) return;
    const int ia = r0 + blockIdx.y * rows_per_block;
    const int ib = min(r1, ia + rows_per_block - 1);
    for (int i = ia; i <= ib; ++i) {
        const size_t o = (size_t)i * g.pitch + j;
        F[o] = var3(F[o], 0.0f, 1.0f);
    }
}

// ======================================================================================
// set_init_F  (2dvof.py:137-159) + find_area (102-134).  xs/ys are the fp32 node arrays.
// ======================================================================================
struct InitConsts {
    float x2, y2;            // ic 1: Lx/3, Ly/2
    float r, cx, cy;         // ic 2/3 circle
    float ycut;              // ic 3: Ly*0.37
    float hdx, hdy;          // dx/2, dy/2 (double-folded)
    float sqrt2dx;           // sqrt(2)*dx (double-folded)
};

__device__ __forceinline__ float find_area(int gi, int j, const Consts& c, const InitConsts& ic) {
    const float xc = (float)(gi - 1) * c.dx + ic.hdx, yc = (float)(j - 1) * c.dy + ic.hdy;
    const float xl = xc - ic.hdx, xr = xc + ic.hdx, yd = yc - ic.hdy, yu = yc + ic.hdy;
    auto dist = [&](float xx, float yy) {
        const float ddx = xx - ic.cx, ddy = yy - ic.cy;
        return sqrtf(ddx * ddx + ddy * ddy);
    };
    const float d_ct = dist(xc, yc), d_lu = dist(xl, yu), d_ld = dist(xl, yd), d_ru = dist(xr, yu), d_rd = dist(xr, yd);
    const float r = ic.r;
    if (d_lu > r && d_ld > r && d_ru > r && d_rd > r) return 1.0f;
    if (d_lu < r && d_ld < r && d_ru < r && d_rd < r) return 0.0f;
    const float a = 0.5f + (0.5f * (d_ct - r)) / ic.sqrt2dx;
    return var3(a, 0.0f, 1.0f);
}

__global__ void __launch_bounds__(kBlockJ)
k_set_init_F(Grid g, Consts c, InitConsts ic, int which, const float* __restrict__ xs,
             const float* __restrict__ ys, float* __restrict__ F, int r0, int r1) {
    const int j = blockIdx.x * kBlockJ + threadIdx.x;
    if (j > g.ny + 1) return;
    const float yj = ys[j];
    for (int i = r0 + blockIdx.y; i <= r1; i += gridDim.y) {
        const int gi = g.gi0 + i;
        if (gi < 0 || gi > g.nx + 1) continue;
        const size_t o = (size_t)i * g.pitch + j;
        if (which == 1) {
            const float xi = xs[gi];
            if (xi >= 0.0f && xi <= ic.x2 && yj >= 0.0f && yj <= ic.y2) F[o] = 1.0f;
        } else if (which == 2) {
            F[o] = find_area(gi, j, c, ic);
        } else {
            float f = 1.0f - find_area(gi, j, c, ic);
            if (yj < ic.ycut) f = 1.0f;
            F[o] = f;
        }
    }
}

// ======================================================================================
// diagnostics (new): warp-shuffle reductions, one atomic per block.
// out[0] = sum F (double), out[1] = bit pattern of max cfl (float >= 0), out[2] = residual L-inf
// ======================================================================================
struct Diag {
    double mass;
    unsigned int max_cfl_bits;
    unsigned int resid_bits;
    unsigned long long courant_count;
    unsigned int queue;          // work-queue head of the persistent kernels
};

__global__ void __launch_bounds__(256)
k_diag(Grid g, Consts c, const float* __restrict__ F, const float* __restrict__ u,
       const float* __restrict__ v, const float* __restrict__ p, const float* __restrict__ rhs,
       Diag* __restrict__ d, int r0, int r1, int do_resid) {
    __shared__ double s_mass[8];
    __shared__ float s_cfl[8], s_res[8];
    double mass = 0.0;
    float cfl = 0.0f, res = 0.0f;
    const int P = g.pitch;
    for (int i = r0 + blockIdx.y; i <= r1; i += gridDim.y) {
        const int gi = g.gi0 + i;
        if (gi < 1 || gi > g.nx) continue;
        for (int j = 1 + blockIdx.x * 256 + threadIdx.x; j <= g.ny; j += gridDim.x * 256) {
            const size_t o = (size_t)i * P + j;
            mass += (double)F[o];
            cfl = fmaxf(cfl, fmaxf(fabsf(u[o]) * c.dt * c.dxi, fabsf(v[o]) * c.dt * c.dyi));
            if (do_resid) {
                const float ae = (gi != g.nx) ? c.dxi2 : 0.0f, aw = (gi != 1) ? c.dxi2 : 0.0f;
                const float an = (j != g.ny) ? c.dyi2 : 0.0f, as = (j != 1) ? c.dyi2 : 0.0f;
                const float pc = p[o];
                // residual of  sum_k a_k (p_k - p_c) = rhs   (the system the Jacobi sweep relaxes)
                const float lap = ae * (p[o + P] - pc) + aw * (p[o - P] - pc) + an * (p[o + 1] - pc) + as * (p[o - 1] - pc);
                res = fmaxf(res, fabsf(rhs[o] - lap));
            }
        }
    }
    mass = warp_sum(mass); cfl = warp_max(cfl); res = warp_max(res);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_mass[w] = mass; s_cfl[w] = cfl; s_res[w] = res; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { mass += s_mass[k]; cfl = fmaxf(cfl, s_cfl[k]); res = fmaxf(res, s_res[k]); }
        atomicAdd(&d->mass, mass);
        atomicMax(&d->max_cfl_bits, __float_as_uint(cfl));
        atomicMax(&d->resid_bits, __float_as_uint(res));
    }
}

}  // namespace vof
