// sm_100a kernels of the 2-D VOF step, part 1: the simple streaming kernels (properties, boundary fill, Poisson
// rhs, the one-sweep Jacobi used for small grids / sequence mode, post-process, initial condition, diagnostics).
// The heavy kernels live in vof2d_kappa.cuh, vof2d_momentum.cuh, vof2d_jacobi_tb.cuh and vof2d_fct.cuh.
// Threads map to the contiguous axis j (coalesced 128-byte warp accesses, interior column 1 on a 128-byte
// boundary), blocks tile i, values reused along i roll through registers.
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
         const float* __restrict__ vs, int r0, int r1, int rows_per_block, float b0, float rcp0, float b1, float rcp1, int bare) {
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
            // bare: the context proved q = RN(t r), q' = fma(fma(-q, b, t), r, q) equal to t / b for every fp32 t and both
            // interior-row diagonals (vof2d_create); rows next to an i-wall keep the IEEE division
            if (bare && gi != 1 && gi != g.nx) {
                const bool wj = j == 1 || j == g.ny;
                const float b = wj ? b1 : b0, r = wj ? rcp1 : rcp0;
                const float q = t * r;
                out = __fmaf_rn(__fmaf_rn(-q, b, t), r, q);
            } else out = div_nz(t, ap);         // p = 0 ahead of the pressure front: skip nvcc's slow path for 0 / ap
        }
        pn[o] = out;
        p_m = p_c; p_c = p_p; us_c = us_p;
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
// display kernels (2dvof.py:458-492): x2 nearest-neighbour upsample of F / u / v / |V| into rgb_buf (2nx, 2ny),
// velocities scaled by L / 0.2; cell-centred velocity V.  Monitoring only -- never on the step's path.
// ======================================================================================
template <int VIEW>      // 0 F, 1 u / (Lx/0.2), 2 v / (Ly/0.2), 3 sqrt(u^2 + v^2) / (Ly/0.2)
__global__ void __launch_bounds__(256)
k_display(Grid g, const float* __restrict__ F, const float* __restrict__ u, const float* __restrict__ v, float vmax,
          float* __restrict__ rgb) {
    const int J = blockIdx.x * 256 + threadIdx.x, I = blockIdx.y;
    if (J >= 2 * g.ny) return;
    const size_t o = (size_t)(I >> 1) * g.pitch + (J >> 1);                 // rgb_buf[I] = field[I // r], r = 2
    float x;
    if (VIEW == 0) x = F[o];
    else if (VIEW == 1) x = u[o] / vmax;
    else if (VIEW == 2) x = v[o] / vmax;
    else x = sqrtf(u[o] * u[o] + v[o] * v[o]) / vmax;
    rgb[(size_t)I * (2 * g.ny) + J] = x;
}
// V[i, j] = ((u[i,j] + u[i+1,j]) / 2, (v[i,j] + v[i,j+1]) / 2) for i in [1, nx+1], j in [1, ny] (2dvof.py:489-492).  The
// reference's range reads u[nx+2, j], one row past the field (undefined there); that row of V is left at zero here.
__global__ void __launch_bounds__(256)
k_interp_velocity(Grid g, const float* __restrict__ u, const float* __restrict__ v, float2* __restrict__ V) {
    const int j = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
    if (j > g.ny + 1) return;
    float2 r = make_float2(0.0f, 0.0f);
    if (i >= 1 && i <= g.nx && j >= 1 && j <= g.ny) {
        const size_t o = (size_t)i * g.pitch + j;
        r.x = (u[o] + u[o + g.pitch]) / 2.0f;
        r.y = (v[o] + v[o + 1]) / 2.0f;
    }
    V[(size_t)i * (g.ny + 2) + j] = r;
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
    unsigned int queue;          // work-queue head of the persistent Jacobi kernel (zeroed before each launch)
    unsigned int wq[2];          // WorkQueue counters of the streaming kernels (self re-arming, vof2d_stream.cuh)
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
