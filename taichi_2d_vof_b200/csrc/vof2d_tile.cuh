// Whole-step tile kernel for small grids: ONE launch per time step (2dvof.py:513-528).
//
// At 200 x 200 -- the reference's own configuration -- the streaming kernels are launch bound: a step is ~19 dependent
// launches of a few microseconds each (100 us per step, 3.3x a 16-thread CPU).  This kernel applies the communication-
// avoiding idea of the multi-GPU slabs at thread-block level: a block owns TI x TJ cells and holds
// them together with a halo of H = n_jacobi + 5 cells -- the dependency radius of one step -- in its 227 KB of shared
// memory; it then runs the entire step on the tile (normals, curvature, predictor, set_BC, rhs, all Jacobi sweeps,
// projection, set_BC, both FCT sweeps, post-process, set_BC) with block barriers only, and writes back the cells it owns.
// Halo cells are computed redundantly ((TI + 2H)(TJ + 2H) / (TI TJ) times the work, irrelevant at this size); garbage
// from the tile edge travels inwards one stencil radius per phase and by construction stops short of the owned cells.
// Every phase computes each cell whose stencil lies inside the tile with the reference's expression, operand order and
// loop bounds (the per-cell forms are those of oracle/vof2d_oracle.c, which is pinned to the reference run), so the
// result is bit-identical to the streaming path -- tests/test_tile_gpu.py.
//
// Blocks read the old state (u, v, p, F) from one set of buffers and write the new state into another (F / p ping-pong
// as everywhere; u, v into the rho / nu buffers, which the fused step does not use -- the host swaps the pointers).
// u*, v* and kappa are updated in place: a block stores only the entries it owns and loads only entries that no block
// ever stores (the rows / columns outside advect_upwind's loop ranges), so there is no inter-block hazard and no
// grid-wide barrier.
#pragma once
#include "vof2d_jacobi_tb.cuh"
#include "vof_common.cuh"

namespace vof {

constexpr int kTileW = 64;          // tile width in cells (a power of two: cell c = (c >> 6, c & 63)); owned columns = kTileW - 2 H
constexpr int kTileThreads = 1024;

struct TileArgs {
    const float* u; const float* v; const float* p; const float* F;     // old state
    float* un; float* vn; float* pn; float* Fn;                          // new state (other buffers)
    float* us; float* vs; float* kappa;                                  // in place (see above)
    unsigned long long* courant_count;
    int ti, tj;                 // owned cells per block along i and j (of the padded index range 0 .. n + 1)
    int H, th;                  // halo; tile height th = ti + 2 H (tile width is kTileW = tj + 2 H)
    int n_jacobi, istep;
    ConstDiv d_dx, d_dy, d_dxdy, d_ap0, d_ap1;     // exact reciprocal division by dx, dy, dx dy and the Poisson diagonal of an
    int fast, bare;                                // interior row (away from / next to a j-wall); proven at vof2d_create; bare: the
                                                   // Poisson diagonals need no sub-normal fix-up either (JacTB::bare_div_ok)
};

// cell classes, computed once per step (one byte per tile cell) instead of index arithmetic in each of the ~30 phases
enum : unsigned { T_IN = 1u,      // interior cell (1 <= i <= nx, 1 <= j <= ny) whose radius-1 stencil lies inside the tile
                  T_I2 = 2u,      // i >= 2 (the u loops)
                  T_J2 = 4u,      // j >= 2 (the v loops)
                  T_OWN = 8u,     // owned by this block
                  T_WI = 16u,     // next to an i-wall (i == 1 or i == nx)
                  T_WJ = 32u };   // next to a j-wall

__device__ __forceinline__ float tdiv(float t, const ConstDiv& d, int fast) { return fast ? div_by_const(t, d) : div_nz(t, d.b); }
__device__ __forceinline__ float tdiv_ap(float t, const ConstDiv& d, int fast, int bare) {
    return bare ? div_by_const_core(t, d.b, d.r) : tdiv(t, d, fast);
}

// The ten tile arrays.  Scratch use: mx / my live in pB / FB until the curvature is done; Ftd / rp / rm live in
// rhs / kap / pB during the FCT sweeps (kappa has been written out, the pressure iteration is over).
struct TileMem {
    float *u, *v, *p, *pB, *F, *FB, *us, *vs, *kap, *rhs;
};


template <class Fn>
__device__ __forceinline__ void tile_for(int th, Fn fn) {       // the tile's cells dealt out evenly; cell c = (c >> 6, c & 63), its neighbours c +- kTileW, c +- 1
    const int ncell = th * kTileW;
    for (int c = threadIdx.x; c < ncell; c += kTileThreads) fn(c);
}

template <class Fn>
__device__ __forceinline__ void tile_rows(int r0, int r1, Fn fn) {   // the cells of tile rows r0 .. r1: 16 rows per pass, a thread keeps its column
    for (int li = r0 + (int)(threadIdx.x >> 6); li <= r1; li += kTileThreads / kTileW) fn(li * kTileW + (int)(threadIdx.x & 63));
}

__global__ void __launch_bounds__(kTileThreads, 1)
k_step_tile(Grid g, Consts k, JacTB jc, TileArgs a) {
    extern __shared__ __align__(16) float tile_smem[];
    const int th = a.th, H = a.H, nx = g.nx, ny = g.ny, P = g.pitch;
    const int cells = th * kTileW;
    TileMem m;
    m.u = tile_smem; m.v = m.u + cells; m.p = m.v + cells; m.pB = m.p + cells; m.F = m.pB + cells; m.FB = m.F + cells;
    m.us = m.FB + cells; m.vs = m.us + cells; m.kap = m.vs + cells; m.rhs = m.kap + cells;
    unsigned char* cls = reinterpret_cast<unsigned char*>(m.rhs + cells);
    const int oi0 = blockIdx.y * a.ti, oj0 = blockIdx.x * a.tj;                  // first owned cell
    int oi1 = min(oi0 + a.ti - 1, nx + 1), oj1 = min(oj0 + a.tj - 1, ny + 1);
    // A block never owns the far ghost row / column alone: F, p, u, v there are set_BC copies of row nx / column ny, whose own
    // dependency radius H would then reach one cell beyond the tile.  The block that owns row nx (column ny) owns the
    // ghost line next to it as well (the host launches no block for it, tile_geometry); nothing exists beyond it, so the
    // tile still covers everything its owned cells depend on.
    if (oi1 == nx) oi1 = nx + 1;
    if (oj1 == ny) oj1 = ny + 1;
    const int gi0 = oi0 - H, gj0 = oj0 - H;                                      // global index of tile cell (0, 0)
    // a cell takes part in a phase when its stencil (radius r) lies inside the tile array
    auto inside = [&](int li, int lj, int r) { return li >= r && li < th - r && lj >= r && lj < kTileW - r; };
    // Late phases run on the rows that can still reach an owned cell: `within(m, fn)` = the owned rows and m rows either side.
    // Margins, backwards from the write-back (0): final set_BC column loop 0 / row loop 1, post-process 1, second FCT sweep
    // 1 / 2 / 3 (loops 3+4, 2, 1: each reads its predecessor one cell further out; across the sweep direction the limiter
    // ratios of the neighbours), first sweep 4 / 5 / 5, so F, u, v are needed 5 rows out after set_BC (525): its column
    // loop 5, row loop 6, projection 6, Jacobi sweep s (of n) 5 + (n - 1 - s) (see there).  The margins are those of the
    // sweep order that needs more (second sweep along i); rows further out hold garbage either way -- what the halo is for.
    const int own_r0 = H, own_r1 = H + (oi1 - oi0), own_c0 = H, own_c1 = H + (oj1 - oj0);
    auto within = [&](int mi, auto fn) { tile_rows(max(0, own_r0 - mi), min(th - 1, own_r1 + mi), fn); };
    // The phases after the sweeps are few cells with long expressions: the same margins along j as well, and the cells of the
    // rectangle dealt out compactly (cell = idx / w, idx % w through one multiplication: exact for idx < 2^13, w <= 64) --
    // at 200^2 the 19 x 44 cells of the first FCT sweep are one pass of the block instead of two.
    auto within2 = [&](int mi, int mj, auto fn) {
        const int r0 = max(0, own_r0 - mi), r1 = min(th - 1, own_r1 + mi), c0 = max(0, own_c0 - mj), c1 = min(kTileW - 1, own_c1 + mj);
        const int w = c1 - c0 + 1, n = (r1 - r0 + 1) * w;
        const float rw = 1.0f / (float)w;
        for (int idx = threadIdx.x; idx < n; idx += kTileThreads) {
            const int q = (int)(((float)idx + 0.5f) * rw);
            fn((r0 + q) * kTileW + c0 + (idx - q * w));
        }
    };

    // ---- load the old state (cells outside the field do not exist: never read by a cell inside a loop range)
    tile_for(th, [&](int c) {
        const int li = c >> 6, lj = c & 63;
        const int gi = gi0 + li, gj = gj0 + lj;
        const bool ex = gi >= 0 && gi <= nx + 1 && gj >= 0 && gj <= ny + 1;
        const size_t o = ex ? (size_t)gi * P + gj : 0;
        m.u[c] = ex ? a.u[o] : 0.0f; m.v[c] = ex ? a.v[o] : 0.0f; m.p[c] = ex ? a.p[o] : 0.0f; m.F[c] = ex ? a.F[o] : 0.0f;
        // u*, v* are updated in place: only their never-written entries (outside the loop ranges of 2dvof.py:208, 221 --
        // stable, nobody stores to them) are loaded; every entry inside the ranges is recomputed here before it is used
        const bool us_loop = gi >= 2 && gi <= nx && gj >= 1 && gj <= ny, vs_loop = gi >= 1 && gi <= nx && gj >= 2 && gj <= ny;
        m.us[c] = (ex && !us_loop) ? a.us[o] : 0.0f; m.vs[c] = (ex && !vs_loop) ? a.vs[o] : 0.0f;
        m.pB[c] = 0.0f; m.FB[c] = 0.0f;          // mx, my: never-written entries are 0 (2dvof.py:80-81)
        m.kap[c] = 0.0f; m.rhs[c] = 0.0f;
        unsigned q = 0;
        if (gi >= 1 && gi <= nx && gj >= 1 && gj <= ny && inside(li, lj, 1)) q |= T_IN;
        if (gi >= 2) q |= T_I2;
        if (gj >= 2) q |= T_J2;
        if (gi >= oi0 && gi <= oi1 && gj >= oj0 && gj <= oj1) q |= T_OWN;
        if (gi == 1 || gi == nx) q |= T_WI;
        if (gj == 1 || gj == ny) q |= T_WJ;
        cls[c] = (unsigned char)q;
    });
    __syncthreads();

    // ---- get_normal_young, loop 1 (2dvof.py:285-305): unit normals of the interior cells -> mx (pB), my (FB)
    tile_for(th, [&](int c) {
        if (!(cls[c] & T_IN)) return;
        const float* F = m.F;
        const float f_pp = F[c + kTileW + 1], f_pc = F[c + kTileW], f_pm = F[c + kTileW - 1];
        const float f_cp = F[c + 1], f_cc = F[c], f_cm = F[c - 1];
        const float f_mp = F[c - kTileW + 1], f_mc = F[c - kTileW], f_mm = F[c - kTileW - 1];
        const float kx = k.m1_2dx, ky = k.m1_2dy;
        const float mx1 = kx * (((f_pp + f_pc) - f_cp) - f_cc), my1 = ky * (((f_pp - f_pc) + f_cp) - f_cc);
        const float mx2 = kx * (((f_pc + f_pm) - f_cc) - f_cm), my2 = ky * (((f_pc - f_pm) + f_cc) - f_cm);
        const float mx3 = kx * (((f_cc + f_cm) - f_mc) - f_mm), my3 = ky * (((f_cc - f_cm) + f_mc) - f_mm);
        const float mx4 = kx * (((f_cp + f_cc) - f_mp) - f_mc), my4 = ky * (((f_cp - f_cc) + f_mp) - f_mc);
        const float sx = (((mx1 + mx2) + mx3) + mx4) / 4.0f, sy = (((my1 + my2) + my3) + my4) / 4.0f;
        float ox = sx, oy = sy;
        if (!(fabsf(sx) < 1e-10f && fabsf(sy) < 1e-10f)) {
            const float mag = sqrtf(sx * sx + sy * sy);
            ox = sx / mag; oy = sy / mag;
        }
        m.pB[c] = ox; m.FB[c] = oy;
    });
    __syncthreads();
    // ---- loop 2 (306-309): kappa
    tile_for(th, [&](int c) {
        const unsigned q = cls[c];
        if (!(q & T_IN)) return;
        const float kp = -(k.i_dx_2 * (m.pB[c + kTileW] - m.pB[c - kTileW]) +
                           k.i_dy_2 * (m.FB[c + 1] - m.FB[c - 1]));
        m.kap[c] = kp;
        if (q & T_OWN) a.kappa[(size_t)(gi0 + (c >> 6)) * P + gj0 + (c & 63)] = kp;
    });
    __syncthreads();

    // ---- advect_upwind (206-233); rho, nu from F (cal_nu_rho 198-203 inlined: F does not change before the FCT sweeps)
    tile_for(th, [&](int c) {
        const unsigned q = cls[c];
        if (!(q & T_IN)) return;
        const float* u = m.u; const float* v = m.v;
        const float F_c = m.F[c], nu_c = nu_of(F_c, k), k_c = m.kap[c];
        if (q & T_I2) {
            const float uc = u[c], u_m = u[c - kTileW], u_p = u[c + kTileW];
            const float u_jm = u[c - 1], u_jp = u[c + 1];
            const float v_here = 0.25f * (((v[c - kTileW] + v[c - kTileW + 1]) + v[c]) + v[c + 1]);
            const float dudx = uc > 0.0f ? (uc - u_m) * k.dxi : (u_p - uc) * k.dxi;
            const float dudy = v_here > 0.0f ? (uc - u_jm) * k.dyi : (u_jp - uc) * k.dyi;
            const float F_m = m.F[c - kTileW];
            float acc = (nu_c * ((u_m - 2.0f * uc) + u_p)) * k.dxi2;
            acc = acc + (nu_c * ((u_jm - 2.0f * uc) + u_jp)) * k.dyi2;
            acc = acc - uc * dudx;
            acc = acc - v_here * dudy;
            acc = acc + k.gx;
            const float dF = F_c - F_m;
            if (dF != 0.0f) {                     // otherwise the CSF term is +-0 and acc + (+-0) = acc (acc is never -0 + ... : gx, gy
                const float kappa_ave = (k_c + m.kap[c - kTileW]) / 2.0f;         //  enter before it, as in k_advect5)
                const float fx_kappa = tdiv((k.neg_sigma * dF) * kappa_ave, a.d_dx, a.fast);
                acc = acc + (fx_kappa * 2.0f) / (rho_of(F_c, k) + rho_of(F_m, k));
            }
            m.us[c] = uc + k.dt * acc;
        }
        if (q & T_J2) {
            const float vc = v[c], v_m = v[c - kTileW], v_p = v[c + kTileW];
            const float v_jm = v[c - 1], v_jp = v[c + 1];
            const float u_here = 0.25f * (((u[c - 1] + u[c]) + u[c + kTileW - 1]) + u[c + kTileW]);
            const float dvdx = u_here > 0.0f ? (vc - v_m) * k.dxi : (v_p - vc) * k.dxi;
            const float dvdy = vc > 0.0f ? (vc - v_jm) * k.dyi : (v_jp - vc) * k.dyi;
            const float F_jm = m.F[c - 1];
            float acc = (nu_c * ((v_m - 2.0f * vc) + v_p)) * k.dxi2;
            acc = acc + (nu_c * ((v_jm - 2.0f * vc) + v_jp)) * k.dyi2;
            acc = acc - u_here * dvdx;
            acc = acc - vc * dvdy;
            acc = acc + k.gy;
            const float dF = F_c - F_jm;
            if (dF != 0.0f) {
                const float kappa_ave = (k_c + m.kap[c - 1]) / 2.0f;
                const float fy_kappa = tdiv((k.neg_sigma * dF) * kappa_ave, a.d_dy, a.fast);
                acc = acc + (fy_kappa * 2.0f) / (rho_of(F_c, k) + rho_of(F_jm, k));
            }
            m.vs[c] = vc + k.dt * acc;
        }
    });
    __syncthreads();

    // set_BC (162-189) on the tile: the row loop, a barrier, the column loop; u, v, F, p (rho is taken from F).
    auto set_bc = [&](int mi) {                  // mi: rows / columns needed afterwards (the row loop runs one row further out)
        within2(mi + 1, mi, [&](int c) {
            const int li = c >> 6, lj = c & 63;
            const int gi = gi0 + li, gj = gj0 + lj;
            if (gi < 0 || gi > nx + 1) return;
            if (gj == 0 && lj + 1 < kTileW) {
                m.u[c] = m.u[c + 1]; m.F[c] = m.F[c + 1]; m.p[c] = m.p[c + 1];
            }
            if (gj == 1) m.v[c] = 0.0f;
            if (gj == ny + 1 && lj >= 1) {
                m.u[c] = m.u[c - 1]; m.F[c] = m.F[c - 1]; m.p[c] = m.p[c - 1];
                m.v[c] = 0.0f;
            }
        });
        __syncthreads();
        within2(mi, mi, [&](int c) {
            const int li = c >> 6, lj = c & 63;
            const int gi = gi0 + li, gj = gj0 + lj;
            if (gj < 0 || gj > ny + 1) return;
            if (gi == 1) m.u[c] = 0.0f;
            if (gi == 0 && li + 1 < th) {
                m.v[c] = m.v[c + kTileW]; m.F[c] = m.F[c + kTileW]; m.p[c] = m.p[c + kTileW];
            }
            if (gi == nx + 1 && li >= 1) {
                m.u[c] = 0.0f;
                m.v[c] = m.v[c - kTileW]; m.F[c] = m.F[c - kTileW]; m.p[c] = m.p[c - kTileW];
            }
        });
        __syncthreads();
    };
    set_bc(th);                                                                  // 518: the whole tile

    // ---- solve_p_jacobi x n (236-266, 521-522): rhs once (its value is the same in every sweep), then the sweeps
    tile_for(th, [&](int c) {
        m.pB[c] = m.p[c];           // ghost cells pass through the sweeps
        if (!(cls[c] & T_IN)) return;
        m.rhs[c] = (rho_of(m.F[c], k) / k.dt) *
                              ((m.us[c + kTileW] - m.us[c]) * k.dxi + (m.vs[c + 1] - m.vs[c]) * k.dyi);
    });
    __syncthreads();
    float* pc = m.p; float* pn = m.pB;
    for (int s = 0; s < a.n_jacobi; ++s) {
        auto sweep = [&](int c) {
            const unsigned q = cls[c];
            if (!(q & T_IN)) return;
            const float b = m.rhs[c];
            float out;
            if (!(q & (T_WI | T_WJ))) {               // all four coefficients are dxi^2 / dyi^2
                float t = b - k.dxi2 * pc[c + kTileW];
                t = t - k.dxi2 * pc[c - kTileW];
                t = t - k.dyi2 * pc[c + 1];
                t = t - k.dyi2 * pc[c - 1];
                out = tdiv_ap(t, a.d_ap0, a.fast, a.bare);
            } else {
                const int gi = gi0 + (c >> 6), gj = gj0 + (c & 63);
                const float ae = gi != nx ? k.dxi2 : 0.0f, aw = gi != 1 ? k.dxi2 : 0.0f;
                const float an = gj != ny ? k.dyi2 : 0.0f, as = gj != 1 ? k.dyi2 : 0.0f;
                const bool wi = (q & T_WI) != 0, wj = (q & T_WJ) != 0;       // diagonal by wall class (selects: no indexed parameter)
                const float ap = wi ? (wj ? jc.ap[1][1] : jc.ap[1][0]) : (wj ? jc.ap[0][1] : jc.ap[0][0]);
                float t = b - ae * pc[c + kTileW];
                t = t - aw * pc[c - kTileW];
                t = t - an * pc[c + 1];
                t = t - as * pc[c - 1];
                out = (!wi) ? tdiv_ap(t, a.d_ap1, a.fast, a.bare) : div_nz(t, ap);
            }
            pn[c] = out;
        };
        // The FCT sweeps read the new u, v at most 3 cells beyond the owned ones (u[i-2 .. i+3], 2dvof.py:323-369; the
        // cross-direction limiter ratios, multiplied by 0.0, one further: they only have to be finite), set_BC copies
        // inwards; the regions of the later phases are the conservative chain (5 / 6).  The last sweep's p is therefore
        // needed at most 5 cells out, each earlier sweep's one more; what the projection reads further out is one sweep
        // behind (finite) and feeds no needed cell.  The last sweeps are few enough cells for the compact rectangle
        // (one pass of the block instead of two at 200^2).
        const int mg = 5 + (a.n_jacobi - 1 - s);
        if (mg <= 6) within2(mg, mg, sweep); else within(mg, sweep);
        __syncthreads();
        float* sw = pc; pc = pn; pn = sw;
    }
    // pc: the new pressure; the other pressure array is free from here on

    // ---- update_uv (269-280), Courant offenders counted over the owned cells
    unsigned flags = 0;
    within2(6, 6, [&](int c) {
        const unsigned q = cls[c];
        if (!(q & T_IN)) return;
        const bool own = (q & T_OWN) != 0;
        const float rho_c = rho_of(m.F[c], k), p_c = pc[c];
        if (q & T_I2) {
            const float r = (rho_c + rho_of(m.F[c - kTileW], k)) * 0.5f;
            const float un = m.us[c] - ((k.dt / r) * (p_c - pc[c - kTileW])) * k.dxi;
            m.u[c] = un;
            flags += own && (un * k.dt > k.cflx);
        }
        if (q & T_J2) {
            const float r = (rho_c + rho_of(m.F[c - 1], k)) * 0.5f;
            const float vn = m.vs[c] - ((k.dt / r) * (p_c - pc[c - 1])) * k.dyi;
            m.v[c] = vn;
            flags += own && (vn * k.dt > k.cfly);
        }
    });
    if (flags) atomicAdd(a.courant_count, (unsigned long long)flags);
    __syncthreads();
    if (pc != m.p) {                          // keep the pressure in m.p: set_bc and the write-back address it there
        within2(7, 7, [&](int c) { m.p[c] = pc[c]; });
        __syncthreads();
    }
    set_bc(5);                                                                   // 525

    // ---- solve_VOF_rudman (312-318): two FCT sweeps (321-448), order by the step's parity.
    // Scratch: Ftd -> rhs, rp -> kap, rm -> pB.  Never-written entries of Ftd, rp, rm, cx, cy are 0.
    float* Ftd = m.rhs; float* rp = m.kap; float* rm = m.pB;
    float* Fc = m.F; float* Fo = m.FB;
    for (int half = 0; half < 2; ++half) {
        const bool along_x = (a.istep % 2 == 0) ? (half == 1) : (half == 0);
        const int sd = along_x ? kTileW : 1, so = along_x ? 1 : kTileW;      // cell strides along / across the sweep
        const float* vel = along_x ? m.u : m.v;
        const float dtd = along_x ? k.dtdy : k.dtdx;
        const int m3 = half == 0 ? 4 : 1;  // rows / columns this sweep's result is needed on beyond the owned ones
        const int m1 = half == 0 ? 5 : 3;
        within2(m1, m1, [&](int c) {       // loop 1: the transported-diffused value
            Ftd[c] = 0.0f; rp[c] = 0.0f; rm[c] = 0.0f;
            if (!(cls[c] & T_IN)) return;
            const float vc = vel[c], vp = vel[c + sd];
            const float f_m = Fc[c - sd], f_c = Fc[c], f_p = Fc[c + sd];
            const float dv = k.dxdy - dtd * (vp - vc);
            const float lo_L = vc >= 0.0f ? (vc * k.dt) * f_m : (vc * k.dt) * f_c;
            const float hi_L = vp >= 0.0f ? (vp * k.dt) * f_c : (vp * k.dt) * f_p;
            const float sum = along_x ? ((lo_L - hi_L) + 0.0f) - 0.0f : ((0.0f - 0.0f) + lo_L) - hi_L;
            float t = div_nz((((f_c + tdiv(sum * k.dy, a.d_dxdy, a.fast)) * k.dx) * k.dy), dv);
            if (t > 1.0f || t < 0.0f) t = var3(0.0f, 1.0f, t);
            Ftd[c] = t;
        });
        __syncthreads();
        within2(m3 + 1, m3 + 1, [&](int c) {   // loop 2: limiter ratios
            if (!(cls[c] & T_IN)) return;
            const float t_c = Ftd[c], t_m = Ftd[c - sd], t_p = Ftd[c + sd];
            const float fmax = fmaxf(fmaxf(t_c, t_m), t_p), fmin = fminf(fminf(t_c, t_m), t_p);
            const float vc = vel[c], vp = vel[c + sd];
            const float f_m = Fc[c - sd], f_c = Fc[c], f_p = Fc[c + sd];
            const float lo_L = vc >= 0.0f ? (vc * k.dt) * f_m : (vc * k.dt) * f_c;
            const float hi_L = vp >= 0.0f ? (vp * k.dt) * f_c : (vp * k.dt) * f_p;
            const float lo_H = vc <= 0.0f ? (vc * k.dt) * f_m : (vc * k.dt) * f_c;
            const float hi_H = vp <= 0.0f ? (vp * k.dt) * f_c : (vp * k.dt) * f_p;
            const float axc = along_x ? lo_H - lo_L : 0.0f, axp = along_x ? hi_H - hi_L : 0.0f;
            const float ayc = along_x ? 0.0f : lo_H - lo_L, ayp = along_x ? 0.0f : hi_H - hi_L;
            const float pp = ((fmaxf(0.0f, axc) - fminf(0.0f, axp)) + fmaxf(0.0f, ayc)) - fminf(0.0f, ayp);
            const float qp = (fmax - t_c) * k.dx;
            rp[c] = pp > 0.0f ? fminf(1.0f, div_nz(qp, pp)) : 0.0f;
            const float pm = ((fmaxf(0.0f, axp) - fminf(0.0f, axc)) + fmaxf(0.0f, ayp)) - fminf(0.0f, ayc);
            const float qm = (t_c - fmin) * k.dx;
            rm[c] = pm > 0.0f ? fminf(1.0f, div_nz(qm, pm)) : 0.0f;
        });
        __syncthreads();
        within2(m3, m3, [&](int c) {       // loops 3 + 4: face limiters and the corrective update
            Fo[c] = Fc[c];                                // ghost cells keep their value through a sweep
            const unsigned q = cls[c];
            if (!(q & T_IN)) return;
            const float vc = vel[c], vp = vel[c + sd];
            const float f_m = Fc[c - sd], f_c = Fc[c], f_p = Fc[c + sd];
            const float lo_L = vc >= 0.0f ? (vc * k.dt) * f_m : (vc * k.dt) * f_c;
            const float hi_L = vp >= 0.0f ? (vp * k.dt) * f_c : (vp * k.dt) * f_p;
            const float lo_H = vc <= 0.0f ? (vc * k.dt) * f_m : (vc * k.dt) * f_c;
            const float hi_H = vp <= 0.0f ? (vp * k.dt) * f_c : (vp * k.dt) * f_p;
            const float a_lo = lo_H - lo_L, a_hi = hi_H - hi_L;          // the swept direction's faces of this cell
            const float rp_c = rp[c], rm_c = rm[c];
            const float rp_m = rp[c - sd], rm_m = rm[c - sd];
            const float rp_p = rp[c + sd], rm_p = rm[c + sd];
            // limiter of the upper face: written by this cell's iteration of loop 3; of the lower face: by the cell
            // below, or never (first interior cell: 0)
            const float c_hi = a_hi >= 0.0f ? fminf(rp_p, rm_c) : fminf(rp_c, rm_p);
            const bool k2 = (q & (along_x ? T_I2 : T_J2)) != 0, o2 = (q & (along_x ? T_J2 : T_I2)) != 0;
            const float c_lo = k2 ? (a_lo >= 0.0f ? fminf(rp_c, rm_m) : fminf(rp_m, rm_c)) : 0.0f;
            // the other direction's antidiffusive fluxes are 0 and its limiters are finite: + 0 * c - 0 * c
            const float rp_op = rp[c + so], rm_om = rm[c - so];
            const float c_ohi = fminf(rp_op, rm_c);                       // 0 >= 0: min(rp[other + 1], rm[c])
            const float c_olo = o2 ? fminf(rp_c, rm_om) : 0.0f;
            float t;
            if (along_x) {
                t = a_hi * c_hi - a_lo * c_lo;
                t = t + 0.0f * c_ohi;
                t = t - 0.0f * c_olo;
            } else {
                t = 0.0f * c_ohi - 0.0f * c_olo;
                t = t + a_hi * c_hi;
                t = t - a_lo * c_lo;
            }
            const float dv = k.dxdy - dtd * (vp - vc);
            const float fn = Ftd[c] - div_nz(((tdiv(t, a.d_dy, a.fast) * k.dx) * k.dy), dv);
            Fo[c] = var3(0.0f, 1.0f, fn);
        });
        __syncthreads();
        float* sw = Fc; Fc = Fo; Fo = sw;
    }
    // two sweeps: the new F is back in m.F.  post_process_f (452-455) on every cell, then set_BC (528)
    within2(1, 1, [&](int c) { m.F[c] = var3(m.F[c], 0.0f, 1.0f); });
    __syncthreads();
    set_bc(0);

    // ---- write back the owned cells
    within2(0, 0, [&](int c) {
        if (!(cls[c] & T_OWN)) return;
        const size_t o = (size_t)(gi0 + (c >> 6)) * P + gj0 + (c & 63);
        a.un[o] = m.u[c]; a.vn[o] = m.v[c]; a.pn[o] = m.p[c]; a.Fn[o] = m.F[c];
        const unsigned q = cls[c];
        if ((q & (T_IN | T_I2)) == (T_IN | T_I2)) a.us[o] = m.us[c];          // the entries advect_upwind writes, no others
        if ((q & (T_IN | T_J2)) == (T_IN | T_J2)) a.vs[o] = m.vs[c];
    });
}

}  // namespace vof
