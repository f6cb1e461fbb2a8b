// Momentum predictor and projection (2dvof.py:206-233, 269-280), instruction-lean versions.
//
// Same pattern as the other kernels: a lane owns 4 consecutive columns (one LDG.128 per field per row),
// a warp owns 128 aligned columns and marches up a chunk of rows; the i-neighbours of a row live in a
// three-slot register ring (the loop is unrolled by 3, nothing is moved), the j-neighbours come from the
// adjacent lanes by shuffle (the two lanes at the strip edge read their missing neighbour directly).
// Arithmetic is the reference's, term for term; the only short-cut is exact: the CSF term is skipped
// where F has no jump across the face (it is then +-0, and x + (+-0) = x).
#pragma once
#include "vof2d_fct.cuh"
#include "vof2d_jacobi_tb.cuh"
#include "vof2d_stream.cuh"
#include "vof_common.cuh"

namespace vof {

constexpr int kMomWarps = 4;

struct MomC {
    Consts k;
    ConstDiv d_dx, d_dy;   // exact division by the constants dx and dy (2dvof.py:213, 226)
    int fast_div_ok;
};

// one row of a field as seen by a lane: [0] = left neighbour (column jl-1), [1..NC] = own columns,
// [NC+1] = right neighbour (jl+NC)
template <int NC> struct RowN { float x[NC + 2]; };

template <int NC>
__device__ __forceinline__ void load_row(RowN<NC>& r, const float* __restrict__ base, size_t off, int lane, bool active,
                                         bool need_left, bool need_right) {
    float v[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) v[q] = 0.0f;
    if (active) VecN<NC>::ld(base + off, v);
#pragma unroll
    for (int q = 0; q < NC; ++q) r.x[q + 1] = v[q];
    if (need_left) {
        r.x[0] = __shfl_up_sync(0xffffffffu, v[NC - 1], 1);
        if (lane == 0 && active) r.x[0] = base[off - 1];
    }
    if (need_right) {
        r.x[NC + 1] = __shfl_down_sync(0xffffffffu, v[0], 1);
        if (lane == 31 && active) r.x[NC + 1] = base[off + NC];
    }
}

template <bool INLINE_PROPS, int NC>
struct AdvState {
    RowN<NC> u[3], v[3];       // ring over rows i-1, i, i+1
    RowN<NC> F[3], kp[3];      // F and kappa: own columns + left neighbour
    float nu[3][NC];           // rho is only needed by the CSF term, i.e. at the interface: fetched / recomputed there
};

template <bool INLINE_PROPS, int NC, int PH>
__device__ __forceinline__ void adv_load(AdvState<INLINE_PROPS, NC>& S, const MomC& c, const float* __restrict__ u,
                                         const float* __restrict__ v, const float* __restrict__ F,
                                         const float* __restrict__ kappa, const float* __restrict__ rho,
                                         const float* __restrict__ nu, size_t off, int lane, bool active) {
    load_row<NC>(S.u[PH], u, off, lane, active, true, true);
    load_row<NC>(S.v[PH], v, off, lane, active, true, true);
    load_row<NC>(S.F[PH], F, off, lane, active, true, false);
    load_row<NC>(S.kp[PH], kappa, off, lane, active, true, false);
    if (INLINE_PROPS) {
#pragma unroll
        for (int q = 0; q < NC; ++q) S.nu[PH][q] = nu_of(S.F[PH].x[q + 1], c.k);
    } else {
#pragma unroll
        for (int q = 0; q < NC; ++q) S.nu[PH][q] = 0.0f;
        if (active) VecN<NC>::ld(nu + off, S.nu[PH]);
    }
}

// row i = slot MID; writes u*(i, .) and v*(i, .)
template <bool INLINE_PROPS, int NC, int PH>
__device__ __forceinline__ void adv_row(const AdvState<INLINE_PROPS, NC>& S, const MomC& c, const float* __restrict__ rho,
                                        int P, float* __restrict__ us, float* __restrict__ vs, size_t off, int gi, int jl,
                                        int nx, int ny) {
    constexpr int P1 = PH, C = (PH + 2) % 3, M = (PH + 1) % 3;   // rows i+1, i, i-1
    const Consts& k = c.k;
    float ou[NC], ov[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const int e = q + 1;                      // index into RowN
        const float u_c = S.u[C].x[e], v_c = S.v[C].x[e];
        const float F_c = S.F[C].x[e], k_c = S.kp[C].x[e], nu_c = S.nu[C][q];
        {   // ---- u*  (2dvof.py:208-220)
            const float u_m = S.u[M].x[e], u_p = S.u[P1].x[e], u_jm = S.u[C].x[e - 1], u_jp = S.u[C].x[e + 1];
            const float v_here = 0.25f * (((S.v[M].x[e] + S.v[M].x[e + 1]) + v_c) + S.v[C].x[e + 1]);
            const float dudx = u_c > 0.0f ? (u_c - u_m) * k.dxi : (u_p - u_c) * k.dxi;
            const float dudy = v_here > 0.0f ? (u_c - u_jm) * k.dyi : (u_jp - u_c) * k.dyi;
            const float two_u = 2.0f * u_c;
            float acc = (nu_c * ((u_m - two_u) + u_p)) * k.dxi2;
            acc = acc + (nu_c * ((u_jm - two_u) + u_jp)) * k.dyi2;
            acc = acc - u_c * dudx;
            acc = acc - v_here * dudy;
            acc = acc + k.gx;
            const float dF = F_c - S.F[M].x[e];
            if (dF != 0.0f) {                     // otherwise the CSF term is +-0
                const float kappa_ave = (k_c + S.kp[M].x[e]) / 2.0f;
                const float t = (k.neg_sigma * dF) * kappa_ave;
                const float fx_kappa = c.fast_div_ok ? div_by_const(t, c.d_dx) : t / k.dx;
                const float rho_c = INLINE_PROPS ? rho_of(F_c, c.k) : rho[off + q];
                const float rho_m = INLINE_PROPS ? rho_of(S.F[M].x[e], c.k) : rho[off + q - P];
                acc = acc + (fx_kappa * 2.0f) / (rho_c + rho_m);
            }
            ou[q] = u_c + k.dt * acc;
        }
        {   // ---- v*  (2dvof.py:221-233)
            const float v_m = S.v[M].x[e], v_p = S.v[P1].x[e], v_jm = S.v[C].x[e - 1], v_jp = S.v[C].x[e + 1];
            const float u_here = 0.25f * (((S.u[C].x[e - 1] + u_c) + S.u[P1].x[e - 1]) + S.u[P1].x[e]);
            const float dvdx = u_here > 0.0f ? (v_c - v_m) * k.dxi : (v_p - v_c) * k.dxi;
            const float dvdy = v_c > 0.0f ? (v_c - v_jm) * k.dyi : (v_jp - v_c) * k.dyi;
            const float two_v = 2.0f * v_c;
            float acc = (nu_c * ((v_m - two_v) + v_p)) * k.dxi2;
            acc = acc + (nu_c * ((v_jm - two_v) + v_jp)) * k.dyi2;
            acc = acc - u_here * dvdx;
            acc = acc - v_c * dvdy;
            acc = acc + k.gy;
            const float dF = F_c - S.F[C].x[e - 1];
            if (dF != 0.0f) {
                const float kappa_ave = (k_c + S.kp[C].x[e - 1]) / 2.0f;
                const float t = (k.neg_sigma * dF) * kappa_ave;
                const float fy_kappa = c.fast_div_ok ? div_by_const(t, c.d_dy) : t / k.dy;
                const float rho_c = INLINE_PROPS ? rho_of(F_c, c.k) : rho[off + q];
                const float rho_jm = INLINE_PROPS ? rho_of(S.F[C].x[e - 1], c.k) : rho[off + q - 1];
                acc = acc + (fy_kappa * 2.0f) / (rho_c + rho_jm);
            }
            ov[q] = v_c + k.dt * acc;
        }
    }
    // u*: gi in [2, nx], j in [1, ny];  v*: gi in [1, nx], j in [2, ny].  Everything else is never written.
    const bool urow = gi >= 2 && gi <= nx, vrow = gi >= 1 && gi <= nx;
    if (jl + NC - 1 <= ny) {
        if (urow) VecN<NC>::st(us + off, ou);
        if (vrow) {
            if (jl >= 2) VecN<NC>::st(vs + off, ov);
            else {
#pragma unroll
                for (int q = 1; q < NC; ++q) vs[off + q] = ov[q];
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int j = jl + q;
            if (urow && j <= ny) us[off + q] = ou[q];
            if (vrow && j >= 2 && j <= ny) vs[off + q] = ov[q];
        }
    }
}

// --------------------------------------------------------------------------------------
// adv_row for two columns per lane in Blackwell packed fp32 (FFMA2, vof_common.cuh): the lane's two cells are the two
// halves of every operand, each packed operation is one separately rounded operation of the reference in both halves,
// in the reference's order (2dvof.py:208-233).  Selects (upwind direction) and the CSF term -- non-zero only where F
// jumps across the face -- stay per element.  Half the arithmetic issue slots of adv_row<true, 2, PH>, same bits.
// --------------------------------------------------------------------------------------
struct AdvPkC {
    f32x2 dxi, dyi, dxi2, dyi2, gx, gy, dt, quarter, two;
    Pk2 o;
};
__device__ __forceinline__ AdvPkC adv_pk_consts(const Consts& k) {
    AdvPkC c;
    c.dxi = pk2(k.dxi); c.dyi = pk2(k.dyi); c.dxi2 = pk2(k.dxi2); c.dyi2 = pk2(k.dyi2);
    c.gx = pk2(k.gx); c.gy = pk2(k.gy); c.dt = pk2(k.dt); c.quarter = pk2(0.25f); c.two = pk2(2.0f);
    c.o = pk2_ops(k);
    return c;
}
__device__ __forceinline__ f32x2 sel2(float c0, float c1, f32x2 pos, f32x2 neg) {      // (c > 0 ? pos : neg) per half
    float p0, p1, n0, n1;
    unpk2(pos, p0, p1);
    unpk2(neg, n0, n1);
    return pk2(c0 > 0.0f ? p0 : n0, c1 > 0.0f ? p1 : n1);
}

template <int PH>
__device__ __forceinline__ void adv_row_pk(const AdvState<true, 2>& S, const MomC& c, const AdvPkC& K, int P,
                                           float* __restrict__ us, float* __restrict__ vs, size_t off, int gi, int jl, int nx,
                                           int ny) {
    constexpr int P1 = PH, C = (PH + 2) % 3, M = (PH + 1) % 3;   // rows i+1, i, i-1
    const Consts& k = c.k;
    const float* uC = S.u[C].x; const float* uM = S.u[M].x; const float* uP = S.u[P1].x;
    const float* vC = S.v[C].x; const float* vM = S.v[M].x; const float* vP = S.v[P1].x;
    const f32x2 u_c = pk2(uC[1], uC[2]), u_m = pk2(uM[1], uM[2]), u_p = pk2(uP[1], uP[2]);
    const f32x2 u_jm = pk2(uC[0], uC[1]), u_jp = pk2(uC[2], uC[3]);
    const f32x2 v_c = pk2(vC[1], vC[2]), v_m = pk2(vM[1], vM[2]), v_p = pk2(vP[1], vP[2]);
    const f32x2 v_jm = pk2(vC[0], vC[1]), v_jp = pk2(vC[2], vC[3]);
    const f32x2 nu_c = pk2(S.nu[C][0], S.nu[C][1]);
    const f32x2 F_c = pk2(S.F[C].x[1], S.F[C].x[2]);
    float ou[2], ov[2];
    {   // ---- u*  (2dvof.py:208-220)
        f32x2 v_here = K.o.add(v_m, pk2(vM[2], vM[3]));
        v_here = K.o.add(v_here, v_c);
        v_here = K.o.mul(K.quarter, K.o.add(v_here, v_jp));
        const f32x2 dudx = sel2(uC[1], uC[2], K.o.mul(K.o.sub(u_c, u_m), K.dxi), K.o.mul(K.o.sub(u_p, u_c), K.dxi));
        float vh0, vh1;
        unpk2(v_here, vh0, vh1);
        const f32x2 dudy = sel2(vh0, vh1, K.o.mul(K.o.sub(u_c, u_jm), K.dyi), K.o.mul(K.o.sub(u_jp, u_c), K.dyi));
        const f32x2 two_u = K.o.mul(K.two, u_c);
        f32x2 acc = K.o.mul(K.o.mul(nu_c, K.o.add(K.o.sub(u_m, two_u), u_p)), K.dxi2);
        acc = K.o.add(acc, K.o.mul(K.o.mul(nu_c, K.o.add(K.o.sub(u_jm, two_u), u_jp)), K.dyi2));
        acc = K.o.sub(acc, K.o.mul(u_c, dudx));
        acc = K.o.sub(acc, K.o.mul(v_here, dudy));
        acc = K.o.add(acc, K.gx);
        const f32x2 dF = K.o.sub(F_c, pk2(S.F[M].x[1], S.F[M].x[2]));
        float d0, d1;
        unpk2(dF, d0, d1);
        if (d0 != 0.0f || d1 != 0.0f) {           // otherwise the CSF term is +-0 in both cells
            float a[2];
            unpk2(acc, a[0], a[1]);
            const float d[2] = {d0, d1};
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (d[q] != 0.0f) {
                    const int e = q + 1;
                    const float kappa_ave = (S.kp[C].x[e] + S.kp[M].x[e]) / 2.0f;
                    const float t = (k.neg_sigma * d[q]) * kappa_ave;
                    const float fx_kappa = c.fast_div_ok ? div_by_const(t, c.d_dx) : t / k.dx;
                    a[q] = a[q] + (fx_kappa * 2.0f) / (rho_of(S.F[C].x[e], k) + rho_of(S.F[M].x[e], k));
                }
            acc = pk2(a[0], a[1]);
        }
        unpk2(K.o.add(u_c, K.o.mul(K.dt, acc)), ou[0], ou[1]);
    }
    {   // ---- v*  (2dvof.py:221-233)
        f32x2 u_here = K.o.add(u_jm, u_c);
        u_here = K.o.add(u_here, pk2(uP[0], uP[1]));
        u_here = K.o.mul(K.quarter, K.o.add(u_here, u_p));
        float uh0, uh1;
        unpk2(u_here, uh0, uh1);
        const f32x2 dvdx = sel2(uh0, uh1, K.o.mul(K.o.sub(v_c, v_m), K.dxi), K.o.mul(K.o.sub(v_p, v_c), K.dxi));
        const f32x2 dvdy = sel2(vC[1], vC[2], K.o.mul(K.o.sub(v_c, v_jm), K.dyi), K.o.mul(K.o.sub(v_jp, v_c), K.dyi));
        const f32x2 two_v = K.o.mul(K.two, v_c);
        f32x2 acc = K.o.mul(K.o.mul(nu_c, K.o.add(K.o.sub(v_m, two_v), v_p)), K.dxi2);
        acc = K.o.add(acc, K.o.mul(K.o.mul(nu_c, K.o.add(K.o.sub(v_jm, two_v), v_jp)), K.dyi2));
        acc = K.o.sub(acc, K.o.mul(u_here, dvdx));
        acc = K.o.sub(acc, K.o.mul(v_c, dvdy));
        acc = K.o.add(acc, K.gy);
        const f32x2 dF = K.o.sub(F_c, pk2(S.F[C].x[0], S.F[C].x[1]));
        float d0, d1;
        unpk2(dF, d0, d1);
        if (d0 != 0.0f || d1 != 0.0f) {
            float a[2];
            unpk2(acc, a[0], a[1]);
            const float d[2] = {d0, d1};
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (d[q] != 0.0f) {
                    const int e = q + 1;
                    const float kappa_ave = (S.kp[C].x[e] + S.kp[C].x[e - 1]) / 2.0f;
                    const float t = (k.neg_sigma * d[q]) * kappa_ave;
                    const float fy_kappa = c.fast_div_ok ? div_by_const(t, c.d_dy) : t / k.dy;
                    a[q] = a[q] + (fy_kappa * 2.0f) / (rho_of(S.F[C].x[e], k) + rho_of(S.F[C].x[e - 1], k));
                }
            acc = pk2(a[0], a[1]);
        }
        unpk2(K.o.add(v_c, K.o.mul(K.dt, acc)), ov[0], ov[1]);
    }
    // u*: gi in [2, nx], j in [1, ny];  v*: gi in [1, nx], j in [2, ny].  Everything else is never written.
    const bool urow = gi >= 2 && gi <= nx, vrow = gi >= 1 && gi <= nx;
    if (jl + 1 <= ny) {
        if (urow) VecN<2>::st(us + off, ou);
        if (vrow) {
            if (jl >= 2) VecN<2>::st(vs + off, ov);
            else vs[off + 1] = ov[1];
        }
    } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = jl + q;
            if (urow && j <= ny) us[off + q] = ou[q];
            if (vrow && j >= 2 && j <= ny) vs[off + q] = ov[q];
        }
    }
}

template <bool INLINE_PROPS, int NC>
__global__ void __launch_bounds__(32 * kMomWarps)
k_advect4(Grid g, MomC c, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ F,
          const float* __restrict__ kappa, const float* __restrict__ rho, const float* __restrict__ nu,
          float* __restrict__ us, float* __restrict__ vs, int r0, int r1, int rows_per_chunk, int nstrips) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kMomWarps + (threadIdx.x >> 5);
    const int strip = w % nstrips, chunk = w / nstrips;
    const int ia = r0 + chunk * rows_per_chunk;
    if (ia > r1) return;
    const int ib = min(r1, ia + rows_per_chunk - 1);
    const int jl = 1 + 32 * NC * strip + NC * lane;
    const bool active = jl <= g.ny + 1;
    const int P = g.pitch;
    const size_t col = (size_t)jl;
    AdvState<INLINE_PROPS, NC> S;
    // rows ia-1 and ia enter slots 1 and 2 (OLD and MID of phase 0)
    adv_load<INLINE_PROPS, NC, 1>(S, c, u, v, F, kappa, rho, nu, (size_t)(ia - 1) * P + col, lane, active);
    adv_load<INLINE_PROPS, NC, 2>(S, c, u, v, F, kappa, rho, nu, (size_t)ia * P + col, lane, active);
    for (int i = ia; i <= ib; i += 3) {
        // every lane takes part in the shuffles of the loads; stores are guarded by `active` and the row range
        adv_load<INLINE_PROPS, NC, 0>(S, c, u, v, F, kappa, rho, nu, (size_t)min(i + 1, g.nrows - 1) * P + col, lane, active);
        if (active) adv_row<INLINE_PROPS, NC, 0>(S, c, rho, P, us, vs, (size_t)i * P + col, g.gi0 + i, jl, g.nx, g.ny);
        if (i + 1 > ib) break;
        adv_load<INLINE_PROPS, NC, 1>(S, c, u, v, F, kappa, rho, nu, (size_t)min(i + 2, g.nrows - 1) * P + col, lane, active);
        if (active) adv_row<INLINE_PROPS, NC, 1>(S, c, rho, P, us, vs, (size_t)(i + 1) * P + col, g.gi0 + i + 1, jl, g.nx, g.ny);
        if (i + 2 > ib) break;
        adv_load<INLINE_PROPS, NC, 2>(S, c, u, v, F, kappa, rho, nu, (size_t)min(i + 3, g.nrows - 1) * P + col, lane, active);
        if (active) adv_row<INLINE_PROPS, NC, 2>(S, c, rho, P, us, vs, (size_t)(i + 2) * P + col, g.gi0 + i + 2, jl, g.nx, g.ny);
    }
}

// --------------------------------------------------------------------------------------
// Momentum predictor, second generation: the same adv_row arithmetic; rows of u, v, F, kappa (and the columns next to
// the strip) arrive through the per-lane cp.async ring, items through the work queue (vof2d_stream.cuh).  The
// first-generation kernel waited on global loads issued one row ahead (ncu: long-scoreboard stalls dominate).
// Properties are always taken from F (the fused step); the materialised rho / nu variant stays k_advect4<false>.
// --------------------------------------------------------------------------------------
constexpr int kAdvSlots = 8;

template <int NC, int PH, class Ring>
__device__ __forceinline__ void adv_load_ring(AdvState<true, NC>& S, const MomC& c, Ring& ring, const float* const (&src)[4]) {
    float X[4][NC + 2];
    ring.next(X, src);
#pragma unroll
    for (int q = 0; q < NC + 2; ++q) { S.u[PH].x[q] = X[0][q]; S.v[PH].x[q] = X[1][q]; }
#pragma unroll
    for (int q = 0; q < NC + 1; ++q) { S.F[PH].x[q] = X[2][q]; S.kp[PH].x[q] = X[3][q]; }
#pragma unroll
    for (int q = 0; q < NC; ++q) S.nu[PH][q] = nu_of(S.F[PH].x[q + 1], c.k);
}

template <int NC, bool PACKED>
__global__ void __launch_bounds__(32 * kMomWarps)
k_advect5(Grid g, MomC c, WorkQueue wq, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ F,
          const float* __restrict__ kappa, float* __restrict__ us, float* __restrict__ vs, int r0, int r1, int rows_per_chunk,
          int nstrips) {
    using Ring = RowRingH<4, NC, kAdvSlots, 32 * kMomWarps, 0xfu, 0x3u>;     // left: u, v, F, kappa; right: u, v
    __shared__ __align__(16) unsigned char ring_mem[Ring::kBytes];
    const int lane = threadIdx.x & 31;
    const int P = g.pitch;
    Ring ring;
    ring.init(ring_mem, threadIdx.x);
    const AdvPkC K = adv_pk_consts(c.k);
    for (;;) {
        const int item = wq_claim(wq, lane);
        if (item >= wq.nitems) break;
        const int strip = item % nstrips, chunk = item / nstrips;
        const int ia = r0 + chunk * rows_per_chunk;
        const int ib = min(r1, ia + rows_per_chunk - 1);
        const int jl = 1 + 32 * NC * strip + NC * lane;
        const bool active = jl <= g.ny + 1;
        const size_t col = (size_t)jl;
        const float* const src[4] = {u + jl, v + jl, F + jl, kappa + jl};
        ring.start(active, ia - 1, ib + 1, g.nrows - 1, P, src);
        AdvState<true, NC> S;
#define VOF_ADV_ROW(PH, I)                                                                                            \
    if (active) {                                                                                                     \
        if constexpr (NC == 2) { if (PACKED) adv_row_pk<PH>(S, c, K, P, us, vs, (size_t)(I) * P + col, g.gi0 + (I), jl, g.nx, g.ny); \
                                 else adv_row<true, NC, PH>(S, c, nullptr, P, us, vs, (size_t)(I) * P + col, g.gi0 + (I), jl, g.nx, g.ny); } \
        else adv_row<true, NC, PH>(S, c, nullptr, P, us, vs, (size_t)(I) * P + col, g.gi0 + (I), jl, g.nx, g.ny);     \
    }
        // rows ia-1 and ia enter slots 1 and 2 (OLD and MID of phase 0)
        adv_load_ring<NC, 1>(S, c, ring, src);
        adv_load_ring<NC, 2>(S, c, ring, src);
        for (int i = ia; i <= ib; i += 3) {
            adv_load_ring<NC, 0>(S, c, ring, src);
            VOF_ADV_ROW(0, i)
            if (i + 1 > ib) break;
            adv_load_ring<NC, 1>(S, c, ring, src);
            VOF_ADV_ROW(1, i + 1)
            if (i + 2 > ib) break;
            adv_load_ring<NC, 2>(S, c, ring, src);
            VOF_ADV_ROW(2, i + 2)
        }
#undef VOF_ADV_ROW
    }
    ring.drain();
    wq_leave(wq, lane, gridDim.x * kMomWarps);
}

// ======================================================================================
// projection, 2dvof.py:269-280: u = u* - dt/r * (p[i,j]-p[i-1,j]) * dxi, r = (rho[i,j]+rho[i-1,j])*0.5 (v alike)
// ======================================================================================
template <bool INLINE_PROPS>
__global__ void __launch_bounds__(32 * kMomWarps)
k_project4(Grid g, Consts k, const float* __restrict__ rhoF, const float* __restrict__ p, const float* __restrict__ us,
           const float* __restrict__ vs, float* __restrict__ u, float* __restrict__ v,
           unsigned long long* __restrict__ courant_count, int r0, int r1, int rows_per_chunk, int nstrips, int own_a,
           int own_b) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kMomWarps + (threadIdx.x >> 5);
    const int strip = w % nstrips, chunk = w / nstrips;
    const int ia = r0 + chunk * rows_per_chunk;
    if (ia > r1) return;
    const int ib = min(r1, ia + rows_per_chunk - 1);
    const int jl = 1 + 128 * strip + 4 * lane;
    const bool active = jl <= g.ny;
    const int P = g.pitch;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ld4 = [&](const float* base, int i) { return active ? *reinterpret_cast<const float4*>(base + (size_t)i * P + jl) : zero4; };
    auto props = [&](float4 f) {
        return INLINE_PROPS ? make_float4(rho_of(f.x, k), rho_of(f.y, k), rho_of(f.z, k), rho_of(f.w, k)) : f;
    };
    float4 p_m = ld4(p, ia - 1), r_m = props(ld4(rhoF, ia - 1));
    unsigned flags = 0;
    float4 p_n = ld4(p, ia), f_n = ld4(rhoF, ia), us_n = ld4(us, ia), vs_n = ld4(vs, ia);
    for (int i = ia; i <= ib; ++i) {
        const float4 p_c = p_n, us_c = us_n, vs_c = vs_n;
        const float4 r_c = props(f_n);
        if (i < ib) { p_n = ld4(p, i + 1); f_n = ld4(rhoF, i + 1); us_n = ld4(us, i + 1); vs_n = ld4(vs, i + 1); }
        const int gi = g.gi0 + i;
        // left neighbours (column jl-1) of p and rho
        float p_l = __shfl_up_sync(0xffffffffu, p_c.w, 1), r_l = __shfl_up_sync(0xffffffffu, r_c.w, 1);
        if (lane == 0 && active) {
            const size_t o = (size_t)i * P + jl - 1;
            p_l = p[o];
            r_l = INLINE_PROPS ? rho_of(rhoF[o], k) : rhoF[o];
        }
        if (!active) continue;
        const float pc[5] = {p_l, p_c.x, p_c.y, p_c.z, p_c.w}, rc[5] = {r_l, r_c.x, r_c.y, r_c.z, r_c.w};
        const float pm[4] = {p_m.x, p_m.y, p_m.z, p_m.w}, rm[4] = {r_m.x, r_m.y, r_m.z, r_m.w};
        const float usv[4] = {us_c.x, us_c.y, us_c.z, us_c.w}, vsv[4] = {vs_c.x, vs_c.y, vs_c.z, vs_c.w};
        float ou[4], ov[4];
        const bool own = i >= own_a && i <= own_b;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float ru = (rc[q + 1] + rm[q]) * 0.5f;
            ou[q] = usv[q] - ((k.dt / ru) * (pc[q + 1] - pm[q])) * k.dxi;
            const float rv = (rc[q + 1] + rc[q]) * 0.5f;
            ov[q] = vsv[q] - ((k.dt / rv) * (pc[q + 1] - pc[q])) * k.dyi;
        }
        const bool urow = gi >= 2 && gi <= g.nx, vrow = gi >= 1 && gi <= g.nx;
        const size_t off = (size_t)i * P + jl;
        if (jl + 3 <= g.ny) {
            if (urow) {
                *reinterpret_cast<float4*>(u + off) = make_float4(ou[0], ou[1], ou[2], ou[3]);
                if (own) flags += (ou[0] * k.dt > k.cflx) + (ou[1] * k.dt > k.cflx) + (ou[2] * k.dt > k.cflx) + (ou[3] * k.dt > k.cflx);
            }
            if (vrow) {
                if (jl >= 2) *reinterpret_cast<float4*>(v + off) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                else { v[off + 1] = ov[1]; v[off + 2] = ov[2]; v[off + 3] = ov[3]; }
                if (own) flags += (jl >= 2 && ov[0] * k.dt > k.cfly) + (ov[1] * k.dt > k.cfly) + (ov[2] * k.dt > k.cfly) + (ov[3] * k.dt > k.cfly);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = jl + q;
                if (urow && j <= g.ny) { u[off + q] = ou[q]; flags += own && (ou[q] * k.dt > k.cflx); }
                if (vrow && j >= 2 && j <= g.ny) { v[off + q] = ov[q]; flags += own && (ov[q] * k.dt > k.cfly); }
            }
        }
        p_m = p_c; r_m = r_c;
    }
    if (flags) atomicAdd(courant_count, (unsigned long long)flags);
}

}  // namespace vof
