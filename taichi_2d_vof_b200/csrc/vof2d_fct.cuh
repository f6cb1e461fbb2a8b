// Zalesak/Rudman FCT sweeps (2dvof.py:321-448), instruction-lean versions.
//
// Both sweeps are issue-bound, not HBM-bound, when written naively (exact IEEE divisions: six per
// cell).  These kernels keep the arithmetic bit-identical to the reference's expressions and cut the
// instruction count by
//   * 4 columns per lane (one LDG.128 / STG.128 per lane per row and per field);
//   * computing each face flux once (the reference recomputes it in loops 1 and 2 and from both sides);
//   * exact short-cuts that skip divisions whose result is known: a face whose two cells hold the same
//     F has zero antidiffusive flux, so away from the interface the limiter ratios, the face limiter
//     and the corrective update are skipped (x - 0 = x, 0 / y = 0 exactly);  min(1, q/p) = 1 when
//     q >= p > 0;
//   * division by the constants dx*dy and dy through their correctly rounded reciprocals plus one FMA
//     residual correction (verified exhaustively at context creation, see vof2d_jacobi_tb.cuh).
// x-sweep: a lane owns 4 columns and marches up a chunk of rows with the whole dependency chain in
// registers (radius 3 along i).  y-sweep: a warp owns a strip of 128 columns of one row at a time and
// exchanges the radius-3 chain with its neighbour lanes by shuffles (120 of 128 columns are stored).
#pragma once
#include "vof2d_jacobi_tb.cuh"
#include "vof2d_stream.cuh"
#include "vof_common.cuh"

namespace vof {

struct FctC {
    float dt, dx, dy, dxdy, dtd;   // dtd = dt*dy (x-sweep) or dt*dx (y-sweep), double-folded
    ConstDiv d_dxdy, d_dy;         // exact division by the constants dx*dy and dy
    int fast_div_ok;
};

__device__ __forceinline__ float fdiv_const(float t, const ConstDiv& d, bool fast) {
    return fast ? div_by_const(t, d) : t / d.b;
}

// one face: low-order (donor) flux and antidiffusive flux a = downwind - donor   (2dvof.py:325-326, 342-348)
__device__ __forceinline__ void fct_face(float vel, float F_m, float F_c, const FctC& c, float& lo, float& a) {
    const float vd = vel * c.dt;
    const float Fl = vel >= 0.0f ? F_m : F_c;
    const float Fh = vel <= 0.0f ? F_m : F_c;
    lo = vd * Fl;
    a = vd * Fh - lo;
}

// transported-diffused value (2dvof.py:329-331); `interior` = the cell is in the global interior
__device__ __forceinline__ float fct_ftd2(float F_c, float lo_c, float lo_p, float dv, bool interior, const FctC& c) {
    const float s = lo_c - lo_p;
    float t = F_c;
    if (s != 0.0f) t = F_c + fdiv_const(s * c.dy, c.d_dxdy, c.fast_div_ok);   // F + (+-0) = F otherwise
    float r = 0.0f;
    if (t != 0.0f) {                                   // 0 * dx * dy / dv = 0
        r = ((t * c.dx) * c.dy) / dv;
        if (r > 1.0f || r < 0.0f) r = var01(r);
    }
    return interior ? r : 0.0f;
}

// The same with one more exact short-cut: a full cell (t == 1) whose faces move alike (dv == dx*dy bit for bit -- the
// liquid at rest, or in uniform motion) evaluates ((1 * dx) * dy) / (dx*dy), a constant: `td_one` is that expression,
// range check included, evaluated once per work item by the same instructions.
__device__ __forceinline__ float fct_td_one(const FctC& c) {
    float r = ((1.0f * c.dx) * c.dy) / c.dxdy;
    if (r > 1.0f || r < 0.0f) r = var01(r);
    return r;
}
__device__ __forceinline__ float fct_ftd3(float F_c, float lo_c, float lo_p, float dv, bool interior, const FctC& c, float td_one) {
    const float s = lo_c - lo_p;
    float t = F_c;
    if (s != 0.0f) t = F_c + fdiv_const(s * c.dy, c.d_dxdy, c.fast_div_ok);   // F + (+-0) = F otherwise
    float r = 0.0f;
    if (t == 1.0f && dv == c.dxdy) r = td_one;
    else if (t != 0.0f) {                              // 0 * dx * dy / dv = 0
        r = ((t * c.dx) * c.dy) / dv;
        if (r > 1.0f || r < 0.0f) r = var01(r);
    }
    return interior ? r : 0.0f;
}

// limiter ratios of one cell (2dvof.py:334-335, 352-363); all-zero antidiffusive fluxes give 0, 0
__device__ __forceinline__ void fct_ratios2(float td_m, float td_c, float td_p, float a_c, float a_p, bool interior,
                                            const FctC& c, float& rp, float& rm) {
    rp = 0.0f; rm = 0.0f;
    if (interior && (a_c != 0.0f || a_p != 0.0f)) {
        const float pp = fmaxf(0.0f, a_c) - fminf(0.0f, a_p);
        const float pm = fmaxf(0.0f, a_p) - fminf(0.0f, a_c);
        if (pp > 0.0f) {
            const float fmax = fmaxf(fmaxf(td_c, td_m), td_p);
            const float qp = (fmax - td_c) * c.dx;
            rp = qp >= pp ? 1.0f : div_nz(qp, pp);     // min(1, q/p) = 1 whenever q >= p > 0; 0 / p without the slow path
        }
        if (pm > 0.0f) {
            const float fmin = fminf(fminf(td_c, td_m), td_p);
            const float qm = (td_c - fmin) * c.dx;
            rm = qm >= pm ? 1.0f : div_nz(qm, pm);
        }
    }
}

// face limiter (2dvof.py:366-369); irrelevant (multiplied by a = 0) when the face carries no flux
__device__ __forceinline__ float fct_cface2(float a_f, float rp_m, float rm_m, float rp_c, float rm_c, bool valid_face) {
    return valid_face ? (a_f >= 0.0f ? fminf(rp_c, rm_m) : fminf(rp_m, rm_c)) : 0.0f;
}

// corrective update + clamp (2dvof.py:377-382) [+ post_process_f, 452-455]
template <bool POST>
__device__ __forceinline__ float fct_update2(float td_c, float a_c, float c_c, float a_p, float c_p, float dv, const FctC& c) {
    float fn = td_c;
    if (a_c != 0.0f || a_p != 0.0f) {
        const float t = a_p * c_p - a_c * c_c;
        if (t != 0.0f) fn = td_c - ((fdiv_const(t, c.d_dy, c.fast_div_ok) * c.dx) * c.dy) / dv;
    }
    float f = var01(fn);
    if (POST) f = var01(f);
    return f;
}

// ======================================================================================
// x-sweep: lane = 4 columns, marches rows ia-2 .. ib+3, output row = row loaded 3 iterations ago
// ======================================================================================
constexpr int kFctXWarps = 4;

// NC consecutive floats as one vector load / store (NC = 2 or 4)
template <int NC> struct VecN;
template <> struct VecN<4> {
    static __device__ __forceinline__ void ld(const float* p, float (&x)[4]) { const float4 v = *reinterpret_cast<const float4*>(p); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
    static __device__ __forceinline__ void st(float* p, const float (&x)[4]) { *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]); }
};
template <> struct VecN<2> {
    static __device__ __forceinline__ void ld(const float* p, float (&x)[2]) { const float2 v = *reinterpret_cast<const float2*>(p); x[0] = v.x; x[1] = v.y; }
    static __device__ __forceinline__ void st(float* p, const float (&x)[2]) { *reinterpret_cast<float2*>(p) = make_float2(x[0], x[1]); }
};

// NC columns per lane: 4 = fewest instructions per cell, 2 = half the registers (twice the resident warps);
// the sweep is latency bound, so the narrower variant wins on B200 (profiles/)
template <bool POST, int NC>
__global__ void __launch_bounds__(32 * kFctXWarps)
k_fct_x4(Grid g, FctC c, const float* __restrict__ Fin, const float* __restrict__ u, float* __restrict__ Fout,
         int r0, int r1, int rows_per_chunk, int nstrips) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kFctXWarps + (threadIdx.x >> 5);
    const int strip = w % nstrips, chunk = w / nstrips;
    const int ia = r0 + chunk * rows_per_chunk;
    if (ia > r1) return;
    const int ib = min(r1, ia + rows_per_chunk - 1);
    const int jl = 1 + 32 * NC * strip + NC * lane;
    if (jl > g.ny + 1) return;
    const int P = g.pitch, last = g.nrows - 1;
    const float* Fc = Fin + jl;
    const float* uc = u + jl;
    float* Fo = Fout + jl;
    auto ldv = [&](const float* base, int i, float (&x)[NC]) { VecN<NC>::ld(base + (size_t)min(max(i, 0), last) * P, x); };
    const bool lo_wall = g.gi0 + ia == 1, hi_wall = g.gi0 + ib == g.nx;
    // ghost rows / ghost column 0 pass through (the sweep never writes them; out-of-place needs the copy)
    if (lo_wall) { float t[NC]; ldv(Fc, ia - 1, t); VecN<NC>::st(Fo + (size_t)(ia - 1) * P, t); }
    if (hi_wall) { float t[NC]; ldv(Fc, ib + 1, t); VecN<NC>::st(Fo + (size_t)(ib + 1) * P, t); }
    if (strip == 0 && lane == 0) {
        for (int i = ia - (lo_wall ? 1 : 0); i <= ib + (hi_wall ? 1 : 0); ++i) Fout[(size_t)i * P] = Fin[(size_t)i * P];
    }
    bool colin[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) colin[k] = jl + k <= g.ny;

    float F1[NC], u1[NC], lo1[NC], a1[NC], a2[NC], a3[NC], td1[NC], td2[NC], td3[NC], dv1[NC], dv2[NC], dv3[NC], rp2[NC], rm2[NC], c2[NC];
    float Fk[NC], uk[NC];
    ldv(Fc, ia - 3, F1); ldv(Fc, ia - 2, Fk); ldv(uc, ia - 2, uk);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        u1[k] = 0.f; lo1[k] = 0.f; a1[k] = a2[k] = a3[k] = 0.f; td1[k] = td2[k] = td3[k] = 0.f;
        dv1[k] = dv2[k] = dv3[k] = 1.f; rp2[k] = rm2[k] = 0.f; c2[k] = 0.f;
    }
    for (int k = ia - 2; k <= ib + 3; ++k) {
        float Fn[NC], un[NC];
        ldv(Fc, k + 1, Fn); ldv(uc, k + 1, un);                         // prefetch the next row
        const int gk = g.gi0 + k;
        const bool in1 = gk - 1 >= 1 && gk - 1 <= g.nx;                 // cell k-1 interior
        const bool in2 = gk - 2 >= 1 && gk - 2 <= g.nx;                 // cell k-2 interior
        const bool face2 = gk - 2 >= 2 && gk - 2 <= g.nx + 1;           // face k-2 has a limiter (cx[1] is never written)
        const int io = k - 3;
        const bool store = io >= ia && io <= ib;                        // rows ia..ib are interior by construction
        float out[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            float lo0, a0;
            fct_face(uk[q], F1[q], Fk[q], c, lo0, a0);                                   // face k
            const float dv_n = c.dxdy - c.dtd * (uk[q] - u1[q]);                         // cell k-1
            const float td_n = fct_ftd2(F1[q], lo1[q], lo0, dv_n, in1, c);
            td3[q] = td2[q]; td2[q] = td1[q]; td1[q] = td_n;
            dv3[q] = dv2[q]; dv2[q] = dv1[q]; dv1[q] = dv_n;
            float rp_n, rm_n;
            fct_ratios2(td3[q], td2[q], td1[q], a2[q], a1[q], in2, c, rp_n, rm_n);       // cell k-2
            const float c_n = fct_cface2(a2[q], rp2[q], rm2[q], rp_n, rm_n, face2);      // face k-2
            out[q] = fct_update2<POST>(td3[q], a3[q], c2[q], a2[q], c_n, dv3[q], c);     // cell k-3
            rp2[q] = rp_n; rm2[q] = rm_n; c2[q] = c_n;
            a3[q] = a2[q]; a2[q] = a1[q]; a1[q] = a0; lo1[q] = lo0;
            F1[q] = Fk[q]; u1[q] = uk[q];
        }
        if (store) {
            if (!colin[NC - 1]) {                 // pass-through for columns past ny (ghost ny+1, padding)
                float old[NC];
                ldv(Fc, io, old);
#pragma unroll
                for (int q = 0; q < NC; ++q) out[q] = colin[q] ? out[q] : old[q];
            }
            VecN<NC>::st(Fo + (size_t)io * P, out);
        }
#pragma unroll
        for (int q = 0; q < NC; ++q) { Fk[q] = Fn[q]; uk[q] = un[q]; }
    }
}

// The expressions of one FCT sweep as a policy, so that the second-generation y-sweep kernel (stencil along the
// contiguous axis) also serves the z-sweep of the 3-D solver (FctOps3 in vof3d_kernels.cuh).
struct FctOps2 {                                   // 2dvof.py:321-448
    using C = FctC;
    static __device__ __forceinline__ void face(float vel, float F_m, float F_c, const C& c, float& lo, float& a) { fct_face(vel, F_m, F_c, c, lo, a); }
    static __device__ __forceinline__ float lo_uniform(float vel, float c0, const C& c) { return (vel * c.dt) * c0; }
    static __device__ __forceinline__ float dv(const C& c, float dvel) { return c.dxdy - c.dtd * dvel; }
    static __device__ __forceinline__ float td_one(const C& c) { return fct_td_one(c); }
    static __device__ __forceinline__ float ftd(float F_c, float lo_c, float lo_p, float dv, bool interior, const C& c, float td_one) {
        return fct_ftd3(F_c, lo_c, lo_p, dv, interior, c, td_one);
    }
    static __device__ __forceinline__ void ratios(float td_m, float td_c, float td_p, float a_c, float a_p, bool interior, const C& c,
                                                  float& rp, float& rm) { fct_ratios2(td_m, td_c, td_p, a_c, a_p, interior, c, rp, rm); }
    static __device__ __forceinline__ float cface(float a_f, float rp_m, float rm_m, float rp_c, float rm_c, bool valid) {
        return fct_cface2(a_f, rp_m, rm_m, rp_c, rm_c, valid);
    }
    template <bool POST>
    static __device__ __forceinline__ float update(float td_c, float a_c, float c_c, float a_p, float c_p, float dv, const C& c) {
        return fct_update2<POST>(td_c, a_c, c_c, a_p, c_p, dv, c);
    }
};

// --------------------------------------------------------------------------------------
// x-sweep, second generation: the same per-cell arithmetic (fct_face .. fct_update2), fed by the per-lane cp.async
// ring of vof2d_stream.cuh, scheduled through a work queue, with two warp-uniform short-cuts for the bulk of a
// two-phase field.
//   * Gas mode: while the 7 rows F[k-6 .. k] of the strip are all 0 (and u finite), cell k-3 stays 0: 0 * dx * dy
//     / dv = 0, all fluxes +-0.  The pipeline is not advanced; when a non-zero row arrives its state is what 7
//     zero rows leave behind: Ftd = 0, a = 0, ratios 0, lo = (u dt) * 0 -- the chunk's initial state.
//   * Flux-free rows (any uniform value, typically the liquid at exactly 1): when no lane has a non-zero
//     antidiffusive flux on faces k-3 .. k-1, ratios, face limiter and corrective update are skipped warp-wide
//     instead of lane by lane (no divergence bookkeeping): F' = var(Ftd).
// --------------------------------------------------------------------------------------
constexpr int kFctXSlots = 8;                  // ring slots per lane and field (6 rows in flight)

// X = the sweep's expressions (FctOps2 / FctOps3).  BATCH = false: one 2-D field.  BATCH = true (3-D x- and y-sweeps):
// nbatch independent 2-D problems `batch_stride` floats apart (x-sweep: the j-lines of every plane, marching i with
// pitch = plane stride; y-sweep: the planes, marching j with pitch = line stride); problems outside the interior
// [batch_lo, batch_hi] (ghost lines / ghost planes) are copied through.
template <bool POST, int NC, class X, bool BATCH>
__global__ void __launch_bounds__(32 * kFctXWarps)
k_fct_x5(Grid g, typename X::C c, WorkQueue wq, const float* __restrict__ Fin0, const float* __restrict__ u0, float* __restrict__ Fout0,
         int r0, int r1, int rows_per_chunk, int nstrips, int nbatch, long long batch_stride, int batch_lo, int batch_hi) {
    using Ring = RowRing<2, NC, kFctXSlots, 32 * kFctXWarps>;
    __shared__ __align__(16) unsigned char ring_mem[Ring::kBytes];
    const int lane = threadIdx.x & 31;
    const int P = g.pitch, last = g.nrows - 1;
    const int P4i = P * 4;
    const long long P4 = P4i;
    const float td_one = X::td_one(c);
    Ring ring;
    ring.init(ring_mem, threadIdx.x);
    for (;;) {
    const int item = wq_claim(wq, lane);
    if (item >= wq.nitems) break;
    const int strip = item % nstrips;
    int chunk = item / nstrips;
    const float* Fin = Fin0;
    const float* u = u0;
    float* Fout = Fout0;
    bool batch_in = true;
    if (BATCH) {
        const int bi = chunk % nbatch;
        chunk /= nbatch;
        Fin += bi * batch_stride; u += bi * batch_stride; Fout += bi * batch_stride;
        batch_in = bi >= batch_lo && bi <= batch_hi;
    }
    const int ia = r0 + chunk * rows_per_chunk;
    const int ib = min(r1, ia + rows_per_chunk - 1);
    const int jl = 1 + 32 * NC * strip + NC * lane;
    const bool active = jl <= g.ny + 1;                                 // idle lanes stay for the votes
    const float* Fc = Fin + jl;
    const float* uc = u + jl;
    char* Fo = reinterpret_cast<char*>(Fout + jl);
    const float* const src[2] = {Fc, uc};
    auto ldv = [&](const float* base, int i, float (&x)[NC]) { VecN<NC>::ld(base + (size_t)min(max(i, 0), last) * P, x); };
    const bool lo_wall = g.gi0 + ia == 1, hi_wall = g.gi0 + ib == g.nx;
    // ghost rows / ghost column 0 pass through (the sweep never writes them; out-of-place needs the copy)
    if (active && lo_wall) { float t[NC]; ldv(Fc, ia - 1, t); VecN<NC>::st(reinterpret_cast<float*>(Fo + (ia - 1) * P4), t); }
    if (active && hi_wall) { float t[NC]; ldv(Fc, ib + 1, t); VecN<NC>::st(reinterpret_cast<float*>(Fo + (ib + 1) * P4), t); }
    if (strip == 0 && lane == 0) {
        for (int i = ia - (lo_wall ? 1 : 0); i <= ib + (hi_wall ? 1 : 0); ++i) Fout[(size_t)i * P] = Fin[(size_t)i * P];
    }
    if (BATCH && !batch_in) {                     // ghost line / ghost plane: the sweep leaves it alone
        if (active) {
            for (int i = ia; i <= ib; ++i) { float t[NC]; ldv(Fc, i, t); VecN<NC>::st(reinterpret_cast<float*>(Fo + i * P4), t); }
        }
        continue;
    }
    bool colin[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) colin[k] = jl + k <= g.ny;
    if (active && !colin[NC - 1]) {               // last strip only: ghost column ny+1 and padding pass through
        for (int i = ia; i <= ib; ++i) {
#pragma unroll
            for (int q = 0; q < NC; ++q) if (!colin[q]) Fout[(size_t)i * P + jl + q] = Fin[(size_t)i * P + jl + q];
        }
    }
    char* po = Fo + (long long)ia * P4i;          // row io = ia is the first one stored

    ring.start(active, ia - 3, ib + 3, last, P, src);
    float X[2][NC];
    float F1[NC], u1[NC], lo1[NC], a1[NC], a2[NC], a3[NC], td1[NC], td2[NC], td3[NC], dv1[NC], dv2[NC], dv3[NC], rp2[NC], rm2[NC], c2[NC];
    ring.next(X, src);                                                  // row ia-3
#pragma unroll
    for (int q = 0; q < NC; ++q) { F1[q] = X[0][q]; u1[q] = X[1][q]; }
    auto reset_state = [&]() {          // what 7 all-zero rows leave in the pipeline (= the start of a chunk)
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            lo1[q] = X::lo_uniform(u1[q], 0.0f, c);
            a1[q] = a2[q] = a3[q] = 0.f; td1[q] = td2[q] = td3[q] = 0.f;
            dv1[q] = dv2[q] = dv3[q] = 1.f; rp2[q] = rm2[q] = 0.f; c2[q] = 0.f;
        }
    };
    reset_state();
    // zbits: bit b = row k-b of the strip is all zero with finite u.  Rows before ia-3 count as zero: no stored cell
    // depends on them (radius 3), and the pipeline's initial state above is the state zero rows leave behind.
    // abits: bit b = some lane has a non-zero antidiffusive flux on face k-b
    unsigned zbits, abits = 0;
    {
        bool z = all_finite_n<NC>(u1);
#pragma unroll
        for (int q = 0; q < NC; ++q) z = z && F1[q] == 0.0f;
        zbits = 0xfffffffeu | (__all_sync(0xffffffffu, z) ? 1u : 0u);
    }
    bool gas = true;                                                    // state == reset_state()
    for (int k = ia - 2; k <= ib + 3; ++k) {
        ring.next(X, src);                                              // row k
        const float (&Fk)[NC] = X[0];
        const float (&uk)[NC] = X[1];
        const int io = k - 3;
        const bool store = active && io >= ia;                          // io <= ib by the loop bound; rows ia..ib are interior
        float out[NC];
        bool z = all_finite_n<NC>(uk);
#pragma unroll
        for (int q = 0; q < NC; ++q) z = z && Fk[q] == 0.0f;
        zbits = (zbits << 1) | (__all_sync(0xffffffffu, z) ? 1u : 0u);
        if ((zbits & 0x7fu) == 0x7fu) {                                 // rows k-6 .. k are zero: cell k-3 stays zero
            gas = true;
#pragma unroll
            for (int q = 0; q < NC; ++q) out[q] = 0.0f;
        } else {
            if (gas) { reset_state(); abits = 0; gas = false; }        // u1 = u[k-1]; F1 = row k-1
            const int gk = g.gi0 + k;
            const bool in1 = gk - 1 >= 1 && gk - 1 <= g.nx;             // cell k-1 interior
            const bool in2 = gk - 2 >= 1 && gk - 2 <= g.nx;             // cell k-2 interior
            const bool face2 = gk - 2 >= 2 && gk - 2 <= g.nx + 1;       // face k-2 has a limiter (cx[1] is never written)
            float a0[NC];
            bool anya = false;
#pragma unroll
            for (int q = 0; q < NC; ++q) {
                float lo0;
                X::face(uk[q], F1[q], Fk[q], c, lo0, a0[q]);                                 // face k
                const float dv_n = X::dv(c, uk[q] - u1[q]);                                  // cell k-1
                const float td_n = X::ftd(F1[q], lo1[q], lo0, dv_n, in1, c, td_one);
                td3[q] = td2[q]; td2[q] = td1[q]; td1[q] = td_n;
                dv3[q] = dv2[q]; dv2[q] = dv1[q]; dv1[q] = dv_n;
                lo1[q] = lo0;
                anya = anya || a0[q] != 0.0f;
            }
            abits = (abits << 1) | (__any_sync(0xffffffffu, anya) ? 1u : 0u);
            if ((abits & 0xeu) == 0u) {                                 // faces k-1, k-2, k-3 carry no flux in any lane
#pragma unroll
                for (int q = 0; q < NC; ++q) {
                    float f = var01(td3[q]);
                    if (POST) f = var01(f);
                    out[q] = f;
                    rp2[q] = 0.0f; rm2[q] = 0.0f; c2[q] = 0.0f;
                }
            } else {
#pragma unroll
                for (int q = 0; q < NC; ++q) {
                    float rp_n, rm_n;
                    X::ratios(td3[q], td2[q], td1[q], a2[q], a1[q], in2, c, rp_n, rm_n);         // cell k-2
                    const float c_n = X::cface(a2[q], rp2[q], rm2[q], rp_n, rm_n, face2);        // face k-2
                    out[q] = X::template update<POST>(td3[q], a3[q], c2[q], a2[q], c_n, dv3[q], c);   // cell k-3
                    rp2[q] = rp_n; rm2[q] = rm_n; c2[q] = c_n;
                }
            }
#pragma unroll
            for (int q = 0; q < NC; ++q) { a3[q] = a2[q]; a2[q] = a1[q]; a1[q] = a0[q]; }
        }
        if (store) {
            float* dst = reinterpret_cast<float*>(po);
            if (colin[NC - 1]) VecN<NC>::st(dst, out);
            else {                                // last strip: columns past ny were copied through above
#pragma unroll
                for (int q = 0; q < NC; ++q) if (colin[q]) dst[q] = out[q];
            }
        }
        if (io >= ia) po += P4;
#pragma unroll
        for (int q = 0; q < NC; ++q) { F1[q] = Fk[q]; u1[q] = uk[q]; }
    }
    }
    ring.drain();
    wq_leave(wq, lane, gridDim.x * kFctXWarps);
}

// ======================================================================================
// y-sweep: warp = strip of 128 columns (lanes 1..30 store), one row per iteration
// ======================================================================================
constexpr int kFctYWarps = 4;
constexpr int kFctYValid = 120;

// every bit pattern with an all-ones exponent is inf or NaN
__device__ __forceinline__ bool all_finite4(const float (&x)[4]) {
    const unsigned m = 0x7fffffffu;
    const unsigned a = max(max(__float_as_uint(x[0]) & m, __float_as_uint(x[1]) & m),
                           max(__float_as_uint(x[2]) & m, __float_as_uint(x[3]) & m));
    return a < 0x7f800000u;
}

// ADAPT: warp-uniform bulk rows.  A two-phase field is exactly 0 or exactly 1 away from the interface; when every
// cell a warp's strip touches holds the same value c, every face has Fl == Fh == c, so its antidiffusive flux is
// vd*c - vd*c = +0 (finite vd) and limiter, face limiter and corrective update are no-ops: F' = var(Ftd).  For c = 0
// also Ftd = 0 (0 * dx * dy / dv), so the row is 0.  The test costs five compares and one vote per lane and row.
template <bool POST, bool ADAPT>
__global__ void __launch_bounds__(32 * kFctYWarps)
k_fct_y4(Grid g, FctC c, const float* __restrict__ Fin, const float* __restrict__ v, float* __restrict__ Fout,
         int r0, int r1, int rows_per_warp, int nstrips) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kFctYWarps + (threadIdx.x >> 5);
    const int strip = w % nstrips, chunk = w / nstrips;
    const int ia = r0 + chunk * rows_per_warp;
    if (ia > r1) return;
    const int ib = min(r1, ia + rows_per_warp - 1);
    const int jl = 1 - 4 + kFctYValid * strip + 4 * lane;     // == 1 (mod 4)
    const bool active = jl <= g.ny + 1;
    const bool store_lane = active && lane >= 1 && lane <= 30;
    const int P = g.pitch;
    const float* Fc = Fin + jl;
    const float* vc = v + jl;
    float* Fo = Fout + jl;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    bool cin[6];      // cells jl-1 .. jl+4 in the global interior
#pragma unroll
    for (int k = 0; k < 6; ++k) cin[k] = jl - 1 + k >= 1 && jl - 1 + k <= g.ny;
    bool fvalid[5];   // faces jl .. jl+4 carry a limiter (cy on face 1 is never written)
#pragma unroll
    for (int k = 0; k < 5; ++k) fvalid[k] = jl + k >= 2 && jl + k <= g.ny + 1;
    bool rel[4];      // columns that exist (ghosts included); padding is exempt from the uniformity test
#pragma unroll
    for (int k = 0; k < 4; ++k) rel[k] = active && jl + k >= 0 && jl + k <= g.ny + 1;

    constexpr int PF = 2;                                              // rows in flight per warp
    float4 Fq[PF], vq[PF];
#pragma unroll
    for (int d = 0; d < PF; ++d) {
        const bool ok = active && ia + d <= ib;
        Fq[d] = ok ? *reinterpret_cast<const float4*>(Fc + (size_t)(ia + d) * P) : zero4;
        vq[d] = ok ? *reinterpret_cast<const float4*>(vc + (size_t)(ia + d) * P) : zero4;
    }
    for (int ibase = ia; ibase <= ib; ibase += PF) {
#pragma unroll
      for (int d = 0; d < PF; ++d) {
        const int i = ibase + d;
        if (i > ib) break;
        const float F[4] = {Fq[d].x, Fq[d].y, Fq[d].z, Fq[d].w}, vv[4] = {vq[d].x, vq[d].y, vq[d].z, vq[d].w};
        if (i + PF <= ib && active) {                                  // prefetch
            Fq[d] = *reinterpret_cast<const float4*>(Fc + (size_t)(i + PF) * P);
            vq[d] = *reinterpret_cast<const float4*>(vc + (size_t)(i + PF) * P);
        }
        const int gi = g.gi0 + i;
        const bool rowin = gi >= 1 && gi <= g.nx;
        const float v_p4 = __shfl_down_sync(0xffffffffu, vv[0], 1);    // v[jl+4]
        if (strip == 0 && lane == 0) Fout[(size_t)i * P] = F[3];       // ghost column 0 (jl + 3 == 0)
        if (ADAPT) {
            const float c0 = __shfl_sync(0xffffffffu, F[0], 1);        // lane 1's first column always exists
            bool uni = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) uni = uni && (!rel[k] || F[k] == c0);
            if (c0 == 0.0f) uni = uni && all_finite4(vv);              // 0 * inf must stay NaN: general path
            if (__all_sync(0xffffffffu, uni)) {
                if (store_lane) {
                    float out[4];
                    if (c0 == 0.0f) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) out[k] = (rowin && cin[k + 1]) ? 0.0f : F[k];
                    } else {
                        float lo[5];
#pragma unroll
                        for (int k = 0; k < 4; ++k) lo[k] = (vv[k] * c.dt) * c0;
                        lo[4] = (v_p4 * c.dt) * c0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float dv = c.dxdy - c.dtd * ((k < 3 ? vv[k + 1] : v_p4) - vv[k]);
                            const float td = fct_ftd2(F[k], lo[k], lo[k + 1], dv, cin[k + 1], c);
                            float f = var01(td);
                            if (POST) f = var01(f);
                            out[k] = (rowin && cin[k + 1]) ? f : F[k];
                        }
                    }
                    *reinterpret_cast<float4*>(Fo + (size_t)i * P) = make_float4(out[0], out[1], out[2], out[3]);
                }
                continue;
            }
        }
        const float F_m1 = __shfl_up_sync(0xffffffffu, F[3], 1);       // F[jl-1]
        float lo[5], a[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) fct_face(vv[k], k ? F[k - 1] : F_m1, F[k], c, lo[k], a[k]);
        lo[4] = __shfl_down_sync(0xffffffffu, lo[0], 1);
        a[4] = __shfl_down_sync(0xffffffffu, a[0], 1);
        float dv[4], td[6];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dv[k] = c.dxdy - c.dtd * ((k < 3 ? vv[k + 1] : v_p4) - vv[k]);
            td[k + 1] = fct_ftd2(F[k], lo[k], lo[k + 1], dv[k], cin[k + 1], c);
        }
        td[0] = __shfl_up_sync(0xffffffffu, td[4], 1);
        td[5] = __shfl_down_sync(0xffffffffu, td[1], 1);
        float rp[5], rm[5];   // index k+1 = cell jl+k; index 0 = cell jl-1
#pragma unroll
        for (int k = 0; k < 4; ++k) fct_ratios2(td[k], td[k + 1], td[k + 2], a[k], a[k + 1], cin[k + 1], c, rp[k + 1], rm[k + 1]);
        rp[0] = __shfl_up_sync(0xffffffffu, rp[4], 1);
        rm[0] = __shfl_up_sync(0xffffffffu, rm[4], 1);
        float cf[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) cf[k] = fct_cface2(a[k], rp[k], rm[k], rp[k + 1], rm[k + 1], fvalid[k]);
        cf[4] = __shfl_down_sync(0xffffffffu, cf[0], 1);
        if (store_lane) {
            float out[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float f = fct_update2<POST>(td[k + 1], a[k], cf[k], a[k + 1], cf[k + 1], dv[k], c);
                out[k] = (rowin && cin[k + 1]) ? f : F[k];             // ghost rows / columns pass through
            }
            *reinterpret_cast<float4*>(Fo + (size_t)i * P) = make_float4(out[0], out[1], out[2], out[3]);
        }
      }
    }
}

// Second generation (ring + queue, vof2d_stream.cuh) -- ADAPT: warp-uniform bulk rows.  A two-phase field is exactly 0 or exactly 1 away from the interface; when every
// cell a warp's strip touches holds the same value c, every face has Fl == Fh == c, so its antidiffusive flux is
// vd*c - vd*c = +0 (finite vd) and limiter, face limiter and corrective update are no-ops: F' = var(Ftd).  For c = 0
// also Ftd = 0 (0 * dx * dy / dv), so the row is 0.  The test costs five compares and one vote per lane and row.
constexpr int kFctYSlots = 8;
// X = the sweep's expressions (FctOps2 / FctOps3).  PLANES = false: the rows are the rows i of a 2-D field.  PLANES = true
// (3-D z-sweep): row r is the k-line (plane r / rows_per_plane, j = r % rows_per_plane) of a field whose planes are
// rows_per_plane = ny3 + 2 lines apart, g.ny is the number of cells along k and a row is interior when both its plane
// and its j are.
template <bool POST, class X, bool PLANES>
__global__ void __launch_bounds__(32 * kFctYWarps)
k_fct_y5(Grid g, typename X::C c, WorkQueue wq, const float* __restrict__ Fin, const float* __restrict__ v, float* __restrict__ Fout,
         int r0, int r1, int rows_per_warp, int nstrips, int rows_per_plane) {
    constexpr bool ADAPT = true;
    using Ring = RowRing<2, 4, kFctYSlots, 32 * kFctYWarps>;
    __shared__ __align__(16) unsigned char ring_mem[Ring::kBytes];
    const int lane = threadIdx.x & 31;
    const float td_one = X::td_one(c);
    Ring ring;
    ring.init(ring_mem, threadIdx.x);
    for (;;) {
    const int item = wq_claim(wq, lane);
    if (item >= wq.nitems) break;
    const int strip = item % nstrips, chunk = item / nstrips;
    const int ia = r0 + chunk * rows_per_warp;
    const int ib = min(r1, ia + rows_per_warp - 1);
    const int jl = 1 - 4 + kFctYValid * strip + 4 * lane;     // == 1 (mod 4)
    const bool active = jl <= g.ny + 1;
    const bool store_lane = active && lane >= 1 && lane <= 30;
    const int P = g.pitch;
    const float* Fc = Fin + jl;
    const float* vc = v + jl;
    float* Fo = Fout + jl;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    bool cin[6];      // cells jl-1 .. jl+4 in the global interior
#pragma unroll
    for (int k = 0; k < 6; ++k) cin[k] = jl - 1 + k >= 1 && jl - 1 + k <= g.ny;
    bool fvalid[5];   // faces jl .. jl+4 carry a limiter (cy on face 1 is never written)
#pragma unroll
    for (int k = 0; k < 5; ++k) fvalid[k] = jl + k >= 2 && jl + k <= g.ny + 1;
    bool rel[4];      // columns that exist (ghosts included); padding is exempt from the uniformity test
#pragma unroll
    for (int k = 0; k < 4; ++k) rel[k] = active && jl + k >= 0 && jl + k <= g.ny + 1;

    const float* const src[2] = {Fc, vc};
    ring.start(active, ia, ib, g.nrows - 1, P, src);
    {
      for (int i = ia; i <= ib; ++i) {
        float X[2][4];
        ring.next(X, src);
        const float (&F)[4] = X[0];
        const float (&vv)[4] = X[1];
        bool rowin;
        if (!PLANES) {
            const int gi = g.gi0 + i;
            rowin = gi >= 1 && gi <= g.nx;
        } else {
            const int plane = i / rows_per_plane, jj = i - plane * rows_per_plane, gi = g.gi0 + plane;
            rowin = gi >= 1 && gi <= g.nx && jj >= 1 && jj <= rows_per_plane - 2;
        }
        const float v_p4 = __shfl_down_sync(0xffffffffu, vv[0], 1);    // v[jl+4]
        if (strip == 0 && lane == 0) Fout[(size_t)i * P] = F[3];       // ghost column 0 (jl + 3 == 0)
        if (ADAPT) {
            const float c0 = __shfl_sync(0xffffffffu, F[0], 1);        // lane 1's first column always exists
            bool uni = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) uni = uni && (!rel[k] || F[k] == c0);
            if (c0 == 0.0f) uni = uni && all_finite4(vv);              // 0 * inf must stay NaN: general path
            if (__all_sync(0xffffffffu, uni)) {
                if (store_lane) {
                    float out[4];
                    if (c0 == 0.0f) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) out[k] = (rowin && cin[k + 1]) ? 0.0f : F[k];
                    } else {
                        float lo[5];
#pragma unroll
                        for (int k = 0; k < 4; ++k) lo[k] = X::lo_uniform(vv[k], c0, c);
                        lo[4] = X::lo_uniform(v_p4, c0, c);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float dv = X::dv(c, (k < 3 ? vv[k + 1] : v_p4) - vv[k]);
                            const float td = X::ftd(F[k], lo[k], lo[k + 1], dv, cin[k + 1], c, td_one);
                            float f = var01(td);
                            if (POST) f = var01(f);
                            out[k] = (rowin && cin[k + 1]) ? f : F[k];
                        }
                    }
                    *reinterpret_cast<float4*>(Fo + (size_t)i * P) = make_float4(out[0], out[1], out[2], out[3]);
                }
                continue;
            }
        }
        const float F_m1 = __shfl_up_sync(0xffffffffu, F[3], 1);       // F[jl-1]
        float lo[5], a[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) X::face(vv[k], k ? F[k - 1] : F_m1, F[k], c, lo[k], a[k]);
        lo[4] = __shfl_down_sync(0xffffffffu, lo[0], 1);
        a[4] = __shfl_down_sync(0xffffffffu, a[0], 1);
        float dv[4], td[6];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dv[k] = X::dv(c, (k < 3 ? vv[k + 1] : v_p4) - vv[k]);
            td[k + 1] = X::ftd(F[k], lo[k], lo[k + 1], dv[k], cin[k + 1], c, td_one);
        }
        td[0] = __shfl_up_sync(0xffffffffu, td[4], 1);
        td[5] = __shfl_down_sync(0xffffffffu, td[1], 1);
        float rp[5], rm[5];   // index k+1 = cell jl+k; index 0 = cell jl-1
#pragma unroll
        for (int k = 0; k < 4; ++k) X::ratios(td[k], td[k + 1], td[k + 2], a[k], a[k + 1], cin[k + 1], c, rp[k + 1], rm[k + 1]);
        rp[0] = __shfl_up_sync(0xffffffffu, rp[4], 1);
        rm[0] = __shfl_up_sync(0xffffffffu, rm[4], 1);
        float cf[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) cf[k] = X::cface(a[k], rp[k], rm[k], rp[k + 1], rm[k + 1], fvalid[k]);
        cf[4] = __shfl_down_sync(0xffffffffu, cf[0], 1);
        if (store_lane) {
            float out[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float f = X::template update<POST>(td[k + 1], a[k], cf[k], a[k + 1], cf[k + 1], dv[k], c);
                out[k] = (rowin && cin[k + 1]) ? f : F[k];             // ghost rows / columns pass through
            }
            *reinterpret_cast<float4*>(Fo + (size_t)i * P) = make_float4(out[0], out[1], out[2], out[3]);
        }
      }
    }
    }
    ring.drain();
    wq_leave(wq, lane, gridDim.x * kFctYWarps);
}


}  // namespace vof
