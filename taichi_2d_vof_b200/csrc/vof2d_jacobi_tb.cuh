// Temporally blocked Jacobi: T sweeps of solve_p_jacobi (2dvof.py:236-266) per pass over HBM.
//
// Warp-autonomous register pipeline -- no shared memory, no block barriers:
//   * a warp owns a strip of 128 columns (4 per lane, one LDG.128/STG.128 per lane per row) and
//     marches down a chunk of rows;
//   * sweep s of row r needs sweep s-1 of rows r-1, r, r+1: the T sweeps run as a software
//     pipeline skewed by one row per sweep, each sweep keeping its last three rows in registers
//     (slots rotate with period 3, so the row loop is unrolled by 3 and rows are never moved);
//   * j-neighbours inside a row come from the adjacent lanes with two shuffles per row per sweep;
//   * a strip loses one column per sweep at each edge, so strips overlap by 2 x 8 columns
//     (112 of 128 columns are stored; 120 when T <= 4) and chunks by 2 x T rows -- redundant work instead of
//     any inter-warp communication.
// Walls are zeroed coefficients (2dvof.py:258-262).  Strips that touch no j-wall use literal
// constants; the two edge strips carry per-lane coefficients (EDGE = true).  Rows that touch an
// i-wall (2 of nx) and ghost rows go through one out-of-line general routine.  Persistent warps
// pull (strip, chunk) items from a queue, because the cost of an item is data dependent.
// Results are bit-identical to T single sweeps of k_jacobi: same expressions, same order; the
// division by the (constant) diagonal uses its correctly rounded reciprocal plus one exact
// FMA residual correction, which returns the correctly rounded quotient (verified against
// __fdiv_rn for EVERY fp32 numerator up to 2^100 when the context is created; numerators below
// 2^-100 run the same scheme in fp64).
#pragma once
#include "vof_common.cuh"

namespace vof {

constexpr float kTinyNumerator = 7.8886090522101181e-31f;   // 2^-100

// A constant divisor with its reciprocals.  Fast path (fp32): q = RN(t r), q' = RN(q + (t - q b) r) is the
// correctly rounded quotient [Markstein] unless the residual underflows (|t| < 2^-100).  Slow path (tiny t,
// usually a subnormal quotient): the same two-step scheme in fp64 gives RN64(t / b) -- nothing underflows
// there -- and RN32(RN64(x)) == RN32(x) for a quotient because 53 >= 2*24 + 2 (double rounding is innocuous;
// holds for subnormal fp32 results too).  Both paths are checked against __fdiv_rn for every fp32 t.
struct ConstDiv {
    float b, r;        // divisor, RN32(1/b)
    double bd, rd;     // divisor, RN64(1/b)
};
__device__ __forceinline__ float div_slow(float t, const ConstDiv& d) {
    const double td = (double)t;
    const double q = td * d.rd;
    const double rem = __fma_rn(-q, d.bd, td);
    return __double2float_rn(__fma_rn(rem, d.rd, q));
}

// correctly rounded t / b from r = RN(1/b): q = RN(t r); q' = RN(q + (t - q b) r)   [Markstein]
// valid for t = 0 (sign included) and 2^-100 <= |t| <= 2^100; see div_needs_ieee
__device__ __forceinline__ float div_by_const_core(float t, float b, float r) {
    const float q = t * r;
    const float rem = __fmaf_rn(-q, b, t);
    return __fmaf_rn(rem, r, q);
}
__device__ __forceinline__ bool div_needs_ieee(float t) { return fabsf(t) < kTinyNumerator && t != 0.0f; }
__device__ __forceinline__ float div_by_const(float t, const ConstDiv& d) {
    float q = div_by_const_core(t, d.b, d.r);
    if (div_needs_ieee(t)) q = div_slow(t, d);   // the fp32 residual would underflow
    return q;
}

struct JacTB {
    float cx, cy;          // dxi^2, dyi^2 (2dvof.py:258-261)
    float ap[2][2];        // -(ae+aw+an+as) by [row touches an i-wall][column touches a j-wall]
    ConstDiv dv[2];        // division by ap[0][jc]
    int fast_div_ok;       // reciprocal division validated for ap[0][0] and ap[0][1]
    int bare_div_ok;       // ... and the bare three-operation form (no sub-normal fix-up) proven exact for EVERY fp32 numerator
};

// exhaustive check of div_by_const against IEEE division: every fp32 bit pattern.  bare: the three operations alone, with no
// fp64 path for tiny numerators -- for most divisors they are exact there too (the residual t - q b is a multiple of the
// smallest sub-normal and therefore always representable; what can fail is the final rounding of a sub-normal quotient
// next to a tie), and a kernel that has this proof for its divisor needs neither the test nor the fix-up.
static __global__ void k_check_div_by_const(ConstDiv d, unsigned long long* mismatches, int bare) {
    const unsigned long long n = 1ull << 32;
    unsigned long long bad = 0;
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < n;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        const float t = __uint_as_float((unsigned)k);
        const float at = fabsf(t);
        if (!(at <= 1.2676506002282294e30f)) continue;       // 2^100; also skips NaN
        const float a = bare ? div_by_const_core(t, d.b, d.r) : div_by_const(t, d), e = __fdiv_rn(t, d.b);
        bad += (__float_as_uint(a) != __float_as_uint(e));
    }
    if (bad) atomicAdd(mismatches, bad);
}

template <int T>
struct JacPipe {
    float st[T][3][4];   // st[s][slot][col]: sweep-s values (s = 0: the input field) of the last three rows
    float dl[T][4];      // dl[s-1] = rhs of the row sweep s computes this step
};

// per-lane wall coefficients of an edge strip (unused by interior strips)
struct JacEdge {
    float an[4], as[4];
    int cls[4];            // 1: the column touches a j-wall (diagonal ap[0][1])
    bool colin[4];
};

// rows that touch an i-wall, ghost rows and rows outside the field: the reference's expression with
// every coefficient selected, IEEE division, pass-through for non-interior cells.  Out of line: rare.
static __device__ __noinline__ float4 jac_general_row(float4 up, float4 md, float4 dn, float left, float right, float4 b,
                                               int gi, int jl, int nx, int ny, JacTB jc) {
    const bool rowin = gi >= 1 && gi <= nx;
    const float ae = (gi != nx) ? jc.cx : 0.0f;
    const float aw = (gi != 1) ? jc.cx : 0.0f;
    const int ic = (gi == 1 || gi == nx) ? 1 : 0;
    const float u4[4] = {up.x, up.y, up.z, up.w}, m4[4] = {md.x, md.y, md.z, md.w}, d4[4] = {dn.x, dn.y, dn.z, dn.w};
    const float b4[4] = {b.x, b.y, b.z, b.w};
    float out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = jl + k;
        const float pn = k < 3 ? m4[k + 1] : right;
        const float ps = k > 0 ? m4[k - 1] : left;
        const float an = (j != ny) ? jc.cy : 0.0f;
        const float as = (j != 1) ? jc.cy : 0.0f;
        const float ap = jc.ap[ic][(j == 1 || j == ny) ? 1 : 0];
        float t = b4[k] - ae * u4[k];
        t = t - aw * d4[k];
        t = t - an * pn;
        t = t - as * ps;
        const float q = div_nz(t, ap);
        out[k] = (rowin && j >= 1 && j <= ny) ? q : m4[k];
    }
    return make_float4(out[0], out[1], out[2], out[3]);
}

// one row step: input row R (stage-0 values `pin`, and rhs of row R-1 in `rin`) enters the pipeline,
// sweep s produces row R - s, the last sweep's row R - T is stored.
template <int T, int PH, bool EDGE, bool FAST_DIV, bool WALLROWS>
__device__ __forceinline__ void jac_step(JacPipe<T>& S, const JacEdge& E, const float4 pin, const float4 rin, const int R,
                                         const Grid& g, const JacTB& jc, const int jl, float* __restrict__ pout,
                                         const int ra, const int rb, const bool store_lane) {
    constexpr int NEW = PH, MID = (PH + 2) % 3, OLD = (PH + 1) % 3;
    S.st[0][NEW][0] = pin.x; S.st[0][NEW][1] = pin.y; S.st[0][NEW][2] = pin.z; S.st[0][NEW][3] = pin.w;
#pragma unroll
    for (int s = T - 1; s >= 1; --s) {
#pragma unroll
        for (int k = 0; k < 4; ++k) S.dl[s][k] = S.dl[s - 1][k];
    }
    S.dl[0][0] = rin.x; S.dl[0][1] = rin.y; S.dl[0][2] = rin.z; S.dl[0][3] = rin.w;
#pragma unroll
    for (int s = 1; s <= T; ++s) {
        const int r = R - s;                 // local row this sweep produces
        const int gi = g.gi0 + r;
        const float* up = S.st[s - 1][NEW];  // row r+1  -> p[i+1, j]  (ae)
        const float* md = S.st[s - 1][MID];  // row r
        const float* dn = S.st[s - 1][OLD];  // row r-1  -> p[i-1, j]  (aw)
        const float* bb = S.dl[s - 1];
        const float left = __shfl_up_sync(0xffffffffu, md[3], 1);     // p[i, j-1] of column 0
        const float right = __shfl_down_sync(0xffffffffu, md[0], 1);  // p[i, j+1] of column 3
        float out[4];
        // items whose rows stay away from the i-walls (WALLROWS = false: almost all of them) carry no fallback code at
        // all, which keeps the hot loop body small enough for the instruction cache
        if (!WALLROWS || (gi >= 2 && gi <= g.nx - 1)) {     // ae = aw = cx (warp-uniform test)
            float t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float pn = k < 3 ? md[k + 1] : right;
                const float ps = k > 0 ? md[k - 1] : left;
                t[k] = bb[k] - jc.cx * up[k];
                t[k] = t[k] - jc.cx * dn[k];
                t[k] = t[k] - (EDGE ? E.an[k] : jc.cy) * pn;
                t[k] = t[k] - (EDGE ? E.as[k] : jc.cy) * ps;
            }
            if (FAST_DIV) {
                bool slow = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const ConstDiv& d = jc.dv[EDGE ? E.cls[k] : 0];
                    out[k] = div_by_const_core(t[k], d.b, d.r);
                }
                if (!jc.bare_div_ok) {          // bare_div_ok: the three operations are proven exact for every numerator
#pragma unroll
                    for (int k = 0; k < 4; ++k) slow = slow || div_needs_ieee(t[k]);
                }
                if (slow) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (div_needs_ieee(t[k])) out[k] = div_slow(t[k], jc.dv[EDGE ? E.cls[k] : 0]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) out[k] = t[k] / jc.dv[EDGE ? E.cls[k] : 0].b;
            }
            if (EDGE) {
#pragma unroll
                for (int k = 0; k < 4; ++k) out[k] = E.colin[k] ? out[k] : md[k];   // ghost / pad columns pass through
            }
        } else {
            const float4 o = jac_general_row(make_float4(up[0], up[1], up[2], up[3]), make_float4(md[0], md[1], md[2], md[3]),
                                             make_float4(dn[0], dn[1], dn[2], dn[3]), left, right,
                                             make_float4(bb[0], bb[1], bb[2], bb[3]), gi, jl, g.nx, g.ny, jc);
            out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
        }
        if (s < T) {
#pragma unroll
            for (int k = 0; k < 4; ++k) S.st[s][NEW][k] = out[k];
        } else if (store_lane && r >= ra && r <= rb) {
            *reinterpret_cast<float4*>(pout + (size_t)r * g.pitch + jl) = make_float4(out[0], out[1], out[2], out[3]);
        }
    }
}

constexpr int kJacStripCols = 128;      // columns a warp loads
// columns given up at each strip edge: >= T (a strip loses one column per sweep), a multiple of 4 (float4 lanes)
template <int T> struct JacStrip {
    static constexpr int margin = T <= 4 ? 4 : 8;
    static constexpr int valid = kJacStripCols - 2 * margin;
};
constexpr int kJacWarpsPerBlock = 4;

struct JacSched {          // work item -> (strip, chunk); warps pull items from a global counter
    int nstrips, nchunks, rpc;
    unsigned int* counter; // zeroed by the host before the launch
};

template <int T, bool EDGE, bool FAST_DIV, bool WALLROWS>
__device__ __forceinline__ void jac_run(const Grid& g, const JacTB& jc, const float* __restrict__ p, float* __restrict__ pout,
                                        const float* __restrict__ rhs, const int ra, const int rb, const int jstrip,
                                        const int lane) {
    const int jl = jstrip + 4 * lane;
    const bool active = jl <= g.ny + 1;                                // lanes past the right ghost column idle
    const bool store_lane = active && lane >= JacStrip<T>::margin / 4 && lane < 32 - JacStrip<T>::margin / 4;
    const int P = g.pitch, last = g.nrows - 1;
    const float* pc = p + jl;
    const float* rc = rhs + jl;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ldp = [&](int R) { return active ? *reinterpret_cast<const float4*>(pc + (size_t)min(max(R, 0), last) * P) : zero4; };
    auto ldr = [&](int R) { return active ? *reinterpret_cast<const float4*>(rc + (size_t)min(max(R, 0), last) * P) : zero4; };

    JacEdge E;
    if (EDGE) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = jl + k;
            E.an[k] = (j != g.ny) ? jc.cy : 0.0f;
            E.as[k] = (j != 1) ? jc.cy : 0.0f;
            E.cls[k] = (j == 1 || j == g.ny) ? 1 : 0;
            E.colin[k] = j >= 1 && j <= g.ny;
        }
    }
    JacPipe<T> S;
#pragma unroll
    for (int s = 0; s < T; ++s) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { S.st[s][0][k] = S.st[s][1][k] = S.st[s][2][k] = 0.0f; S.dl[s][k] = 0.0f; }
    }
    // rows enter in groups of three (one slot rotation); the next group is in flight while this one computes
    int R = ra - T;
    float4 pa0 = ldp(R), pa1 = ldp(R + 1), pa2 = ldp(R + 2);
    float4 qa0 = ldr(R - 1), qa1 = ldr(R), qa2 = ldr(R + 1);
    for (; R <= rb + T; R += 3) {
        const float4 pb0 = ldp(R + 3), pb1 = ldp(R + 4), pb2 = ldp(R + 5);
        const float4 qb0 = ldr(R + 2), qb1 = ldr(R + 3), qb2 = ldr(R + 4);
        jac_step<T, 0, EDGE, FAST_DIV, WALLROWS>(S, E, pa0, qa0, R, g, jc, jl, pout, ra, rb, store_lane);
        jac_step<T, 1, EDGE, FAST_DIV, WALLROWS>(S, E, pa1, qa1, R + 1, g, jc, jl, pout, ra, rb, store_lane);
        jac_step<T, 2, EDGE, FAST_DIV, WALLROWS>(S, E, pa2, qa2, R + 2, g, jc, jl, pout, ra, rb, store_lane);
        pa0 = pb0; pa1 = pb1; pa2 = pb2; qa0 = qb0; qa1 = qb1; qa2 = qb2;
    }
}

template <int T, bool FAST_DIV>
__global__ void __launch_bounds__(32 * kJacWarpsPerBlock)
k_jacobi_tb(Grid g, JacTB jc, JacSched sc, const float* __restrict__ p, float* __restrict__ pout,
            const float* __restrict__ rhs, int r0, int r1) {
    static_assert(T >= 1 && T <= JacStrip<T>::margin, "strip margin must cover the sweeps of one pass");
    const int lane = threadIdx.x & 31;
    const int nitems = sc.nstrips * sc.nchunks;
    // persistent warps + work queue: the cost of an item is data dependent (tiny numerators take the fp64
    // division) and edge strips are slower, so a static one-item-per-warp wave would wait for the slowest warp
    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(sc.counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) break;
        const int strip = item % sc.nstrips, chunk = item / sc.nstrips;
        const int ra = r0 + chunk * sc.rpc;
        const int rb = min(r1, ra + sc.rpc - 1);
        const int jstrip = 1 - JacStrip<T>::margin + strip * JacStrip<T>::valid;   // == 1 (mod 4): float4-aligned
        const bool strip_interior = jstrip >= 2 && jstrip + kJacStripCols - 1 <= g.ny - 1;
        // every row the pipeline touches (ra - T .. rb + T, minus the sweeps' skew) strictly inside the i-walls?
        const bool wallrows = g.gi0 + ra - T - T < 2 || g.gi0 + rb + T > g.nx - 1;
        if (strip_interior) {
            if (wallrows) jac_run<T, false, FAST_DIV, true>(g, jc, p, pout, rhs, ra, rb, jstrip, lane);
            else jac_run<T, false, FAST_DIV, false>(g, jc, p, pout, rhs, ra, rb, jstrip, lane);
        } else {
            if (wallrows) jac_run<T, true, FAST_DIV, true>(g, jc, p, pout, rhs, ra, rb, jstrip, lane);
            else jac_run<T, true, FAST_DIV, false>(g, jc, p, pout, rhs, ra, rb, jstrip, lane);
        }
    }
}

// ghost cells of the field pass through a sweep unchanged; the pipeline above only stores columns of
// output lanes, so the API entry (not the fused step, where set_BC rewrites them) copies the frame.
static __global__ void __launch_bounds__(128)
k_copy_frame(Grid g, const float* __restrict__ src, float* __restrict__ dst, int row_a, int row_b, int copy_lo_row,
             int copy_hi_row) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    const int nrow = row_b - row_a + 1, ncol = g.ny + 2;
    if (t < nrow) {
        const size_t o = (size_t)(row_a + t) * g.pitch;
        dst[o] = src[o];
        dst[o + g.ny + 1] = src[o + g.ny + 1];
        return;
    }
    int k = t - nrow;
    if (copy_lo_row >= 0) { if (k < ncol) { dst[(size_t)copy_lo_row * g.pitch + k] = src[(size_t)copy_lo_row * g.pitch + k]; return; } k -= ncol; }
    if (copy_hi_row >= 0 && k < ncol) dst[(size_t)copy_hi_row * g.pitch + k] = src[(size_t)copy_hi_row * g.pitch + k];
}

}  // namespace vof
