// Youngs normal + curvature (2dvof.py:283-309), instruction-lean version: a warp owns a strip of 128
// columns (4 per lane, lanes 1..30 store = 120 columns) and marches up a chunk of rows.  Per new row of F
// it forms the 5 corner gradients of the lane once (the four mx1..mx4 / my1..my4 of a cell are the same
// expression at its four corners), the unit normals of the previous row, and kappa of the row before that;
// everything rolls through registers, j-neighbours come by shuffle.  R F, W kappa = 8 B/cell.
#pragma once
#include "vof2d_stream.cuh"
#include "vof_common.cuh"

namespace vof {

constexpr int kKapWarps = 4;
constexpr int kKapValid = 120;
constexpr int kKapPrefetch = 1;

// ADAPT: warp-uniform bulk rows.  Where two consecutive rows of the strip hold one and the same value (F is exactly
// 0 or 1 away from the interface) every corner gradient between them is c * ((f + f) - f - f) = c * (+0), and where
// two consecutive corner rows are such zeros the normal is that zero itself (|m| < 1e-10: not normalised), so the
// 40 + 48 flops, two divisions and the square root per lane and row are skipped -- same bits, sign of zero included.
template <bool ADAPT>
__global__ void __launch_bounds__(32 * kKapWarps)
k_kappa4(Grid g, Consts c, const float* __restrict__ F, float* __restrict__ kappa, int r0, int r1, int rows_per_chunk,
         int nstrips) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kKapWarps + (threadIdx.x >> 5);
    const int strip = w % nstrips, chunk = w / nstrips;
    const int ia = r0 + chunk * rows_per_chunk;
    if (ia > r1) return;
    const int ib = min(r1, ia + rows_per_chunk - 1);
    const int jl = 1 - 4 + kKapValid * strip + 4 * lane;       // == 1 (mod 4)
    const bool active = jl <= g.ny + 1;
    const bool store_lane = active && lane >= 1 && lane <= 30 && jl <= g.ny;
    const int P = g.pitch, last = g.nrows - 1;
    const float* Fc = F + jl;
    bool colin[4], rel[4];        // interior columns; columns that exist at all (ghosts included, padding not)
#pragma unroll
    for (int q = 0; q < 4; ++q) { colin[q] = jl + q >= 1 && jl + q <= g.ny; rel[q] = active && jl + q >= 0 && jl + q <= g.ny + 1; }
    const float zx = c.m1_2dx * 0.0f, zy = c.m1_2dy * 0.0f;    // the zero a flat corner evaluates to
    bool flat_p = false, zc_p = false;                           // previous row flat / previous corner row all zero
    float c_p = 0.0f;

    float Fp[6];                  // previous row of F: columns jl-1 .. jl+4
    float Gp[5], Hp[5];           // previous corner row: corners jl-1 .. jl+3
    float mx_a[4], mx_b[4];       // unit normal x of rows n-2, n-1 (n = row whose normals are formed now)
    float my_b[6];                // unit normal y of row n-1: columns jl-1 .. jl+4
#pragma unroll
    for (int q = 0; q < 6; ++q) { Fp[q] = 0.f; my_b[q] = 0.f; }
#pragma unroll
    for (int q = 0; q < 5; ++q) { Gp[q] = 0.f; Hp[q] = 0.f; }
#pragma unroll
    for (int q = 0; q < 4; ++q) { mx_a[q] = 0.f; mx_b[q] = 0.f; }

    // rows are requested kKapPrefetch iterations ahead: with ~28 resident warps per SM one row in flight per
    // warp is only ~14 KB per SM, far below the ~45 KB that HBM latency x bandwidth needs
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ldF = [&](int r) { return (active && r <= ib + 2) ? *reinterpret_cast<const float4*>(Fc + (size_t)min(max(r, 0), last) * P) : zero4; };
    float4 pf[kKapPrefetch];
#pragma unroll
    for (int d = 0; d < kKapPrefetch; ++d) pf[d] = ldF(ia - 2 + d);
    for (int rbase = ia - 2; rbase <= ib + 2; rbase += kKapPrefetch) {
#pragma unroll
      for (int d = 0; d < kKapPrefetch; ++d) {
        const int r = rbase + d;
        if (r > ib + 2) break;
        const float4 f4 = pf[d];
        pf[d] = ldF(r + kKapPrefetch);
        float Fr[6];
        Fr[1] = f4.x; Fr[2] = f4.y; Fr[3] = f4.z; Fr[4] = f4.w;
        Fr[0] = __shfl_up_sync(0xffffffffu, f4.w, 1);
        Fr[5] = __shfl_down_sync(0xffffffffu, f4.x, 1);
        // corner row r-1 (between rows r-1 and r): corner q sits between columns jl-1+q and jl+q
        float G[5], H[5];
        bool zc = false;                                       // corner row r-1 is all zero across the strip
        if (ADAPT) {
            const float c_r = __shfl_sync(0xffffffffu, f4.x, 1);               // lane 1's first column always exists
            bool e = true;
#pragma unroll
            for (int q = 0; q < 4; ++q) e = e && (!rel[q] || Fr[q + 1] == c_r);
            const bool flat = __all_sync(0xffffffffu, e);
            zc = flat && flat_p && c_r == c_p && fabsf(c_r) < 1e30f;       // inf - inf must stay NaN
            flat_p = flat; c_p = c_r;
        }
        if (ADAPT && zc) {
#pragma unroll
            for (int q = 0; q < 5; ++q) { G[q] = zx; H[q] = zy; }
        } else {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float f00 = Fp[q], f01 = Fp[q + 1], f10 = Fr[q], f11 = Fr[q + 1];
                G[q] = c.m1_2dx * (((f11 + f10) - f01) - f00);     // 2dvof.py:287 and its aliases 289, 291, 293
                H[q] = c.m1_2dy * (((f11 - f10) + f01) - f00);     // 2dvof.py:288 and 290, 292, 294
            }
        }
        // unit normals of row n = r-1 (corner rows n-1 = previous, n = this one)
        const int gn = g.gi0 + r - 1;
        const bool rowin = gn >= 1 && gn <= g.nx;
        float mx_n[4], my_n[6];
        if (ADAPT && zc && zc_p) {                             // (((z + z) + z) + z) / 4 = z, kept unnormalised
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool in = rowin && colin[q];
                mx_n[q] = in ? zx : 0.0f; my_n[q + 1] = in ? zy : 0.0f;
            }
        } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float mx = 0.0f, my = 0.0f;
            if (rowin && colin[q]) {
                // mx1 = G(i,j), mx2 = G(i,j-1), mx3 = G(i-1,j-1), mx4 = G(i-1,j)      (2dvof.py:296-297)
                const float mxs = (((G[q + 1] + G[q]) + Gp[q]) + Gp[q + 1]) / 4.0f;
                const float mys = (((H[q + 1] + H[q]) + Hp[q]) + Hp[q + 1]) / 4.0f;
                mx = mxs; my = mys;
                if (!(fabsf(mxs) < 1e-10f && fabsf(mys) < 1e-10f)) {   // 2dvof.py:300-306
                    const float mag = sqrtf(mxs * mxs + mys * mys);
                    mx = div_nz(mxs, mag); my = div_nz(mys, mag);   // a flat interface has one zero component
                }
            }
            mx_n[q] = mx; my_n[q + 1] = my;
        }
        }
        zc_p = zc;
        my_n[0] = __shfl_up_sync(0xffffffffu, my_n[4], 1);
        my_n[5] = __shfl_down_sync(0xffffffffu, my_n[1], 1);
        // kappa of row k = r-2: mx of rows k-1 (mx_a) and k+1 (mx_n), my of row k (my_b)        (2dvof.py:308-309)
        const int k = r - 2;
        if (store_lane && k >= ia && k <= ib) {
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                o[q] = -(c.i_dx_2 * (mx_n[q] - mx_a[q]) + c.i_dy_2 * (my_b[q + 2] - my_b[q]));
            float* dst = kappa + (size_t)k * P + jl;
            if (jl + 3 <= g.ny) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (colin[q]) dst[q] = o[q];
            }
        }
        // roll
#pragma unroll
        for (int q = 0; q < 6; ++q) { Fp[q] = Fr[q]; my_b[q] = my_n[q]; }
#pragma unroll
        for (int q = 0; q < 5; ++q) { Gp[q] = G[q]; Hp[q] = H[q]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) { mx_a[q] = mx_b[q]; mx_b[q] = mx_n[q]; }
      }
    }
}

// Second generation -- ADAPT: warp-uniform bulk rows.  Where two consecutive rows of the strip hold one and the same value (F is exactly
// 0 or 1 away from the interface) every corner gradient between them is c * ((f + f) - f - f) = c * (+0), and where
// two consecutive corner rows are such zeros the normal is that zero itself (|m| < 1e-10: not normalised), so the
// 40 + 48 flops, two divisions and the square root per lane and row are skipped -- same bits, sign of zero included.
// Rows stream through the per-lane cp.async ring and (strip, chunk) items come from the work queue (vof2d_stream.cuh).
constexpr int kKapSlots = 8;
__global__ void __launch_bounds__(32 * kKapWarps)
k_kappa5(Grid g, Consts c, WorkQueue wq, const float* __restrict__ F, float* __restrict__ kappa, int r0, int r1,
         int rows_per_chunk, int nstrips) {
    constexpr bool ADAPT = true;
    using Ring = RowRing<1, 4, kKapSlots, 32 * kKapWarps>;
    __shared__ __align__(16) unsigned char ring_mem[Ring::kBytes];
    const int lane = threadIdx.x & 31;
    Ring ring;
    ring.init(ring_mem, threadIdx.x);
    for (;;) {
    const int item = wq_claim(wq, lane);
    if (item >= wq.nitems) break;
    const int strip = item % nstrips, chunk = item / nstrips;
    const int ia = r0 + chunk * rows_per_chunk;
    const int ib = min(r1, ia + rows_per_chunk - 1);
    const int jl = 1 - 4 + kKapValid * strip + 4 * lane;       // == 1 (mod 4)
    const bool active = jl <= g.ny + 1;
    const bool store_lane = active && lane >= 1 && lane <= 30 && jl <= g.ny;
    const int P = g.pitch, last = g.nrows - 1;
    const float* Fc = F + jl;
    bool colin[4], rel[4];        // interior columns; columns that exist at all (ghosts included, padding not)
#pragma unroll
    for (int q = 0; q < 4; ++q) { colin[q] = jl + q >= 1 && jl + q <= g.ny; rel[q] = active && jl + q >= 0 && jl + q <= g.ny + 1; }
    const float zx = c.m1_2dx * 0.0f, zy = c.m1_2dy * 0.0f;    // the zero a flat corner evaluates to
    bool flat_p = false, zc_p = false;                           // previous row flat / previous corner row all zero
    float c_p = 0.0f;

    float Fp[6];                  // previous row of F: columns jl-1 .. jl+4
    float Gp[5], Hp[5];           // previous corner row: corners jl-1 .. jl+3
    float mx_a[4], mx_b[4];       // unit normal x of rows n-2, n-1 (n = row whose normals are formed now)
    float my_b[6];                // unit normal y of row n-1: columns jl-1 .. jl+4
#pragma unroll
    for (int q = 0; q < 6; ++q) { Fp[q] = 0.f; my_b[q] = 0.f; }
#pragma unroll
    for (int q = 0; q < 5; ++q) { Gp[q] = 0.f; Hp[q] = 0.f; }
#pragma unroll
    for (int q = 0; q < 4; ++q) { mx_a[q] = 0.f; mx_b[q] = 0.f; }

    // Steady flat state: once four consecutive corner rows are zero and the three rows of normals involved are interior
    // rows, everything in the pipeline is a per-lane constant (normals zx / zy inside the domain, +0 outside), so
    // kappa repeats itself row after row: the pipeline is left alone and re-created from those constants when the
    // flat run ends.
    float mxf[4], myf[6], o_flat[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) mxf[q] = colin[q] ? zx : 0.0f;
#pragma unroll
    for (int q = 0; q < 6; ++q) { const int j = jl - 1 + q; myf[q] = (j >= 1 && j <= g.ny) ? zy : 0.0f; }
#pragma unroll
    for (int q = 0; q < 4; ++q) o_flat[q] = -(c.i_dx_2 * (mxf[q] - mxf[q]) + c.i_dy_2 * (myf[q + 2] - myf[q]));
    int zrun = 0;                 // consecutive zero corner rows ending at corner row r-1 (saturates at 4)
    bool steady_p = false;

    const float* const src[1] = {Fc};
    ring.start(active, ia - 2, ib + 2, last, P, src);
    {
      for (int r = ia - 2; r <= ib + 2; ++r) {
        float X[1][4];
        ring.next(X, src);
        const float4 f4 = make_float4(X[0][0], X[0][1], X[0][2], X[0][3]);
        float Fr[6];
        Fr[1] = f4.x; Fr[2] = f4.y; Fr[3] = f4.z; Fr[4] = f4.w;
        bool zc = false;                                       // corner row r-1 is all zero across the strip
        const float c_before = c_p;                            // the value of row r-1 if that row was flat
        {
            const float c_r = __shfl_sync(0xffffffffu, f4.x, 1);               // lane 1's first column always exists
            bool e = true;
#pragma unroll
            for (int q = 0; q < 4; ++q) e = e && (!rel[q] || Fr[q + 1] == c_r);
            const bool flat = __all_sync(0xffffffffu, e);
            zc = flat && flat_p && c_r == c_p && fabsf(c_r) < 1e30f;       // inf - inf must stay NaN
            flat_p = flat; c_p = c_r;
        }
        zrun = zc ? min(zrun + 1, 4) : 0;
        {
            const int g3 = g.gi0 + r - 3;                       // rows r-3 .. r-1 carry the normals kappa(r-2) reads
            if (zrun >= 4 && g3 >= 1 && g3 + 2 <= g.nx) {
                const int k = r - 2;
                if (store_lane && k >= ia && k <= ib) {
                    float* dst = kappa + (size_t)k * P + jl;
                    if (jl + 3 <= g.ny) *reinterpret_cast<float4*>(dst) = make_float4(o_flat[0], o_flat[1], o_flat[2], o_flat[3]);
                    else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) if (colin[q]) dst[q] = o_flat[q];
                    }
                }
                steady_p = true; zc_p = true;
                continue;
            }
        }
        if (steady_p) {           // the flat run ended with row r-1 (all c_before): what the pipeline holds after it
            steady_p = false;
#pragma unroll
            for (int q = 0; q < 6; ++q) { Fp[q] = c_before; my_b[q] = myf[q]; }
#pragma unroll
            for (int q = 0; q < 5; ++q) { Gp[q] = zx; Hp[q] = zy; }
#pragma unroll
            for (int q = 0; q < 4; ++q) { mx_a[q] = mxf[q]; mx_b[q] = mxf[q]; }
        }
        Fr[0] = __shfl_up_sync(0xffffffffu, f4.w, 1);
        Fr[5] = __shfl_down_sync(0xffffffffu, f4.x, 1);
        // corner row r-1 (between rows r-1 and r): corner q sits between columns jl-1+q and jl+q
        float G[5], H[5];
        if (ADAPT && zc) {
#pragma unroll
            for (int q = 0; q < 5; ++q) { G[q] = zx; H[q] = zy; }
        } else {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float f00 = Fp[q], f01 = Fp[q + 1], f10 = Fr[q], f11 = Fr[q + 1];
                G[q] = c.m1_2dx * (((f11 + f10) - f01) - f00);     // 2dvof.py:287 and its aliases 289, 291, 293
                H[q] = c.m1_2dy * (((f11 - f10) + f01) - f00);     // 2dvof.py:288 and 290, 292, 294
            }
        }
        // unit normals of row n = r-1 (corner rows n-1 = previous, n = this one)
        const int gn = g.gi0 + r - 1;
        const bool rowin = gn >= 1 && gn <= g.nx;
        float mx_n[4], my_n[6];
        if (ADAPT && zc && zc_p) {                             // (((z + z) + z) + z) / 4 = z, kept unnormalised
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool in = rowin && colin[q];
                mx_n[q] = in ? zx : 0.0f; my_n[q + 1] = in ? zy : 0.0f;
            }
        } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float mx = 0.0f, my = 0.0f;
            if (rowin && colin[q]) {
                // mx1 = G(i,j), mx2 = G(i,j-1), mx3 = G(i-1,j-1), mx4 = G(i-1,j)      (2dvof.py:296-297)
                const float mxs = (((G[q + 1] + G[q]) + Gp[q]) + Gp[q + 1]) / 4.0f;
                const float mys = (((H[q + 1] + H[q]) + Hp[q]) + Hp[q + 1]) / 4.0f;
                mx = mxs; my = mys;
                if (!(fabsf(mxs) < 1e-10f && fabsf(mys) < 1e-10f)) {   // 2dvof.py:300-306
                    const float mag = sqrtf(mxs * mxs + mys * mys);
                    mx = div_nz(mxs, mag); my = div_nz(mys, mag);   // a flat interface has one zero component
                }
            }
            mx_n[q] = mx; my_n[q + 1] = my;
        }
        }
        zc_p = zc;
        my_n[0] = __shfl_up_sync(0xffffffffu, my_n[4], 1);
        my_n[5] = __shfl_down_sync(0xffffffffu, my_n[1], 1);
        // kappa of row k = r-2: mx of rows k-1 (mx_a) and k+1 (mx_n), my of row k (my_b)        (2dvof.py:308-309)
        const int k = r - 2;
        if (store_lane && k >= ia && k <= ib) {
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                o[q] = -(c.i_dx_2 * (mx_n[q] - mx_a[q]) + c.i_dy_2 * (my_b[q + 2] - my_b[q]));
            float* dst = kappa + (size_t)k * P + jl;
            if (jl + 3 <= g.ny) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (colin[q]) dst[q] = o[q];
            }
        }
        // roll
#pragma unroll
        for (int q = 0; q < 6; ++q) { Fp[q] = Fr[q]; my_b[q] = my_n[q]; }
#pragma unroll
        for (int q = 0; q < 5; ++q) { Gp[q] = G[q]; Hp[q] = H[q]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) { mx_a[q] = mx_b[q]; mx_b[q] = mx_n[q]; }
      }
    }
    }
    ring.drain();
    wq_leave(wq, lane, gridDim.x * kKapWarps);
}


}  // namespace vof
