// Temporally blocked Jacobi, third generation (round 2): the warp-autonomous register pipeline of
// vof2d_jacobi_tb.cuh with three changes aimed at what ncu showed that kernel to be bound by -- the fp32 pipe
// (13.2 FMA-pipe instructions per cell-update, 43 % busy) behind a 71 % issue rate, not HBM (31 %):
//
//  * Blackwell packed fp32 arithmetic.  sm_100 has two-lane fp32 instructions (PTX add/mul/fma.rn.f32x2, SASS FFMA2):
//    one issue slot does two cell-updates.  A lane's four columns are held as two register pairs and every
//    operation of the sweep is written as an explicit fma.rn.f32x2 whose result equals the reference's separately
//    rounded operation bit for bit: a * b = fma(a, b, -0), a - b = fma(b, -1, a).  (ptxas contracts mul.rn.f32x2 +
//    add.rn.f32x2 into one FFMA2 even under -fmad=false -- measured, profiles/r2_micro -- so the mul/add forms
//    are not used.)
//  * Products instead of values.  With square cells (dxi^2 == dyi^2 bitwise, every BASELINE config) the four
//    products ae p[i+1], aw p[i-1], an p[j+1], as p[j-1] of 2dvof.py:263 are the same number c * p seen from four
//    neighbours; the pipeline keeps c * p of each row (one multiply per produced value instead of four).  The
//    subtraction order b - up - dn - pn - ps is the reference's.  Walls are a property of the cell whose p is
//    multiplied: the product of a ghost cell is 0 * p (ae / aw / an / a_s = 0.0 of 2dvof.py:258-261 as seen from the
//    wall cell); edge strips and chunks that reach an i-wall run a general variant of the same pipeline (GEN, pk_step).
//    Grids with dx != dy run the second generation.
//  * No register traffic for operands in flight: p and rhs rows travel global -> shared with cp.async (LDGSTS,
//    per-lane private slots, wait_group is the only synchronisation) and the rhs of the last T rows is re-read
//    from its slot by every sweep (one LDS.128) instead of being shifted through 4 (T - 1) registers per row.
//
// Arithmetic per cell-update: 1 mul + 4 sub + 3 (Markstein division by the constant diagonal) = 8 fp32 operations
// = 4 FFMA2 issue slots per update (was 11 + 2.2 register moves on the same pipe).  Results are bit-identical to T
// single sweeps of k_jacobi (tests/test_parity_gpu.py::test_jacobi_temporal_blocking_equals_single_sweeps).
#pragma once
#include "vof2d_jacobi_tb.cuh"
#include "vof2d_stream.cuh"

namespace vof {

struct PkConsts {
    f32x2 c, r, nb, m1, nz;     // (c, c), (1/ap, 1/ap), (-ap, -ap), (-1, -1), (-0, -0)
};
// RN(a * b) in both halves: the exact product plus -0 rounds once, and x + (-0) = x for every x including +-0
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b, const PkConsts& k) { return fma2(a, b, k.nz); }
// RN(a - b) in both halves: b * (-1) is exact, the sum rounds once; signs of zero as in a - b
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b, const PkConsts& k) { return fma2(b, k.m1, a); }

constexpr int kPkWarps = 4;
constexpr int kPkBlocksPerSM = 5;   // 40 KB of shared memory each; caps the registers at 96 so that the edge variants cannot cost occupancy
constexpr int kPkAhead = 3;        // rows in flight (a row step of one warp is ~200 issue slots: 3 rows cover the DRAM latency)
constexpr int kPkPSlots = 4;       // p rows in shared memory: in flight + the one being read
constexpr int kPkRSlots = 8;       // rhs rows: in flight + the last T <= 5 rows every sweep re-reads.  Each row is stored
                                   // twice, 8 slots apart, so that the rows R-1 .. R-T sit at base - s * row without a wrap
constexpr int kPkRowBytes = 32 * kPkWarps * 16;                              // one row of one field for the whole block
constexpr int kPkSmem = (kPkPSlots + 2 * kPkRSlots) * kPkRowBytes;           // 40 KB per block

template <int T> struct JacPk {
    f32x2 st[T][3][2];   // st[s][slot][pair] = coefficient * (sweep-s value) of the last three rows, s = 0: the input field
};

// per-lane description of a general item (GEN): a lane's four columns jl .. jl + 3
struct PkLane {
    unsigned colin, cls;       // bit k: column k is interior (1 <= j <= ny) / touches a j-wall
};

// tiny non-zero numerator: |t| < 2^-100 and t != 0, as one unsigned compare on (bits << 1) - 1
__device__ __forceinline__ unsigned tiny_key(float t) { return (__float_as_uint(t) << 1) - 1u; }
constexpr unsigned kTinyKey = (0x0D800000u << 1) - 1u;      // key of 2^-100

// fp64 division by the diagonal of an interior row; wall = the column touches a j-wall (selects instead of an indexed
// kernel parameter, which would be copied to local memory)
template <bool EDGE>
__device__ __forceinline__ float pk_div_slow(float t, const JacTB& jc, unsigned wall) {
    ConstDiv d = jc.dv[0];
    if (EDGE && wall) { d.bd = jc.dv[1].bd; d.rd = jc.dv[1].rd; }
    return div_slow(t, d);
}

// the quotients of the tiny numerators among t, in fp64 (bulk variant; see div_slow)
__device__ __noinline__ float4 pk_fix_tiny(float4 t, float4 q, double bd, double rd) {
    ConstDiv d;
    d.b = 0.0f; d.r = 0.0f; d.bd = bd; d.rd = rd;
    if (div_needs_ieee(t.x)) q.x = div_slow(t.x, d);
    if (div_needs_ieee(t.y)) q.y = div_slow(t.y, d);
    if (div_needs_ieee(t.z)) q.z = div_slow(t.z, d);
    if (div_needs_ieee(t.w)) q.w = div_slow(t.w, d);
    return q;
}

// One row step of the packed pipeline: row R of p enters (its rhs row R-1 is already in the shared-memory ring),
// sweep s produces row R - s, the last sweep's row R - T is stored.
// EDGE = WALL = false: the bulk -- a strip that touches no j-wall, rows strictly inside the i-walls: literal constants only.
// EDGE: the strip may touch a j-wall; WALL (with EDGE): the rows may reach an i-wall.  Walls are a property of the cell whose p is multiplied: the
// product of a ghost cell is 0 * p (ae / aw / an / a_s = 0.0 of 2dvof.py:258-261 seen from the wall cell) and passes
// through every sweep unchanged (ghost cells are not swept), wall-adjacent columns divide by their own diagonal, the
// two wall rows go through the IEEE division with theirs.  Per-lane masks and selects instead of per-lane constants:
// these items are ~5 % of the work and must not cost the bulk loop registers.
// BARE: the context proved the three-operation division exact for every fp32 numerator of this divisor (create: all 2^32
// patterns, JacTB::bare_div_ok), so the sub-normal test and its fp64 fix-up are not compiled in.
template <int T, int PH, bool EDGE, bool WALL, bool BARE>
__device__ __forceinline__ void pk_step(JacPk<T>& S, const PkLane& L, const float4 pin, const int R, const unsigned rcur,
                                        const PkConsts& k, const JacTB& jc, const Grid& g, const int jl,
                                        float* __restrict__ pout, const int ra, const int rb, const bool store_lane) {
    constexpr int NEW = PH, MID = (PH + 2) % 3, OLD = (PH + 1) % 3;
    {
        f32x2 nA = mul2(pk2(pin.x, pin.y), k.c, k), nB = mul2(pk2(pin.z, pin.w), k.c, k);
        if (EDGE) {
            const int gi = g.gi0 + R;
            const unsigned in = (WALL && (gi < 1 || gi > g.nx)) ? 0u : L.colin;
            const f32x2 zA = mul2(pk2(pin.x, pin.y), 0ull, k), zB = mul2(pk2(pin.z, pin.w), 0ull, k);     // 0.0 * p
            float n0, n1, n2, n3, z0, z1, z2, z3;
            unpk2(nA, n0, n1); unpk2(nB, n2, n3);
            unpk2(zA, z0, z1); unpk2(zB, z2, z3);
            nA = pk2((in & 1u) ? n0 : z0, (in & 2u) ? n1 : z1);
            nB = pk2((in & 4u) ? n2 : z2, (in & 8u) ? n3 : z3);
        }
        S.st[0][NEW][0] = nA;
        S.st[0][NEW][1] = nB;
    }
    const int P = g.pitch;
#pragma unroll
    for (int s = 1; s <= T; ++s) {
        const int r = R - s;
        const int gi = g.gi0 + r;
        const f32x2 upA = S.st[s - 1][NEW][0], upB = S.st[s - 1][NEW][1];
        const f32x2 mdA = S.st[s - 1][MID][0], mdB = S.st[s - 1][MID][1];
        const f32x2 dnA = S.st[s - 1][OLD][0], dnB = S.st[s - 1][OLD][1];
        const float4 b = lds_f4(rcur - (unsigned)s * kPkRowBytes);      // rhs of row R - s (upper copy of its slot or the lower one)
        float m0, m1, m2, m3;
        unpk2(mdA, m0, m1);
        unpk2(mdB, m2, m3);
        const float left = __shfl_up_sync(0xffffffffu, m3, 1);      // coefficient * p[i, j-1] of column 0
        const float right = __shfl_down_sync(0xffffffffu, m0, 1);   // coefficient * p[i, j+1] of column 3
        const f32x2 mid = pk2(m1, m2);                              // pn of (0, 1) and ps of (2, 3)
        f32x2 tA = sub2(pk2(b.x, b.y), upA, k);                     // 2dvof.py:263, left to right
        f32x2 tB = sub2(pk2(b.z, b.w), upB, k);
        tA = sub2(tA, dnA, k);
        tB = sub2(tB, dnB, k);
        tA = sub2(tA, mid, k);
        tB = sub2(tB, pk2(m3, right), k);
        tA = sub2(tA, pk2(left, m0), k);
        tB = sub2(tB, mid, k);
        f32x2 qA, qB;
        if (WALL && (gi == 1 || gi == g.nx)) {      // a wall row (warp-uniform, 2 rows of nx): its own diagonal, IEEE division
            float t4[4], q4[4];
            unpk2(tA, t4[0], t4[1]);
            unpk2(tB, t4[2], t4[3]);
#pragma unroll
            for (int c = 0; c < 4; ++c) q4[c] = div_nz(t4[c], ((L.cls >> c) & 1u) ? jc.ap[1][1] : jc.ap[1][0]);
            qA = pk2(q4[0], q4[1]);
            qB = pk2(q4[2], q4[3]);
        } else {
            // t / ap, correctly rounded: q = RN(t r), q' = RN(q + (t - q ap) r)   [Markstein]
            f32x2 rA = k.r, rB = k.r, nbA = k.nb, nbB = k.nb;
            if (EDGE) {
                const float r0 = jc.dv[0].r, r1 = jc.dv[1].r, b0 = -jc.dv[0].b, b1 = -jc.dv[1].b;
                rA = pk2((L.cls & 1u) ? r1 : r0, (L.cls & 2u) ? r1 : r0);
                rB = pk2((L.cls & 4u) ? r1 : r0, (L.cls & 8u) ? r1 : r0);
                nbA = pk2((L.cls & 1u) ? b1 : b0, (L.cls & 2u) ? b1 : b0);
                nbB = pk2((L.cls & 4u) ? b1 : b0, (L.cls & 8u) ? b1 : b0);
            }
            qA = mul2(tA, rA, k);
            qB = mul2(tB, rB, k);
            const f32x2 eA = fma2(qA, nbA, tA), eB = fma2(qB, nbB, tB);
            qA = fma2(eA, rA, qA);
            qB = fma2(eB, rB, qB);
            float t0, t1, t2, t3;
            unpk2(tA, t0, t1);
            unpk2(tB, t2, t3);
            const unsigned key = BARE ? 0xffffffffu : min(min(tiny_key(t0), tiny_key(t1)), min(tiny_key(t2), tiny_key(t3)));
            if (!BARE && key < kTinyKey) {  // the fp32 residual would underflow: same scheme in fp64 (rare: the pressure front)
                float q0, q1, q2, q3;
                unpk2(qA, q0, q1);
                unpk2(qB, q2, q3);
                if (!EDGE) {                // out of line: keeps the bulk loop small (instruction fetch is a visible stall)
                    const float4 f = pk_fix_tiny(make_float4(t0, t1, t2, t3), make_float4(q0, q1, q2, q3), jc.dv[0].bd, jc.dv[0].rd);
                    q0 = f.x; q1 = f.y; q2 = f.z; q3 = f.w;
                } else {
                    if (div_needs_ieee(t0)) q0 = pk_div_slow<EDGE>(t0, jc, L.cls & 1u);
                    if (div_needs_ieee(t1)) q1 = pk_div_slow<EDGE>(t1, jc, L.cls & 2u);
                    if (div_needs_ieee(t2)) q2 = pk_div_slow<EDGE>(t2, jc, L.cls & 4u);
                    if (div_needs_ieee(t3)) q3 = pk_div_slow<EDGE>(t3, jc, L.cls & 8u);
                }
                qA = pk2(q0, q1);
                qB = pk2(q2, q3);
            }
        }
        if (s < T) {
            f32x2 nA = mul2(qA, k.c, k), nB = mul2(qB, k.c, k);
            if (EDGE) {                     // ghost rows, ghost and pad columns are not swept: their product passes through
                const unsigned in = (WALL && (gi < 1 || gi > g.nx)) ? 0u : L.colin;
                float n0, n1, n2, n3;
                unpk2(nA, n0, n1);
                unpk2(nB, n2, n3);
                nA = pk2((in & 1u) ? n0 : m0, (in & 2u) ? n1 : m1);
                nB = pk2((in & 4u) ? n2 : m2, (in & 8u) ? n3 : m3);
            }
            S.st[s][NEW][0] = nA;
            S.st[s][NEW][1] = nB;
        } else if (store_lane && r >= ra && r <= rb) {      // ra .. rb are interior rows
            float q0, q1, q2, q3;
            unpk2(qA, q0, q1);
            unpk2(qB, q2, q3);
            float* dst = pout + (size_t)r * P + jl;
            if (!EDGE || (L.colin & 15u) == 15u) {
                *reinterpret_cast<float4*>(dst) = make_float4(q0, q1, q2, q3);
            } else {                        // a lane that straddles the right wall: interior columns only
                if (L.colin & 1u) dst[0] = q0;
                if (L.colin & 2u) dst[1] = q1;
                if (L.colin & 4u) dst[2] = q2;
                if (L.colin & 8u) dst[3] = q3;
            }
        }
    }
}

// Tolerance mode of the bulk (VOF_OPT_FAST_MATH, opt-in, NOT bit-exact): the same pipeline on raw values with the update
// written the way a compiler with fast_math would: p' = (b - c (pE + pW + pN + pS)) * (1 / ap) -- three packed adds, one fma,
// one multiply by the reciprocal: 5 operations per cell-update instead of 8, no sub-normal fix-up, no per-sweep branch.
// Within ~1 ulp per sweep of the exact path; the parity test holds it to the north-star tolerances (1e-5 after one
// step, 1e-3 after 100).  Edge strips and wall rows keep the exact general variant.
template <int T, int PH>
__device__ __forceinline__ void pk_step_fast(JacPk<T>& S, const float4 pin, const int R, const unsigned rcur, const f32x2 negc,
                                             const f32x2 rinv, const int P, const int jl, float* __restrict__ pout, const int ra,
                                             const int rb, const bool store_lane) {
    constexpr int NEW = PH, MID = (PH + 2) % 3, OLD = (PH + 1) % 3;
    S.st[0][NEW][0] = pk2(pin.x, pin.y);
    S.st[0][NEW][1] = pk2(pin.z, pin.w);
#pragma unroll
    for (int s = 1; s <= T; ++s) {
        const int r = R - s;
        const f32x2 mdA = S.st[s - 1][MID][0], mdB = S.st[s - 1][MID][1];
        const float4 b = lds_f4(rcur - (unsigned)s * kPkRowBytes);
        float m0, m1, m2, m3;
        unpk2(mdA, m0, m1);
        unpk2(mdB, m2, m3);
        const float left = __shfl_up_sync(0xffffffffu, m3, 1), right = __shfl_down_sync(0xffffffffu, m0, 1);
        const f32x2 mid = pk2(m1, m2);
        f32x2 sA = loose_add2(S.st[s - 1][NEW][0], S.st[s - 1][OLD][0]), sB = loose_add2(S.st[s - 1][NEW][1], S.st[s - 1][OLD][1]);
        sA = loose_add2(sA, mid);
        sB = loose_add2(sB, pk2(m3, right));
        sA = loose_add2(sA, pk2(left, m0));
        sB = loose_add2(sB, mid);
        const f32x2 qA = loose_mul2(fma2(sA, negc, pk2(b.x, b.y)), rinv), qB = loose_mul2(fma2(sB, negc, pk2(b.z, b.w)), rinv);
        if (s < T) {
            S.st[s][NEW][0] = qA;
            S.st[s][NEW][1] = qB;
        } else if (store_lane && r >= ra && r <= rb) {
            float q0, q1, q2, q3;
            unpk2(qA, q0, q1);
            unpk2(qB, q2, q3);
            *reinterpret_cast<float4*>(pout + (size_t)r * P + jl) = make_float4(q0, q1, q2, q3);
        }
    }
}

// one (strip, rows) item
template <int T, bool EDGE, bool WALL, bool FAST, bool BARE>
__device__ __forceinline__ void pk_run_impl(const Grid& g, const JacTB& jc, const float* __restrict__ p, float* __restrict__ pout,
                                            const float* __restrict__ rhs, const int ra, const int rb, const int jstrip, const int lane,
                                            const unsigned pbase, const unsigned rbase) {
    const int jl = jstrip + 4 * lane;
    const bool active = jl <= g.ny + 1;          // lanes past the right ghost column re-read the strip's first columns (never used)
    const bool store_lane = active && lane >= JacStrip<T>::margin / 4 && lane < 32 - JacStrip<T>::margin / 4;
    const int P = g.pitch, last = g.nrows - 1;
    const float* pc = p + (active ? jl : jstrip);
    const float* rc = rhs + (active ? jl : jstrip);
    PkConsts k;
    k.c = pk2(jc.cx, jc.cx);
    k.r = pk2(jc.dv[0].r, jc.dv[0].r);
    k.nb = pk2(-jc.dv[0].b, -jc.dv[0].b);
    k.m1 = pk2(-1.0f, -1.0f);
    k.nz = pk2(-0.0f, -0.0f);
    PkLane L;
    L.colin = 15u; L.cls = 0u;
    if (EDGE) {
        L.colin = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = jl + c;
            L.colin |= ((j >= 1 && j <= g.ny) ? 1u : 0u) << c;
            L.cls |= ((j == 1 || j == g.ny) ? 1u : 0u) << c;
        }
    }
    JacPk<T> S;
#pragma unroll
    for (int s = 0; s < T; ++s)
#pragma unroll
        for (int q = 0; q < 3; ++q) S.st[s][q][0] = S.st[s][q][1] = 0ull;
    // row R of p and row R - 1 of rhs travel together; slots are addressed by the row number
    auto issue = [&](int R) {
        cp_async_16(pbase + (unsigned)(R & (kPkPSlots - 1)) * kPkRowBytes, pc + (size_t)min(max(R, 0), last) * P);
        const float* rsrc = rc + (size_t)min(max(R - 1, 0), last) * P;
        const unsigned rdst = rbase + (unsigned)((R - 1) & (kPkRSlots - 1)) * kPkRowBytes;
        cp_async_16(rdst, rsrc);
        cp_async_16(rdst + kPkRSlots * kPkRowBytes, rsrc);
        cp_async_commit();
    };
    int R = ra - T;
#pragma unroll
    for (int d = 0; d < kPkAhead; ++d) issue(R + d);
#define VOF_PK_STEP(PH)                                                                                          \
    {                                                                                                            \
        cp_async_wait<kPkAhead - 1>();                                                                           \
        const float4 pin = lds_f4(pbase + (unsigned)(R & (kPkPSlots - 1)) * kPkRowBytes);                        \
        issue(R + kPkAhead);                                                                                     \
        if constexpr (FAST)                                                                                      \
            pk_step_fast<T, PH>(S, pin, R, rbase + (unsigned)((R & (kPkRSlots - 1)) + kPkRSlots) * kPkRowBytes,  \
                                pk2(-jc.cx), pk2(jc.dv[0].r), P, jl, pout, ra, rb, store_lane);                  \
        else                                                                                                     \
            pk_step<T, PH, EDGE, WALL, BARE>(S, L, pin, R, rbase + (unsigned)((R & (kPkRSlots - 1)) + kPkRSlots) * kPkRowBytes, \
                                       k, jc, g, jl, pout, ra, rb, store_lane);                                  \
        ++R;                                                                                                     \
    }
    while (R <= rb + T) {
        VOF_PK_STEP(0)
        VOF_PK_STEP(1)
        VOF_PK_STEP(2)
    }
#undef VOF_PK_STEP
    cp_async_wait<0>();      // the next item reuses the slots
}

// Work items of the packed kernel, in queue order (slow items first, so that none of them is the last one running):
//   1. edge strips (strip 0 and those from right_first on): every row, in chunks of rpcE rows        [EDGE or EDGE + WALL]
//   2. interior strips, the wlo rows next to the low i-wall and the whi rows next to the high one     [EDGE + WALL]
//   3. interior strips, the rows between: nA long chunks of rpcA rows ...                            [bulk]
//   4. ... followed by nB short ones of rpcB rows                                                    [bulk]
// A resident warp gets about one long item (little warm-up redundancy) and then short ones until the queue is empty: the
// tail behind the last item is a short item, not a long one.  The general variants cost 2 - 3 times the bulk per row,
// so their items are short.
struct PkSched {
    int nstrips, right_first;   // strips; first strip (>= 1) that touches the right j-wall or lies beyond it
    int nE, rpcE;               // chunks per edge strip
    int wlo, whi;               // wall rows of the interior strips (0: the slab does not touch that wall)
    int nA, rpcA, nB, rpcB;
    unsigned int* counter;      // zeroed by the host before the launch
};

// Persistent warps pull items from a queue (the cost of an item is data dependent).  Square cells and reciprocal
// divisions proven exact are the host's precondition (launch_jacobi_tb).
template <int T, bool FAST, bool BARE>
__global__ void __launch_bounds__(32 * kPkWarps, kPkBlocksPerSM)
k_jacobi_pk(Grid g, JacTB jc, PkSched sc, const float* __restrict__ p, float* __restrict__ pout, const float* __restrict__ rhs,
            int r0, int r1) {
    static_assert(T >= 1 && T <= JacStrip<T>::margin && T + kPkAhead <= kPkRSlots, "strip margin / rhs ring must cover the sweeps of one pass");
    extern __shared__ __align__(16) unsigned char pk_smem[];
    const int lane = threadIdx.x & 31;
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(pk_smem) + threadIdx.x * 16;
    const unsigned pbase = s0, rbase = s0 + kPkPSlots * kPkRowBytes;
    const int nES = 1 + sc.nstrips - sc.right_first, nI = sc.nstrips - nES;      // edge strips, interior strips
    const int n1 = nES * sc.nE, n2 = nI * ((sc.wlo > 0) + (sc.whi > 0)), n3 = nI * sc.nA, n4 = nI * sc.nB;
    const int m0 = r0 + sc.wlo;                                                  // first row of the bulk range
    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(sc.counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n1 + n2 + n3 + n4) break;
        int strip, ra, rb;
        if (item < n1) {
            const int e = item % nES, chunk = item / nES;
            strip = e == 0 ? 0 : sc.right_first + e - 1;
            ra = r0 + chunk * sc.rpcE;
            rb = min(r1, ra + sc.rpcE - 1);
        } else if (item < n1 + n2) {
            const int k = item - n1;
            strip = 1 + k % nI;
            const bool lo = sc.wlo > 0 && k / nI == 0;
            ra = lo ? r0 : r1 - sc.whi + 1;
            rb = lo ? r0 + sc.wlo - 1 : r1;
        } else if (item < n1 + n2 + n3) {
            const int k = item - n1 - n2;
            strip = 1 + k % nI;
            ra = m0 + (k / nI) * sc.rpcA;
            rb = ra + sc.rpcA - 1;
        } else {
            const int k = item - n1 - n2 - n3;
            strip = 1 + k % nI;
            ra = m0 + sc.nA * sc.rpcA + (k / nI) * sc.rpcB;
            rb = min(r1 - sc.whi, ra + sc.rpcB - 1);
        }
        const int jstrip = 1 - JacStrip<T>::margin + strip * JacStrip<T>::valid;   // == 1 (mod 4): float4-aligned
        const bool strip_interior = jstrip >= 2 && jstrip + kJacStripCols - 1 <= g.ny - 1;
        // every row the pipeline touches (ra - T .. rb + T, minus the sweeps' skew) strictly inside the i-walls?
        const bool wallrows = g.gi0 + ra - T - T < 2 || g.gi0 + rb + T > g.nx - 1;
        if (wallrows) pk_run_impl<T, true, true, false, BARE>(g, jc, p, pout, rhs, ra, rb, jstrip, lane, pbase, rbase);
        else if (!strip_interior) pk_run_impl<T, true, false, false, BARE>(g, jc, p, pout, rhs, ra, rb, jstrip, lane, pbase, rbase);
        else pk_run_impl<T, false, false, FAST, BARE>(g, jc, p, pout, rhs, ra, rb, jstrip, lane, pbase, rbase);
    }
}

}  // namespace vof
