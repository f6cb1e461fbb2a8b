// Building blocks shared by the second-generation streaming kernels (FCT sweeps, curvature):
//
//  * WorkQueue -- persistent warps claim (strip, chunk) items from a global counter.  A static grid of
//    one-item warps ran 2.5 waves at 8192^2 (ncu: SMs active 75 % of the kernel), and items that cross
//    the interface cost several times a bulk item; with a queue every SM stays busy until the items run
//    out.  The last warp to leave re-arms the counters, so no memset launch is needed.
//  * RowRing   -- per-lane cp.async ring in shared memory: each lane copies its own NC columns of the
//    next D rows of NF fields into private slots (LDGSTS: nothing is held in registers while the rows
//    are in flight) and reads them back with one LDS per field.  A lane only ever reads what it copied
//    itself, so cp.async.wait_group is the only synchronisation.  The first-generation kernels kept one
//    row per warp in flight (register prefetch) and were latency bound in the bulk.
#pragma once
#include "vof_common.cuh"

namespace vof {

struct WorkQueue {
    unsigned int* ctr;     // [0] next item, [1] warps that have left; both zero between launches
    int nitems;
};
__device__ __forceinline__ int wq_claim(const WorkQueue& q, int lane) {
    int item = 0;
    if (lane == 0) item = (int)atomicAdd(q.ctr, 1u);
    return __shfl_sync(0xffffffffu, item, 0);
}
// call once per warp after its last (failed) claim
__device__ __forceinline__ void wq_leave(const WorkQueue& q, int lane, unsigned total_warps) {
    if (lane == 0) {
        const unsigned d = atomicAdd(q.ctr + 1, 1u);
        if (d == total_warps - 1) { q.ctr[0] = 0u; q.ctr[1] = 0u; }   // every warp has made its last claim
    }
}

__device__ __forceinline__ void cp_async_16(unsigned smem_addr, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(unsigned smem_addr, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds_f4(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds_f2(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}

// Ring geometry: slot s, field f, thread t -> byte ((s * NF + f) * THREADS + t) * NC * 4.  SLOTS is a power of
// two; AHEAD = SLOTS - 2 rows are in flight, so a slot is rewritten two reads after it was read.
template <int NF, int NC, int SLOTS, int THREADS>
struct RowRing {
    static constexpr int kLane = NC * 4;                         // bytes per lane, field and row
    static constexpr int kField = THREADS * kLane;
    static constexpr int kSlot = NF * kField;
    static constexpr int kBytes = SLOTS * kSlot;
    static constexpr int kAhead = SLOTS - 2;
    static_assert((SLOTS & (SLOTS - 1)) == 0, "SLOTS must be a power of two");
    unsigned base;        // shared-memory address of this thread's lane in slot 0, field 0
    unsigned wr, rd;      // byte offsets of the slot to fill / to read next
    int row, row_end;     // next local row to request, last row to request
    int last, pitch4;     // clamp bound (rows), row pitch in bytes
    bool active;

    __device__ __forceinline__ void init(void* smem, int tid) {
        base = (unsigned)__cvta_generic_to_shared(smem) + tid * kLane;
        wr = 0; rd = 0; active = false;
    }
    // start streaming rows first .. end (clamped into [0, last_row]) and put AHEAD rows in flight.  Every row of the
    // previous item has been read (its trailing groups are empty), so reading resumes at the write position.
    __device__ __forceinline__ void start(bool active_, int first, int end, int last_row, int pitch_floats, const float* const (&src)[NF]) {
        active = active_; rd = wr; row = first; row_end = end; last = last_row; pitch4 = pitch_floats * 4;
#pragma unroll
        for (int d = 0; d < kAhead; ++d) issue(src);
    }
    __device__ __forceinline__ void issue(const float* const (&src)[NF]) {
        if (active && row <= row_end) {
            const int rr = min(max(row, 0), last);
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const char* g = reinterpret_cast<const char*>(src[f]) + (long long)rr * pitch4;
                if (NC == 4) cp_async_16(base + wr + f * kField, g);
                else cp_async_8(base + wr + f * kField, g);
            }
        }
        cp_async_commit();                                       // empty groups keep the group count uniform
        ++row;
        wr = (wr + kSlot) & (kBytes - 1);
    }
    // the oldest row in flight -> registers (zeros on idle lanes), and request one more row
    __device__ __forceinline__ void next(float (&x)[NF][NC], const float* const (&src)[NF]) {
        cp_async_wait<kAhead - 1>();
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            if constexpr (NC == 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active) v = lds_f4(base + rd + f * kField);
                x[f][0] = v.x; x[f][1] = v.y; x[f][2] = v.z; x[f][3] = v.w;
            } else {
                float2 v = make_float2(0.f, 0.f);
                if (active) v = lds_f2(base + rd + f * kField);
                x[f][0] = v.x; x[f][1] = v.y;
            }
        }
        rd = (rd + kSlot) & (kBytes - 1);
        issue(src);
    }
    __device__ __forceinline__ void drain() { cp_async_wait<0>(); }
};

__device__ __forceinline__ void cp_async_4(unsigned smem_addr, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float lds_f1(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a) : "memory");
    return v;
}

// RowRing plus the two columns next to a warp's strip: lane 0 also copies column jl-1 of the fields in LMASK, lane 31
// column jl+NC of the fields in RMASK (one 4-byte cp.async each), so stencils along j need no synchronous edge loads.
// Slot layout: NF * THREADS lanes of NC floats, then NF * WARPS pairs (left, right).
template <int NF, int NC, int SLOTS, int THREADS, unsigned LMASK, unsigned RMASK>
struct RowRingH {
    static constexpr int kLane = NC * 4;
    static constexpr int kField = THREADS * kLane;
    static constexpr int kHalo = NF * (THREADS / 32) * 8;        // bytes of halo pairs per slot
    static constexpr int kSlot = NF * kField + kHalo;
    static constexpr int kBytes = SLOTS * kSlot;
    static constexpr int kAhead = SLOTS - 2;
    static_assert(kSlot % 16 == 0, "slots must stay 16-byte aligned");
    unsigned base, hbase;  // shared address of this thread's lane (slot 0, field 0) / of this warp's halo pairs (slot 0)
    unsigned wr, rd;       // slot indices
    int row, row_end, last, pitch4;
    bool active;
    int lane;

    __device__ __forceinline__ void init(void* smem, int tid) {
        const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem);
        base = s0 + tid * kLane;
        hbase = s0 + NF * kField + (tid >> 5) * 8;
        lane = tid & 31;
        wr = 0; rd = 0; active = false;
    }
    __device__ __forceinline__ void start(bool active_, int first, int end, int last_row, int pitch_floats, const float* const (&src)[NF]) {
        active = active_; rd = wr; row = first; row_end = end; last = last_row; pitch4 = pitch_floats * 4;
#pragma unroll
        for (int d = 0; d < kAhead; ++d) issue(src);
    }
    __device__ __forceinline__ void issue(const float* const (&src)[NF]) {
        if (active && row <= row_end) {
            const int rr = min(max(row, 0), last);
            const unsigned so = wr * kSlot;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const char* gp = reinterpret_cast<const char*>(src[f]) + (long long)rr * pitch4;
                if constexpr (NC == 4) cp_async_16(base + so + f * kField, gp);
                else cp_async_8(base + so + f * kField, gp);
                if (((LMASK >> f) & 1u) && lane == 0) cp_async_4(hbase + so + f * (THREADS / 32) * 8, gp - 4);
                if (((RMASK >> f) & 1u) && lane == 31) cp_async_4(hbase + so + f * (THREADS / 32) * 8 + 4, gp + kLane);
            }
        }
        cp_async_commit();
        ++row;
        wr = (wr + 1) % SLOTS;
    }
    // x[f][1..NC] = own columns, x[f][0] = column jl-1, x[f][NC+1] = column jl+NC (fields outside the masks: unspecified)
    __device__ __forceinline__ void next(float (&x)[NF][NC + 2], const float* const (&src)[NF]) {
        cp_async_wait<kAhead - 1>();
        const unsigned so = rd * kSlot;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            float v[NC];
#pragma unroll
            for (int q = 0; q < NC; ++q) v[q] = 0.0f;
            if (active) {
                if constexpr (NC == 4) { const float4 t = lds_f4(base + so + f * kField); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
                else { const float2 t = lds_f2(base + so + f * kField); v[0] = t.x; v[1] = t.y; }
            }
#pragma unroll
            for (int q = 0; q < NC; ++q) x[f][q + 1] = v[q];
            if ((LMASK >> f) & 1u) {
                x[f][0] = __shfl_up_sync(0xffffffffu, v[NC - 1], 1);
                if (lane == 0 && active) x[f][0] = lds_f1(hbase + so + f * (THREADS / 32) * 8);
            }
            if ((RMASK >> f) & 1u) {
                x[f][NC + 1] = __shfl_down_sync(0xffffffffu, v[0], 1);
                if (lane == 31 && active) x[f][NC + 1] = lds_f1(hbase + so + f * (THREADS / 32) * 8 + 4);
            }
        }
        rd = (rd + 1) % SLOTS;
        issue(src);
    }
    __device__ __forceinline__ void drain() { cp_async_wait<0>(); }
};

// every bit pattern with an all-ones exponent is inf or NaN
template <int N> __device__ __forceinline__ bool all_finite_n(const float (&x)[N]) {
    unsigned a = 0;
#pragma unroll
    for (int q = 0; q < N; ++q) a = max(a, __float_as_uint(x[q]) & 0x7fffffffu);
    return a < 0x7f800000u;
}

}  // namespace vof
