// NVLink peer-to-peer halo exchange shared by the 2-D and 3-D contexts (new capability; the reference is
// single-address-space).  Every rank maps its neighbours' arenas (CUDA IPC across processes, or a
// plain pointer inside one process); one exchange is ONE kernel launch on the context's stream:
//
//   block 0 : tells both neighbours "my previous step no longer reads my halo rows" (done flag, epoch e), together
//             with which ping-pong buffers hold my live F / p (the push below addresses the neighbour's buffers
//             by MY indices, i.e. it assumes lockstep; a violation is reported instead of served stale rows);
//   all     : wait for the neighbours' done(e), then STORE my boundary rows of every exchanged field straight into
//             the neighbours' halo rows over NVLink (128-bit peer stores);
//   last block to finish its stores: system-scope fence, "data(e)" flag into both neighbours, then waits for
//             their data(e) -- the kernel ends only when my own halo rows are complete.
//
// No NCCL call, no host synchronisation, no separate signal / wait launches (round 1 used five launches per exchange,
// ~56 us per step on 8 GPUs; the fused kernel is one).  A 20 s watchdog in every wait keeps a dead neighbour from
// hanging the GPU; the timeout and lockstep flags are read by p2p_check().
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "vof_host_common.h"

namespace vof {

struct P2PFlags {            // lives at arena + arena_bytes - 128; written by the neighbours, read by me
    unsigned int done_from[2];
    unsigned int data_from[2];
    unsigned int timeout;    // epoch of the first wait that gave up
    unsigned int mismatch;   // epoch of the first exchange whose neighbour held F / p in the other buffer
    unsigned int cur_from[2];
    unsigned int arrive;     // blocks of my exchange kernel that finished their stores (reset by the last one)
    unsigned int pad[7];
};
static_assert(sizeof(P2PFlags) <= 128, "the flag block shares the last 256 bytes of the arena with the diagnostics");

constexpr int kP2PMaxEntries = 16;      // fields x sides
struct P2PTable {
    const float4* src[kP2PMaxEntries];
    float4* dst[kP2PMaxEntries];
    long long count4[kP2PMaxEntries];
    int n;
};
struct P2PPeers {
    P2PFlags* mine;
    P2PFlags* nbr[2];        // null on a physical wall
    unsigned int epoch, cur;
};

__device__ __forceinline__ bool p2p_wait_flag(const unsigned int* flag, unsigned int epoch, unsigned int* timeout_flag) {
    const volatile unsigned int* f = flag;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(*f - epoch) < 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 20000000000ull) { *timeout_flag = epoch; return false; }   // 20 s: the neighbour is gone
        __nanosleep(100);
    }
    return true;
}

static __global__ void __launch_bounds__(256)
k_p2p_exchange(P2PTable tab, P2PPeers pr) {
    __shared__ int s_last;
    const unsigned int e = pr.epoch;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // I am the lower neighbour's upper side (index 1) and the upper neighbour's lower side (index 0)
        for (int sd = 0; sd < 2; ++sd)
            if (pr.nbr[sd]) *reinterpret_cast<volatile unsigned int*>(&pr.nbr[sd]->cur_from[1 - sd]) = (e << 2) | pr.cur;
        __threadfence_system();
        for (int sd = 0; sd < 2; ++sd)
            if (pr.nbr[sd]) *reinterpret_cast<volatile unsigned int*>(&pr.nbr[sd]->done_from[1 - sd]) = e;
    }
    if (threadIdx.x == 0) {
        for (int sd = 0; sd < 2; ++sd)
            if (pr.nbr[sd]) p2p_wait_flag(&pr.mine->done_from[sd], e, &pr.mine->timeout);
        __threadfence_system();
    }
    __syncthreads();
    for (int k = 0; k < tab.n; ++k) {
        const float4* __restrict__ src = tab.src[k];
        float4* __restrict__ dst = tab.dst[k];
        const long long n4 = tab.count4[k];
        for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&pr.mine->arrive, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    pr.mine->arrive = 0;
    __threadfence_system();
    for (int sd = 0; sd < 2; ++sd)
        if (pr.nbr[sd]) *reinterpret_cast<volatile unsigned int*>(&pr.nbr[sd]->data_from[1 - sd]) = e;
    for (int sd = 0; sd < 2; ++sd) {
        if (!pr.nbr[sd]) continue;
        p2p_wait_flag(&pr.mine->data_from[sd], e, &pr.mine->timeout);
        const unsigned int v = *reinterpret_cast<const volatile unsigned int*>(&pr.mine->cur_from[sd]);
        if ((v >> 2) == e && (v & 3u) != pr.cur) pr.mine->mismatch = e;     // the neighbour's live buffers are not mine
    }
    __threadfence_system();
}

// host side: one endpoint per context
struct P2PEndpoint {
    char* peer_arena[2];
    size_t peer_arena_bytes[2];
    long long peer_nrows[2];
    bool peer_ipc[2];
    unsigned int epoch;
};

static inline P2PFlags* p2p_flags_of(char* arena, size_t arena_bytes) { return (P2PFlags*)(arena + arena_bytes - 128); }

static inline int p2p_connect(P2PEndpoint& ep, int side, const void* handle64, void* same_process_arena, long long peer_nrows,
                       size_t peer_arena_bytes, cudaStream_t stream) {
    if (same_process_arena) { ep.peer_arena[side] = (char*)same_process_arena; ep.peer_ipc[side] = false; }
    else {
        if (!handle64) return vofhost::fail(VOF_EINVAL, "null IPC handle");
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        void* ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ep.peer_arena[side] = (char*)ptr; ep.peer_ipc[side] = true;
    }
    ep.peer_nrows[side] = peer_nrows;
    ep.peer_arena_bytes[side] = peer_arena_bytes;
    // CUDA loads kernels lazily and a first-use load may wait for the device to go idle: with a flag-waiting kernel
    // already spinning that is a deadlock.  Load the exchange kernel now, while nothing waits.
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, k_p2p_exchange));
    P2PTable none; memset(&none, 0, sizeof(none));
    P2PPeers nobody; memset(&nobody, 0, sizeof(nobody));
    static unsigned int* dummy = nullptr;                      // a private flag block for the warm-up launch
    if (!dummy) { CU(cudaMalloc((void**)&dummy, sizeof(P2PFlags))); CU(cudaMemset(dummy, 0, sizeof(P2PFlags))); }
    nobody.mine = (P2PFlags*)dummy;
    k_p2p_exchange<<<1, 256, 0, stream>>>(none, nobody);
    CU(cudaStreamSynchronize(stream));
    return VOF_OK;
}

static inline void p2p_close(P2PEndpoint& ep) {
    for (int sd = 0; sd < 2; ++sd)
        if (ep.peer_arena[sd] && ep.peer_ipc[sd]) cudaIpcCloseMemHandle(ep.peer_arena[sd]);
}

// launch one exchange; `tab` holds (src, dst, count4) of every (field, side) pair that has a neighbour
static inline int p2p_exchange(P2PEndpoint& ep, char* arena, size_t arena_bytes, bool nbr_lo, bool nbr_hi, unsigned int cur,
                        const P2PTable& tab, cudaStream_t stream) {
    if ((nbr_lo && !ep.peer_arena[0]) || (nbr_hi && !ep.peer_arena[1]))
        return vofhost::fail(VOF_ESTATE, "p2p_connect was not called for every neighbour");
    P2PPeers pr;
    pr.mine = p2p_flags_of(arena, arena_bytes);
    pr.nbr[0] = nbr_lo ? p2p_flags_of(ep.peer_arena[0], ep.peer_arena_bytes[0]) : nullptr;
    pr.nbr[1] = nbr_hi ? p2p_flags_of(ep.peer_arena[1], ep.peer_arena_bytes[1]) : nullptr;
    pr.epoch = ++ep.epoch;
    pr.cur = cur;
    long long total4 = 0;
    for (int k = 0; k < tab.n; ++k) total4 += tab.count4[k];
    const int blocks = (int)std::min<long long>(256, std::max<long long>(1, (total4 + 2047) / 2048));
    k_p2p_exchange<<<blocks, 256, 0, stream>>>(tab, pr);
    return vofhost::launch_ok("k_p2p_exchange");
}

// VOF_ESTATE if any exchange so far timed out or found a neighbour out of lockstep.  Synchronises the stream.
static inline int p2p_check(char* arena, size_t arena_bytes, cudaStream_t stream, int* timed_out_epoch = nullptr, bool report = true) {
    unsigned int t[2] = {0, 0};
    CU(cudaMemcpyAsync(t, &p2p_flags_of(arena, arena_bytes)->timeout, sizeof(t), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    if (timed_out_epoch) *timed_out_epoch = (int)t[0];
    if (!report) return VOF_OK;
    if (t[0]) return vofhost::fail(VOF_ESTATE, "halo exchange %u timed out waiting for a neighbour (20 s); the halo rows since then are stale", t[0]);
    if (t[1]) return vofhost::fail(VOF_ESTATE, "halo exchange %u: a neighbour holds F / p in the other ping-pong buffer (ranks are not in lockstep)", t[1]);
    return VOF_OK;
}

}  // namespace vof
