// libvof C ABI (include/vof.h) -- 2-D context, launches, field access, diagnostics.
// Host code only decides ranges and launch shapes; all arithmetic is in vof2d_kernels.cuh.
#include <map>
#include <mutex>

#include "vof_host_common.h"
#include "vof2d_kernels.cuh"
#include "vof2d_jacobi_tb.cuh"
#include "vof2d_jacobi_pk.cuh"
#include "vof2d_fct.cuh"
#include "vof2d_momentum.cuh"
#include "vof2d_kappa.cuh"
#include "vof2d_extras.cuh"
#include "vof2d_tile.cuh"
#include "vof_p2p.cuh"

using namespace vof;

using vofhost::cdiv;
using vofhost::fail;
using vofhost::launch_ok;
using vofhost::make_const_div;
using vofhost::node_coords;

extern "C" const char* vof_last_error(void) { return vofhost::g_err; }
extern "C" int vof_abi_version(void) { return VOF_ABI_VERSION; }

extern "C" void vof_default_params(VofParams* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->nx = 200; p->ny = 200; p->nz = 0;                 // 2dvof.py:19-20
    p->Lx = 0.1; p->Ly = 0.1; p->Lz = 0.1;               // 2dvof.py:22-23
    p->dx = p->dy = p->dz = 0.0;                         // derive the reference way
    p->dt = 4e-6;                                        // 2dvof.py:33
    p->rho_l = 1000.0; p->rho_g = 50.0;                  // 2dvof.py:24-25
    p->nu_l = 1.0e-6; p->nu_g = 1.5e-5;                  // 2dvof.py:26-27
    p->sigma = 0.007;                                    // 2dvof.py:29
    p->gx = 0; p->gy = -5; p->gz = 0;                    // 2dvof.py:30-31
    p->n_jacobi = 10;                                    // 2dvof.py:521
    p->slab_lo = 0; p->slab_hi = 0; p->halo = 0; p->device = -1;
}

// ------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------
enum { BUF_F0 = 0, BUF_F1, BUF_U, BUF_V, BUF_P0, BUF_P1, BUF_US, BUF_VS, BUF_RHS, BUF_KAPPA, BUF_RHO, BUF_NU, BUF_COUNT };

struct ProfSpan { int kind; cudaEvent_t a, b; };

struct VofCtx {
    VofParams P;
    Grid g;
    Consts k;
    InitConsts ic2, ic3, ic1;
    int device;
    cudaStream_t stream;
    bool own_stream;
    bool own_arena;
    char* arena;
    size_t arena_bytes, field_bytes;
    float* buf[BUF_COUNT];     // pointers to logical (row 0, j = 0)
    int F_cur, p_cur;          // which of the ping-pong buffers is live
    float *xs, *ys;            // node coordinates (fp32), 2dvof.py:41-46
    Diag* diag;                // device
    bool rhs_valid;
    int lo, hi, H;             // owned global interior rows, halo depth
    bool has_lo, has_hi;       // this context holds the physical wall at i = 1 / i = nx
    // local row ranges (inclusive)
    int all_a, all_b;          // rows whose global index is in [0, nx+1]
    int in_a, in_b;            // rows whose global index is in [1, nx]
    // CUDA graphs of two consecutive steps, keyed by parity of the first istep and flags
    cudaGraphExec_t graph[2][4];
    long long graph_launches[2][4];
    int graph_cur[2][4];       // (F_cur, p_cur) the graph was captured with: it hard-codes the ping-pong buffers
    vofhost::AsyncGet* aget;   // non-stalling field read for the output path (vof2d_field_get_async)
    float* scratch;            // persistent device scratch of the display kernels (grown on demand)
    size_t scratch_bytes;
    FctC fctx, fcty;           // constants of the FCT sweeps
    MomC mom;                  // constants of the momentum predictor
    JacTB jac;                 // constants of the temporally blocked Jacobi
    int jac_resident_warps[6]; // warps of k_jacobi_tb<T> resident on the whole GPU, by T
    int jac_resident_warps_pk[6];
    int opt_jac_long_pct;      // third-generation Jacobi: share of the rows (percent) cut into one long item per resident warp
    int opt_pressure_solver;   // 0 (default): the reference's Jacobi sweeps; 1: Chebyshev-accelerated Jacobi (changes p: outside parity mode)
    int opt_bare_div;          // 1 (default): use the bare division where it is proven (0: always test for tiny numerators; A/B)
    int opt_fast_math;         // 1: tolerance mode of the blocked Jacobi (opt-in, not bit-exact; vof2d_jacobi_pk.cuh: pk_step_fast)
    int opt_tile;              // whole-step tile kernel (vof2d_tile.cuh): 0 never, 1 by grid size (default), 2 whenever it fits
    int tile_smem_set;
    int opt_packed;            // 1 (default): packed fp32x2 arithmetic in the streaming kernels that have it (same bits)
    int opt_jacobi_pk;         // 1 (default): third-generation blocked Jacobi (packed fp32x2, vof2d_jacobi_pk.cuh), 0: second generation
    int opt_jacobi_tb;         // 1: temporal blocking (default), 0: one launch per sweep
    int opt_jacobi_maxt;       // sweeps per HBM pass at most: 0 = by grid size, else 1..5 (5: 10 sweeps = 5 + 5; 3: 3 + 3 + 2 + 2)
    int opt_fct_x_cols;        // columns per lane of the x-sweep (2 or 4)
    int opt_advect_cols;       // columns per lane of the momentum predictor (2 or 4)
    int resident[16];          // resident blocks (whole device) of the persistent streaming kernels, by variant; 0 = not asked yet
    int opt_jac_rows;          // > 0: rows per item of the blocked Jacobi (default: max(16 T, 48))
    int opt_chunk_cap;         // > 0: upper bound on the rows one warp marches in the streaming kernels (load-balance experiments)
    int opt_adaptive;          // 1: interface-adaptive kernels (warp-uniform bulk rows short-cut, cp.async ring), 0: first generation
    P2PEndpoint p2p;           // neighbour arenas mapped into this process (lower / upper), NVLink P2P
    int sm_count;
    // launch accounting + optional per-kernel-kind CUDA-event timing (vof2d_profile)
    long long launches;
    bool profiling;            // spans are recorded (and graph replay is off)
    int prof_every, prof_count; // ... on every prof_every-th whole step only (sampling keeps the instrumentation out of the timing)
    bool prof_now;
    std::vector<cudaEvent_t>* ev_pool;        // recycled events
    std::vector<ProfSpan>* spans;             // (kind, start, stop) recorded on c->stream
    float* F() { return buf[F_cur ? BUF_F1 : BUF_F0]; }
    float* F_alt() { return buf[F_cur ? BUF_F0 : BUF_F1]; }
    float* p() { return buf[p_cur ? BUF_P1 : BUF_P0]; }
    float* p_alt() { return buf[p_cur ? BUF_P0 : BUF_P1]; }
};

static size_t field_stride_bytes(int nrows, int pitch) {
    size_t b = (size_t)nrows * pitch * sizeof(float);
    return (b + 255) / 256 * 256;
}

static int resolve(const VofParams* in, VofParams* P, Grid* g, int* lo, int* hi, int* H) {
    if (!in) return fail(VOF_EINVAL, "null params");
    *P = *in;
    if (P->nx < 4 || P->ny < 4) return fail(VOF_EINVAL, "nx, ny must be >= 4 (got %d x %d)", P->nx, P->ny);
    if (P->nz != 0) return fail(VOF_EINVAL, "vof2d_* needs nz == 0 (use vof3d_* for 3-D)");
    if (!(P->Lx > 0) || !(P->Ly > 0) || !(P->dt > 0)) return fail(VOF_EINVAL, "Lx, Ly, dt must be positive");
    if (P->n_jacobi < 0) return fail(VOF_EINVAL, "n_jacobi must be >= 0");
    if (P->slab_lo == 0 && P->slab_hi == 0) { P->slab_lo = 1; P->slab_hi = P->nx; }
    if (P->halo == 0) P->halo = 1;
    if (P->slab_lo < 1 || P->slab_hi > P->nx || P->slab_lo > P->slab_hi)
        return fail(VOF_EINVAL, "slab rows [%d, %d] outside [1, %d]", P->slab_lo, P->slab_hi, P->nx);
    const bool full = (P->slab_lo == 1 && P->slab_hi == P->nx);
    if (!full) {
        const int need = P->n_jacobi + 5;   // dependency radius of one step along i
        if (P->halo < need) return fail(VOF_EINVAL, "slab halo %d < n_jacobi + 5 = %d", P->halo, need);
        if (P->slab_hi - P->slab_lo + 1 < P->halo)
            return fail(VOF_EINVAL, "slab of %d rows is thinner than its halo %d", P->slab_hi - P->slab_lo + 1, P->halo);
    }
    if (P->halo < 1) return fail(VOF_EINVAL, "halo must be >= 1");
    *lo = P->slab_lo; *hi = P->slab_hi; *H = P->halo;
    g->nx = P->nx; g->ny = P->ny;
    g->gi0 = *lo - *H;
    g->nrows = (*hi - *lo + 1) + 2 * *H;
    g->pitch = round_up(kColOff + P->ny + 2, kPitchAlign);
    return VOF_OK;
}

extern "C" size_t vof2d_arena_bytes(const VofParams* p) {
    VofParams P; Grid g{}; int lo = 0, hi = 0, H = 0;
    if (resolve(p, &P, &g, &lo, &hi, &H) != VOF_OK) return 0;
    size_t xy = ((size_t)(P.nx + 3 + P.ny + 3) * sizeof(float) + 255) / 256 * 256;
    return field_stride_bytes(g.nrows, g.pitch) * BUF_COUNT + xy + 256;
}

// Proof that the reciprocal division by `d` is exact: every fp32 numerator against __fdiv_rn (4.8 ms per divisor).
// The verdict depends on the divisor (and the device executing it) only, so it is cached: the 16 slab contexts of a
// streamer, or repeated solver creation in a test session, prove each constant once.
static int const_div_exact(VofCtx* c, const ConstDiv& d, bool* ok, int bare = 0) {
    static std::mutex mu;
    static std::map<std::pair<int, unsigned long long>, bool> verdicts;
    unsigned int bits;
    memcpy(&bits, &d.b, sizeof(bits));
    const std::pair<int, unsigned long long> key(c->device, ((unsigned long long)bare << 32) | bits);
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = verdicts.find(key);
        if (it != verdicts.end()) { *ok = it->second; return VOF_OK; }
    }
    unsigned long long* bad = &c->diag->courant_count;
    unsigned long long h = 1;
    CU(cudaMemsetAsync(bad, 0, sizeof(*bad), c->stream));
    k_check_div_by_const<<<c->sm_count * 8, 256, 0, c->stream>>>(d, bad, bare);
    CU(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemsetAsync(bad, 0, sizeof(*bad), c->stream));
    *ok = (h == 0);
    std::lock_guard<std::mutex> lock(mu);
    verdicts[key] = *ok;
    return VOF_OK;
}

static int create_finish(VofCtx* c, void* arena, size_t arena_bytes, const std::vector<float>& x, const std::vector<float>& y);

static int create_impl(const VofParams* in, void* arena, size_t arena_bytes, VofCtx** out) {
    if (!out) return fail(VOF_EINVAL, "null out pointer");
    *out = nullptr;
    VofParams P; Grid g{}; int lo = 0, hi = 0, H = 0;
    int rc = resolve(in, &P, &g, &lo, &hi, &H);
    if (rc != VOF_OK) return rc;

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(VOF_ENODEV, "no CUDA device (%s); libvof has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    int dev = P.device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(VOF_EINVAL, "device %d out of range (%d devices)", dev, ndev);
    CU(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(VOF_ENODEV, "device %d is sm_%d%d; libvof is built for sm_100a only", dev, prop.major, prop.minor);

    VofCtx* c = new (std::nothrow) VofCtx();
    if (!c) return fail(VOF_ENOMEM, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->device = dev;
    c->lo = lo; c->hi = hi; c->H = H;
    c->has_lo = (lo == 1); c->has_hi = (hi == P.nx);
    c->g = g;

    // node coordinates and dx, dy (2dvof.py:41-50)
    std::vector<float> x, y;
    node_coords(x, P.nx, P.Lx);
    node_coords(y, P.ny, P.Ly);
    if (!(P.dx > 0)) P.dx = (double)x[3] - (double)x[2];
    if (!(P.dy > 0)) P.dy = (double)y[3] - (double)y[2];
    c->P = P;
    const double dx = P.dx, dy = P.dy, dxi = 1 / dx, dyi = 1 / dy;
    Consts& k = c->k;
    k.dt = (float)P.dt; k.dx = (float)dx; k.dy = (float)dy; k.dxi = (float)dxi; k.dyi = (float)dyi;
    k.dxi2 = (float)(dxi * dxi); k.dyi2 = (float)(dyi * dyi);
    k.dxdy = (float)(dx * dy); k.dtdy = (float)(P.dt * dy); k.dtdx = (float)(P.dt * dx);
    k.m1_2dx = (float)(-1 / (2 * dx)); k.m1_2dy = (float)(-1 / (2 * dy));
    k.i_dx_2 = (float)(1 / dx / 2); k.i_dy_2 = (float)(1 / dy / 2);
    k.neg_sigma = -(float)P.sigma;
    k.rho_l = (float)P.rho_l; k.rho_g = (float)P.rho_g; k.nu_l = (float)P.nu_l; k.nu_g = (float)P.nu_g;
    k.gx = (float)P.gx; k.gy = (float)P.gy;
    k.cflx = (float)(0.25 * dx); k.cfly = (float)(0.25 * dy);
    k.one = 1.0f; k.neg_zero = -0.0f;
    // initial-condition constants (2dvof.py:141-158)
    InitConsts ic{};
    ic.hdx = (float)(dx / 2); ic.hdy = (float)(dy / 2); ic.sqrt2dx = (float)(std::sqrt(2.0) * dx);
    ic.x2 = (float)(P.Lx / 3); ic.y2 = (float)(P.Ly / 2);
    ic.r = (float)(P.Lx / 12); ic.cx = (float)(P.Lx / 2);
    ic.ycut = (float)(P.Ly * 0.37);
    c->ic1 = ic;
    c->ic2 = ic; c->ic2.cy = 2.0f * ic.r;                 // cy = 2 * r in fp32 (r is a kernel local)
    c->ic3 = ic; c->ic3.cy = (float)P.Ly - 3.0f * ic.r;   // Ly - 3 * r in fp32

    {   // diagonal of the Poisson stencil by wall class, summed in the reference's order (2dvof.py:258-262)
        JacTB& j = c->jac;
        j.cx = k.dxi2; j.cy = k.dyi2;
        for (int ic = 0; ic < 2; ++ic)
            for (int jc2 = 0; jc2 < 2; ++jc2) {
                const float ae = j.cx, aw = ic ? 0.0f : j.cx, an = j.cy, as = jc2 ? 0.0f : j.cy;
                volatile float sum = ae + aw; sum = sum + an; sum = sum + as;
                j.ap[ic][jc2] = -1.0f * sum;
            }
        j.dv[0] = make_const_div(j.ap[0][0]);
        j.dv[1] = make_const_div(j.ap[0][1]);
        j.fast_div_ok = 0;
    }
    {
        FctC f{};
        f.dt = k.dt; f.dx = k.dx; f.dy = k.dy; f.dxdy = k.dxdy;
        f.d_dxdy = make_const_div(k.dxdy); f.d_dy = make_const_div(k.dy);
        f.fast_div_ok = 0;
        c->fctx = f; c->fctx.dtd = k.dtdy;     // 2dvof.py:324
        c->fcty = f; c->fcty.dtd = k.dtdx;     // 2dvof.py:388
    }
    c->mom.k = k; c->mom.d_dx = make_const_div(k.dx); c->mom.d_dy = make_const_div(k.dy); c->mom.fast_div_ok = 0;
    c->opt_jacobi_tb = 1;
    c->opt_jacobi_pk = 1;
    c->opt_packed = 1;
    c->opt_tile = 1;
    c->opt_bare_div = 1;
    if (const char* e = getenv("VOF_TILE")) { const int v = atoi(e); if (v >= 0 && v <= 2) c->opt_tile = v; }   // A/B and test default
    c->opt_jac_long_pct = 75;
    c->opt_jacobi_maxt = 0;
    c->opt_fct_x_cols = 2;
    c->opt_advect_cols = 2;
    c->opt_adaptive = 1;
    c->opt_chunk_cap = 0;
    c->sm_count = prop.multiProcessorCount;
    c->all_a = std::max(0, -g.gi0);
    c->all_b = std::min(g.nrows - 1, P.nx + 1 - g.gi0);
    c->in_a = std::max(0, 1 - g.gi0);
    c->in_b = std::min(g.nrows - 1, P.nx - g.gi0);

    // everything below can fail half-way: one cleanup path (vof2d_destroy copes with a partially built context)
    rc = create_finish(c, arena, arena_bytes, x, y);
    if (rc != VOF_OK) { vof2d_destroy(c); return rc; }
    *out = c;
    return VOF_OK;
}

static int create_finish(VofCtx* c, void* arena, size_t arena_bytes, const std::vector<float>& x, const std::vector<float>& y) {
    const VofParams& P = c->P;
    const Grid& g = c->g;
    const size_t need = vof2d_arena_bytes(&P);
    c->field_bytes = field_stride_bytes(g.nrows, g.pitch);
    if (arena) {
        if (((uintptr_t)arena & 255) != 0) return fail(VOF_EINVAL, "arena must be 256-byte aligned");
        if (arena_bytes < need) return fail(VOF_EINVAL, "arena too small: %zu < %zu", arena_bytes, need);
        c->arena = (char*)arena; c->own_arena = false;
    } else {
        cudaError_t e = cudaMalloc((void**)&c->arena, need);
        if (e != cudaSuccess) { c->arena = nullptr; return fail(VOF_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", need, cudaGetErrorString(e)); }
        c->own_arena = true;
    }
    c->arena_bytes = need;
    CU(cudaMemset(c->arena, 0, need));   // the reference's fields start at zero (2dvof.py:53-89)
    for (int b = 0; b < BUF_COUNT; ++b) c->buf[b] = (float*)(c->arena + c->field_bytes * b) + kColOff;
    c->xs = (float*)(c->arena + c->field_bytes * BUF_COUNT);
    c->ys = c->xs + (P.nx + 3);
    c->diag = (Diag*)(c->arena + need - 256);
    CU(cudaMemcpy(c->xs, x.data(), x.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->ys, y.data(), y.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    {   // prove the reciprocal divisions exact for this grid's constants (falls back to IEEE division otherwise)
        bool a = false, b = false;
        TRY(const_div_exact(c, c->jac.dv[0], &a)); TRY(const_div_exact(c, c->jac.dv[1], &b));
        c->jac.fast_div_ok = a && b;
        c->jac.bare_div_ok = 0;
        if (a && b) {       // the bare form for the packed Jacobi (no sub-normal test, no fix-up): every numerator again
            bool a2 = false, b2 = false;
            TRY(const_div_exact(c, c->jac.dv[0], &a2, 1)); TRY(const_div_exact(c, c->jac.dv[1], &b2, 1));
            c->jac.bare_div_ok = a2 && b2;
        }
        TRY(const_div_exact(c, c->fctx.d_dxdy, &a)); TRY(const_div_exact(c, c->fctx.d_dy, &b));
        c->fctx.fast_div_ok = c->fcty.fast_div_ok = a && b;
        TRY(const_div_exact(c, c->mom.d_dx, &a)); TRY(const_div_exact(c, c->mom.d_dy, &b));
        c->mom.fast_div_ok = a && b;
    }
    c->ev_pool = new std::vector<cudaEvent_t>();
    c->spans = new std::vector<ProfSpan>();
    return VOF_OK;
}

extern "C" int vof2d_create(const VofParams* p, VofCtx** out) { return create_impl(p, nullptr, 0, out); }
extern "C" int vof2d_create_in(const VofParams* p, void* arena, size_t arena_bytes, VofCtx** out) {
    if (!arena) return fail(VOF_EINVAL, "null arena");
    return create_impl(p, arena, arena_bytes, out);
}

extern "C" int vof2d_destroy(VofCtx* c) {
    if (!c) return VOF_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 4; ++b)
            if (c->graph[a][b]) cudaGraphExecDestroy(c->graph[a][b]);
    p2p_close(c->p2p);
    if (c->spans) { for (auto& sp : *c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); } delete c->spans; }
    if (c->ev_pool) { for (auto e : *c->ev_pool) cudaEventDestroy(e); delete c->ev_pool; }
    if (c->scratch) cudaFree(c->scratch);
    if (c->aget) { vofhost::async_get_free(*c->aget); delete c->aget; }
    if (c->own_arena && c->arena) cudaFree(c->arena);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);   // a caller-provided stream is left alone
    delete c;
    return VOF_OK;
}

extern "C" int vof2d_set_stream(VofCtx* c, void* cuda_stream) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 4; ++b)
            if (c->graph[a][b]) { cudaGraphExecDestroy(c->graph[a][b]); c->graph[a][b] = nullptr; }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return VOF_OK;
}

extern "C" int vof2d_synchronize(VofCtx* c) {
    CHECK_CTX(c);
    CU(cudaStreamSynchronize(c->stream));
    return VOF_OK;
}

extern "C" int vof2d_get_params(const VofCtx* c, VofParams* out) {
    CHECK_CTX(c);
    if (!out) return fail(VOF_EINVAL, "null out");
    *out = c->P;
    return VOF_OK;
}

// ------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------
// RAII span: counts the launch and, when profiling, brackets it with events on the ctx stream
struct Span {
    VofCtx* c; int kind; cudaEvent_t b;
    Span(VofCtx* c_, int kind_, int nlaunch = 1) : c(c_), kind(kind_), b(nullptr) {
        c->launches += nlaunch;
        if (!c->profiling || !c->prof_now) return;
        cudaEvent_t a;
        auto get = [&]() { cudaEvent_t e; if (!c->ev_pool->empty()) { e = c->ev_pool->back(); c->ev_pool->pop_back(); } else cudaEventCreate(&e); return e; };
        a = get(); b = get();
        cudaEventRecord(a, c->stream);
        c->spans->push_back({kind, a, b});
    }
    ~Span() { if (b) cudaEventRecord(b, c->stream); }
};

// Rows marched by one warp / block of the streaming kernels.  Long chunks amortise the warm-up rows that a chunk
// re-reads, but a grid needs full SMs to hide latency (a first version aimed at 16 warps per SM and left 2048^2 at 64 %
// of what short items reach): on big grids the cap `max_rows` applies,
// on small ones the chunks shrink (down to `min_rows`) until there are enough of them.  Small and medium grids are
// latency bound -- a step is ~19 dependent kernels, each a serial march over its rows -- so short items win there even
// when the warm-up rows double the work (200^2: 7.1k -> 10.2k steps/s, 1024^2: 4.6k -> 7.5k with 4-row items).
static int chunk_rows(const VofCtx* c, int rows, int columns_of_units, int min_rows, int max_rows, int units_per_warp = 1) {
    const int target_warps = c->sm_count * 128;     // two full waves of 64 warps per SM
    const int warps_per_row_chunk = std::max(1, columns_of_units / units_per_warp);
    const int nch = std::max(1, cdiv(target_warps, warps_per_row_chunk));
    int r = cdiv(rows, nch);
    r = std::max(r, min_rows);
    r = std::min(r, max_rows);
    if (c->opt_chunk_cap > 0) r = std::min(r, c->opt_chunk_cap);
    return std::max(1, std::min(r, rows));
}

// persistent launch: blocks that are resident at once (asked once per kernel variant)
template <typename K>
static int resident_blocks(VofCtx* c, K kern, int threads, int slot) {
    if (!c->resident[slot]) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, 0) != cudaSuccess) nb = 1;
        c->resident[slot] = std::max(1, nb) * c->sm_count;
    }
    return c->resident[slot];
}
template <typename K, typename... Args>
static void launch_queue(VofCtx* c, K kern, int slot, int warps_per_block, int nitems, Args... args) {
    const int blocks = std::min(cdiv(nitems, warps_per_block), resident_blocks(c, kern, 32 * warps_per_block, slot));
    kern<<<blocks, 32 * warps_per_block, 0, c->stream>>>(args...);
}

static unsigned bc_mask_all = 31u;

static int run_set_bc(VofCtx* c, unsigned mask) {
    Span span_(c, VOF_K_BC);
    const int nA = c->all_b - c->all_a + 1;
    const int nB = c->g.ny + 2;
    const int n = nA + (c->has_lo ? nB : 0) + (c->has_hi ? nB : 0);
    k_set_bc<<<cdiv(n, 128), 128, 0, c->stream>>>(c->g, c->buf[BUF_U], c->buf[BUF_V], c->F(), c->p(), c->buf[BUF_RHO],
                                                   c->all_a, c->all_b, c->has_lo, c->has_hi, mask);
    return launch_ok("k_set_bc");
}

static int run_cal_nu_rho(VofCtx* c) {
    Span span_(c, VOF_K_PROPS);
    const int rows = c->all_b - c->all_a + 1;
    const int rpb = chunk_rows(c, rows, cdiv(c->g.ny + 2, 32), 4, 32);
    dim3 grid(cdiv(c->g.ny + 2, kBlockJ), cdiv(rows, rpb));
    k_cal_nu_rho<<<grid, kBlockJ, 0, c->stream>>>(c->g, c->k, c->F(), c->buf[BUF_RHO], c->buf[BUF_NU], c->all_a, c->all_b, rpb);
    return launch_ok("k_cal_nu_rho");
}

static int run_kappa(VofCtx* c) {
    Span span_(c, VOF_K_KAPPA);
    const int rows = c->in_b - c->in_a + 1;
    const int nstrips = cdiv(c->g.ny, kKapValid);
    // adaptive kernel: items come from a queue and bulk rows are nearly free, so short items (less tail behind the
    // few expensive interface items) cost little; first generation: long chunks amortise the 4 warm-up rows
    int rpc = c->opt_adaptive ? chunk_rows(c, rows, nstrips, 6, 24) : chunk_rows(c, rows, nstrips, 6, 64);
    dim3 grid(cdiv(nstrips * cdiv(rows, rpc), kKapWarps));
    if (c->opt_adaptive) {
        const int nitems = nstrips * cdiv(rows, rpc);
        launch_queue(c, k_kappa5, 4, kKapWarps, nitems, c->g, c->k, WorkQueue{c->diag->wq, nitems}, c->F(), c->buf[BUF_KAPPA], c->in_a, c->in_b, rpc, nstrips);
    } else k_kappa4<false><<<grid, 32 * kKapWarps, 0, c->stream>>>(c->g, c->k, c->F(), c->buf[BUF_KAPPA], c->in_a, c->in_b, rpc, nstrips);
    return launch_ok("k_kappa4");
}


static int run_advect(VofCtx* c, bool inline_props) {
    Span span_(c, VOF_K_ADVECT);
    const int a = std::max(c->in_a, 1), b = std::min(c->in_b, c->g.nrows - 2);
    const int rows = b - a + 1;
    const int nc = c->opt_advect_cols;
    const int nstrips = cdiv(c->g.ny, 32 * nc);
    // queue items of at most 24 rows: an item costs two warm-up rows only, and with 64-row items the SMs idled 15 % of
    // the kernel behind the last items (ncu: sm__cycles_active min / max 652k / 849k; 8192^2: 0.458 -> 0.403 ms)
    const int rpc = chunk_rows(c, rows, nstrips, 4, (c->opt_adaptive && inline_props && nc == 2) ? 24 : 64);
    dim3 grid(cdiv(nstrips * cdiv(rows, rpc), kMomWarps));
    if (c->opt_adaptive && inline_props && nc == 2) {
        const int nitems = nstrips * cdiv(rows, rpc);
        WorkQueue wq{c->diag->wq, nitems};
#define ADQ c->g, c->mom, wq, c->buf[BUF_U], c->buf[BUF_V], c->F(), c->buf[BUF_KAPPA], c->buf[BUF_US], c->buf[BUF_VS], a, b, rpc, nstrips
        if (c->opt_packed) launch_queue(c, k_advect5<2, true>, 8, kMomWarps, nitems, ADQ);    // packed fp32x2 arithmetic (FFMA2)
        else launch_queue(c, k_advect5<2, false>, 7, kMomWarps, nitems, ADQ);     // 4 columns per lane would need > 48 KB of ring
#undef ADQ
        return launch_ok("k_advect5");
    }
#define ADA c->g, c->mom, c->buf[BUF_U], c->buf[BUF_V], c->F(), c->buf[BUF_KAPPA], c->buf[BUF_RHO], c->buf[BUF_NU], c->buf[BUF_US], c->buf[BUF_VS], a, b, rpc, nstrips
    if (nc == 2) {
        if (inline_props) k_advect4<true, 2><<<grid, 32 * kMomWarps, 0, c->stream>>>(ADA);
        else k_advect4<false, 2><<<grid, 32 * kMomWarps, 0, c->stream>>>(ADA);
    } else {
        if (inline_props) k_advect4<true, 4><<<grid, 32 * kMomWarps, 0, c->stream>>>(ADA);
        else k_advect4<false, 4><<<grid, 32 * kMomWarps, 0, c->stream>>>(ADA);
    }
#undef ADA
    return launch_ok("k_advect4");
}

static int run_rhs(VofCtx* c, bool inline_props) {
    Span span_(c, VOF_K_RHS);
    const int a = c->in_a, b = std::min(c->in_b, c->g.nrows - 2);
    const int rows = b - a + 1;
    const int rpb = chunk_rows(c, rows, cdiv(c->g.ny, 32), 4, 32);
    dim3 grid(cdiv(c->g.ny, kBlockJ), cdiv(rows, rpb));
    if (inline_props)
        k_rhs<true><<<grid, kBlockJ, 0, c->stream>>>(c->g, c->k, c->F(), c->buf[BUF_US], c->buf[BUF_VS], c->buf[BUF_RHS], a, b, rpb);
    else
        k_rhs<false><<<grid, kBlockJ, 0, c->stream>>>(c->g, c->k, c->buf[BUF_RHO], c->buf[BUF_US], c->buf[BUF_VS], c->buf[BUF_RHS], a, b, rpb);
    c->rhs_valid = true;
    return launch_ok("k_rhs");
}

// one sweep, rhs_mode as in k_jacobi
static int run_jacobi_sweep(VofCtx* c, int rhs_mode) {
    Span span_(c, VOF_K_JACOBI);
    const int rows = c->all_b - c->all_a + 1;
    const int rpb = chunk_rows(c, rows, cdiv(c->g.ny + 2, 32), 4, 32);
    dim3 grid(cdiv(c->g.ny + 2, kBlockJ), cdiv(rows, rpb));
    const float* rhoF = rhs_mode == 2 ? c->F() : c->buf[BUF_RHO];
    const int bare = c->jac.bare_div_ok && c->opt_bare_div;
#define JARGS c->g, c->k, c->p(), c->p_alt(), c->buf[BUF_RHS], rhoF, c->buf[BUF_US], c->buf[BUF_VS], c->all_a, c->all_b, rpb, \
              c->jac.dv[0].b, c->jac.dv[0].r, c->jac.dv[1].b, c->jac.dv[1].r, bare
    if (rhs_mode == 0) k_jacobi<0><<<grid, kBlockJ, 0, c->stream>>>(JARGS);
    else if (rhs_mode == 1) k_jacobi<1><<<grid, kBlockJ, 0, c->stream>>>(JARGS);
    else k_jacobi<2><<<grid, kBlockJ, 0, c->stream>>>(JARGS);
#undef JARGS
    c->p_cur ^= 1;
    return launch_ok("k_jacobi");
}

// Third generation (packed fp32x2 arithmetic, c*p products, cp.async rings; vof2d_jacobi_pk.cuh)
template <int T>
static int launch_jacobi_pk(VofCtx* c, const float* pin, float* pout) {
    // <tolerance mode, bare division>: the bare variant when the context proved it for both interior-row diagonals
    const bool bare = c->jac.bare_div_ok && c->opt_bare_div;
    auto kern_pk = c->opt_fast_math ? (bare ? k_jacobi_pk<T, true, true> : k_jacobi_pk<T, true, false>)
                                    : (bare ? k_jacobi_pk<T, false, true> : k_jacobi_pk<T, false, false>);
    if (!c->jac_resident_warps_pk[T]) {
        int nb = 0;
        CU(cudaFuncSetAttribute(k_jacobi_pk<T, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPkSmem));
        CU(cudaFuncSetAttribute(k_jacobi_pk<T, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPkSmem));
        CU(cudaFuncSetAttribute(k_jacobi_pk<T, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPkSmem));
        CU(cudaFuncSetAttribute(k_jacobi_pk<T, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPkSmem));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern_pk, 32 * kPkWarps, kPkSmem));
        c->jac_resident_warps_pk[T] = std::max(1, nb) * kPkWarps * c->sm_count;
    }
    const int resident = c->jac_resident_warps_pk[T];
    const int ra = c->in_a, rb = c->in_b, rows = rb - ra + 1;
    PkSched sc;
    sc.nstrips = cdiv(c->g.ny, JacStrip<T>::valid);
    sc.right_first = sc.nstrips;
    for (int st = 1; st < sc.nstrips; ++st) {
        const int jstrip = 1 - JacStrip<T>::margin + st * JacStrip<T>::valid;
        if (!(jstrip >= 2 && jstrip + kJacStripCols - 1 <= c->g.ny - 1)) { sc.right_first = st; break; }
    }
    // Interior strips: the 2T + 1 rows next to an i-wall are items of their own (general variant); between them one
    // long item per resident warp over the first `long_pct` of the rows and short items over the rest -- a grid too
    // small for that (long items of fewer than two short ones) gets short items only.  Edge strips: short items.
    {
        const int short_rows = c->opt_jac_rows > 0 ? c->opt_jac_rows : std::max(8 * T, 24);
        sc.wlo = c->has_lo ? std::min(rows / 2, 2 * T + 1) : 0;
        sc.whi = c->has_hi ? std::min(rows / 2, 2 * T + 1) : 0;
        const int mid = rows - sc.wlo - sc.whi;
        const int wps = std::max(1, resident / sc.nstrips);                 // resident warps per strip
        const int rowsA = (int)((long long)mid * c->opt_jac_long_pct / 100);
        sc.rpcA = cdiv(rowsA, wps);
        if (sc.rpcA < 2 * short_rows) { sc.nA = 0; sc.rpcA = 0; }
        else sc.nA = rowsA / sc.rpcA;
        const int rest = mid - sc.nA * sc.rpcA;
        sc.rpcB = std::min(std::max(rest, 1), short_rows);
        sc.nB = cdiv(rest, sc.rpcB);
        sc.rpcE = std::min(rows, std::max(short_rows, sc.rpcA / 3));
        sc.nE = cdiv(rows, sc.rpcE);
    }
    sc.counter = &c->diag->queue;
    const int nES = 1 + sc.nstrips - sc.right_first, nI = sc.nstrips - nES;
    const int nitems = nES * sc.nE + nI * ((sc.wlo > 0) + (sc.whi > 0) + sc.nA + sc.nB);
    CU(cudaMemsetAsync(sc.counter, 0, sizeof(unsigned int), c->stream));
    const int nwarps = std::min(nitems, resident);
    kern_pk<<<cdiv(nwarps, kPkWarps), 32 * kPkWarps, kPkSmem, c->stream>>>(c->g, c->jac, sc, pin, pout, c->buf[BUF_RHS], ra, rb);
    return launch_ok("k_jacobi_pk");
}

template <int T>
static int launch_jacobi_tb(VofCtx* c, const float* pin, float* pout) {
    const int rows = c->in_b - c->in_a + 1;
    // third generation when the cells are square, the reciprocal division is proven exact and the slab is tall enough
    if (c->opt_jacobi_pk && c->jac.fast_div_ok && c->jac.cx == c->jac.cy && rows >= 8 * T) return launch_jacobi_pk<T>(c, pin, pout);
    auto kern = c->jac.fast_div_ok ? k_jacobi_tb<T, true> : k_jacobi_tb<T, false>;
    if (!c->jac_resident_warps[T]) {
        int nb = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32 * kJacWarpsPerBlock, 0));
        c->jac_resident_warps[T] = std::max(1, nb) * kJacWarpsPerBlock * c->sm_count;
    }
    JacSched sc;
    sc.nstrips = cdiv(c->g.ny, JacStrip<T>::valid);
    // items small enough that the queue balances data-dependent costs (several items per resident warp),
    // large enough to amortise the 2T warm-up rows of a chunk (16T rows: 11 % redundant); on grids that cannot fill the
    // device that way the items shrink (down to 4T rows: 1.5x the work, but every SM has some)
    {
        const int rpc_max = c->opt_jac_rows > 0 ? c->opt_jac_rows : std::max(16 * T, 48), rpc_min = 4 * T;
        const long long want_items = 2LL * c->jac_resident_warps[T];
        const int fill = (int)std::min<long long>(rpc_max, (long long)rows * sc.nstrips / want_items);
        sc.rpc = std::min(rows, std::max(rpc_min, fill));
    }
    sc.nchunks = cdiv(rows, sc.rpc);
    sc.counter = &c->diag->queue;
    const int nitems = sc.nstrips * sc.nchunks;
    const int nwarps = std::min(nitems, c->jac_resident_warps[T]);
    CU(cudaMemsetAsync(sc.counter, 0, sizeof(unsigned int), c->stream));
    JacTB jc = c->jac;
    jc.bare_div_ok = jc.bare_div_ok && c->opt_bare_div;
    kern<<<cdiv(nwarps, kJacWarpsPerBlock), 32 * kJacWarpsPerBlock, 0, c->stream>>>(c->g, jc, sc, pin, pout, c->buf[BUF_RHS],
                                                                                  c->in_a, c->in_b);
    return launch_ok("k_jacobi_tb");
}

// Blocked or single-sweep Jacobi, and how many sweeps per HBM pass (profiles/exp_tb_sizes.py, graph-replayed steps/s):
// up to ~1500^2 the plain one-sweep kernel wins (it runs from L2 and has far more parallelism than strips x chunks of a
// small grid); from 2048^2 the blocked kernel does, with at most 3 sweeps per pass while the three fields are a few
// hundred MB (short register pipelines: the device is not full and per-warp latency counts), 5 beyond (HBM traffic counts).
// opt_jacobi_tb: 0 = never, 1 = by grid size (default), 2 = always.  opt_jacobi_maxt: 0 = by grid size (default), 1..5.
static double jacobi_field_mb(const VofCtx* c) { return 3.0 * (double)c->g.nrows * c->g.pitch * sizeof(float) / 1e6; }   // p, p', rhs
static bool packed_jacobi_eligible(const VofCtx* c) { return c->opt_jacobi_pk && c->jac.fast_div_ok && c->jac.cx == c->jac.cy; }
static bool use_jacobi_tb(const VofCtx* c) {
    if (c->opt_jacobi_tb == 0) return false;
    if (c->opt_jacobi_tb == 2) return true;
    // the packed kernel breaks even with ten single sweeps at ~1024^2 (0.089 ms both) and wins from there (1536^2: 0.089 vs 0.129)
    return jacobi_field_mb(c) > (packed_jacobi_eligible(c) ? 15.0 : 40.0);
}
static int jacobi_max_sweeps(const VofCtx* c) {
    if (c->opt_jacobi_maxt > 0) return c->opt_jacobi_maxt;
    // the packed kernel (96 registers, 20 warps per SM) has no short-pipeline advantage to lose: 5 sweeps per pass at every
    // size it runs at (10 sweeps, 3 + 3 + 2 + 2 -> 5 + 5: 2048^2 0.111 -> 0.098 ms, 3072^2 0.148 -> 0.121, 4096^2 0.242 -> 0.146)
    if (packed_jacobi_eligible(c)) return 5;
    return jacobi_field_mb(c) > 400.0 ? 5 : 3;
}

// nsweeps sweeps from the hoisted rhs, at most 5 per HBM pass.  The blocked kernels store interior cells only and ghost
// cells pass through a sweep unchanged (2dvof.py:265-266 copies the interior): one copy of the frame into the other
// ping-pong buffer before the first pass keeps the ghosts of both buffers equal to the reference's single array for
// every pass (they enter the sweeps as 0 * p[ghost], whose sign is the ghost's).
static int run_jacobi_tb(VofCtx* c, int nsweeps) {
    const int npass = cdiv(nsweeps, jacobi_max_sweeps(c));
    const int base = nsweeps / npass, extra = nsweeps % npass;
    for (int k = 0; k < npass; ++k) {
        const int T = base + (k < extra ? 1 : 0);
        Span span_(c, VOF_K_JACOBI, k == 0 ? 2 : 1);
        if (k == 0) {
            const int nrow = c->all_b - c->all_a + 1;
            const int lo_row = c->has_lo ? 0 - c->g.gi0 : -1, hi_row = c->has_hi ? c->g.nx + 1 - c->g.gi0 : -1;
            const int n = nrow + (lo_row >= 0 ? c->g.ny + 2 : 0) + (hi_row >= 0 ? c->g.ny + 2 : 0);
            k_copy_frame<<<cdiv(n, 128), 128, 0, c->stream>>>(c->g, c->p(), c->p_alt(), c->all_a, c->all_b, lo_row, hi_row);
            TRY(launch_ok("k_copy_frame"));
        }
        int rc = VOF_OK;
        switch (T) {
            case 1: rc = launch_jacobi_tb<1>(c, c->p(), c->p_alt()); break;
            case 2: rc = launch_jacobi_tb<2>(c, c->p(), c->p_alt()); break;
            case 3: rc = launch_jacobi_tb<3>(c, c->p(), c->p_alt()); break;
            case 4: rc = launch_jacobi_tb<4>(c, c->p(), c->p_alt()); break;
            default: rc = launch_jacobi_tb<5>(c, c->p(), c->p_alt()); break;
        }
        if (rc != VOF_OK) return rc;
        c->p_cur ^= 1;
    }
    return VOF_OK;
}

// Opt-in stronger pressure solver: nsweeps iterations of the Chebyshev semi-iterative
// method on the reference's Jacobi iteration, from the hoisted rhs.  rho = (1 + cos(pi / max(nx, ny))) / 2 bounds the
// spectrum of the Jacobi matrix on the Neumann grid away from the constant mode (which the iteration leaves alone:
// the polynomial is 1 at eigenvalue 1); w(1) = 1, w(2) = 1 / (1 - rho^2 / 2), w(k+1) = 1 / (1 - rho^2 w(k) / 4).
static int run_jacobi_cheb(VofCtx* c, int nsweeps) {
    const double rho = 0.5 * (1.0 + std::cos(M_PI / std::max(c->g.nx, c->g.ny)));
    double w = 1.0;
    for (int k = 1; k <= nsweeps; ++k) {
        if (k == 1) { TRY(run_jacobi_sweep(c, 0)); continue; }      // x(1) = J x(0); also makes the other buffer's frame current
        w = k == 2 ? 1.0 / (1.0 - 0.5 * rho * rho) : 1.0 / (1.0 - 0.25 * rho * rho * w);
        Span span_(c, VOF_K_JACOBI);
        const int rows = c->all_b - c->all_a + 1;
        const int rpb = chunk_rows(c, rows, cdiv(c->g.ny + 2, 32), 4, 32);
        dim3 grid(cdiv(c->g.ny + 2, kBlockJ), cdiv(rows, rpb));
        k_jacobi_cheb<<<grid, kBlockJ, 0, c->stream>>>(c->g, c->k, c->p(), c->p_alt(), c->buf[BUF_RHS], (float)w, c->all_a, c->all_b, rpb);
        TRY(launch_ok("k_jacobi_cheb"));
        c->p_cur ^= 1;
    }
    return VOF_OK;
}

// test/forward_fct.py:254-264: solve_VOF_rudman(t, eps) of the stand-alone FCT variant, set_BC(F) after each sweep
static int run_fct_forward(VofCtx* c, int istep, float eps) {
    if (c->g.gi0 != 0 || c->g.nrows != c->g.nx + 2) return fail(VOF_ESTATE, "the forward-FCT variant needs a full-domain context");
    FwdC fx{c->k.dt, c->k.dx, c->k.dy, c->k.dxdy, c->k.dtdy, eps}, fy = fx;
    fy.dtd = c->k.dtdx;
    const dim3 grid(cdiv(c->g.ny, 128), c->g.nx);
    for (int half = 0; half < 2; ++half) {
        const bool y = (istep % 2 == 0) == (half == 0);
        ++c->launches;
        if (y) k_fct_forward<1><<<grid, 128, 0, c->stream>>>(c->g, fy, c->F(), c->buf[BUF_V], c->F_alt());
        else k_fct_forward<0><<<grid, 128, 0, c->stream>>>(c->g, fx, c->F(), c->buf[BUF_U], c->F_alt());
        TRY(launch_ok("k_fct_forward"));
        // ghost cells of the new level start from the old level's (the reference's F[level + 1] starts from zeros and
        // set_BC then writes every ghost cell: same result)
        c->F_cur ^= 1;
        TRY(run_set_bc(c, 4u));
    }
    return VOF_OK;
}

static int run_project(VofCtx* c, bool inline_props) {
    Span span_(c, VOF_K_PROJECT);
    const int a = std::max(c->in_a, 1), b = c->in_b;
    const int rows = b - a + 1;
    const int nstrips = cdiv(c->g.ny, 128);
    const int rpc = chunk_rows(c, rows, nstrips, 4, 64);
    dim3 grid(cdiv(nstrips * cdiv(rows, rpc), kMomWarps));
    unsigned long long* cc = &c->diag->courant_count;
    CU(cudaMemsetAsync(cc, 0, sizeof(*cc), c->stream));
    const float* rhoF = inline_props ? c->F() : c->buf[BUF_RHO];
    if (inline_props)
        k_project4<true><<<grid, 32 * kMomWarps, 0, c->stream>>>(c->g, c->k, rhoF, c->p(), c->buf[BUF_US], c->buf[BUF_VS], c->buf[BUF_U], c->buf[BUF_V],
                                                               cc, a, b, rpc, nstrips, c->lo - c->g.gi0, c->hi - c->g.gi0);
    else
        k_project4<false><<<grid, 32 * kMomWarps, 0, c->stream>>>(c->g, c->k, rhoF, c->p(), c->buf[BUF_US], c->buf[BUF_VS], c->buf[BUF_U], c->buf[BUF_V],
                                                                cc, a, b, rpc, nstrips, c->lo - c->g.gi0, c->hi - c->g.gi0);
    return launch_ok("k_project4");
}

static int run_fct_x(VofCtx* c, bool post) {
    Span span_(c, VOF_K_FCT_X);
    const int rows = c->in_b - c->in_a + 1;
    const int nc = c->opt_fct_x_cols;
    const int nstrips = cdiv(c->g.ny + 1, 32 * nc);
    int rpc = c->opt_adaptive ? chunk_rows(c, rows, nstrips, 6, 48)             // queue-scheduled: shorter items, less tail
                              : chunk_rows(c, rows, nstrips, 6, 96);            // 6 warm-up rows are re-read per chunk
    const int nwarps = nstrips * cdiv(rows, rpc);
    dim3 grid(cdiv(nwarps, kFctXWarps));
#define FXA c->g, c->fctx, c->F(), c->buf[BUF_U], c->F_alt(), c->in_a, c->in_b, rpc, nstrips
    if (c->opt_adaptive) {
        WorkQueue wq{c->diag->wq, nwarps};
#define FXQ c->g, c->fctx, wq, c->F(), c->buf[BUF_U], c->F_alt(), c->in_a, c->in_b, rpc, nstrips, 1, 0LL, 0, 0
        if (nc == 2) {
            if (post) launch_queue(c, k_fct_x5<true, 2, FctOps2, false>, 0, kFctXWarps, nwarps, FXQ);
            else launch_queue(c, k_fct_x5<false, 2, FctOps2, false>, 1, kFctXWarps, nwarps, FXQ);
        } else {
            if (post) launch_queue(c, k_fct_x5<true, 4, FctOps2, false>, 2, kFctXWarps, nwarps, FXQ);
            else launch_queue(c, k_fct_x5<false, 4, FctOps2, false>, 3, kFctXWarps, nwarps, FXQ);
        }
#undef FXQ
    } else if (nc == 2) {
        if (post) k_fct_x4<true, 2><<<grid, 32 * kFctXWarps, 0, c->stream>>>(FXA);
        else k_fct_x4<false, 2><<<grid, 32 * kFctXWarps, 0, c->stream>>>(FXA);
    } else {
        if (post) k_fct_x4<true, 4><<<grid, 32 * kFctXWarps, 0, c->stream>>>(FXA);
        else k_fct_x4<false, 4><<<grid, 32 * kFctXWarps, 0, c->stream>>>(FXA);
    }
#undef FXA
    c->F_cur ^= 1;
    return launch_ok("k_fct_x4");
}

static int run_fct_y(VofCtx* c, bool post) {
    Span span_(c, VOF_K_FCT_Y);
    const int rows = c->all_b - c->all_a + 1;
    const int nstrips = cdiv(c->g.ny + 1, kFctYValid);
    int rpw = chunk_rows(c, rows, nstrips, 2, 16);
    const int nwarps = nstrips * cdiv(rows, rpw);
    dim3 grid(cdiv(nwarps, kFctYWarps));
#define FYA c->g, c->fcty, c->F(), c->buf[BUF_V], c->F_alt(), c->all_a, c->all_b, rpw, nstrips
    if (c->opt_adaptive) {
        WorkQueue wq{c->diag->wq, nwarps};
        if (post) launch_queue(c, k_fct_y5<true, FctOps2, false>, 5, kFctYWarps, nwarps, c->g, c->fcty, wq, c->F(), c->buf[BUF_V], c->F_alt(), c->all_a, c->all_b, rpw, nstrips, 0);
        else launch_queue(c, k_fct_y5<false, FctOps2, false>, 6, kFctYWarps, nwarps, c->g, c->fcty, wq, c->F(), c->buf[BUF_V], c->F_alt(), c->all_a, c->all_b, rpw, nstrips, 0);
    } else {
        if (post) k_fct_y4<true, false><<<grid, 32 * kFctYWarps, 0, c->stream>>>(FYA);
        else k_fct_y4<false, false><<<grid, 32 * kFctYWarps, 0, c->stream>>>(FYA);
    }
#undef FYA
    c->F_cur ^= 1;
    return launch_ok("k_fct_y4");
}

static int run_post(VofCtx* c) {
    Span span_(c, VOF_K_POST);
    const int rows = c->all_b - c->all_a + 1;
    const int rpb = chunk_rows(c, rows, cdiv(c->g.ny + 2, 32), 4, 32);
    dim3 grid(cdiv(c->g.ny + 2, kBlockJ), cdiv(rows, rpb));
    k_post_process_f<<<grid, kBlockJ, 0, c->stream>>>(c->g, c->F(), c->all_a, c->all_b, rpb);
    return launch_ok("k_post_process_f");
}

// ------------------------------------------------------------------------------------
// one entry per reference kernel
// ------------------------------------------------------------------------------------
extern "C" int vof2d_set_init_F(VofCtx* c, int ic) {
    CHECK_CTX(c);
    if (ic < 1 || ic > 3) return fail(VOF_EINVAL, "ic must be 1, 2 or 3 (got %d)", ic);
    const InitConsts& k = ic == 1 ? c->ic1 : (ic == 2 ? c->ic2 : c->ic3);
    const int rows = c->all_b - c->all_a + 1;
    dim3 grid(cdiv(c->g.ny + 2, kBlockJ), std::min(rows, 1024));
    k_set_init_F<<<grid, kBlockJ, 0, c->stream>>>(c->g, c->k, k, ic, c->xs, c->ys, c->F(), c->all_a, c->all_b);
    return launch_ok("k_set_init_F");
}
extern "C" int vof2d_set_BC(VofCtx* c) { CHECK_CTX(c); return run_set_bc(c, bc_mask_all); }
extern "C" int vof2d_cal_nu_rho(VofCtx* c) { CHECK_CTX(c); return run_cal_nu_rho(c); }
extern "C" int vof2d_get_normal_young(VofCtx* c) { CHECK_CTX(c); return run_kappa(c); }
extern "C" int vof2d_advect_upwind(VofCtx* c) { CHECK_CTX(c); c->rhs_valid = false; return run_advect(c, false); }

extern "C" int vof2d_solve_p_jacobi(VofCtx* c, int nsweeps) {
    CHECK_CTX(c);
    if (nsweeps < 0) return fail(VOF_EINVAL, "nsweeps must be >= 0");
    if (nsweeps == 1) return run_jacobi_sweep(c, 1);   // the reference's structure: rhs recomputed inside the sweep
    if (nsweeps == 0) return VOF_OK;
    TRY(run_rhs(c, false));
    if (c->opt_pressure_solver == 1) return run_jacobi_cheb(c, nsweeps);
    if (use_jacobi_tb(c)) return run_jacobi_tb(c, nsweeps);
    for (int s = 0; s < nsweeps; ++s) TRY(run_jacobi_sweep(c, 0));
    return VOF_OK;
}
extern "C" int vof2d_fct_forward(VofCtx* c, int istep, float eps) { CHECK_CTX(c); return run_fct_forward(c, istep, eps); }
extern "C" int vof2d_update_uv(VofCtx* c) { CHECK_CTX(c); return run_project(c, false); }
extern "C" int vof2d_fct_x_sweep(VofCtx* c) { CHECK_CTX(c); return run_fct_x(c, false); }
extern "C" int vof2d_fct_y_sweep(VofCtx* c) { CHECK_CTX(c); return run_fct_y(c, false); }
extern "C" int vof2d_solve_VOF_rudman(VofCtx* c, int istep) {
    CHECK_CTX(c);
    if (istep % 2 == 0) { TRY(run_fct_y(c, false)); TRY(run_fct_x(c, false)); }
    else { TRY(run_fct_x(c, false)); TRY(run_fct_y(c, false)); }
    return VOF_OK;
}
extern "C" int vof2d_post_process_f(VofCtx* c) { CHECK_CTX(c); return run_post(c); }

// ------------------------------------------------------------------------------------
// display kernels, 2dvof.py:458-492 (+ rgb_buf.to_numpy(), 535): monitoring output, full-domain contexts only
// ------------------------------------------------------------------------------------
static int display_scratch(VofCtx* c, size_t bytes, float** out) {
    if (c->scratch_bytes < bytes) {
        CU(cudaStreamSynchronize(c->stream));
        if (c->scratch) cudaFree(c->scratch);
        c->scratch = nullptr; c->scratch_bytes = 0;
        if (cudaMalloc((void**)&c->scratch, bytes) != cudaSuccess) return fail(VOF_ENOMEM, "cudaMalloc(%zu bytes) for the display buffer failed", bytes);
        c->scratch_bytes = bytes;
    }
    *out = c->scratch;
    return VOF_OK;
}

static int display_dev(VofCtx* c, int view, float* rgb) {
    if (c->g.gi0 != 0 || c->g.nrows != c->g.nx + 2) return fail(VOF_ESTATE, "display kernels need a full-domain context");
    ++c->launches;
    dim3 grid(cdiv(2 * c->g.ny, 256), 2 * c->g.nx);
    const float mx = (float)(c->P.Lx / 0.2), my = (float)(c->P.Ly / 0.2);      // Python scalars, folded in double
    switch (view) {
        case VOF_VIEW_VOF: k_display<0><<<grid, 256, 0, c->stream>>>(c->g, c->F(), c->buf[BUF_U], c->buf[BUF_V], 1.0f, rgb); break;
        case VOF_VIEW_U: k_display<1><<<grid, 256, 0, c->stream>>>(c->g, c->F(), c->buf[BUF_U], c->buf[BUF_V], mx, rgb); break;
        case VOF_VIEW_V: k_display<2><<<grid, 256, 0, c->stream>>>(c->g, c->F(), c->buf[BUF_U], c->buf[BUF_V], my, rgb); break;
        case VOF_VIEW_VNORM: k_display<3><<<grid, 256, 0, c->stream>>>(c->g, c->F(), c->buf[BUF_U], c->buf[BUF_V], my, rgb); break;
        default: return fail(VOF_EINVAL, "unknown view %d", view);
    }
    return launch_ok("k_display");
}
extern "C" int vof2d_display_field_dev(VofCtx* c, int view, float* rgb_dev) {
    CHECK_CTX(c);
    if (!rgb_dev) return fail(VOF_EINVAL, "null output");
    CU(cudaSetDevice(c->device));
    return display_dev(c, view, rgb_dev);
}
extern "C" int vof2d_display_field(VofCtx* c, int view, float* rgb_host) {
    CHECK_CTX(c);
    if (!rgb_host) return fail(VOF_EINVAL, "null output");
    CU(cudaSetDevice(c->device));
    const size_t bytes = (size_t)4 * c->g.nx * c->g.ny * sizeof(float);
    float* d = nullptr;
    TRY(display_scratch(c, bytes, &d));          // persistent: no cudaMalloc / cudaFree (device-wide syncs) per frame
    int rc = display_dev(c, view, d);
    if (rc == VOF_OK) {
        cudaError_t e = cudaMemcpyAsync(rgb_host, d, bytes, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail((int)e, "rgb_buf copy failed: %s", cudaGetErrorString(e));
    }
    return rc;
}
extern "C" int vof2d_interp_velocity(VofCtx* c, float* V_host) {
    CHECK_CTX(c);
    if (!V_host) return fail(VOF_EINVAL, "null output");
    if (c->g.gi0 != 0 || c->g.nrows != c->g.nx + 2) return fail(VOF_ESTATE, "display kernels need a full-domain context");
    CU(cudaSetDevice(c->device));
    const size_t bytes = (size_t)2 * (c->g.nx + 2) * (c->g.ny + 2) * sizeof(float);
    float* df = nullptr;
    TRY(display_scratch(c, bytes, &df));
    float2* d = (float2*)df;
    ++c->launches;
    k_interp_velocity<<<dim3(cdiv(c->g.ny + 2, 256), c->g.nx + 2), 256, 0, c->stream>>>(c->g, c->buf[BUF_U], c->buf[BUF_V], d);
    int rc = launch_ok("k_interp_velocity");
    if (rc == VOF_OK) {
        cudaError_t e = cudaMemcpyAsync(V_host, d, bytes, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail((int)e, "V copy failed: %s", cudaGetErrorString(e));
    }
    return rc;
}

// ------------------------------------------------------------------------------------
// Whole-step tile kernel (vof2d_tile.cuh): one launch per step on grids small enough to be launch bound.
// ------------------------------------------------------------------------------------
static bool tile_geometry(const VofCtx* c, TileArgs* a, dim3* grid, size_t* smem) {
    if (c->opt_tile == 0 || c->opt_pressure_solver != 0) return false;
    if (c->g.gi0 != 0 || c->g.nrows != c->g.nx + 2) return false;                 // full-domain contexts only
    const int H = c->P.n_jacobi + 5;                                                // dependency radius of one step
    const int tj = kTileW - 2 * H;
    if (tj < 8) return false;
    const int max_th = (int)((227 * 1024) / (10 * kTileW * sizeof(float) + kTileW));   // ten tile arrays + one class byte per cell
    int bj = cdiv(c->g.ny + 2, tj);
    if (c->g.ny + 2 - (bj - 1) * tj == 1 && bj > 1) --bj;                          // the far ghost column goes with column ny (see k_step_tile)
    int bi = std::max(1, c->sm_count / bj);                                         // about one block per SM ...
    int ti = cdiv(c->g.nx + 2, bi);
    ti = std::max(4, std::min(ti, max_th - 2 * H));                                 // ... as long as the tile fits
    if (ti + 2 * H > max_th) return false;
    bi = cdiv(c->g.nx + 2, ti);
    if (c->g.nx + 2 - (bi - 1) * ti == 1 && bi > 1) --bi;                          // the far ghost row goes with row nx
    // by grid size: the tile kernel wins while its blocks are one wave (200^2 2.1x ... 512^2 1.4x); with a second wave
    // the redundant halo work costs more than the launches save (640^2: 0.8x) -- profiles/exp_tile.py
    if (c->opt_tile == 1 && (long long)bi * bj > c->sm_count) return false;
    a->ti = ti; a->tj = tj; a->H = H; a->th = ti + 2 * H; a->n_jacobi = c->P.n_jacobi;
    *grid = dim3(bj, bi);
    *smem = (size_t)a->th * kTileW * (10 * sizeof(float) + 1);
    return true;
}

static int run_step_tile(VofCtx* c, int istep, TileArgs a, dim3 grid, size_t smem) {
    Span span_(c, VOF_K_TILE, 2);
    if (!c->tile_smem_set) {
        CU(cudaFuncSetAttribute(k_step_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        c->tile_smem_set = 1;
    }
    a.u = c->buf[BUF_U]; a.v = c->buf[BUF_V]; a.p = c->p(); a.F = c->F();
    a.un = c->buf[BUF_RHO]; a.vn = c->buf[BUF_NU]; a.pn = c->p_alt(); a.Fn = c->F_alt();
    a.us = c->buf[BUF_US]; a.vs = c->buf[BUF_VS]; a.kappa = c->buf[BUF_KAPPA];
    a.courant_count = &c->diag->courant_count;
    a.istep = istep;
    a.d_dx = c->mom.d_dx; a.d_dy = c->mom.d_dy; a.d_dxdy = c->fctx.d_dxdy; a.d_ap0 = c->jac.dv[0]; a.d_ap1 = c->jac.dv[1];
    a.fast = c->jac.fast_div_ok && c->mom.fast_div_ok && c->fctx.fast_div_ok;
    a.bare = a.fast && c->jac.bare_div_ok && c->opt_bare_div;
    CU(cudaMemsetAsync(a.courant_count, 0, sizeof(unsigned long long), c->stream));
    k_step_tile<<<grid, kTileThreads, smem, c->stream>>>(c->g, c->k, c->jac, a);
    TRY(launch_ok("k_step_tile"));
    // the new state lives in the other buffers: F / p ping-pong, u / v exchanged with the (dead) rho / nu buffers
    c->F_cur ^= 1; c->p_cur ^= 1;
    std::swap(c->buf[BUF_U], c->buf[BUF_RHO]);
    std::swap(c->buf[BUF_V], c->buf[BUF_NU]);
    c->rhs_valid = false;
    return VOF_OK;
}

// ------------------------------------------------------------------------------------
// the loop body 2dvof.py:513-528
// ------------------------------------------------------------------------------------
static int step_impl(VofCtx* c, int istep, unsigned flags) {
    if (c->profiling) c->prof_now = (c->prof_count++ % c->prof_every) == 0;
    if (flags & VOF_STEP_NO_FUSION) {
        TRY(run_cal_nu_rho(c));                                   // 513
        TRY(run_kappa(c));                                        // 514
        c->rhs_valid = false;
        TRY(run_advect(c, false));                                // 517
        TRY(run_set_bc(c, bc_mask_all));                          // 518
        for (int s = 0; s < c->P.n_jacobi; ++s) TRY(run_jacobi_sweep(c, 1));   // 521-522
        TRY(run_project(c, false));                               // 524
        TRY(run_set_bc(c, bc_mask_all));                          // 525
        if (istep % 2 == 0) { TRY(run_fct_y(c, false)); TRY(run_fct_x(c, false)); }   // 526
        else { TRY(run_fct_x(c, false)); TRY(run_fct_y(c, false)); }
        TRY(run_post(c));                                         // 527
        TRY(run_set_bc(c, bc_mask_all));                          // 528
        return VOF_OK;
    }
    const bool props = (flags & VOF_STEP_MATERIALIZE_PROPS) != 0;
    const unsigned mask = props ? 31u : 15u;
    if (!props) {
        TileArgs ta; dim3 tgrid; size_t tsmem;
        if (tile_geometry(c, &ta, &tgrid, &tsmem)) return run_step_tile(c, istep, ta, tgrid, tsmem);
    }
    if (props) TRY(run_cal_nu_rho(c));
    TRY(run_kappa(c));
    TRY(run_advect(c, true));
    TRY(run_set_bc(c, mask));
    TRY(run_rhs(c, true));
    if (c->opt_pressure_solver == 1 && c->P.n_jacobi > 0) TRY(run_jacobi_cheb(c, c->P.n_jacobi));
    else if (use_jacobi_tb(c) && c->P.n_jacobi > 0) TRY(run_jacobi_tb(c, c->P.n_jacobi));
    else for (int s = 0; s < c->P.n_jacobi; ++s) TRY(run_jacobi_sweep(c, 0));
    TRY(run_project(c, true));
    TRY(run_set_bc(c, mask));
    if (istep % 2 == 0) { TRY(run_fct_y(c, false)); TRY(run_fct_x(c, true)); }
    else { TRY(run_fct_x(c, false)); TRY(run_fct_y(c, true)); }
    TRY(run_set_bc(c, mask));
    return VOF_OK;
}

extern "C" int vof2d_step(VofCtx* c, int istep, unsigned flags) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    return step_impl(c, istep, flags);
}

extern "C" int vof2d_run(VofCtx* c, int istep0, int nsteps, unsigned flags) {
    CHECK_CTX(c);
    if (nsteps < 0) return fail(VOF_EINVAL, "nsteps must be >= 0");
    CU(cudaSetDevice(c->device));
    int istep = istep0;
    int left = nsteps;
    const int par = istep0 & 1, fk = (int)(flags & 3u);
    if (left >= 4 && !c->profiling) {
        // capture two consecutive steps (the FCT sweep order alternates with istep parity; after two
        // steps every ping-pong buffer is back where it started, so the graph can be replayed)
        // the captured launches hard-code the ping-pong buffers of capture time: a graph taken with another
        // (F_cur, p_cur) -- e.g. after a single vof2d_fct_x_sweep or one Jacobi sweep between two runs -- is stale
        const int cur = c->F_cur | (c->p_cur << 1);
        if (c->graph[par][fk] && c->graph_cur[par][fk] != cur) {
            cudaGraphExecDestroy(c->graph[par][fk]);
            c->graph[par][fk] = nullptr;
        }
        if (!c->graph[par][fk]) {
            const int F0 = c->F_cur, p0 = c->p_cur;
            const long long l0 = c->launches;
            cudaGraph_t gr = nullptr;
            CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            int rc = step_impl(c, istep, flags);
            if (rc == VOF_OK) rc = step_impl(c, istep + 1, flags);
            cudaError_t e = cudaStreamEndCapture(c->stream, &gr);
            c->F_cur = F0; c->p_cur = p0;   // capture did not execute anything
            c->graph_launches[par][fk] = c->launches - l0;
            c->launches = l0;
            if (rc != VOF_OK) { if (gr) cudaGraphDestroy(gr); return rc; }
            if (e != cudaSuccess) return fail((int)e, "stream capture failed: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&c->graph[par][fk], gr, 0);
            cudaGraphDestroy(gr);
            if (e != cudaSuccess) { c->graph[par][fk] = nullptr; return fail((int)e, "graph instantiate failed: %s", cudaGetErrorString(e)); }
            c->graph_cur[par][fk] = cur;
        }
        while (left >= 2) {
            CU(cudaGraphLaunch(c->graph[par][fk], c->stream));
            c->launches += c->graph_launches[par][fk];
            left -= 2; istep += 2;
        }
        c->rhs_valid = !(flags & VOF_STEP_NO_FUSION);
    }
    for (; left > 0; --left, ++istep) TRY(step_impl(c, istep, flags));
    return VOF_OK;
}

// ------------------------------------------------------------------------------------
// field access
// ------------------------------------------------------------------------------------
static float* field_dev(VofCtx* c, int field) {
    switch (field) {
        case VOF_F: return c->F();
        case VOF_U: return c->buf[BUF_U];
        case VOF_V: return c->buf[BUF_V];
        case VOF_P: return c->p();
        case VOF_RHO: return c->buf[BUF_RHO];
        case VOF_NU: return c->buf[BUF_NU];
        case VOF_KAPPA: return c->buf[BUF_KAPPA];
        case VOF_USTAR: return c->buf[BUF_US];
        case VOF_VSTAR: return c->buf[BUF_VS];
    }
    return nullptr;
}

extern "C" int vof2d_field_ptr(VofCtx* c, int field, float** dev, int64_t* pitch_elems, int64_t* rows) {
    CHECK_CTX(c);
    float* d = field_dev(c, field);
    if (!d) return fail(VOF_EINVAL, "unknown 2-D field id %d", field);
    if (dev) *dev = d;
    if (pitch_elems) *pitch_elems = c->g.pitch;
    if (rows) *rows = c->g.nrows;
    return VOF_OK;
}

extern "C" int vof2d_field_get(VofCtx* c, int field, float* host_dst) {
    CHECK_CTX(c);
    float* d = field_dev(c, field);
    if (!d || !host_dst) return fail(VOF_EINVAL, "bad field id %d or null destination", field);
    CU(cudaSetDevice(c->device));
    const size_t w = (size_t)(c->g.ny + 2) * sizeof(float);
    CU(cudaMemcpy2DAsync(host_dst, w, d, (size_t)c->g.pitch * sizeof(float), w, c->g.nrows, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VOF_OK;
}

// Non-stalling read for the output path (-s, 2dvof.py:563-571): snapshot on the compute stream, D2H on a side stream.
// `host_dst` (dense (rows, ny+2) fp32; pinned memory for a true overlap, see vof_pinned_alloc) is valid after
// vof2d_field_get_wait.  One read in flight per context: a second call waits for the first one's copy.
extern "C" int vof2d_field_get_async(VofCtx* c, int field, float* host_dst) {
    CHECK_CTX(c);
    float* d = field_dev(c, field);
    if (!d || !host_dst) return fail(VOF_EINVAL, "bad field id %d or null destination", field);
    CU(cudaSetDevice(c->device));
    if (!c->aget) c->aget = new vofhost::AsyncGet();
    return vofhost::async_get_begin(*c->aget, c->stream, d - kColOff, c->field_bytes, (size_t)c->g.pitch * sizeof(float),
                                    (size_t)(c->g.ny + 2) * sizeof(float), (size_t)c->g.nrows, (size_t)kColOff * sizeof(float), host_dst);
}
extern "C" int vof2d_field_get_wait(VofCtx* c) {
    CHECK_CTX(c);
    if (!c->aget) return VOF_OK;
    CU(cudaSetDevice(c->device));
    return vofhost::async_get_wait(*c->aget);
}
extern "C" int vof_pinned_alloc(size_t bytes, void** out) {
    if (!out) return fail(VOF_EINVAL, "null out pointer");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(VOF_ENOMEM, "cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return VOF_OK;
}
extern "C" int vof_pinned_free(void* ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return VOF_OK;
}

static int field_set_async(VofCtx* c, int field, const float* host_src) {
    float* d = field_dev(c, field);
    if (!d || !host_src) return fail(VOF_EINVAL, "bad field id %d or null source", field);
    const size_t w = (size_t)(c->g.ny + 2) * sizeof(float);
    CU(cudaMemcpy2DAsync(d, (size_t)c->g.pitch * sizeof(float), host_src, w, w, c->g.nrows, cudaMemcpyHostToDevice, c->stream));
    return VOF_OK;
}

extern "C" int vof2d_field_set(VofCtx* c, int field, const float* host_src) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    TRY(field_set_async(c, field, host_src));
    CU(cudaStreamSynchronize(c->stream));
    return VOF_OK;
}

extern "C" int vof2d_field_fill(VofCtx* c, int field, float value) {
    CHECK_CTX(c);
    float* d = field_dev(c, field);
    if (!d) return fail(VOF_EINVAL, "unknown 2-D field id %d", field);
    CU(cudaSetDevice(c->device));
    std::vector<float> row((size_t)c->g.ny + 2, value);
    std::vector<float> all((size_t)c->g.nrows * (c->g.ny + 2));
    for (int i = 0; i < c->g.nrows; ++i) memcpy(&all[(size_t)i * (c->g.ny + 2)], row.data(), row.size() * sizeof(float));
    return vof2d_field_set(c, field, all.data());
}

extern "C" int vof2d_step_host(VofCtx* c, int istep, unsigned flags, const float* u_in, const float* v_in,
                               const float* p_in, const float* F_in, float* u_out, float* v_out, float* p_out,
                               float* F_out) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    if (F_in) TRY(field_set_async(c, VOF_F, F_in));
    if (u_in) TRY(field_set_async(c, VOF_U, u_in));
    if (v_in) TRY(field_set_async(c, VOF_V, v_in));
    if (p_in) TRY(field_set_async(c, VOF_P, p_in));
    TRY(step_impl(c, istep, flags));
    const size_t w = (size_t)(c->g.ny + 2) * sizeof(float), dp = (size_t)c->g.pitch * sizeof(float);
    const int ids[4] = {VOF_U, VOF_V, VOF_P, VOF_F};
    float* outs[4] = {u_out, v_out, p_out, F_out};
    for (int k = 0; k < 4; ++k)
        if (outs[k]) CU(cudaMemcpy2DAsync(outs[k], w, field_dev(c, ids[k]), dp, w, c->g.nrows, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VOF_OK;
}

// ------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------
extern "C" int vof2d_diagnostics(VofCtx* c, double* mass, float* max_cfl, float* residual, int64_t* courant_count) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    const int do_resid = residual != nullptr;
    if (do_resid && !c->rhs_valid) TRY(run_rhs(c, false));   // sequence mode: rho array holds the step's densities
    // mass / cfl / residual words are reset; courant_count is kept (written by the last projection)
    CU(cudaMemsetAsync(c->diag, 0, offsetof(Diag, courant_count), c->stream));
    const int a = std::max(c->lo - c->g.gi0, 0), b = c->hi - c->g.gi0;   // owned rows only
    dim3 grid(std::min(cdiv(c->g.ny, 256), 64), std::min(b - a + 1, 256));
    k_diag<<<grid, 256, 0, c->stream>>>(c->g, c->k, c->F(), c->buf[BUF_U], c->buf[BUF_V], c->p(), c->buf[BUF_RHS], c->diag, a, b, do_resid);
    TRY(launch_ok("k_diag"));
    Diag h;
    CU(cudaMemcpyAsync(&h, c->diag, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (mass) *mass = h.mass;
    if (max_cfl) memcpy(max_cfl, &h.max_cfl_bits, sizeof(float));
    if (residual) memcpy(residual, &h.resid_bits, sizeof(float));
    if (courant_count) *courant_count = (int64_t)h.courant_count;
    return VOF_OK;
}

// ------------------------------------------------------------------------------------
// NVLink peer-to-peer halo exchange: see vof_p2p.cuh (one fused kernel per exchange).
// ------------------------------------------------------------------------------------
// `cur_from`: which ping-pong buffers (F_cur | p_cur << 1, tagged with the epoch) the neighbour on that side holds its
// live F and p in.  The push addresses the neighbour's buffers by MY F_cur / p_cur, i.e. it assumes the ranks run in
// lockstep; the tag turns a violation (a rank that called an extra sweep) into an error instead of silent garbage.
// ------------------------------------------------------------------------------------
// slabs: halo rows are contiguous (rows * pitch floats, pad columns included)
// ------------------------------------------------------------------------------------
extern "C" int vof2d_halo_rows(const VofCtx* c, int* rows_per_side, int64_t* floats_per_field_side) {
    CHECK_CTX(c);
    if (rows_per_side) *rows_per_side = c->H;
    if (floats_per_field_side) *floats_per_field_side = (int64_t)c->H * c->g.pitch;
    return VOF_OK;
}

extern "C" int vof2d_halo_ptr(VofCtx* c, int field, int side, int send, float** dev, int64_t* count) {
    CHECK_CTX(c);
    float* d = field_dev(c, field);
    if (!d || (side != 0 && side != 1)) return fail(VOF_EINVAL, "bad field id %d or side %d", field, side);
    if ((side == 0 && c->has_lo) || (side == 1 && c->has_hi))
        return fail(VOF_ESTATE, "side %d of this context is a physical wall, not a slab interface", side);
    const int H = c->H, n = c->g.nrows;
    int row;
    if (side == 0) row = send ? H : 0;                 // lower: send owned rows [H, 2H), receive into [0, H)
    else row = send ? n - 2 * H : n - H;               // upper: send [n-2H, n-H), receive into [n-H, n)
    if (dev) *dev = d + (size_t)row * c->g.pitch - kColOff;   // whole pitched rows, from the row start
    if (count) *count = (int64_t)H * c->g.pitch;
    return VOF_OK;
}

extern "C" int vof2d_halo_push(VofCtx* c, int field, int side, float* peer_halo_dst) {
    CHECK_CTX(c);
    float* src; int64_t n;
    TRY(vof2d_halo_ptr(c, field, side, 1, &src, &n));
    if (!peer_halo_dst) return fail(VOF_EINVAL, "null peer destination");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(peer_halo_dst, src, (size_t)n * sizeof(float), cudaMemcpyDefault, c->stream));
    return VOF_OK;
}

// ------------------------------------------------------------------------------------
// launch accounting / per-kernel timing (measurement support; not on the hot path)
// ------------------------------------------------------------------------------------
extern "C" int64_t vof2d_launch_count(const VofCtx* c) { return c ? (int64_t)c->launches : -1; }

extern "C" int vof2d_profile(VofCtx* c, int enable) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (auto& sp : *c->spans) { c->ev_pool->push_back(sp.a); c->ev_pool->push_back(sp.b); }
    c->spans->clear();
    c->profiling = enable != 0;
    c->prof_every = enable > 1 ? enable : 1;     // enable = k > 1: record the kernels of every k-th vof2d_step only
    c->prof_count = 0; c->prof_now = true;
    return VOF_OK;
}

extern "C" int vof2d_profile_read(VofCtx* c, int kind, double* ms_total, int64_t* spans) {
    CHECK_CTX(c);
    if (kind < 0 || kind >= VOF_K_COUNT) return fail(VOF_EINVAL, "bad kernel kind %d", kind);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    double tot = 0.0; int64_t n = 0;
    for (auto& sp : *c->spans) {
        if (sp.kind != kind) continue;
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, sp.a, sp.b));
        tot += ms; ++n;
    }
    if (ms_total) *ms_total = tot;
    if (spans) *spans = n;
    return VOF_OK;
}

// ------------------------------------------------------------------------------------
// tuning knobs (A/B measurements; defaults are the fast paths)
// ------------------------------------------------------------------------------------
extern "C" int vof2d_set_option(VofCtx* c, int option, int value) {
    CHECK_CTX(c);
    switch (option) {
        case VOF_OPT_JACOBI_TB: if (value < 0 || value > 2) return fail(VOF_EINVAL, "jacobi_tb must be 0, 1 or 2"); c->opt_jacobi_tb = value; break;
        case VOF_OPT_ADVECT_COLS: if (value != 2 && value != 4) return fail(VOF_EINVAL, "advect columns per lane must be 2 or 4"); c->opt_advect_cols = value; break;
        case VOF_OPT_JACOBI_MAXT: if (value < 0 || value > 5) return fail(VOF_EINVAL, "jacobi sweeps per pass must be 0 (by grid size) or 1..5"); c->opt_jacobi_maxt = value; break;
        case VOF_OPT_JACOBI_ROWS: if (value < 0) return fail(VOF_EINVAL, "jacobi rows per item must be >= 0"); c->opt_jac_rows = value; break;
        case VOF_OPT_CHUNK_CAP: if (value < 0) return fail(VOF_EINVAL, "chunk cap must be >= 0"); c->opt_chunk_cap = value; break;
        case VOF_OPT_ADAPTIVE: if (value != 0 && value != 1) return fail(VOF_EINVAL, "adaptive must be 0 or 1"); c->opt_adaptive = value; break;
        case VOF_OPT_JACOBI_LONG_PCT: if (value < 0 || value > 100) return fail(VOF_EINVAL, "jacobi long-item share must be 0 .. 100"); c->opt_jac_long_pct = value; break;
        case VOF_OPT_PRESSURE_SOLVER: if (value != 0 && value != 1) return fail(VOF_EINVAL, "pressure solver must be 0 (Jacobi) or 1 (Chebyshev)"); c->opt_pressure_solver = value; break;
        case VOF_OPT_BARE_DIV: if (value != 0 && value != 1) return fail(VOF_EINVAL, "bare_div must be 0 or 1"); c->opt_bare_div = value; break;
        case VOF_OPT_FAST_MATH: if (value != 0 && value != 1) return fail(VOF_EINVAL, "fast_math must be 0 or 1"); c->opt_fast_math = value; break;
        case VOF_OPT_TILE: if (value < 0 || value > 2) return fail(VOF_EINVAL, "tile must be 0, 1 or 2"); c->opt_tile = value; break;
        case VOF_OPT_PACKED: if (value != 0 && value != 1) return fail(VOF_EINVAL, "packed must be 0 or 1"); c->opt_packed = value; break;
        case VOF_OPT_JACOBI_PK: if (value != 0 && value != 1) return fail(VOF_EINVAL, "jacobi_pk must be 0 or 1"); c->opt_jacobi_pk = value; break;
        case VOF_OPT_FCT_X_COLS: if (value != 2 && value != 4) return fail(VOF_EINVAL, "fct_x columns per lane must be 2 or 4"); c->opt_fct_x_cols = value; break;
        default: return fail(VOF_EINVAL, "unknown option %d", option);
    }
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 4; ++b)
            if (c->graph[a][b]) { cudaGraphExecDestroy(c->graph[a][b]); c->graph[a][b] = nullptr; }
    return VOF_OK;
}

// ------------------------------------------------------------------------------------
// P2P halo exchange API
// ------------------------------------------------------------------------------------
extern "C" int vof2d_p2p_export(VofCtx* c, void* handle64, int64_t* nrows, int64_t* arena_bytes) {
    CHECK_CTX(c);
    if (!c->own_arena) return fail(VOF_ESTATE, "p2p export needs a library-owned (cudaMalloc) arena");
    CU(cudaSetDevice(c->device));
    if (handle64) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, c->arena));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(handle64, &h, 64);
    }
    if (nrows) *nrows = c->g.nrows;
    if (arena_bytes) *arena_bytes = (int64_t)c->arena_bytes;
    return VOF_OK;
}

static size_t arena_bytes_for_rows(const VofCtx* c, long long nrows) {
    const size_t fb = field_stride_bytes((int)nrows, c->g.pitch);
    const size_t xy = ((size_t)(c->P.nx + 3 + c->P.ny + 3) * sizeof(float) + 255) / 256 * 256;
    return fb * BUF_COUNT + xy + 256;
}

extern "C" int vof2d_p2p_connect(VofCtx* c, int side, const void* handle64, void* same_process_arena, int64_t peer_nrows) {
    CHECK_CTX(c);
    if (side != 0 && side != 1) return fail(VOF_EINVAL, "side must be 0 or 1");
    if ((side == 0 && c->has_lo) || (side == 1 && c->has_hi)) return fail(VOF_ESTATE, "side %d is a physical wall", side);
    CU(cudaSetDevice(c->device));
    return p2p_connect(c->p2p, side, handle64, same_process_arena, peer_nrows, arena_bytes_for_rows(c, peer_nrows), c->stream);
}

extern "C" int vof2d_p2p_arena(VofCtx* c, void** arena) { CHECK_CTX(c); if (arena) *arena = c->arena; return VOF_OK; }

extern "C" int vof2d_halo_exchange_p2p(VofCtx* c) {
    CHECK_CTX(c);
    const bool nlo = !c->has_lo, nhi = !c->has_hi;
    if (!nlo && !nhi) return VOF_OK;
    CU(cudaSetDevice(c->device));
    Span span_(c, VOF_K_HALO, 1);
    P2PTable t;
    memset(&t, 0, sizeof(t));
    const int H = c->H, n = c->g.nrows, P = c->g.pitch;
    const int fields[4] = {BUF_U, BUF_V, c->p_cur ? BUF_P1 : BUF_P0, c->F_cur ? BUF_F1 : BUF_F0};
    const long long count4 = (long long)H * P / 4;     // pitch is a multiple of 32 floats
    for (int f = 0; f < 4; ++f) {
        const size_t my_off = c->field_bytes * fields[f];
        if (nlo) {   // my rows [H, 2H) -> lower neighbour's rows [n' - H, n')
            const size_t pfb = field_stride_bytes((int)c->p2p.peer_nrows[0], P);
            t.src[t.n] = (const float4*)(c->arena + my_off + (size_t)H * P * sizeof(float));
            t.dst[t.n] = (float4*)(c->p2p.peer_arena[0] + pfb * fields[f] + (size_t)(c->p2p.peer_nrows[0] - H) * P * sizeof(float));
            t.count4[t.n++] = count4;
        }
        if (nhi) {   // my rows [n - 2H, n - H) -> upper neighbour's rows [0, H)
            const size_t pfb = field_stride_bytes((int)c->p2p.peer_nrows[1], P);
            t.src[t.n] = (const float4*)(c->arena + my_off + (size_t)(n - 2 * H) * P * sizeof(float));
            t.dst[t.n] = (float4*)(c->p2p.peer_arena[1] + pfb * fields[f]);
            t.count4[t.n++] = count4;
        }
    }
    if ((nlo && !c->p2p.peer_arena[0]) || (nhi && !c->p2p.peer_arena[1])) return fail(VOF_ESTATE, "vof2d_p2p_connect was not called for every neighbour");
    return p2p_exchange(c->p2p, c->arena, c->arena_bytes, nlo, nhi, (unsigned int)(c->F_cur | (c->p_cur << 1)), t, c->stream);
}

extern "C" int vof2d_p2p_status(VofCtx* c, int* timed_out_epoch) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    return p2p_check(c->arena, c->arena_bytes, c->stream, timed_out_epoch, false);
}

// Error (VOF_ESTATE) if any halo exchange so far gave up waiting for a neighbour (20 s watchdog) or found the
// neighbour's ping-pong buffers out of step with this rank's.  Synchronises the stream: call it every N steps.
extern "C" int vof2d_p2p_check(VofCtx* c) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    return p2p_check(c->arena, c->arena_bytes, c->stream);
}

// ------------------------------------------------------------------------------------
// Streamed host-buffer step: the state lives in HOST memory (as it does for a caller of the reference who
// round-trips through to_numpy()/from_numpy(), 2dvof.py:535/565); the domain is cut into row slabs with the
// same deep halo the multi-GPU decomposition uses, and slab s+1 is uploaded while slab s is computed and
// slab s-1 is downloaded -- three streams, PCIe used in both directions at once.  Each slab's step is the
// ordinary slab step (bit-identical to the full-domain step, see tests), fed its halo rows from the host
// state instead of from a neighbour.
// ------------------------------------------------------------------------------------
// PCIe moves contiguous blocks faster than pitched rows (both directions at once on a B200 box: 45.0 against 40.5 GB/s each
// way, profiles/r2_micro/pcie_2d.cu -- a host row of ny + 2 floats starts 8 bytes further off a 64-byte boundary than the
// row before it), so the copies go between the dense host rows and a dense device staging block, and a kernel on the
// compute stream moves the rows between the staging block and the slab's pitched fields (device memory bandwidth: ~3 %
// of the PCIe time).  Without room for the staging blocks the copies are 2-D and direct.
struct VofStreamer {
    int device = 0, nx = 0, ny = 0, H = 0;
    std::vector<VofCtx*> slab;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_run;
    float* stage = nullptr;                    // per slab: 4 fields x uploaded rows, then 4 fields x downloaded rows, dense
    size_t stage_bytes = 0;
    std::vector<size_t> stage_in, stage_out;   // offsets (floats) of a slab's upload / download block
};

struct StageRows { float* dev[4]; float* stage[4]; };
// rows of four fields between dense staging rows (w floats) and pitched field rows; to_stage: field -> staging
__global__ void __launch_bounds__(256)
k_stage_rows(StageRows a, int pitch, int w, int to_stage) {
    const int f = blockIdx.z, row = blockIdx.y;
    float* d = a.dev[f] + (size_t)row * pitch;
    float* s = a.stage[f] + (size_t)row * w;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < w; j += gridDim.x * 256) {
        if (to_stage) s[j] = d[j]; else d[j] = s[j];
    }
}

extern "C" int vof2d_streamer_destroy(VofStreamer* st) {
    if (!st) return VOF_OK;
    cudaSetDevice(st->device);
    if (st->s_in) cudaStreamSynchronize(st->s_in);
    if (st->s_out) cudaStreamSynchronize(st->s_out);
    for (VofCtx* c : st->slab) vof2d_destroy(c);       // synchronises s_run; a caller-provided stream is left alone
    for (auto e : st->ev_in) cudaEventDestroy(e);
    for (auto e : st->ev_run) cudaEventDestroy(e);
    if (st->s_in) cudaStreamDestroy(st->s_in);
    if (st->s_run) cudaStreamDestroy(st->s_run);
    if (st->s_out) cudaStreamDestroy(st->s_out);
    if (st->stage) cudaFree(st->stage);
    delete st;
    return VOF_OK;
}

extern "C" int vof2d_streamer_create(const VofParams* p, int n_slabs, VofStreamer** out) {
    if (!out) return fail(VOF_EINVAL, "null out pointer");
    *out = nullptr;
    if (!p) return fail(VOF_EINVAL, "null params");
    if (p->slab_lo != 0 || p->slab_hi != 0) return fail(VOF_EINVAL, "the streamer takes full-domain params (slab_lo = slab_hi = 0)");
    const int H = std::max(p->halo, p->n_jacobi + 5);
    if (n_slabs < 1 || (n_slabs > 1 && p->nx / n_slabs < H))
        return fail(VOF_EINVAL, "%d slabs of a %d-row domain are thinner than the halo %d", n_slabs, p->nx, H);
    VofStreamer* st = new VofStreamer;
    st->nx = p->nx; st->ny = p->ny; st->H = H;
    int rc = VOF_OK;
    for (int s = 0; s < n_slabs && rc == VOF_OK; ++s) {
        VofParams q = *p;
        q.slab_lo = 1 + (int)((long long)p->nx * s / n_slabs);
        q.slab_hi = (int)((long long)p->nx * (s + 1) / n_slabs);
        q.halo = n_slabs > 1 ? H : std::max(p->halo, 1);
        VofCtx* c = nullptr;
        rc = vof2d_create(&q, &c);
        if (rc == VOF_OK) { st->slab.push_back(c); st->device = c->device; }
    }
    if (rc != VOF_OK) { vof2d_streamer_destroy(st); return rc; }
    cudaError_t e = cudaSetDevice(st->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st->s_run, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st->s_out, cudaStreamNonBlocking);
    st->ev_in.resize(n_slabs, nullptr); st->ev_run.resize(n_slabs, nullptr);
    for (int s = 0; s < n_slabs && e == cudaSuccess; ++s) {
        e = cudaEventCreateWithFlags(&st->ev_in[s], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&st->ev_run[s], cudaEventDisableTiming);
    }
    for (size_t s = 0; s < st->slab.size() && e == cudaSuccess; ++s)
        if (vof2d_set_stream(st->slab[s], st->s_run) != VOF_OK) e = cudaErrorUnknown;
    if (e != cudaSuccess) { vof2d_streamer_destroy(st); return fail((int)e, "streamer setup failed: %s", cudaGetErrorString(e)); }
    {   // staging blocks: the rows each slab uploads (halo included) and downloads (owned rows; wall slabs with their ghost row)
        const size_t w = (size_t)p->ny + 2;
        size_t total = 0;
        for (VofCtx* c : st->slab) {
            const size_t rows_in = (size_t)(c->all_b - c->all_a + 1);
            const size_t rows_out = (size_t)((c->has_hi ? st->nx + 1 : c->hi) - (c->has_lo ? 0 : c->lo) + 1);
            st->stage_in.push_back(total); total += 4 * rows_in * w;
            st->stage_out.push_back(total); total += 4 * rows_out * w;
        }
        if (cudaMalloc((void**)&st->stage, total * sizeof(float)) == cudaSuccess) st->stage_bytes = total * sizeof(float);
        else { st->stage = nullptr; (void)cudaGetLastError(); }          // no room: direct 2-D copies
    }
    *out = st;
    return VOF_OK;
}

extern "C" int vof2d_streamer_info(const VofStreamer* st, int* n_slabs, int* halo, size_t* device_bytes) {
    if (!st) return fail(VOF_EINVAL, "null streamer");
    if (n_slabs) *n_slabs = (int)st->slab.size();
    if (halo) *halo = st->H;
    if (device_bytes) { size_t b = st->stage_bytes; for (VofCtx* c : st->slab) b += c->arena_bytes; *device_bytes = b; }
    return VOF_OK;
}

// rows [ga, gb] (global) of one field between the dense host array and slab c
static int streamer_copy(VofCtx* c, int field, const float* h_src, float* h_dst, int ga, int gb, cudaStream_t s) {
    const size_t w = (size_t)(c->g.ny + 2) * sizeof(float), dp = (size_t)c->g.pitch * sizeof(float);
    float* d = field_dev(c, field) + (size_t)(ga - c->g.gi0) * c->g.pitch;
    if (h_src) CU(cudaMemcpy2DAsync(d, dp, h_src + (size_t)ga * (c->g.ny + 2), w, w, gb - ga + 1, cudaMemcpyHostToDevice, s));
    if (h_dst) CU(cudaMemcpy2DAsync(h_dst + (size_t)ga * (c->g.ny + 2), w, d, dp, w, gb - ga + 1, cudaMemcpyDeviceToHost, s));
    return VOF_OK;
}

extern "C" int vof2d_streamer_step_host(VofStreamer* st, int istep, unsigned flags, const float* u_in,
                                        const float* v_in, const float* p_in, const float* F_in, float* u_out,
                                        float* v_out, float* p_out, float* F_out) {
    if (!st) return fail(VOF_EINVAL, "null streamer");
    if (!u_in || !v_in || !p_in || !F_in || !u_out || !v_out || !p_out || !F_out)
        return fail(VOF_EINVAL, "the streamed step needs all four input and all four output arrays");
    CU(cudaSetDevice(st->device));
    const int S = (int)st->slab.size();
    const int ids[4] = {VOF_U, VOF_V, VOF_P, VOF_F};
    const float* ins[4] = {u_in, v_in, p_in, F_in};
    float* outs[4] = {u_out, v_out, p_out, F_out};
    const size_t w = (size_t)st->ny + 2;
    // does any output array overlap any input array?  (separate output arrays: a slab's download need not wait for the next
    // slab's upload, which shortens the tail of the pipeline by one slab's transfer)
    bool aliased = false;
    {
        const size_t bytes = (size_t)(st->nx + 2) * w * sizeof(float);
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) {
                const char* i0 = reinterpret_cast<const char*>(ins[a]);
                const char* o0 = reinterpret_cast<const char*>(outs[b]);
                if (i0 < o0 + bytes && o0 < i0 + bytes) aliased = true;
            }
    }
    // rows [ga, gb] of the four fields between slab c's pitched arrays and a dense staging block, on the compute stream
    auto stage_rows = [&](VofCtx* c, float* block, int ga, int gb, int to_stage) -> int {
        const int rows = gb - ga + 1;
        StageRows a;
        for (int k = 0; k < 4; ++k) {
            a.dev[k] = field_dev(c, ids[k]) + (size_t)(ga - c->g.gi0) * c->g.pitch;
            a.stage[k] = block + (size_t)k * rows * w;
        }
        k_stage_rows<<<dim3(std::min(8, cdiv((int)w, 256)), rows, 4), 256, 0, st->s_run>>>(a, c->g.pitch, (int)w, to_stage);
        return launch_ok("k_stage_rows");
    };
    auto out_rows = [&](VofCtx* c, int* ga, int* gb) { *ga = c->has_lo ? 0 : c->lo; *gb = c->has_hi ? st->nx + 1 : c->hi; };   // wall slabs own their ghost row
    auto download = [&](int s) -> int {
        VofCtx* c = st->slab[s];
        CU(cudaStreamWaitEvent(st->s_out, st->ev_run[s], 0));
        // in-place callers: slab s+1 reads my last H rows as its halo -- they may be overwritten only after that upload
        if (s + 1 < S && aliased) CU(cudaStreamWaitEvent(st->s_out, st->ev_in[s + 1], 0));
        int ga, gb; out_rows(c, &ga, &gb);
        if (st->stage) {
            const size_t n = (size_t)(gb - ga + 1) * w;
            for (int k = 0; k < 4; ++k)
                CU(cudaMemcpyAsync(outs[k] + (size_t)ga * w, st->stage + st->stage_out[s] + k * n, n * sizeof(float), cudaMemcpyDeviceToHost, st->s_out));
        } else {
            for (int k = 0; k < 4; ++k) TRY(streamer_copy(c, ids[k], nullptr, outs[k], ga, gb, st->s_out));
        }
        return VOF_OK;
    };
    for (int s = 0; s < S; ++s) {
        VofCtx* c = st->slab[s];
        const int ga = c->g.gi0 + c->all_a, gb = c->g.gi0 + c->all_b;
        if (st->stage) {
            const size_t n = (size_t)(gb - ga + 1) * w;
            for (int k = 0; k < 4; ++k)
                CU(cudaMemcpyAsync(st->stage + st->stage_in[s] + k * n, ins[k] + (size_t)ga * w, n * sizeof(float), cudaMemcpyHostToDevice, st->s_in));
        } else {
            for (int k = 0; k < 4; ++k) TRY(streamer_copy(c, ids[k], ins[k], nullptr, ga, gb, st->s_in));
        }
        CU(cudaEventRecord(st->ev_in[s], st->s_in));
        CU(cudaStreamWaitEvent(st->s_run, st->ev_in[s], 0));
        if (st->stage) TRY(stage_rows(c, st->stage + st->stage_in[s], ga, gb, 0));
        TRY(step_impl(c, istep, flags));
        if (st->stage) { int oa, ob; out_rows(c, &oa, &ob); TRY(stage_rows(c, st->stage + st->stage_out[s], oa, ob, 1)); }
        CU(cudaEventRecord(st->ev_run[s], st->s_run));
        if (s > 0) TRY(download(s - 1));
    }
    TRY(download(S - 1));
    CU(cudaStreamSynchronize(st->s_out));
    CU(cudaStreamSynchronize(st->s_run));
    return VOF_OK;
}
