// libvof C ABI, 3-D context (include/vof.h, vof3d_*): replaces the loop body of 3dvof.py:598-623.
#include "vof_host_common.h"
#include "vof3d_kernels.cuh"
#include "vof_p2p.cuh"

using namespace vof;
using vofhost::cdiv;
using vofhost::fail;
using vofhost::launch_ok;
using vofhost::node_coords;

enum { B3_F0 = 0, B3_F1, B3_U, B3_V, B3_W, B3_P0, B3_P1, B3_US, B3_VS, B3_WS, B3_RHS, B3_RHO, B3_NU, B3_COUNT };

constexpr int kJacRows3 = 16;  // planes one block of the second-generation 7-point sweep marches (tunable: VOF_OPT_CHUNK_CAP)

struct Vof3Ctx {
    VofParams P;
    Grid3 g;
    Consts3 k;
    Fct3C fct[3];
    Jac3C jac;
    int device;
    cudaStream_t stream;
    bool own_stream;
    char* arena;
    size_t arena_bytes, field_bytes;
    float* buf[B3_COUNT];
    int F_cur, p_cur;
    float *xs, *ys, *zs;
    Diag3* diag;
    int lo, hi, H;
    bool has_lo, has_hi;
    int all_a, all_b, in_a, in_b;
    long long launches;
    int sm_count, resident[8];  // resident blocks (whole device) of the queue-scheduled kernels, by variant
    int opt_jac_rows;          // planes one block of k3_jacobi5 marches
    bool bc_clean;             // every ghost cell is what set_BC would write now (true after a whole step; any other writer clears it)
    int opt_gen2;              // 1 (default): second-generation kernels, 0: first generation (same bits)
    P2PEndpoint p2p;           // neighbour arenas (lower / upper) mapped for the NVLink peer-store halo exchange
    vofhost::AsyncGet* aget;   // non-stalling field read for the VTK export path (vof3d_field_get_async)
    float* F() { return buf[F_cur ? B3_F1 : B3_F0]; }
    float* F_alt() { return buf[F_cur ? B3_F0 : B3_F1]; }
    float* p() { return buf[p_cur ? B3_P1 : B3_P0]; }
    float* p_alt() { return buf[p_cur ? B3_P0 : B3_P1]; }
};

static int resolve3(const VofParams* in, VofParams* P, Grid3* g) {
    if (!in) return fail(VOF_EINVAL, "null params");
    *P = *in;
    if (P->nx < 4 || P->ny < 4 || P->nz < 4) return fail(VOF_EINVAL, "vof3d_* needs nx, ny, nz >= 4 (got %d x %d x %d)", P->nx, P->ny, P->nz);
    if (!(P->Lx > 0) || !(P->Ly > 0) || !(P->Lz > 0) || !(P->dt > 0)) return fail(VOF_EINVAL, "Lx, Ly, Lz, dt must be positive");
    if (P->n_jacobi < 0) return fail(VOF_EINVAL, "n_jacobi must be >= 0");
    if (P->slab_lo == 0 && P->slab_hi == 0) { P->slab_lo = 1; P->slab_hi = P->nx; }
    if (P->halo == 0) P->halo = 1;
    if (P->slab_lo < 1 || P->slab_hi > P->nx || P->slab_lo > P->slab_hi)
        return fail(VOF_EINVAL, "slab planes [%d, %d] outside [1, %d]", P->slab_lo, P->slab_hi, P->nx);
    if (!(P->slab_lo == 1 && P->slab_hi == P->nx)) {
        const int need = P->n_jacobi + 4;   // no curvature in 3-D: F_new(i) <- u_new(i+3) <- rhs(i+2+n) <- u*(i+3+n) <- u(i+4+n)
        if (P->halo < need) return fail(VOF_EINVAL, "slab halo %d < n_jacobi + 4 = %d", P->halo, need);
        if (P->slab_hi - P->slab_lo + 1 < P->halo) return fail(VOF_EINVAL, "slab is thinner than its halo");
    }
    g->nx = P->nx; g->ny = P->ny; g->nz = P->nz;
    g->gi0 = P->slab_lo - P->halo;
    g->nrows = (P->slab_hi - P->slab_lo + 1) + 2 * P->halo;
    g->pk = round_up(kColOff + P->nz + 2, kPitchAlign);
    g->pj = (long long)(P->ny + 2) * g->pk;
    return VOF_OK;
}

static size_t field_bytes3(const Grid3& g) {
    size_t b = ((size_t)g.nrows * (size_t)g.pj + 64) * sizeof(float);   // + slack: shifted reads of the last k-row
    return (b + 255) / 256 * 256;
}

extern "C" size_t vof3d_arena_bytes(const VofParams* p) {
    VofParams P; Grid3 g{};
    if (resolve3(p, &P, &g) != VOF_OK) return 0;
    size_t xyz = ((size_t)(P.nx + P.ny + P.nz + 9) * sizeof(float) + 255) / 256 * 256;
    return field_bytes3(g) * B3_COUNT + xyz + 256;
}

extern "C" int vof3d_create(const VofParams* in, Vof3Ctx** out) {
    if (!out) return fail(VOF_EINVAL, "null out pointer");
    *out = nullptr;
    VofParams P; Grid3 g{};
    TRY(resolve3(in, &P, &g));
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
    if (prop.major < 10) return fail(VOF_ENODEV, "device %d is sm_%d%d; libvof is built for sm_100a only", dev, prop.major, prop.minor);

    Vof3Ctx* c = new (std::nothrow) Vof3Ctx();
    if (!c) return fail(VOF_ENOMEM, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->opt_gen2 = 1;
    c->opt_jac_rows = kJacRows3;
    c->device = dev; c->g = g;
    c->lo = P.slab_lo; c->hi = P.slab_hi; c->H = P.halo;
    c->has_lo = c->lo == 1; c->has_hi = c->hi == P.nx;
    std::vector<float> x, y, z;
    node_coords(x, P.nx, P.Lx); node_coords(y, P.ny, P.Ly); node_coords(z, P.nz, P.Lz);
    if (!(P.dx > 0)) P.dx = (double)x[3] - (double)x[2];
    if (!(P.dy > 0)) P.dy = (double)y[3] - (double)y[2];
    if (!(P.dz > 0)) P.dz = (double)z[3] - (double)z[2];     // 3dvof.py:65
    c->P = P;
    const double dx = P.dx, dy = P.dy, dz = P.dz, dt = P.dt, dxi = 1 / dx, dyi = 1 / dy, dzi = 1 / dz;
    Consts3& k = c->k;
    k.dt = (float)dt; k.dx = (float)dx; k.dy = (float)dy; k.dz = (float)dz;
    k.dxi = (float)dxi; k.dyi = (float)dyi; k.dzi = (float)dzi;
    k.dxi2 = (float)(dxi * dxi); k.dyi2 = (float)(dyi * dyi); k.dzi2 = (float)(dzi * dzi);
    k.vol = (float)(dx * dy * dz); k.dxdy = (float)(dx * dy);
    k.dt_yz = (float)(dt * dy * dz); k.dt_xz = (float)(dt * dx * dz); k.dt_xy = (float)(dt * dx * dy);
    k.rho_l = (float)P.rho_l; k.rho_g = (float)P.rho_g; k.nu_l = (float)P.nu_l; k.nu_g = (float)P.nu_g;
    k.gx = (float)P.gx; k.gy = (float)P.gy; k.gz = (float)P.gz;
    k.cflx = (float)(0.25 * dx); k.cfly = (float)(0.25 * dy);
    for (int iw = 0; iw < 2; ++iw)
        for (int jw = 0; jw < 2; ++jw)
            for (int kw = 0; kw < 2; ++kw) {   // -(ae + aw + an + a_s + ab + af), left to right (3dvof.py:274)
                const float ae = k.dxi2, aw = iw ? 0.0f : k.dxi2, an = k.dyi2, as = jw ? 0.0f : k.dyi2;
                const float ab = kw ? 0.0f : k.dzi2, af = k.dzi2;
                volatile float s = ae + aw; s = s + an; s = s + as; s = s + ab; s = s + af;
                k.ap[iw][jw][kw] = -1.0f * s;
            }
    c->jac.cx = k.dxi2; c->jac.cy = k.dyi2; c->jac.cz = k.dzi2;
    c->jac.dv = vofhost::make_const_div(k.ap[0][0][0]);
    c->jac.fast_div_ok = 0;
    for (int ax = 0; ax < 3; ++ax) {
        Fct3C& f = c->fct[ax];
        f.dt = k.dt; f.dx = k.dx; f.dy = k.dy; f.dz = k.dz; f.vol = k.vol;
        f.dtd = ax == 0 ? k.dt_yz : (ax == 1 ? k.dt_xz : k.dt_xy);
        f.m1 = k.dy;
        f.has_m2 = ax != 1; f.m2 = ax == 0 ? k.dz : k.dx;
        f.d1 = ax == 1 ? k.dxdy : k.vol;
        f.qs = ax == 2 ? k.dz : k.dx;
        f.d2 = ax == 2 ? k.dz : k.dy;
    }
    c->all_a = std::max(0, -g.gi0);
    c->all_b = std::min(g.nrows - 1, P.nx + 1 - g.gi0);
    c->in_a = std::max(0, 1 - g.gi0);
    c->in_b = std::min(g.nrows - 1, P.nx - g.gi0);
    const size_t need = vof3d_arena_bytes(&P);
    c->field_bytes = field_bytes3(g);
    e = cudaMalloc((void**)&c->arena, need);
    if (e != cudaSuccess) { delete c; return fail(VOF_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", need, cudaGetErrorString(e)); }
    c->arena_bytes = need;
    CU(cudaMemset(c->arena, 0, need));
    for (int b = 0; b < B3_COUNT; ++b) c->buf[b] = (float*)(c->arena + c->field_bytes * b) + kColOff;
    c->xs = (float*)(c->arena + c->field_bytes * B3_COUNT);
    c->ys = c->xs + (P.nx + 3);
    c->zs = c->ys + (P.ny + 3);
    c->diag = (Diag3*)(c->arena + need - 256);
    CU(cudaMemcpy(c->xs, x.data(), x.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->ys, y.data(), y.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->zs, z.data(), z.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    {   // prove the reciprocal division by the interior diagonal exact (every fp32 numerator against __fdiv_rn)
        unsigned long long* bad = &c->diag->courant_count;
        c->sm_count = prop.multiProcessorCount;
        k_check_div_by_const<<<prop.multiProcessorCount * 8, 256, 0, c->stream>>>(c->jac.dv, bad, 0);
        unsigned long long h = 1;
        CU(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMemsetAsync(bad, 0, sizeof(*bad), c->stream));
        c->jac.fast_div_ok = (h == 0);
        c->jac.bare_div_ok = 0;
        if (h == 0) {           // the bare three-operation form, every numerator again: no sub-normal test in the sweep if it holds
            k_check_div_by_const<<<prop.multiProcessorCount * 8, 256, 0, c->stream>>>(c->jac.dv, bad, 1);
            unsigned long long h2 = 1;
            CU(cudaMemcpyAsync(&h2, bad, sizeof(h2), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            CU(cudaMemsetAsync(bad, 0, sizeof(*bad), c->stream));
            c->jac.bare_div_ok = (h2 == 0);
        }
    }
    *out = c;
    return VOF_OK;
}

extern "C" int vof3d_destroy(Vof3Ctx* c) {
    if (!c) return VOF_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    p2p_close(c->p2p);
    if (c->aget) { vofhost::async_get_free(*c->aget); delete c->aget; }
    if (c->arena) cudaFree(c->arena);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return VOF_OK;
}

extern "C" int vof3d_set_stream(Vof3Ctx* c, void* cuda_stream) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return VOF_OK;
}
extern "C" int vof3d_synchronize(Vof3Ctx* c) { CHECK_CTX(c); CU(cudaStreamSynchronize(c->stream)); return VOF_OK; }
extern "C" int vof3d_get_params(const Vof3Ctx* c, VofParams* out) {
    CHECK_CTX(c);
    if (!out) return fail(VOF_EINVAL, "null out");
    *out = c->P;
    return VOF_OK;
}
extern "C" int64_t vof3d_launch_count(const Vof3Ctx* c) { return c ? (int64_t)c->launches : -1; }

// ------------------------------------------------------------------------------------ launches
constexpr int kRows3 = 16;     // planes marched by one block of the streaming kernels

static dim3 grid_jk(const Vof3Ctx* c, int kcount, int jcount, int planes, int per) { return dim3(cdiv(kcount, kB3), jcount, cdiv(planes, per)); }

static int run3_cal_nu_rho(Vof3Ctx* c) {
    ++c->launches;
    const int planes = c->all_b - c->all_a + 1;
    k3_cal_nu_rho<<<dim3(cdiv(c->g.nz + 2, kB3), c->g.ny + 2, std::min(planes, 64)), kB3, 0, c->stream>>>(c->g, c->k, c->F(), c->buf[B3_RHO], c->buf[B3_NU], c->all_a, c->all_b);
    return launch_ok("k3_cal_nu_rho");
}
static int run3_advect(Vof3Ctx* c, bool inl) {
    ++c->launches;
    const int a = std::max(c->in_a, 1), b = std::min(c->in_b, c->g.nrows - 2);
    dim3 grid = grid_jk(c, c->g.nz, c->g.ny, b - a + 1, kRows3);
#define A3 c->g, c->k, c->buf[B3_U], c->buf[B3_V], c->buf[B3_W], c->F(), c->buf[B3_RHO], c->buf[B3_NU], c->buf[B3_US], c->buf[B3_VS], c->buf[B3_WS], a, b, kRows3
    if (inl && c->opt_gen2) {
        dim3 g5(cdiv(c->g.nz, 128), cdiv(c->g.ny, 4), cdiv(b - a + 1, kRows3));
        k3_advect5<<<g5, 128, 0, c->stream>>>(c->g, c->k, c->buf[B3_U], c->buf[B3_V], c->buf[B3_W], c->F(), c->buf[B3_US], c->buf[B3_VS], c->buf[B3_WS],
                                              a, b, kRows3);
    } else if (inl) k3_advect<true><<<grid, kB3, 0, c->stream>>>(A3);
    else k3_advect<false><<<grid, kB3, 0, c->stream>>>(A3);
#undef A3
    return launch_ok("k3_advect");
}
static int run3_set_bc(Vof3Ctx* c, unsigned mask) {
    c->launches += 3;
    const int planes = c->all_b - c->all_a + 1;
#define B3ARGS c->g, c->buf[B3_U], c->buf[B3_V], c->buf[B3_W], c->F(), c->p(), c->buf[B3_RHO]
    k3_set_bc<<<dim3(cdiv(c->g.nz + 2, kB3), planes), kB3, 0, c->stream>>>(B3ARGS, 0, c->all_a, c->all_b, c->has_lo, c->has_hi, mask);
    if (c->has_lo || c->has_hi)
        k3_set_bc<<<dim3(cdiv(c->g.nz + 2, kB3), c->g.ny + 2), kB3, 0, c->stream>>>(B3ARGS, 1, c->all_a, c->all_b, c->has_lo, c->has_hi, mask);
    k3_set_bc<<<dim3(cdiv(c->g.ny + 2, kB3), planes), kB3, 0, c->stream>>>(B3ARGS, 2, c->all_a, c->all_b, c->has_lo, c->has_hi, mask);
#undef B3ARGS
    return launch_ok("k3_set_bc");
}
static int run3_rhs(Vof3Ctx* c, bool inl) {
    ++c->launches;
    const int a = c->in_a, b = std::min(c->in_b, c->g.nrows - 2);
    dim3 grid = grid_jk(c, c->g.nz, c->g.ny, b - a + 1, kRows3);
    if (inl && c->opt_gen2) {
        dim3 g5(cdiv(c->g.nz, 128), cdiv(c->g.ny, 4), cdiv(b - a + 1, kRows3));
        k3_rhs5<<<g5, 128, 0, c->stream>>>(c->g, c->k, c->F(), c->buf[B3_US], c->buf[B3_VS], c->buf[B3_WS], c->buf[B3_RHS], a, b, kRows3);
    } else if (inl) k3_rhs<true><<<grid, kB3, 0, c->stream>>>(c->g, c->k, c->F(), c->buf[B3_US], c->buf[B3_VS], c->buf[B3_WS], c->buf[B3_RHS], a, b, kRows3);
    else k3_rhs<false><<<grid, kB3, 0, c->stream>>>(c->g, c->k, c->buf[B3_RHO], c->buf[B3_US], c->buf[B3_VS], c->buf[B3_WS], c->buf[B3_RHS], a, b, kRows3);
    return launch_ok("k3_rhs");
}
static int run3_jacobi(Vof3Ctx* c, int mode) {
    ++c->launches;
    const int planes = c->all_b - c->all_a + 1;
    dim3 grid = grid_jk(c, c->g.nz + 2, c->g.ny + 2, planes, kRows3);
#define J3 c->g, c->k, c->p(), c->p_alt(), c->buf[B3_RHS], c->buf[B3_RHO], c->buf[B3_US], c->buf[B3_VS], c->buf[B3_WS], c->all_a, c->all_b, kRows3
    if (mode == 0) {
        if (c->opt_gen2) {
            const int rows = std::max(1, std::min(c->opt_jac_rows, planes));
            dim3 g5(cdiv(c->g.nz, 128), cdiv(c->g.ny + 2, 4), cdiv(planes, rows));
            if (c->jac.bare_div_ok) k3_jacobi5<true><<<g5, 128, 0, c->stream>>>(c->g, c->k, c->jac, c->p(), c->p_alt(), c->buf[B3_RHS], c->all_a, c->all_b, rows);
            else k3_jacobi5<false><<<g5, 128, 0, c->stream>>>(c->g, c->k, c->jac, c->p(), c->p_alt(), c->buf[B3_RHS], c->all_a, c->all_b, rows);
        } else {
            dim3 g4(cdiv(c->g.nz + 1, 128), cdiv(c->g.ny + 2, 4), cdiv(planes, kRows3));
            k3_jacobi4<<<g4, 128, 0, c->stream>>>(c->g, c->k, c->jac, c->p(), c->p_alt(), c->buf[B3_RHS], c->all_a, c->all_b, kRows3);
        }
    } else k3_jacobi<1><<<grid, kB3, 0, c->stream>>>(J3);
#undef J3
    c->p_cur ^= 1;
    return launch_ok("k3_jacobi");
}
static int run3_project(Vof3Ctx* c, bool inl) {
    ++c->launches;
    const int a = std::max(c->in_a, 1), b = c->in_b;
    dim3 grid = grid_jk(c, c->g.nz, c->g.ny, b - a + 1, kRows3);
    unsigned long long* cc = &c->diag->courant_count;
    CU(cudaMemsetAsync(cc, 0, sizeof(*cc), c->stream));
#define P3 c->g, c->k, inl ? c->F() : c->buf[B3_RHO], c->p(), c->buf[B3_US], c->buf[B3_VS], c->buf[B3_WS], c->buf[B3_U], c->buf[B3_V], c->buf[B3_W], cc, a, b, kRows3, c->lo - c->g.gi0, c->hi - c->g.gi0
    if (inl && c->opt_gen2) {
        dim3 g5(cdiv(c->g.nz, 128), cdiv(c->g.ny, 4), cdiv(b - a + 1, kRows3));
        k3_project5<<<g5, 128, 0, c->stream>>>(c->g, c->k, c->F(), c->p(), c->buf[B3_US], c->buf[B3_VS], c->buf[B3_WS], c->buf[B3_U], c->buf[B3_V],
                                               c->buf[B3_W], cc, a, b, kRows3, c->lo - c->g.gi0, c->hi - c->g.gi0);
    } else if (inl) k3_project<true><<<grid, kB3, 0, c->stream>>>(P3);
    else k3_project<false><<<grid, kB3, 0, c->stream>>>(P3);
#undef P3
    return launch_ok("k3_project");
}
// persistent launch of a queue-scheduled kernel (resident blocks asked once per variant)
template <typename K, typename... Args>
static void launch_queue3(Vof3Ctx* c, K kern, int slot, int warps_per_block, int nitems, Args... args) {
    if (!c->resident[slot]) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32 * warps_per_block, 0) != cudaSuccess) nb = 1;
        c->resident[slot] = std::max(1, nb) * c->sm_count;
    }
    const int blocks = std::min(cdiv(nitems, warps_per_block), c->resident[slot]);
    kern<<<blocks, 32 * warps_per_block, 0, c->stream>>>(args...);
}

static int run3_fct(Vof3Ctx* c, int axis, bool post) {
    ++c->launches;
    const float* vel = c->buf[axis == 0 ? B3_U : (axis == 1 ? B3_V : B3_W)];
    if (axis != 2 && c->opt_gen2) {
        // x- and y-sweeps on the second-generation 2-D x-sweep kernel (marching axis strided, columns = k): a batch of 2-D
        // problems -- the j-lines marching i (x), or the planes marching j (y)
        Grid g2{};
        g2.ny = c->g.nz;
        const int nstrips = cdiv(c->g.nz + 1, 64), rpc = 48;
        int r0, r1, nbatch, blo, bhi;
        long long bstride, boff;
        if (axis == 0) {
            g2.nx = c->g.nx; g2.gi0 = c->g.gi0; g2.nrows = c->g.nrows; g2.pitch = (int)c->g.pj;
            r0 = c->in_a; r1 = c->in_b;
            nbatch = c->g.ny + 2; bstride = c->g.pk; boff = 0; blo = 1; bhi = c->g.ny;
        } else {
            g2.nx = c->g.ny; g2.gi0 = 0; g2.nrows = c->g.ny + 2; g2.pitch = c->g.pk;
            r0 = 1; r1 = c->g.ny;
            nbatch = c->all_b - c->all_a + 1; bstride = c->g.pj; boff = (long long)c->all_a * c->g.pj;
            blo = std::max(0, 1 - (c->g.gi0 + c->all_a)); bhi = c->g.nx - (c->g.gi0 + c->all_a);
        }
        const int nitems = nstrips * nbatch * cdiv(r1 - r0 + 1, rpc);
        WorkQueue wq{c->diag->wq, nitems};
#define FX3 g2, c->fct[axis], wq, c->F() + boff, vel + boff, c->F_alt() + boff, r0, r1, rpc, nstrips, nbatch, bstride, blo, bhi
        if (post) launch_queue3(c, k_fct_x5<true, 2, FctOps3, true>, 2, kFctXWarps, nitems, FX3);
        else launch_queue3(c, k_fct_x5<false, 2, FctOps3, true>, 3, kFctXWarps, nitems, FX3);
#undef FX3
        c->F_cur ^= 1;
        return launch_ok("k_fct_x5 (3-D)");
    }
    if (axis == 0) {
        const int planes = c->in_b - c->in_a + 1, per = 64;
        dim3 grid(cdiv(c->g.nz + 2, kB3), c->g.ny + 2, cdiv(planes, per));
        if (post) k3_fct_strided<0, true><<<grid, kB3, 0, c->stream>>>(c->g, c->fct[0], c->F(), vel, c->F_alt(), c->in_a, c->in_b, per);
        else k3_fct_strided<0, false><<<grid, kB3, 0, c->stream>>>(c->g, c->fct[0], c->F(), vel, c->F_alt(), c->in_a, c->in_b, per);
    } else if (axis == 1) {
        const int planes = c->all_b - c->all_a + 1;
        dim3 grid(cdiv(c->g.nz + 2, kB3), planes);
        if (post) k3_fct_strided<1, true><<<grid, kB3, 0, c->stream>>>(c->g, c->fct[1], c->F(), vel, c->F_alt(), c->all_a, c->all_b, 0);
        else k3_fct_strided<1, false><<<grid, kB3, 0, c->stream>>>(c->g, c->fct[1], c->F(), vel, c->F_alt(), c->all_a, c->all_b, 0);
    } else if (c->opt_gen2) {
        // z-sweep on the second-generation 2-D y-sweep kernel: rows = all (i, j) lines of the local planes
        Grid g2{};
        g2.nx = c->g.nx; g2.ny = c->g.nz; g2.gi0 = c->g.gi0; g2.pitch = c->g.pk;
        const int rpp = c->g.ny + 2;
        g2.nrows = c->g.nrows * rpp;
        const int r0 = c->all_a * rpp, r1 = (c->all_b + 1) * rpp - 1;
        const int nstrips = cdiv(c->g.nz + 1, kFctYValid), rpw = 16;
        const int nitems = nstrips * cdiv(r1 - r0 + 1, rpw);
        WorkQueue wq{c->diag->wq, nitems};
        if (post) launch_queue3(c, k_fct_y5<true, FctOps3, true>, 0, kFctYWarps, nitems, g2, c->fct[2], wq, c->F(), vel, c->F_alt(), r0, r1, rpw, nstrips, rpp);
        else launch_queue3(c, k_fct_y5<false, FctOps3, true>, 1, kFctYWarps, nitems, g2, c->fct[2], wq, c->F(), vel, c->F_alt(), r0, r1, rpw, nstrips, rpp);
    } else {
        constexpr int TR = 8, TK = 128;
        const long long rows = (long long)(c->all_b - c->all_a + 1) * (c->g.ny + 2);
        dim3 grid(cdiv(c->g.nz, TK), (unsigned)((rows + TR - 1) / TR));
        if (grid.y > 65535u) return fail(VOF_EINVAL, "3-D z-sweep: %lld rows exceed the launch grid", rows);
        if (post) k3_fct_z<true, TR, TK><<<grid, 256, 0, c->stream>>>(c->g, c->fct[2], c->F(), vel, c->F_alt(), c->all_a, c->all_b);
        else k3_fct_z<false, TR, TK><<<grid, 256, 0, c->stream>>>(c->g, c->fct[2], c->F(), vel, c->F_alt(), c->all_a, c->all_b);
    }
    c->F_cur ^= 1;
    return launch_ok("k3_fct");
}
static int run3_post(Vof3Ctx* c) {
    ++c->launches;
    const int planes = c->all_b - c->all_a + 1;
    k3_post_process_f<<<dim3(cdiv(c->g.nz + 2, kB3), c->g.ny + 2, std::min(planes, 64)), kB3, 0, c->stream>>>(c->g, c->F(), c->all_a, c->all_b);
    return launch_ok("k3_post_process_f");
}
static void order3(int istep, int (&o)[3]) {   // 3dvof.py:351-363
    static const int tab[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}};
    const int r = ((istep % 3) + 3) % 3;
    for (int q = 0; q < 3; ++q) o[q] = tab[r][q];
}

// ------------------------------------------------------------------------------------ entries
extern "C" int vof3d_set_init_F(Vof3Ctx* c, int ic) {
    CHECK_CTX(c); c->bc_clean = false;
    if (ic < 1 || ic > 3) return fail(VOF_EINVAL, "ic must be 1, 2 or 3 (got %d)", ic);
    if (ic != 1) return VOF_OK;   // 3dvof.py:126-138: -ic 2/3 are accepted and leave F == 0
    ++c->launches;
    const int planes = c->all_b - c->all_a + 1;
    k3_set_init_F<<<dim3(cdiv(c->g.nz + 2, kB3), c->g.ny + 2, std::min(planes, 64)), kB3, 0, c->stream>>>(
        c->g, (float)(c->P.Lx / 3), (float)(c->P.Ly / 2), (float)(c->P.Lz / 3), c->xs, c->ys, c->zs, c->F(), c->all_a, c->all_b);
    return launch_ok("k3_set_init_F");
}
extern "C" int vof3d_set_BC(Vof3Ctx* c) { CHECK_CTX(c); c->bc_clean = false; return run3_set_bc(c, 63u); }
extern "C" int vof3d_cal_nu_rho(Vof3Ctx* c) { CHECK_CTX(c); c->bc_clean = false; return run3_cal_nu_rho(c); }
extern "C" int vof3d_advect_upwind(Vof3Ctx* c) { CHECK_CTX(c); c->bc_clean = false; return run3_advect(c, false); }
extern "C" int vof3d_solve_p_jacobi(Vof3Ctx* c, int nsweeps) {
    CHECK_CTX(c); c->bc_clean = false;
    if (nsweeps < 0) return fail(VOF_EINVAL, "nsweeps must be >= 0");
    if (nsweeps == 1) return run3_jacobi(c, 1);
    if (nsweeps == 0) return VOF_OK;
    TRY(run3_rhs(c, false));
    for (int s = 0; s < nsweeps; ++s) TRY(run3_jacobi(c, 0));
    return VOF_OK;
}
extern "C" int vof3d_update_uv(Vof3Ctx* c) { CHECK_CTX(c); c->bc_clean = false; return run3_project(c, false); }
extern "C" int vof3d_fct_x_sweep(Vof3Ctx* c) { CHECK_CTX(c); c->bc_clean = false; return run3_fct(c, 0, false); }
extern "C" int vof3d_fct_y_sweep(Vof3Ctx* c) { CHECK_CTX(c); c->bc_clean = false; return run3_fct(c, 1, false); }
extern "C" int vof3d_fct_z_sweep(Vof3Ctx* c) { CHECK_CTX(c); c->bc_clean = false; return run3_fct(c, 2, false); }
extern "C" int vof3d_solve_VOF_rudman(Vof3Ctx* c, int istep) {
    CHECK_CTX(c); c->bc_clean = false;
    int o[3]; order3(istep, o);
    for (int q = 0; q < 3; ++q) TRY(run3_fct(c, o[q], false));
    return VOF_OK;
}
extern "C" int vof3d_post_process_f(Vof3Ctx* c) { CHECK_CTX(c); c->bc_clean = false; return run3_post(c); }

extern "C" int vof3d_step(Vof3Ctx* c, int istep, unsigned flags) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    int o[3]; order3(istep, o);
    if (flags & VOF_STEP_NO_FUSION) {     // 3dvof.py:606-623, one launch group per reference kernel
        TRY(run3_cal_nu_rho(c));
        TRY(run3_advect(c, false));
        TRY(run3_set_bc(c, 63u));
        for (int s = 0; s < c->P.n_jacobi; ++s) TRY(run3_jacobi(c, 1));
        TRY(run3_project(c, false));
        TRY(run3_set_bc(c, 63u));
        for (int q = 0; q < 3; ++q) TRY(run3_fct(c, o[q], false));
        TRY(run3_post(c));
        TRY(run3_set_bc(c, 63u));
        c->bc_clean = true;
        return VOF_OK;
    }
    const bool props = (flags & VOF_STEP_MATERIALIZE_PROPS) != 0;
    const unsigned mask = props ? 63u : 31u;
    // set_BC is idempotent per field: a field whose interior has not changed since its ghosts were last filled keeps
    // them.  Inside one step only u, v, w, p change before the 2nd call and only F before the 3rd, and the 1st call
    // (after the predictor, which writes u*, v*, w* only) finds everything as the previous step's last call left it.
    // Full-domain contexts only (halo planes of a slab are written from outside), never with materialised rho, and
    // only while nothing but whole steps has touched the fields (bc_clean).
    const bool lean_bc = c->opt_gen2 && !props && c->lo == 1 && c->hi == c->g.nx;   // the whole domain in one context
    if (props) TRY(run3_cal_nu_rho(c));
    TRY(run3_advect(c, true));
    if (!(lean_bc && c->bc_clean)) TRY(run3_set_bc(c, mask));
    TRY(run3_rhs(c, true));
    for (int s = 0; s < c->P.n_jacobi; ++s) TRY(run3_jacobi(c, 0));
    TRY(run3_project(c, true));
    TRY(run3_set_bc(c, lean_bc ? (1u | 2u | 4u | 16u) : mask));
    for (int q = 0; q < 3; ++q) TRY(run3_fct(c, o[q], q == 2));   // post_process_f fused into the last sweep
    TRY(run3_set_bc(c, lean_bc ? 8u : mask));
    c->bc_clean = true;
    return VOF_OK;
}

extern "C" int vof3d_run(Vof3Ctx* c, int istep0, int nsteps, unsigned flags) {
    CHECK_CTX(c);
    for (int s = 0; s < nsteps; ++s) TRY(vof3d_step(c, istep0 + s, flags));
    return VOF_OK;
}

// ------------------------------------------------------------------------------------ field access
static float* field3(Vof3Ctx* c, int f) {
    switch (f) {
        case VOF_F: return c->F(); case VOF_U: return c->buf[B3_U]; case VOF_V: return c->buf[B3_V]; case VOF_W: return c->buf[B3_W];
        case VOF_P: return c->p(); case VOF_RHO: return c->buf[B3_RHO]; case VOF_NU: return c->buf[B3_NU];
        case VOF_USTAR: return c->buf[B3_US]; case VOF_VSTAR: return c->buf[B3_VS]; case VOF_WSTAR: return c->buf[B3_WS];
    }
    return nullptr;
}
extern "C" int vof3d_field_ptr(Vof3Ctx* c, int field, float** dev, int64_t* pitch_k, int64_t* pitch_j, int64_t* planes) {
    CHECK_CTX(c); c->bc_clean = false;
    float* d = field3(c, field);
    if (!d) return fail(VOF_EINVAL, "unknown 3-D field id %d", field);
    if (dev) *dev = d;
    if (pitch_k) *pitch_k = c->g.pk;
    if (pitch_j) *pitch_j = c->g.pj;
    if (planes) *planes = c->g.nrows;
    return VOF_OK;
}
extern "C" int vof3d_field_get(Vof3Ctx* c, int field, float* host_dst) {
    CHECK_CTX(c);
    float* d = field3(c, field);
    if (!d || !host_dst) return fail(VOF_EINVAL, "bad field id %d or null destination", field);
    CU(cudaSetDevice(c->device));
    const size_t wb = (size_t)(c->g.nz + 2) * sizeof(float);
    CU(cudaMemcpy2DAsync(host_dst, wb, d, (size_t)c->g.pk * sizeof(float), wb, (size_t)c->g.nrows * (c->g.ny + 2), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VOF_OK;
}
// non-stalling read for the VTK export (3dvof.py:624-627): see vof2d_field_get_async
extern "C" int vof3d_field_get_async(Vof3Ctx* c, int field, float* host_dst) {
    CHECK_CTX(c);
    float* d = field3(c, field);
    if (!d || !host_dst) return fail(VOF_EINVAL, "bad field id %d or null destination", field);
    CU(cudaSetDevice(c->device));
    if (!c->aget) c->aget = new vofhost::AsyncGet();
    return vofhost::async_get_begin(*c->aget, c->stream, d - kColOff, c->field_bytes, (size_t)c->g.pk * sizeof(float),
                                    (size_t)(c->g.nz + 2) * sizeof(float), (size_t)c->g.nrows * (c->g.ny + 2),
                                    (size_t)kColOff * sizeof(float), host_dst);
}
extern "C" int vof3d_field_get_wait(Vof3Ctx* c) {
    CHECK_CTX(c);
    if (!c->aget) return VOF_OK;
    CU(cudaSetDevice(c->device));
    return vofhost::async_get_wait(*c->aget);
}
extern "C" int vof3d_field_set(Vof3Ctx* c, int field, const float* host_src) {
    CHECK_CTX(c); c->bc_clean = false;
    float* d = field3(c, field);
    if (!d || !host_src) return fail(VOF_EINVAL, "bad field id %d or null source", field);
    CU(cudaSetDevice(c->device));
    const size_t wb = (size_t)(c->g.nz + 2) * sizeof(float);
    CU(cudaMemcpy2DAsync(d, (size_t)c->g.pk * sizeof(float), host_src, wb, wb, (size_t)c->g.nrows * (c->g.ny + 2), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VOF_OK;
}
extern "C" int vof3d_diagnostics(Vof3Ctx* c, double* mass, float* max_cfl, int64_t* courant_count) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->diag, 0, offsetof(Diag3, courant_count), c->stream));
    const int a = std::max(c->lo - c->g.gi0, 0), b = c->hi - c->g.gi0;
    k3_diag<<<dim3(1, std::min(c->g.ny, 128), std::min(b - a + 1, 128)), 256, 0, c->stream>>>(c->g, c->k, c->F(), c->buf[B3_U], c->buf[B3_V], c->buf[B3_W], c->diag, a, b);
    TRY(launch_ok("k3_diag"));
    Diag3 h;
    CU(cudaMemcpyAsync(&h, c->diag, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (mass) *mass = h.mass;
    if (max_cfl) memcpy(max_cfl, &h.max_cfl_bits, sizeof(float));
    if (courant_count) *courant_count = (int64_t)h.courant_count;
    return VOF_OK;
}

// halo planes are contiguous: `halo` whole i-planes = halo * pj floats per field per side
extern "C" int vof3d_halo_ptr(Vof3Ctx* c, int field, int side, int send, float** dev, int64_t* count) {
    CHECK_CTX(c); c->bc_clean = false;
    float* d = field3(c, field);
    if (!d || (side != 0 && side != 1)) return fail(VOF_EINVAL, "bad field id %d or side %d", field, side);
    if ((side == 0 && c->has_lo) || (side == 1 && c->has_hi)) return fail(VOF_ESTATE, "side %d of this context is a physical wall", side);
    const int H = c->H, n = c->g.nrows;
    const int plane = side == 0 ? (send ? H : 0) : (send ? n - 2 * H : n - H);
    if (dev) *dev = d + (size_t)plane * c->g.pj - kColOff;
    if (count) *count = (int64_t)H * c->g.pj;
    return VOF_OK;
}
extern "C" int vof3d_halo_push(Vof3Ctx* c, int field, int side, float* peer_halo_dst) {
    CHECK_CTX(c); c->bc_clean = false;
    float* src; int64_t n;
    TRY(vof3d_halo_ptr(c, field, side, 1, &src, &n));
    if (!peer_halo_dst) return fail(VOF_EINVAL, "null peer destination");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(peer_halo_dst, src, (size_t)n * sizeof(float), cudaMemcpyDefault, c->stream));
    return VOF_OK;
}

// ---- NVLink peer-to-peer exchange of the halo planes (vof_p2p.cuh): u, v, w, p, F in ONE kernel per step
static size_t arena_bytes_for_planes3(const Vof3Ctx* c, long long nrows) {
    Grid3 g = c->g; g.nrows = (int)nrows;
    const size_t xyz = ((size_t)(c->P.nx + c->P.ny + c->P.nz + 9) * sizeof(float) + 255) / 256 * 256;
    return field_bytes3(g) * B3_COUNT + xyz + 256;
}
extern "C" int vof3d_p2p_export(Vof3Ctx* c, void* handle64, int64_t* nrows, int64_t* arena_bytes) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    if (handle64) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, c->arena));
        memcpy(handle64, &h, 64);
    }
    if (nrows) *nrows = c->g.nrows;
    if (arena_bytes) *arena_bytes = (int64_t)c->arena_bytes;
    return VOF_OK;
}
extern "C" int vof3d_p2p_connect(Vof3Ctx* c, int side, const void* handle64, void* same_process_arena, int64_t peer_nrows) {
    CHECK_CTX(c);
    if (side != 0 && side != 1) return fail(VOF_EINVAL, "side must be 0 or 1");
    if ((side == 0 && c->has_lo) || (side == 1 && c->has_hi)) return fail(VOF_ESTATE, "side %d is a physical wall", side);
    CU(cudaSetDevice(c->device));
    return p2p_connect(c->p2p, side, handle64, same_process_arena, peer_nrows, arena_bytes_for_planes3(c, peer_nrows), c->stream);
}
extern "C" int vof3d_p2p_arena(Vof3Ctx* c, void** arena) { CHECK_CTX(c); if (arena) *arena = c->arena; return VOF_OK; }
extern "C" int vof3d_halo_exchange_p2p(Vof3Ctx* c) {
    CHECK_CTX(c);
    const bool nlo = !c->has_lo, nhi = !c->has_hi;
    if (!nlo && !nhi) return VOF_OK;
    if ((nlo && !c->p2p.peer_arena[0]) || (nhi && !c->p2p.peer_arena[1])) return fail(VOF_ESTATE, "vof3d_p2p_connect was not called for every neighbour");
    CU(cudaSetDevice(c->device));
    c->bc_clean = false;
    ++c->launches;
    P2PTable t;
    memset(&t, 0, sizeof(t));
    const int H = c->H, n = c->g.nrows;
    const size_t plane_bytes = (size_t)c->g.pj * sizeof(float);
    const int fields[5] = {B3_U, B3_V, B3_W, c->p_cur ? B3_P1 : B3_P0, c->F_cur ? B3_F1 : B3_F0};
    const long long count4 = (long long)H * c->g.pj / 4;       // pk (hence pj) is a multiple of 32 floats
    for (int f = 0; f < 5; ++f) {
        const size_t my_off = c->field_bytes * fields[f];
        for (int sd = 0; sd < 2; ++sd) {
            if (!(sd == 0 ? nlo : nhi)) continue;
            Grid3 pg = c->g; pg.nrows = (int)c->p2p.peer_nrows[sd];
            const size_t pfb = field_bytes3(pg);
            const long long src_plane = sd == 0 ? H : n - 2 * H;                     // my boundary planes
            const long long dst_plane = sd == 0 ? c->p2p.peer_nrows[0] - H : 0;      // the neighbour's halo planes
            t.src[t.n] = (const float4*)(c->arena + my_off + (size_t)src_plane * plane_bytes);
            t.dst[t.n] = (float4*)(c->p2p.peer_arena[sd] + pfb * fields[f] + (size_t)dst_plane * plane_bytes);
            t.count4[t.n++] = count4;
        }
    }
    return p2p_exchange(c->p2p, c->arena, c->arena_bytes, nlo, nhi, (unsigned int)(c->F_cur | (c->p_cur << 1)), t, c->stream);
}
extern "C" int vof3d_p2p_check(Vof3Ctx* c) {
    CHECK_CTX(c);
    CU(cudaSetDevice(c->device));
    return p2p_check(c->arena, c->arena_bytes, c->stream);
}

extern "C" int vof3d_set_option(Vof3Ctx* c, int option, int value) {
    if (!c) return fail(VOF_EINVAL, "null context");
    if (option == VOF_OPT_CHUNK_CAP) {
        if (value < 0) return fail(VOF_EINVAL, "chunk cap must be >= 0");
        c->opt_jac_rows = value > 0 ? value : kJacRows3;
        return VOF_OK;
    }
    if (option != VOF_OPT_ADAPTIVE) return fail(VOF_EINVAL, "3-D contexts know VOF_OPT_ADAPTIVE and VOF_OPT_CHUNK_CAP only (got %d)", option);
    if (value != 0 && value != 1) return fail(VOF_EINVAL, "adaptive must be 0 or 1");
    c->opt_gen2 = value;
    return VOF_OK;
}
