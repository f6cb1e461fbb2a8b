/* libvof -- C ABI of the B200-native 2-D/3-D VOF hot path (sm_100a CUDA kernels behind it).
 *
 * The reference (houkensjtu/taichi-2d-vof) has no FFI: its "interface" for the per-timestep
 * path is (i) the zero-argument module-level kernels of 2dvof.py and the order the main loop
 * calls them in (2dvof.py:513-528), (ii) the module-global fp32 fields of shape (nx+2, ny+2)
 * (2dvof.py:53-89) read back with .to_numpy(), and (iii) the constants block 2dvof.py:19-50.
 * Each entry below names the reference interface it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative VOF_E* code for argument/state errors,
 *     or a positive cudaError_t; vof_last_error() gives a thread-local message.
 *   - all compute entries are ASYNCHRONOUS on the context's stream (vof2d_set_stream);
 *     only *_get / *_diagnostics / *_synchronize / *_step_host wait for the device.
 *   - caller-visible arrays are the reference's logical (nx+2, ny+2) fp32, C order (j fastest).
 *     On the device they live in a pitched allocation: element (i, j) of a field is
 *     dev[i * pitch + j] with dev/pitch from vof2d_field_ptr (dev + 1 is 128-byte aligned).
 *   - one context per (device, stream); a context is not re-entrant.
 *   - there is NO CPU fallback: creation fails if no sm_100-class CUDA device is usable.
 */
#ifndef VOF_H_
#define VOF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VOF_ABI_VERSION 1

enum {
    VOF_OK = 0,
    VOF_EINVAL = -1,   /* bad argument */
    VOF_ENODEV = -2,   /* no usable CUDA device (never falls back to the CPU) */
    VOF_ENOMEM = -3,
    VOF_ESTATE = -4    /* call not valid in this state (e.g. slab-only call on a full-domain ctx) */
};

/* Field ids: the live module-global fields of 2dvof.py:53-89 (scratch fields of the reference --
 * mx1..my4, mxsum, mysum, magnitude, Ftd, ax, ay, cx, cy, rp, rm -- are never materialised). */
enum {
    VOF_F = 0,       /* 2dvof.py:53  volume fraction                      */
    VOF_U = 1,       /* 2dvof.py:63  x-velocity on west faces             */
    VOF_V = 2,       /* 2dvof.py:64  y-velocity on south faces            */
    VOF_P = 3,       /* 2dvof.py:67  pressure                             */
    VOF_RHO = 4,     /* 2dvof.py:71                                       */
    VOF_NU = 5,      /* 2dvof.py:72  kinematic viscosity ("mu" in prose)  */
    VOF_KAPPA = 6,   /* 2dvof.py:88                                       */
    VOF_USTAR = 7,   /* 2dvof.py:65                                       */
    VOF_VSTAR = 8,   /* 2dvof.py:66                                       */
    VOF_W = 9,       /* 3dvof.py:88  (3-D contexts only)                  */
    VOF_WSTAR = 10,  /* 3dvof.py:91  (3-D contexts only)                  */
    VOF_FIELD_COUNT = 11
};

/* Constants block of 2dvof.py:19-50 / 3dvof.py:20-68.  Doubles, because the reference keeps them
 * as Python scalars and folds sub-expressions in double before rounding once to fp32. */
typedef struct VofParams {
    int32_t nx, ny, nz;        /* interior cells; nz = 0 for 2-D                              */
    double Lx, Ly, Lz;         /* 2dvof.py:22-23                                              */
    double dx, dy, dz;         /* 2dvof.py:47-48; <= 0 -> derived the reference way from L/n  */
    double dt;                 /* 2dvof.py:33                                                 */
    double rho_l, rho_g;       /* 2dvof.py:24-25                                              */
    double nu_l, nu_g;         /* 2dvof.py:26-27                                              */
    double sigma;              /* 2dvof.py:28-29 (run-time field in the reference)            */
    double gx, gy, gz;         /* 2dvof.py:30-31                                              */
    int32_t n_jacobi;          /* 2dvof.py:521 (10)                                           */
    /* Row-slab decomposition along i (new capability; the reference is single-address-space).
     * Full domain: slab_lo = 1, slab_hi = nx, halo = 1 (or all three 0 = same thing).         */
    int32_t slab_lo, slab_hi;  /* global interior rows owned by this context                  */
    int32_t halo;              /* ghost rows kept each side (>= 1); >= VOF_SLAB_MIN_HALO for slabs */
    int32_t device;            /* CUDA device ordinal, -1 = current                           */
} VofParams;

#define VOF_SLAB_MIN_HALO 15   /* n_jacobi + 5 at n_jacobi = 10: dependency radius of one step along i */

/* vof2d_step / vof2d_run flags */
#define VOF_STEP_MATERIALIZE_PROPS 1u  /* also write rho, nu (they are derived state: 2dvof.py:198-203) */
#define VOF_STEP_NO_FUSION         2u  /* run the reference's kernel sequence one entry at a time       */

typedef struct VofCtx VofCtx;

const char* vof_last_error(void);
int vof_abi_version(void);
void vof_default_params(VofParams* p);          /* the reference's constants, 2dvof.py:19-34 */

/* ---- lifetime (replaces: ti.init + field allocation, 2dvof.py:9, 53-89) ---- */
size_t vof2d_arena_bytes(const VofParams* p);   /* device bytes vof2d_create needs            */
int vof2d_create(const VofParams* p, VofCtx** out);
/* as above but carving the fields out of caller-owned device memory (e.g. a torch /
 * symmetric-memory allocation); arena must be 256-byte aligned and >= vof2d_arena_bytes */
int vof2d_create_in(const VofParams* p, void* arena, size_t arena_bytes, VofCtx** out);
int vof2d_destroy(VofCtx* c);
int vof2d_set_stream(VofCtx* c, void* cuda_stream);
int vof2d_synchronize(VofCtx* c);
int vof2d_get_params(const VofCtx* c, VofParams* out);   /* with dx/dy as resolved */

/* ---- one entry per reference kernel, same observable effect on the live fields ---- */
int vof2d_set_init_F(VofCtx* c, int ic);        /* 2dvof.py:137-159 (+find_area 102-134)      */
int vof2d_set_BC(VofCtx* c);                    /* 2dvof.py:162-189                           */
int vof2d_cal_nu_rho(VofCtx* c);                /* 2dvof.py:198-203                           */
int vof2d_get_normal_young(VofCtx* c);          /* 2dvof.py:283-309                           */
int vof2d_advect_upwind(VofCtx* c);             /* 2dvof.py:206-233                           */
int vof2d_solve_p_jacobi(VofCtx* c, int nsweeps);/* 2dvof.py:236-266, nsweeps calls in one    */
int vof2d_update_uv(VofCtx* c);                 /* 2dvof.py:269-280                           */
int vof2d_fct_x_sweep(VofCtx* c);               /* 2dvof.py:321-382                           */
int vof2d_fct_y_sweep(VofCtx* c);               /* 2dvof.py:385-448                           */
int vof2d_solve_VOF_rudman(VofCtx* c, int istep);/* 2dvof.py:312-318                          */
int vof2d_post_process_f(VofCtx* c);            /* 2dvof.py:452-455                           */
/* The stand-alone FCT variant of test/forward_fct.py: solve_VOF_rudman(t, eps_value) :254-264 = fct_y_sweep :310-351 and
 * fct_x_sweep :267-308 in the order of the step's parity, set_BC(F) :218-229 after each; F advected by the u, v fields as they
 * are (the script's init_uv / set_init_F are the caller's: field_set).  eps regularises the limiter ratios (:286, :293);
 * no var() clamp.  Full-domain contexts.  Bit-identical to the script run (tests/golden/ref_fct_*.npz). */
int vof2d_fct_forward(VofCtx* c, int istep, float eps);

/* ---- the loop body 2dvof.py:513-528 as one call (fused kernels, same result) ---- */
int vof2d_step(VofCtx* c, int istep, unsigned flags);
/* steps istep0 .. istep0+nsteps-1 back to back (CUDA-graph replay for small grids) */
int vof2d_run(VofCtx* c, int istep0, int nsteps, unsigned flags);
/* host-buffer form: H2D of u,v,p,F (each (nx+2)*(ny+2) floats), one step, D2H of u,v,p,F.
 * Synchronous.  Any pointer may be NULL to skip that transfer. */
int vof2d_step_host(VofCtx* c, int istep, unsigned flags,
                    const float* u_in, const float* v_in, const float* p_in, const float* F_in,
                    float* u_out, float* v_out, float* p_out, float* F_out);

/* ---- streamed host-buffer step (new): the state stays in HOST memory; the domain is cut into `n_slabs`
 * row slabs (deep halo, as for multi-GPU) and upload of slab s+1, the step of slab s and download of slab
 * s-1 overlap on three streams.  Same result as vof2d_step_host, bit for bit.  `p` are full-domain params.
 * The arrays are dense (nx+2)*(ny+2) floats; *_out may alias *_in (in-place: a slab's download then waits
 * for the next slab's upload of the rows it overwrites; with separate output arrays it does not, which
 * shortens the pipeline's tail by one slab's transfer).  Pinned host memory is needed for the overlap
 * (pageable memory works, serialised by the driver).  The PCIe copies are contiguous: they go through dense
 * staging blocks on the device (8 more floats per cell, counted in vof2d_streamer_info's device_bytes;
 * without room for them the copies are pitched and direct).  Synchronous.
 * Replaces the to_numpy()/from_numpy() round trip a host-resident caller of the reference pays
 * (2dvof.py:535, 565) around 2dvof.py:513-528. */
typedef struct VofStreamer VofStreamer;
int vof2d_streamer_create(const VofParams* p, int n_slabs, VofStreamer** out);
int vof2d_streamer_destroy(VofStreamer* st);
int vof2d_streamer_info(const VofStreamer* st, int* n_slabs, int* halo, size_t* device_bytes);
int vof2d_streamer_step_host(VofStreamer* st, int istep, unsigned flags,
                             const float* u_in, const float* v_in, const float* p_in, const float* F_in,
                             float* u_out, float* v_out, float* p_out, float* F_out);

/* ---- field access (replaces: field.to_numpy()/from_numpy(), 2dvof.py:535, 565) ---- */
int vof2d_field_ptr(VofCtx* c, int field, float** dev, int64_t* pitch_elems, int64_t* rows);
int vof2d_field_get(VofCtx* c, int field, float* host_dst);        /* logical rows of this ctx */
int vof2d_field_set(VofCtx* c, int field, const float* host_src);
int vof2d_field_fill(VofCtx* c, int field, float value);
/* Non-stalling read for the -s output path (F.to_numpy() at 2dvof.py:565 / 3dvof.py:627 blocks the reference's loop):
 * the field is snapshotted on the compute stream and copied to host_dst (pinned: vof_pinned_alloc) on a side stream
 * while the time loop continues; host_dst is valid after *_field_get_wait.  One read in flight per context. */
int vof2d_field_get_async(VofCtx* c, int field, float* host_dst);
int vof2d_field_get_wait(VofCtx* c);
int vof_pinned_alloc(size_t bytes, void** out);
int vof_pinned_free(void* ptr);

/* ---- diagnostics (new; the reference only prints on Courant violation, 2dvof.py:274-280) ----
 * mass = sum F over owned interior cells (fp64), max_cfl = max(|u|dt/dx, |v|dt/dy),
 * residual = L-inf of (rhs - A p) over owned interior cells, courant_count = number of faces
 * with u*dt > 0.25*dx (the reference's print condition).  Any pointer may be NULL.  Synchronous. */
/* Display kernels of the GUI loop (2dvof.py:458-492, called every 100 steps at 530-561): monitoring output only.
   rgb_buf is (2 nx, 2 ny) fp32, C order: rgb_buf[I] = field[I // 2], velocities divided by L / 0.2.  Full-domain
   contexts only (VOF_ESTATE on a slab). */
enum { VOF_VIEW_VOF = 0,    /* get_vof_field   2dvof.py:458-462 */
       VOF_VIEW_U = 1,      /* get_u_field     2dvof.py:465-470 */
       VOF_VIEW_V = 2,      /* get_v_field     2dvof.py:473-478 */
       VOF_VIEW_VNORM = 3   /* get_vnorm_field 2dvof.py:481-486 */ };
int vof2d_display_field(VofCtx* c, int view, float* rgb_host);      /* kernel + rgb_buf.to_numpy() (2dvof.py:535); synchronous */
int vof2d_display_field_dev(VofCtx* c, int view, float* rgb_dev);   /* the same into a device buffer, asynchronous */
int vof2d_interp_velocity(VofCtx* c, float* V_host);                /* interp_velocity 2dvof.py:489-492: (nx+2, ny+2, 2) fp32 */

int vof2d_diagnostics(VofCtx* c, double* mass, float* max_cfl, float* residual, int64_t* courant_count);

/* ---- measurement support (new): every kernel launch is counted; with profiling on, each launch
 * group is bracketed by CUDA events recorded on the context's stream (not usable under graph replay). */
enum {
    VOF_K_PROPS = 0, VOF_K_KAPPA, VOF_K_ADVECT, VOF_K_BC, VOF_K_RHS, VOF_K_JACOBI, VOF_K_PROJECT,
    VOF_K_FCT_X, VOF_K_FCT_Y, VOF_K_POST, VOF_K_HALO, VOF_K_TILE, VOF_K_COUNT
};
int64_t vof2d_launch_count(const VofCtx* c);
int vof2d_profile(VofCtx* c, int enable);                 /* 0 off, 1 every launch, k > 1 the launches of every k-th vof2d_step; always resets the spans */
int vof2d_profile_read(VofCtx* c, int kind, double* ms_total, int64_t* spans);  /* synchronous */

/* tuning knobs for A/B measurements (defaults = fast paths).  Results are identical either way -- except for the two opt-ins
 * that say otherwise: VOF_OPT_PRESSURE_SOLVER = 1 and VOF_OPT_FAST_MATH = 1. */
enum {
    VOF_OPT_JACOBI_TB = 0,    /* blocked Jacobi (several sweeps per HBM pass): 0 never, 1 from ~1800^2 up (default), 2 always */
    VOF_OPT_FCT_X_COLS = 1,   /* columns per lane of the x-sweep kernel: 2 (default) or 4 */
    VOF_OPT_ADVECT_COLS = 2,  /* columns per lane of the momentum predictor: 2 (default) or 4 */
    VOF_OPT_JACOBI_MAXT = 5,  /* sweeps per HBM pass of the blocked Jacobi at most: 0 (default) by grid size (3 up to ~5800^2,
                                 5 beyond), or 1 .. 5 */
    VOF_OPT_JACOBI_ROWS = 7,  /* > 0: rows per work item of the blocked Jacobi (tuning; default 0 = max(16 T, 48) in the second generation,
                                 the short items of the third: see VOF_OPT_JACOBI_LONG_PCT) */
    VOF_OPT_JACOBI_LONG_PCT = 8, /* third-generation Jacobi: percent of the rows cut into one long work item per resident warp (default 75;
                                  the rest becomes short items of VOF_OPT_JACOBI_ROWS rows, default max(8 T, 24)) */
    VOF_OPT_PRESSURE_SOLVER = 9, /* 0 (default): the reference's Jacobi sweeps (2dvof.py:236-266, 521-522); 1: the same number of
                                  sweeps of the Chebyshev semi-iterative acceleration of that iteration -- a stronger
                                  projection for the same traffic per sweep.  Changes p, u, v: outside parity mode */
    VOF_OPT_BARE_DIV = 13,    /* 1 (default): where vof2d_create proved the three-operation reciprocal division exact for EVERY fp32
                                  numerator of the Poisson diagonal (all 2^32 patterns, sub-normal quotients included), the packed
                                  Jacobi runs without the sub-normal test and its fp64 fix-up; 0: always with them (A/B); same bits */
    VOF_OPT_FAST_MATH = 12,   /* 0 (default): every operation rounded as the reference's fp32 expression (bit-exact).  1: tolerance mode
                                  of the blocked pressure sweeps -- fused multiply-adds and a multiply by the reciprocal of the
                                  diagonal, 5 instead of 8 operations per cell-update; within the north-star tolerances
                                  (rel. L-inf 1e-5 after one step, 1e-3 after 100), NOT bit-exact */
    VOF_OPT_TILE = 11,        /* whole-step tile kernel for small grids (one launch per step, the step's dependency radius as a
                                  shared-memory halo; csrc/vof2d_tile.cuh): 0 never, 1 (default) where the grid is launch
                                  bound (its blocks are one wave: up to ~520^2 on 148 SMs), 2 whenever the tile fits; same bits.
                                  The environment variable VOF_TILE overrides the default at vof2d_create (A/B runs, tests) */
    VOF_OPT_PACKED = 10,      /* 1 (default): Blackwell packed fp32x2 arithmetic (FFMA2) in the streaming kernels that have a packed
                                  variant (momentum predictor); 0: scalar arithmetic; same bits */
    VOF_OPT_JACOBI_PK = 6,    /* 1 (default): blocked Jacobi of the third generation (Blackwell packed fp32x2 arithmetic, c*p products,
                               * cp.async rings; vof2d_jacobi_pk.cuh;
                               * square cells only), 0: second generation; same bits */
    VOF_OPT_CHUNK_CAP = 4,    /* > 0: cap on the rows one warp marches per work item in the streaming kernels (tuning) */
    VOF_OPT_ADAPTIVE = 3      /* 1 (default): interface-adaptive FCT / curvature kernels -- bulk rows where F is uniform
                                 across a warp's strip take an exact short-cut, the x-sweep streams rows through a
                                 cp.async shared-memory ring; 0: the first-generation kernels.  Same bits either way. */
};
int vof2d_set_option(VofCtx* c, int option, int value);

/* ---- slabs (new): halo rows are contiguous runs of `pitch` floats.  Pack/unpack the rows the
 * neighbours need; the transport (NVLink P2P / NCCL) is the caller's.  side: 0 = lower i, 1 = upper. */
int vof2d_halo_rows(const VofCtx* c, int* rows_per_side, int64_t* floats_per_field_side);
int vof2d_halo_ptr(VofCtx* c, int field, int side, int send, float** dev, int64_t* count);
/* direct peer push: write my boundary rows into the neighbour's halo rows (peer-mapped memory) */
int vof2d_halo_push(VofCtx* c, int field, int side, float* peer_halo_dst);

/* NVLink peer-to-peer halo exchange (new): each rank maps its neighbours' arenas (CUDA IPC across processes, or a
 * plain pointer inside one process) and vof2d_halo_exchange_p2p() stores its boundary rows of u, v, p, F straight
 * into their halo rows with device-side flag hand-shakes -- stream-ordered, no NCCL, no host synchronisation.
 * Every rank of the decomposition must call it once per step (before vof2d_step). */
int vof2d_p2p_export(VofCtx* c, void* handle64, int64_t* nrows, int64_t* arena_bytes);   /* cudaIpcMemHandle_t */
int vof2d_p2p_connect(VofCtx* c, int side, const void* handle64, void* same_process_arena, int64_t peer_nrows);
int vof2d_p2p_arena(VofCtx* c, void** arena);
int vof2d_halo_exchange_p2p(VofCtx* c);
int vof2d_p2p_status(VofCtx* c, int* timed_out_epoch);   /* != 0: a wait gave up after 20 s (neighbour gone) */
/* VOF_ESTATE (with the reason in vof_last_error) if any exchange so far timed out, or found a neighbour whose live
 * F / p ping-pong buffers differ from this rank's (the push assumes lockstep).  Synchronises the stream. */
int vof2d_p2p_check(VofCtx* c);

/* =====================================================================================
 * 3-D twin: the loop body of 3dvof.py:598-623.  Fields are the reference's (nx+2, ny+2, nz+2) fp32
 * arrays, index [i, j, k], k contiguous (3dvof.py:70-117); on the device element (i, j, k) is
 * dev[i * pitch_j + j * pitch_k + k].  Same conventions as above.  3dvof.py never computes the
 * curvature (line 607 is commented out) and only -ic 1 sets F (126-138); both are reproduced.
 * ===================================================================================== */
typedef struct Vof3Ctx Vof3Ctx;
size_t vof3d_arena_bytes(const VofParams* p);
int vof3d_create(const VofParams* p, Vof3Ctx** out);       /* ti.init + fields, 3dvof.py:10, 70-117   */
int vof3d_destroy(Vof3Ctx* c);
int vof3d_set_stream(Vof3Ctx* c, void* cuda_stream);
int vof3d_synchronize(Vof3Ctx* c);
int vof3d_get_params(const Vof3Ctx* c, VofParams* out);
int vof3d_set_init_F(Vof3Ctx* c, int ic);                  /* 3dvof.py:126-138                         */
int vof3d_set_BC(Vof3Ctx* c);                              /* 3dvof.py:141-190                         */
int vof3d_cal_nu_rho(Vof3Ctx* c);                          /* 3dvof.py:199-204                         */
int vof3d_advect_upwind(Vof3Ctx* c);                       /* 3dvof.py:207-258                         */
int vof3d_solve_p_jacobi(Vof3Ctx* c, int nsweeps);         /* 3dvof.py:261-283                         */
int vof3d_update_uv(Vof3Ctx* c);                           /* 3dvof.py:286-302                         */
int vof3d_fct_x_sweep(Vof3Ctx* c);                         /* 3dvof.py:366-427                         */
int vof3d_fct_y_sweep(Vof3Ctx* c);                         /* 3dvof.py:430-492                         */
int vof3d_fct_z_sweep(Vof3Ctx* c);                         /* 3dvof.py:495-541                         */
int vof3d_solve_VOF_rudman(Vof3Ctx* c, int istep);         /* 3dvof.py:351-363                         */
int vof3d_post_process_f(Vof3Ctx* c);                      /* 3dvof.py:544-547                         */
int vof3d_step(Vof3Ctx* c, int istep, unsigned flags);     /* 3dvof.py:606-623                         */
int vof3d_run(Vof3Ctx* c, int istep0, int nsteps, unsigned flags);
int vof3d_field_ptr(Vof3Ctx* c, int field, float** dev, int64_t* pitch_k, int64_t* pitch_j, int64_t* planes);
int vof3d_field_get(Vof3Ctx* c, int field, float* host_dst);   /* F.to_numpy(), 3dvof.py:627         */
int vof3d_field_get_async(Vof3Ctx* c, int field, float* host_dst);   /* see vof2d_field_get_async */
int vof3d_field_get_wait(Vof3Ctx* c);
int vof3d_field_set(Vof3Ctx* c, int field, const float* host_src);
int vof3d_set_option(Vof3Ctx* c, int option, int value);      /* VOF_OPT_ADAPTIVE: 1 (default) second-generation kernels, 0 first */
int vof3d_diagnostics(Vof3Ctx* c, double* mass, float* max_cfl, int64_t* courant_count);
int64_t vof3d_launch_count(const Vof3Ctx* c);
int vof3d_halo_ptr(Vof3Ctx* c, int field, int side, int send, float** dev, int64_t* count);
int vof3d_halo_push(Vof3Ctx* c, int field, int side, float* peer_halo_dst);
/* NVLink peer-store exchange of the halo planes of u, v, w, p, F (same protocol as vof2d_*_p2p, one kernel per step) */
int vof3d_p2p_export(Vof3Ctx* c, void* handle64, int64_t* nrows, int64_t* arena_bytes);
int vof3d_p2p_connect(Vof3Ctx* c, int side, const void* handle64, void* same_process_arena, int64_t peer_nrows);
int vof3d_p2p_arena(Vof3Ctx* c, void** arena);
int vof3d_halo_exchange_p2p(Vof3Ctx* c);
int vof3d_p2p_check(Vof3Ctx* c);

#ifdef __cplusplus
}
#endif
#endif /* VOF_H_ */
