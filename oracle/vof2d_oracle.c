/* CPU oracle (C twin) for the 2-D VOF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar restatement of /root/reference/2dvof.py:102-455 and the loop body 506-528, one
 * `#pragma omp parallel for` per top-level `for` of each @ti.kernel -- i.e. the execution
 * model of the reference's ti.cpu backend (one offloaded range-for per loop, barrier after
 * each), including all of its scratch arrays (mx1..my4, Ftd, ax, ay, cx, cy, rp, rm, pt)
 * and the per-sweep rhs recompute + pt->p copy.  It is (a) the second, independent
 * restatement that tests/ cross-check bit-for-bit against oracle/vof2d_oracle.py, and
 * (b) the "port" CPU baseline timed by bench.py (`cpu_baseline`, `--impl reference`).
 *
 * PINNED to the reference run: tests/test_reference_pin_cpu.py::test_c_oracle_equals_reference_run_2d
 * compares it bit for bit with tests/golden/ref_2d_*.npz -- the unmodified text of 2dvof.py executed
 * under oracle/refshim/taichi (taichi==1.4.1 itself cannot be installed in this image).  Arithmetic:
 * IEEE fp32, left-to-right, no FMA contraction (build with -ffp-contract=off), Python-scalar
 * sub-expressions folded in double and rounded once.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t nx, ny;
    double Lx, Ly, dx, dy, dt;
    double rho_l, rho_g, nu_l, nu_g, sigma, gx, gy;
    int32_t n_jacobi;
} OVofParams;

enum { OF_F = 0, OF_U, OF_V, OF_P, OF_RHO, OF_NU, OF_KAPPA, OF_USTAR, OF_VSTAR, OF_COUNT };

typedef struct {
    OVofParams P;
    int nx, ny, pitch;
    size_t n;
    float *x, *y;
    float *F, *Ftd, *ax, *ay, *cx, *cy, *rp, *rm;
    float *u, *v, *us, *vs, *p, *pt, *rho, *nu;
    float *mx1, *my1, *mx2, *my2, *mx3, *my3, *mx4, *my4, *mxsum, *mysum, *mx, *my, *kappa, *mag;
    /* constants as seen by the kernels: double-folded, rounded once */
    float dt, dx, dy, dxi, dyi, dxi2, dyi2, dxdy, dtdy, dtdx, m1_2dx, m1_2dy, i_dx_2, i_dy_2;
    float sigma, rho_l, rho_g, nu_l, nu_g, gx, gy, cflx, cfly;
    int istep;
    long courant_flags;
} OVof;

#define IDX(i, j) ((size_t)(i) * (size_t)pitch + (size_t)(j))
#define MAXF(a, b) ((a) > (b) ? (a) : (b))
#define MINF(a, b) ((a) < (b) ? (a) : (b))

/* 2dvof.py:192-195 */
static inline float var3(float a, float b, float c) {
    float s = (a + b) + c;
    float mx = MAXF(MAXF(a, b), c);
    float mn = MINF(MINF(a, b), c);
    return (s - mx) - mn;
}

static float *zalloc(size_t n) {
    float *p = (float *)aligned_alloc(64, ((n * sizeof(float) + 63) / 64) * 64);
    /* first touch in parallel so pages spread like a threaded runtime would */
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < n; ++k) p[k] = 0.0f;
    return p;
}

OVof *ovof2d_create(const OVofParams *P) {
    OVof *s = (OVof *)calloc(1, sizeof(OVof));
    s->P = *P;
    s->nx = P->nx; s->ny = P->ny; s->pitch = P->ny + 2;
    s->n = (size_t)(P->nx + 2) * (size_t)(P->ny + 2);
    float **all[] = {&s->F, &s->Ftd, &s->ax, &s->ay, &s->cx, &s->cy, &s->rp, &s->rm, &s->u, &s->v, &s->us, &s->vs,
                     &s->p, &s->pt, &s->rho, &s->nu, &s->mx1, &s->my1, &s->mx2, &s->my2, &s->mx3, &s->my3, &s->mx4,
                     &s->my4, &s->mxsum, &s->mysum, &s->mx, &s->my, &s->kappa, &s->mag};
    for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) *all[k] = zalloc(s->n);
    /* 2dvof.py:41-46: hstack(0, linspace(0, L, n+1), L).astype(f32); linspace = k*step, last = stop */
    s->x = (float *)calloc((size_t)P->nx + 3, sizeof(float));
    s->y = (float *)calloc((size_t)P->ny + 3, sizeof(float));
    double stepx = P->Lx / P->nx, stepy = P->Ly / P->ny;
    for (int k = 0; k <= P->nx; ++k) s->x[k + 1] = (float)(k == P->nx ? P->Lx : k * stepx);
    for (int k = 0; k <= P->ny; ++k) s->y[k + 1] = (float)(k == P->ny ? P->Ly : k * stepy);
    s->x[0] = 0.0f; s->x[P->nx + 2] = (float)P->Lx;
    s->y[0] = 0.0f; s->y[P->ny + 2] = (float)P->Ly;
    double dx = P->dx, dy = P->dy;
    if (!(dx > 0)) dx = (double)s->x[3] - (double)s->x[2];   /* 2dvof.py:47 */
    if (!(dy > 0)) dy = (double)s->y[3] - (double)s->y[2];
    s->P.dx = dx; s->P.dy = dy;
    double dxi = 1 / dx, dyi = 1 / dy;
    s->dt = (float)P->dt; s->dx = (float)dx; s->dy = (float)dy; s->dxi = (float)dxi; s->dyi = (float)dyi;
    s->dxi2 = (float)(dxi * dxi); s->dyi2 = (float)(dyi * dyi);
    s->dxdy = (float)(dx * dy); s->dtdy = (float)(P->dt * dy); s->dtdx = (float)(P->dt * dx);
    s->m1_2dx = (float)(-1 / (2 * dx)); s->m1_2dy = (float)(-1 / (2 * dy));
    s->i_dx_2 = (float)(1 / dx / 2); s->i_dy_2 = (float)(1 / dy / 2);
    s->sigma = (float)P->sigma; s->rho_l = (float)P->rho_l; s->rho_g = (float)P->rho_g;
    s->nu_l = (float)P->nu_l; s->nu_g = (float)P->nu_g; s->gx = (float)P->gx; s->gy = (float)P->gy;
    s->cflx = (float)(0.25 * dx); s->cfly = (float)(0.25 * dy);
    return s;
}

void ovof2d_destroy(OVof *s) {
    if (!s) return;
    float *all[] = {s->F, s->Ftd, s->ax, s->ay, s->cx, s->cy, s->rp, s->rm, s->u, s->v, s->us, s->vs, s->p, s->pt,
                    s->rho, s->nu, s->mx1, s->my1, s->mx2, s->my2, s->mx3, s->my3, s->mx4, s->my4, s->mxsum,
                    s->mysum, s->mx, s->my, s->kappa, s->mag, s->x, s->y};
    for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) free(all[k]);
    free(s);
}

float *ovof2d_field(OVof *s, int id) {
    switch (id) {
        case OF_F: return s->F; case OF_U: return s->u; case OF_V: return s->v; case OF_P: return s->p;
        case OF_RHO: return s->rho; case OF_NU: return s->nu; case OF_KAPPA: return s->kappa;
        case OF_USTAR: return s->us; case OF_VSTAR: return s->vs;
    }
    return NULL;
}
int ovof2d_istep(const OVof *s) { return s->istep; }
void ovof2d_set_istep(OVof *s, int istep) { s->istep = istep; }
long ovof2d_courant_flags(const OVof *s) { return s->courant_flags; }
int ovof2d_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* torchrun exports OMP_NUM_THREADS=1 to its workers: the timed CPU legs ask for the host's cores explicitly */
void ovof_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* 2dvof.py:102-134 */
static float find_area(const OVof *s, int i, int j, float cx, float cy, float r) {
    const float dx = s->dx, dy = s->dy;
    const float hdx = (float)(s->P.dx / 2), hdy = (float)(s->P.dy / 2);
    float xc = (float)(i - 1) * dx + hdx, yc = (float)(j - 1) * dy + hdy;
    float xl = xc - hdx, xr = xc + hdx, yd = yc - hdy, yu = yc + hdy;
#define DIST(xx, yy) sqrtf(((xx) - cx) * ((xx) - cx) + ((yy) - cy) * ((yy) - cy))
    float d_ct = DIST(xc, yc), d_lu = DIST(xl, yu), d_ld = DIST(xl, yd), d_ru = DIST(xr, yu), d_rd = DIST(xr, yd);
#undef DIST
    if (d_lu > r && d_ld > r && d_ru > r && d_rd > r) return 1.0f;
    if (d_lu < r && d_ld < r && d_ru < r && d_rd < r) return 0.0f;
    float a = 0.5f + 0.5f * (d_ct - r) / (float)(sqrt(2.0) * s->P.dx);
    return var3(a, 0.0f, 1.0f);
}

/* 2dvof.py:137-159 */
void ovof2d_set_init_F(OVof *s, int ic) {
    const int nx = s->nx, ny = s->ny, pitch = s->pitch;
    const double Lx = s->P.Lx, Ly = s->P.Ly;
    if (ic == 1) {
        const float x1 = 0.0f, x2 = (float)(Lx / 3), y1 = 0.0f, y2 = (float)(Ly / 2);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < nx + 2; ++i)
            for (int j = 0; j < ny + 2; ++j)
                if (s->x[i] >= x1 && s->x[i] <= x2 && s->y[j] >= y1 && s->y[j] <= y2) s->F[IDX(i, j)] = 1.0f;
    } else if (ic == 2) {
        const float r = (float)(Lx / 12), cx = (float)(Lx / 2), cy = 2.0f * r;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < nx + 2; ++i)
            for (int j = 0; j < ny + 2; ++j) s->F[IDX(i, j)] = find_area(s, i, j, cx, cy, r);
    } else if (ic == 3) {
        const float r = (float)(Lx / 12), cx = (float)(Lx / 2), cy = (float)Ly - 3.0f * r;
        const float ycut = (float)(Ly * 0.37);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < nx + 2; ++i)
            for (int j = 0; j < ny + 2; ++j) {
                float f = 1.0f - find_area(s, i, j, cx, cy, r);
                if (s->y[j] < ycut) f = 1.0f;
                s->F[IDX(i, j)] = f;
            }
    }
}

/* 2dvof.py:162-189 */
void ovof2d_set_BC(OVof *s) {
    const int nx = s->nx, ny = s->ny, pitch = s->pitch;
    float *u = s->u, *v = s->v, *F = s->F, *p = s->p, *rho = s->rho;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx + 2; ++i) {
        u[IDX(i, 0)] = u[IDX(i, 1)];
        v[IDX(i, 1)] = 0;
        F[IDX(i, 0)] = F[IDX(i, 1)];
        p[IDX(i, 0)] = p[IDX(i, 1)];
        rho[IDX(i, 0)] = rho[IDX(i, 1)];
        u[IDX(i, ny + 1)] = u[IDX(i, ny)];
        v[IDX(i, ny + 1)] = 0;
        F[IDX(i, ny + 1)] = F[IDX(i, ny)];
        p[IDX(i, ny + 1)] = p[IDX(i, ny)];
        rho[IDX(i, ny + 1)] = rho[IDX(i, ny)];
    }
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny + 2; ++j) {
        u[IDX(1, j)] = 0;
        v[IDX(0, j)] = v[IDX(1, j)];
        F[IDX(0, j)] = F[IDX(1, j)];
        p[IDX(0, j)] = p[IDX(1, j)];
        rho[IDX(0, j)] = rho[IDX(1, j)];
        u[IDX(nx + 1, j)] = 0;
        v[IDX(nx + 1, j)] = v[IDX(nx, j)];
        F[IDX(nx + 1, j)] = F[IDX(nx, j)];
        p[IDX(nx + 1, j)] = p[IDX(nx, j)];
        rho[IDX(nx + 1, j)] = rho[IDX(nx, j)];
    }
}

/* 2dvof.py:198-203 */
void ovof2d_cal_nu_rho(OVof *s) {
    const size_t n = s->n;
    const float rho_g = s->rho_g, rho_l = s->rho_l, nu_l = s->nu_l, nu_g = s->nu_g;
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < n; ++k) {
        float f = var3(0.0f, 1.0f, s->F[k]);
        s->rho[k] = rho_g * (1.0f - f) + rho_l * f;
        s->nu[k] = nu_l * f + nu_g * (1.0f - f);
    }
}

/* 2dvof.py:283-309 */
void ovof2d_get_normal_young(OVof *s) {
    const int nx = s->nx, ny = s->ny, pitch = s->pitch;
    const float kx = s->m1_2dx, ky = s->m1_2dy;
    const float *F = s->F;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) {
            size_t c = IDX(i, j);
            s->mx1[c] = kx * (((F[IDX(i + 1, j + 1)] + F[IDX(i + 1, j)]) - F[IDX(i, j + 1)]) - F[IDX(i, j)]);
            s->my1[c] = ky * (((F[IDX(i + 1, j + 1)] - F[IDX(i + 1, j)]) + F[IDX(i, j + 1)]) - F[IDX(i, j)]);
            s->mx2[c] = kx * (((F[IDX(i + 1, j)] + F[IDX(i + 1, j - 1)]) - F[IDX(i, j)]) - F[IDX(i, j - 1)]);
            s->my2[c] = ky * (((F[IDX(i + 1, j)] - F[IDX(i + 1, j - 1)]) + F[IDX(i, j)]) - F[IDX(i, j - 1)]);
            s->mx3[c] = kx * (((F[IDX(i, j)] + F[IDX(i, j - 1)]) - F[IDX(i - 1, j)]) - F[IDX(i - 1, j - 1)]);
            s->my3[c] = ky * (((F[IDX(i, j)] - F[IDX(i, j - 1)]) + F[IDX(i - 1, j)]) - F[IDX(i - 1, j - 1)]);
            s->mx4[c] = kx * (((F[IDX(i, j + 1)] + F[IDX(i, j)]) - F[IDX(i - 1, j + 1)]) - F[IDX(i - 1, j)]);
            s->my4[c] = ky * (((F[IDX(i, j + 1)] - F[IDX(i, j)]) + F[IDX(i - 1, j + 1)]) - F[IDX(i - 1, j)]);
            s->mxsum[c] = (((s->mx1[c] + s->mx2[c]) + s->mx3[c]) + s->mx4[c]) / 4.0f;
            s->mysum[c] = (((s->my1[c] + s->my2[c]) + s->my3[c]) + s->my4[c]) / 4.0f;
            if (fabsf(s->mxsum[c]) < 1e-10f && fabsf(s->mysum[c]) < 1e-10f) {
                s->mx[c] = s->mxsum[c];
                s->my[c] = s->mysum[c];
            } else {
                s->mag[c] = sqrtf(s->mxsum[c] * s->mxsum[c] + s->mysum[c] * s->mysum[c]);
                s->mx[c] = s->mxsum[c] / s->mag[c];
                s->my[c] = s->mysum[c] / s->mag[c];
            }
        }
    const float ax_ = s->i_dx_2, ay_ = s->i_dy_2;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            s->kappa[IDX(i, j)] = -(ax_ * (s->mx[IDX(i + 1, j)] - s->mx[IDX(i - 1, j)]) +
                                    ay_ * (s->my[IDX(i, j + 1)] - s->my[IDX(i, j - 1)]));
}

/* 2dvof.py:206-233 */
void ovof2d_advect_upwind(OVof *s) {
    const int nx = s->nx, ny = s->ny, pitch = s->pitch;
    const float *u = s->u, *v = s->v, *F = s->F, *kappa = s->kappa, *nu = s->nu, *rho = s->rho;
    const float dt = s->dt, dxi = s->dxi, dyi = s->dyi, dxi2 = s->dxi2, dyi2 = s->dyi2;
    const float msig = -s->sigma, dx = s->dx, dy = s->dy, gx = s->gx, gy = s->gy;
#pragma omp parallel for schedule(static)
    for (int i = 2; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) {
            float uc = u[IDX(i, j)];
            float v_here = 0.25f * (((v[IDX(i - 1, j)] + v[IDX(i - 1, j + 1)]) + v[IDX(i, j)]) + v[IDX(i, j + 1)]);
            float dudx = uc > 0 ? (uc - u[IDX(i - 1, j)]) * dxi : (u[IDX(i + 1, j)] - uc) * dxi;
            float dudy = v_here > 0 ? (uc - u[IDX(i, j - 1)]) * dyi : (u[IDX(i, j + 1)] - uc) * dyi;
            float kappa_ave = (kappa[IDX(i, j)] + kappa[IDX(i - 1, j)]) / 2.0f;
            float fx_kappa = ((msig * (F[IDX(i, j)] - F[IDX(i - 1, j)])) * kappa_ave) / dx;
            float acc = (nu[IDX(i, j)] * ((u[IDX(i - 1, j)] - 2.0f * uc) + u[IDX(i + 1, j)])) * dxi2;
            acc = acc + (nu[IDX(i, j)] * ((u[IDX(i, j - 1)] - 2.0f * uc) + u[IDX(i, j + 1)])) * dyi2;
            acc = acc - uc * dudx;
            acc = acc - v_here * dudy;
            acc = acc + gx;
            acc = acc + (fx_kappa * 2.0f) / (rho[IDX(i, j)] + rho[IDX(i - 1, j)]);
            s->us[IDX(i, j)] = uc + dt * acc;
        }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 2; j <= ny; ++j) {
            float vc = v[IDX(i, j)];
            float u_here = 0.25f * (((u[IDX(i, j - 1)] + u[IDX(i, j)]) + u[IDX(i + 1, j - 1)]) + u[IDX(i + 1, j)]);
            float dvdx = u_here > 0 ? (vc - v[IDX(i - 1, j)]) * dxi : (v[IDX(i + 1, j)] - vc) * dxi;
            float dvdy = vc > 0 ? (vc - v[IDX(i, j - 1)]) * dyi : (v[IDX(i, j + 1)] - vc) * dyi;
            float kappa_ave = (kappa[IDX(i, j)] + kappa[IDX(i, j - 1)]) / 2.0f;
            float fy_kappa = ((msig * (F[IDX(i, j)] - F[IDX(i, j - 1)])) * kappa_ave) / dy;
            float acc = (nu[IDX(i, j)] * ((v[IDX(i - 1, j)] - 2.0f * vc) + v[IDX(i + 1, j)])) * dxi2;
            acc = acc + (nu[IDX(i, j)] * ((v[IDX(i, j - 1)] - 2.0f * vc) + v[IDX(i, j + 1)])) * dyi2;
            acc = acc - u_here * dvdx;
            acc = acc - vc * dvdy;
            acc = acc + gy;
            acc = acc + (fy_kappa * 2.0f) / (rho[IDX(i, j)] + rho[IDX(i, j - 1)]);
            s->vs[IDX(i, j)] = vc + dt * acc;
        }
}

/* 2dvof.py:236-266 -- one sweep */
void ovof2d_solve_p_jacobi(OVof *s) {
    const int nx = s->nx, ny = s->ny, pitch = s->pitch;
    const float *rho = s->rho, *us = s->us, *vs = s->vs;
    float *p = s->p, *pt = s->pt;
    const float dt = s->dt, dxi = s->dxi, dyi = s->dyi, dxi2 = s->dxi2, dyi2 = s->dyi2;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) {
            float rhs = (rho[IDX(i, j)] / dt) *
                        ((us[IDX(i + 1, j)] - us[IDX(i, j)]) * dxi + (vs[IDX(i, j + 1)] - vs[IDX(i, j)]) * dyi);
            float ae = i != nx ? dxi2 : 0.0f;
            float aw = i != 1 ? dxi2 : 0.0f;
            float an = j != ny ? dyi2 : 0.0f;
            float as = j != 1 ? dyi2 : 0.0f;
            float ap = -1.0f * (((ae + aw) + an) + as);
            float t = rhs - ae * p[IDX(i + 1, j)];
            t = t - aw * p[IDX(i - 1, j)];
            t = t - an * p[IDX(i, j + 1)];
            t = t - as * p[IDX(i, j - 1)];
            pt[IDX(i, j)] = t / ap;
        }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) p[IDX(i, j)] = pt[IDX(i, j)];
}

/* 2dvof.py:269-280 */
void ovof2d_update_uv(OVof *s) {
    const int nx = s->nx, ny = s->ny, pitch = s->pitch;
    const float *rho = s->rho, *p = s->p, *us = s->us, *vs = s->vs;
    const float dt = s->dt, dxi = s->dxi, dyi = s->dyi, cflx = s->cflx, cfly = s->cfly;
    long flags = 0;
#pragma omp parallel for schedule(static) reduction(+ : flags)
    for (int i = 2; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) {
            float r = (rho[IDX(i, j)] + rho[IDX(i - 1, j)]) * 0.5f;
            float un = us[IDX(i, j)] - ((dt / r) * (p[IDX(i, j)] - p[IDX(i - 1, j)])) * dxi;
            s->u[IDX(i, j)] = un;
            if (un * dt > cflx) flags++;
        }
#pragma omp parallel for schedule(static) reduction(+ : flags)
    for (int i = 1; i <= nx; ++i)
        for (int j = 2; j <= ny; ++j) {
            float r = (rho[IDX(i, j)] + rho[IDX(i, j - 1)]) * 0.5f;
            float vn = vs[IDX(i, j)] - ((dt / r) * (p[IDX(i, j)] - p[IDX(i, j - 1)])) * dyi;
            s->v[IDX(i, j)] = vn;
            if (vn * dt > cfly) flags++;
        }
    s->courant_flags = flags;
}

/* loops 2-4 are textually identical in both sweeps (2dvof.py:333-382 / 397-448) */
static void fct_limit_update(OVof *s, const float *vel, int along_x) {
    const int nx = s->nx, ny = s->ny, pitch = s->pitch;
    const int di = along_x ? 1 : 0, dj = along_x ? 0 : 1;
    float *F = s->F, *Ftd = s->Ftd, *ax = s->ax, *ay = s->ay, *cx = s->cx, *cy = s->cy, *rp = s->rp, *rm = s->rm;
    const float dt = s->dt, dx = s->dx, dy = s->dy, dxdy = s->dxdy;
    const float dtd = along_x ? s->dtdy : s->dtdx;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) {
            size_t c = IDX(i, j), m = IDX(i - di, j - dj), pl = IDX(i + di, j + dj);
            float fmax = MAXF(MAXF(Ftd[c], Ftd[m]), Ftd[pl]);
            float fmin = MINF(MINF(Ftd[c], Ftd[m]), Ftd[pl]);
            float vc = vel[c], vp = vel[pl];
            float lo_L = vc >= 0 ? (vc * dt) * F[m] : (vc * dt) * F[c];
            float hi_L = vp >= 0 ? (vp * dt) * F[c] : (vp * dt) * F[pl];
            float lo_H = vc <= 0 ? (vc * dt) * F[m] : (vc * dt) * F[c];
            float hi_H = vp <= 0 ? (vp * dt) * F[c] : (vp * dt) * F[pl];
            if (along_x) {
                ax[IDX(i + 1, j)] = hi_H - hi_L; ax[c] = lo_H - lo_L;
                ay[IDX(i, j + 1)] = 0; ay[c] = 0;
            } else {
                ax[IDX(i + 1, j)] = 0; ax[c] = 0;
                ay[IDX(i, j + 1)] = hi_H - hi_L; ay[c] = lo_H - lo_L;
            }
            /* the thread's own writes are what it reads back (same-value benign race in the reference) */
            float axc = along_x ? lo_H - lo_L : 0.0f, axp = along_x ? hi_H - hi_L : 0.0f;
            float ayc = along_x ? 0.0f : lo_H - lo_L, ayp = along_x ? 0.0f : hi_H - hi_L;
            float pp = ((MAXF(0.0f, axc) - MINF(0.0f, axp)) + MAXF(0.0f, ayc)) - MINF(0.0f, ayp);
            float qp = (fmax - Ftd[c]) * dx;
            rp[c] = pp > 0 ? MINF(1.0f, qp / pp) : 0.0f;
            float pm = ((MAXF(0.0f, axp) - MINF(0.0f, axc)) + MAXF(0.0f, ayp)) - MINF(0.0f, ayc);
            float qm = (Ftd[c] - fmin) * dx;
            rm[c] = pm > 0 ? MINF(1.0f, qm / pm) : 0.0f;
        }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) {
            size_t c = IDX(i, j), ip = IDX(i + 1, j), jp = IDX(i, j + 1);
            cx[ip] = ax[ip] >= 0 ? MINF(rp[ip], rm[c]) : MINF(rp[c], rm[ip]);
            cy[jp] = ay[jp] >= 0 ? MINF(rp[jp], rm[c]) : MINF(rp[c], rm[jp]);
        }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) {
            size_t c = IDX(i, j), ip = IDX(i + 1, j), jp = IDX(i, j + 1), pl = IDX(i + di, j + dj);
            float dv = dxdy - dtd * (vel[pl] - vel[c]);
            float t = ax[ip] * cx[ip] - ax[c] * cx[c];
            t = t + ay[jp] * cy[jp];
            t = t - ay[c] * cy[c];
            float fn = Ftd[c] - (((t / dy) * dx) * dy) / dv;
            F[c] = var3(0.0f, 1.0f, fn);
        }
}

static void fct_predict(OVof *s, const float *vel, int along_x) {
    const int nx = s->nx, ny = s->ny, pitch = s->pitch;
    const int di = along_x ? 1 : 0, dj = along_x ? 0 : 1;
    const float *F = s->F;
    float *Ftd = s->Ftd;
    const float dt = s->dt, dx = s->dx, dy = s->dy, dxdy = s->dxdy;
    const float dtd = along_x ? s->dtdy : s->dtdx;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) {
            size_t c = IDX(i, j), m = IDX(i - di, j - dj), pl = IDX(i + di, j + dj);
            float vc = vel[c], vp = vel[pl];
            float dv = dxdy - dtd * (vp - vc);
            float lo_L = vc >= 0 ? (vc * dt) * F[m] : (vc * dt) * F[c];
            float hi_L = vp >= 0 ? (vp * dt) * F[c] : (vp * dt) * F[pl];
            /* x: (fl_L - fr_L + fb_L - ft_L), y: (fl_L - fr_L + fb_L - ft_L) with the other pair = 0 */
            float sum = along_x ? ((lo_L - hi_L) + 0.0f) - 0.0f : ((0.0f - 0.0f) + lo_L) - hi_L;
            float t = ((((F[c] + (sum * dy) / dxdy) * dx) * dy)) / dv;
            if (t > 1.0f || t < 0) t = var3(0.0f, 1.0f, t);
            Ftd[c] = t;
        }
}

/* 2dvof.py:321-382 */
void ovof2d_fct_x_sweep(OVof *s) { fct_predict(s, s->u, 1); fct_limit_update(s, s->u, 1); }
/* 2dvof.py:385-448 */
void ovof2d_fct_y_sweep(OVof *s) { fct_predict(s, s->v, 0); fct_limit_update(s, s->v, 0); }

/* 2dvof.py:452-455 */
void ovof2d_post_process_f(OVof *s) {
    const size_t n = s->n;
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < n; ++k) s->F[k] = var3(s->F[k], 0.0f, 1.0f);
}

/* 2dvof.py:312-318 */
void ovof2d_solve_VOF_rudman(OVof *s) {
    if (s->istep % 2 == 0) { ovof2d_fct_y_sweep(s); ovof2d_fct_x_sweep(s); }
    else { ovof2d_fct_x_sweep(s); ovof2d_fct_y_sweep(s); }
}

/* 2dvof.py:506-528 */
void ovof2d_step(OVof *s) {
    s->istep += 1;
    ovof2d_cal_nu_rho(s);
    ovof2d_get_normal_young(s);
    ovof2d_advect_upwind(s);
    ovof2d_set_BC(s);
    for (int k = 0; k < s->P.n_jacobi; ++k) ovof2d_solve_p_jacobi(s);
    ovof2d_update_uv(s);
    ovof2d_set_BC(s);
    ovof2d_solve_VOF_rudman(s);
    ovof2d_post_process_f(s);
    ovof2d_set_BC(s);
}

void ovof2d_run(OVof *s, int nsteps) {
    for (int k = 0; k < nsteps; ++k) ovof2d_step(s);
}

double ovof2d_mass(const OVof *s) {
    const int nx = s->nx, ny = s->ny, pitch = s->pitch;
    double m = 0.0;
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) m += (double)s->F[IDX(i, j)];
    return m;
}
