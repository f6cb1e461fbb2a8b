/* CPU oracle (C twin) for the 3-D VOF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar restatement of /root/reference/3dvof.py:126-302, 351-547 and the loop body 598-623, one
 * `#pragma omp parallel for` per top-level `for` of each @ti.kernel.  Cross-checked bit-for-bit
 * against oracle/vof3d_oracle.py by tests/; also the CPU baseline of the 3-D path.
 * PINNED to the reference run (tests/golden/ref_3d_*.npz = the unmodified 3dvof.py under oracle/refshim/taichi;
 * tests/test_reference_pin_cpu.py::test_c_oracle_equals_reference_run_3d, bit for bit).  Arithmetic: IEEE fp32,
 * left-to-right, -ffp-contract=off, Python-scalar sub-expressions folded in double.  kappa is never computed in 3dvof.py (607), so it stays 0.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t nx, ny, nz;
    double Lx, Ly, Lz, dx, dy, dz, dt;
    double rho_l, rho_g, nu_l, nu_g, sigma, gx, gy, gz;
    int32_t n_jacobi;
} OVof3Params;

enum { O3_F = 0, O3_U, O3_V, O3_W, O3_P, O3_RHO, O3_NU, O3_USTAR, O3_VSTAR, O3_WSTAR, O3_COUNT };

typedef struct {
    OVof3Params P;
    int nx, ny, nz;
    size_t sj, si, n;   /* strides: k contiguous, j stride = nz+2, i stride = (ny+2)(nz+2) */
    float *x, *y, *z;
    float *F, *Ftd, *a3[3], *cf3[3], *rp, *rm, *u, *v, *w, *us, *vs, *ws, *p, *pt, *rho, *nu, *kappa;
    float dt, dx, dy, dz, dxi, dyi, dzi, dxi2, dyi2, dzi2, vol, dxdy, dt_yz, dt_xz, dt_xy;
    float sigma, rho_l, rho_g, nu_l, nu_g, gx, gy, gz, cflx, cfly;
    int istep;
    long courant_flags;
} OVof3;

#define ID(i, j, k) ((size_t)(i) * s->si + (size_t)(j) * s->sj + (size_t)(k))
#define MAXF(a, b) ((a) > (b) ? (a) : (b))
#define MINF(a, b) ((a) < (b) ? (a) : (b))

static inline float var3(float a, float b, float c) {
    float sum = (a + b) + c;
    return (sum - MAXF(MAXF(a, b), c)) - MINF(MINF(a, b), c);
}

static float *zalloc3(size_t n) {
    float *p = (float *)aligned_alloc(64, ((n * sizeof(float) + 63) / 64) * 64);
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < n; ++k) p[k] = 0.0f;
    return p;
}

static void nodes(float *x, int n, double L) {
    double step = L / n;
    for (int k = 0; k <= n; ++k) x[k + 1] = (float)(k == n ? L : k * step);
    x[0] = 0.0f; x[n + 2] = (float)L;
}

OVof3 *ovof3d_create(const OVof3Params *P) {
    OVof3 *s = (OVof3 *)calloc(1, sizeof(OVof3));
    s->P = *P; s->nx = P->nx; s->ny = P->ny; s->nz = P->nz;
    s->sj = (size_t)P->nz + 2; s->si = ((size_t)P->ny + 2) * s->sj; s->n = ((size_t)P->nx + 2) * s->si;
    float **all[] = {&s->F, &s->Ftd, &s->a3[0], &s->a3[1], &s->a3[2], &s->cf3[0], &s->cf3[1], &s->cf3[2], &s->rp, &s->rm, &s->u, &s->v, &s->w, &s->us, &s->vs, &s->ws,
                     &s->p, &s->pt, &s->rho, &s->nu, &s->kappa};
    for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) *all[k] = zalloc3(s->n);
    s->x = (float *)calloc((size_t)P->nx + 3, sizeof(float));
    s->y = (float *)calloc((size_t)P->ny + 3, sizeof(float));
    s->z = (float *)calloc((size_t)P->nz + 3, sizeof(float));
    nodes(s->x, P->nx, P->Lx); nodes(s->y, P->ny, P->Ly); nodes(s->z, P->nz, P->Lz);
    double dx = P->dx > 0 ? P->dx : (double)s->x[3] - (double)s->x[2];
    double dy = P->dy > 0 ? P->dy : (double)s->y[3] - (double)s->y[2];
    double dz = P->dz > 0 ? P->dz : (double)s->z[3] - (double)s->z[2];
    s->P.dx = dx; s->P.dy = dy; s->P.dz = dz;
    double dxi = 1 / dx, dyi = 1 / dy, dzi = 1 / dz, dt = P->dt;
    s->dt = (float)dt; s->dx = (float)dx; s->dy = (float)dy; s->dz = (float)dz;
    s->dxi = (float)dxi; s->dyi = (float)dyi; s->dzi = (float)dzi;
    s->dxi2 = (float)(dxi * dxi); s->dyi2 = (float)(dyi * dyi); s->dzi2 = (float)(dzi * dzi);
    s->vol = (float)(dx * dy * dz); s->dxdy = (float)(dx * dy);
    s->dt_yz = (float)(dt * dy * dz); s->dt_xz = (float)(dt * dx * dz); s->dt_xy = (float)(dt * dx * dy);
    s->sigma = (float)P->sigma; s->rho_l = (float)P->rho_l; s->rho_g = (float)P->rho_g;
    s->nu_l = (float)P->nu_l; s->nu_g = (float)P->nu_g;
    s->gx = (float)P->gx; s->gy = (float)P->gy; s->gz = (float)P->gz;
    s->cflx = (float)(0.25 * dx); s->cfly = (float)(0.25 * dy);
    return s;
}

void ovof3d_destroy(OVof3 *s) {
    if (!s) return;
    float *all[] = {s->F, s->Ftd, s->a3[0], s->a3[1], s->a3[2], s->cf3[0], s->cf3[1], s->cf3[2], s->rp, s->rm, s->u, s->v, s->w, s->us, s->vs, s->ws, s->p, s->pt,
                    s->rho, s->nu, s->kappa, s->x, s->y, s->z};
    for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) free(all[k]);
    free(s);
}

float *ovof3d_field(OVof3 *s, int id) {
    switch (id) {
        case O3_F: return s->F; case O3_U: return s->u; case O3_V: return s->v; case O3_W: return s->w;
        case O3_P: return s->p; case O3_RHO: return s->rho; case O3_NU: return s->nu;
        case O3_USTAR: return s->us; case O3_VSTAR: return s->vs; case O3_WSTAR: return s->ws;
    }
    return NULL;
}
int ovof3d_istep(const OVof3 *s) { return s->istep; }
void ovof3d_set_istep(OVof3 *s, int v) { s->istep = v; }
long ovof3d_courant_flags(const OVof3 *s) { return s->courant_flags; }

/* 3dvof.py:126-138 */
void ovof3d_set_init_F(OVof3 *s, int ic) {
    if (ic != 1) return;
    const float x2 = (float)(s->P.Lx / 3), y2 = (float)(s->P.Ly / 2), z2 = (float)(s->P.Lz / 3);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s->nx + 2; ++i)
        for (int j = 0; j < s->ny + 2; ++j)
            for (int k = 0; k < s->nz + 2; ++k)
                if (s->x[i] >= 0.0f && s->x[i] <= x2 && s->y[j] >= 0.0f && s->y[j] <= y2 && s->z[k] >= 0.0f && s->z[k] <= z2)
                    s->F[ID(i, j, k)] = 1.0f;
}

/* 3dvof.py:141-190 */
void ovof3d_set_BC(OVof3 *s) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    float *u = s->u, *v = s->v, *w = s->w, *F = s->F, *p = s->p, *rho = s->rho;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx + 2; ++i)
        for (int k = 0; k < nz + 2; ++k) {
            u[ID(i, 0, k)] = u[ID(i, 1, k)]; v[ID(i, 1, k)] = 0; w[ID(i, 0, k)] = w[ID(i, 1, k)];
            F[ID(i, 0, k)] = F[ID(i, 1, k)]; p[ID(i, 0, k)] = p[ID(i, 1, k)]; rho[ID(i, 0, k)] = rho[ID(i, 1, k)];
            u[ID(i, ny + 1, k)] = u[ID(i, ny, k)]; v[ID(i, ny + 1, k)] = 0; w[ID(i, ny + 1, k)] = w[ID(i, ny, k)];
            F[ID(i, ny + 1, k)] = F[ID(i, ny, k)]; p[ID(i, ny + 1, k)] = p[ID(i, ny, k)]; rho[ID(i, ny + 1, k)] = rho[ID(i, ny, k)];
        }
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny + 2; ++j)
        for (int k = 0; k < nz + 2; ++k) {
            u[ID(1, j, k)] = 0; v[ID(0, j, k)] = v[ID(1, j, k)]; w[ID(0, j, k)] = w[ID(1, j, k)];
            F[ID(0, j, k)] = F[ID(1, j, k)]; p[ID(0, j, k)] = p[ID(1, j, k)]; rho[ID(0, j, k)] = rho[ID(1, j, k)];
            u[ID(nx + 1, j, k)] = 0; v[ID(nx + 1, j, k)] = v[ID(nx, j, k)]; w[ID(nx + 1, j, k)] = w[ID(nx, j, k)];
            F[ID(nx + 1, j, k)] = F[ID(nx, j, k)]; p[ID(nx + 1, j, k)] = p[ID(nx, j, k)]; rho[ID(nx + 1, j, k)] = rho[ID(nx, j, k)];
        }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx + 2; ++i)
        for (int j = 0; j < ny + 2; ++j) {
            u[ID(i, j, 0)] = u[ID(i, j, 1)]; v[ID(i, j, 0)] = v[ID(i, j, 1)]; w[ID(i, j, 1)] = 0;
            F[ID(i, j, 0)] = F[ID(i, j, 1)]; p[ID(i, j, 0)] = p[ID(i, j, 1)]; rho[ID(i, j, 0)] = rho[ID(i, j, 1)];
            u[ID(i, j, nz + 1)] = u[ID(i, j, nz)]; v[ID(i, j, nz + 1)] = v[ID(i, j, nz)]; w[ID(i, j, nz + 1)] = 0;
            F[ID(i, j, nz + 1)] = F[ID(i, j, nz)]; p[ID(i, j, nz + 1)] = p[ID(i, j, nz)]; rho[ID(i, j, nz + 1)] = rho[ID(i, j, nz)];
        }
}

/* 3dvof.py:199-204 */
void ovof3d_cal_nu_rho(OVof3 *s) {
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < s->n; ++k) {
        float f = var3(0.0f, 1.0f, s->F[k]);
        s->rho[k] = s->rho_g * (1.0f - f) + s->rho_l * f;
        s->nu[k] = s->nu_l * f + s->nu_g * (1.0f - f);
    }
}

/* 3dvof.py:207-258 */
void ovof3d_advect_upwind(OVof3 *s) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const size_t si = s->si, sj = s->sj;
    const float *u = s->u, *v = s->v, *w = s->w, *F = s->F, *kap = s->kappa, *nu = s->nu, *rho = s->rho;
    const float dt = s->dt, dxi = s->dxi, dyi = s->dyi, dzi = s->dzi, dxi2 = s->dxi2, dyi2 = s->dyi2, dzi2 = s->dzi2;
    const float msig = -s->sigma;
#pragma omp parallel for schedule(static)
    for (int i = 2; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            for (int k = 1; k <= nz; ++k) {
                size_t c = ID(i, j, k);
                float uc = u[c];
                float v_here = 0.25f * (((v[c - si] + v[c - si + sj]) + v[c]) + v[c + sj]);
                float w_here = 0.25f * (((w[c - si] + w[c - si + 1]) + w[c]) + w[c + 1]);
                float dudx = uc > 0 ? (uc - u[c - si]) * dxi : (u[c + si] - uc) * dxi;
                float dudy = v_here > 0 ? (uc - u[c - sj]) * dyi : (u[c + sj] - uc) * dyi;
                float dudz = w_here > 0 ? (uc - u[c - 1]) * dzi : (u[c + 1] - uc) * dzi;
                float kave = (kap[c] + kap[c - si]) / 2.0f;
                float fk = ((msig * (F[c] - F[c - si])) * kave) / s->dx;
                float acc = (nu[c] * ((u[c - si] - 2.0f * uc) + u[c + si])) * dxi2;
                acc = acc + (nu[c] * ((u[c - sj] - 2.0f * uc) + u[c + sj])) * dyi2;
                acc = acc + (nu[c] * ((u[c - 1] - 2.0f * uc) + u[c + 1])) * dzi2;
                acc = acc - uc * dudx; acc = acc - v_here * dudy; acc = acc - w_here * dudz;
                acc = acc + s->gx;
                acc = acc + (fk * 2.0f) / (rho[c] + rho[c - si]);
                s->us[c] = uc + dt * acc;
            }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 2; j <= ny; ++j)
            for (int k = 1; k <= nz; ++k) {
                size_t c = ID(i, j, k);
                float vc = v[c];
                float u_here = 0.25f * (((u[c - sj] + u[c]) + u[c + si - sj]) + u[c + si]);
                float w_here = 0.25f * (((w[c - sj + 1] + w[c - sj]) + w[c]) + w[c + 1]);
                float dvdx = u_here > 0 ? (vc - v[c - si]) * dxi : (v[c + si] - vc) * dxi;
                float dvdy = vc > 0 ? (vc - v[c - sj]) * dyi : (v[c + sj] - vc) * dyi;
                float dvdz = w_here > 0 ? (vc - v[c - 1]) * dzi : (v[c + 1] - vc) * dzi;
                float kave = (kap[c] + kap[c - sj]) / 2.0f;
                float fk = ((msig * (F[c] - F[c - sj])) * kave) / s->dy;
                float acc = (nu[c] * ((v[c - si] - 2.0f * vc) + v[c + si])) * dxi2;
                acc = acc + (nu[c] * ((v[c - sj] - 2.0f * vc) + v[c + sj])) * dyi2;
                acc = acc + (nu[c] * ((v[c - 1] - 2.0f * vc) + v[c + 1])) * dzi2;
                acc = acc - u_here * dvdx; acc = acc - vc * dvdy; acc = acc - w_here * dvdz;
                acc = acc + s->gy;
                acc = acc + (fk * 2.0f) / (rho[c] + rho[c - sj]);
                s->vs[c] = vc + dt * acc;
            }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            for (int k = 2; k <= nz; ++k) {
                size_t c = ID(i, j, k);
                float wc = w[c];
                float u_here = 0.25f * (((u[c + si - 1] + u[c - 1]) + u[c + si]) + u[c]);
                float v_here = 0.25f * (((v[c + sj - 1] + v[c - 1]) + v[c]) + v[c + sj]);
                float dwdx = u_here > 0 ? (wc - w[c - si]) * dxi : (w[c + si] - wc) * dxi;
                float dwdy = v_here > 0 ? (wc - w[c - sj]) * dyi : (w[c + sj] - wc) * dyi;
                float dwdz = wc > 0 ? (wc - w[c - 1]) * dzi : (w[c + 1] - wc) * dzi;
                float kave = (kap[c] + kap[c - 1]) / 2.0f;
                float fk = ((msig * (F[c] - F[c - 1])) * kave) / s->dz;
                float acc = (nu[c] * ((w[c - si] - 2.0f * wc) + w[c + si])) * dxi2;
                acc = acc + (nu[c] * ((w[c - sj] - 2.0f * wc) + w[c + sj])) * dyi2;
                acc = acc + (nu[c] * ((w[c - 1] - 2.0f * wc) + w[c + 1])) * dzi2;
                acc = acc - u_here * dwdx; acc = acc - v_here * dwdy; acc = acc - wc * dwdz;
                acc = acc + s->gz;
                acc = acc + (fk * 2.0f) / (rho[c] + rho[c - 1]);
                s->ws[c] = wc + dt * acc;
            }
}

/* 3dvof.py:261-283 -- one sweep */
void ovof3d_solve_p_jacobi(OVof3 *s) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const size_t si = s->si, sj = s->sj;
    const float *rho = s->rho, *us = s->us, *vs = s->vs, *ws = s->ws;
    float *p = s->p, *pt = s->pt;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            for (int k = 1; k <= nz; ++k) {
                size_t c = ID(i, j, k);
                float rhs = (rho[c] / s->dt) * (((us[c + si] - us[c]) * s->dxi + (vs[c + sj] - vs[c]) * s->dyi) + (ws[c + 1] - ws[c]) * s->dzi);
                float ae = i != nx ? s->dxi2 : 0.0f, aw = i != 1 ? s->dxi2 : 0.0f;
                float an = j != ny ? s->dyi2 : 0.0f, as = j != 1 ? s->dyi2 : 0.0f;
                float af = k != nz ? s->dzi2 : 0.0f, ab = k != 1 ? s->dzi2 : 0.0f;
                float ap = -1.0f * (((((ae + aw) + an) + as) + ab) + af);
                float t = rhs - ae * p[c + si];
                t = t - aw * p[c - si]; t = t - an * p[c + sj]; t = t - as * p[c - sj];
                t = t - af * p[c + 1]; t = t - ab * p[c - 1];
                pt[c] = t / ap;
            }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            for (int k = 1; k <= nz; ++k) p[ID(i, j, k)] = pt[ID(i, j, k)];
}

/* 3dvof.py:286-302 */
void ovof3d_update_uv(OVof3 *s) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const size_t si = s->si, sj = s->sj;
    const float *rho = s->rho, *p = s->p;
    long flags = 0;
#pragma omp parallel for schedule(static) reduction(+ : flags)
    for (int i = 2; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            for (int k = 1; k <= nz; ++k) {
                size_t c = ID(i, j, k);
                float r = (rho[c] + rho[c - si]) * 0.5f;
                float un = s->us[c] - ((s->dt / r) * (p[c] - p[c - si])) * s->dxi;
                s->u[c] = un;
                if (un * s->dt > s->cflx) flags++;
            }
#pragma omp parallel for schedule(static) reduction(+ : flags)
    for (int i = 1; i <= nx; ++i)
        for (int j = 2; j <= ny; ++j)
            for (int k = 1; k <= nz; ++k) {
                size_t c = ID(i, j, k);
                float r = (rho[c] + rho[c - sj]) * 0.5f;
                float vn = s->vs[c] - ((s->dt / r) * (p[c] - p[c - sj])) * s->dyi;
                s->v[c] = vn;
                if (vn * s->dt > s->cfly) flags++;
            }
#pragma omp parallel for schedule(static) reduction(+ : flags)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            for (int k = 2; k <= nz; ++k) {
                size_t c = ID(i, j, k);
                float r = (rho[c] + rho[c - 1]) * 0.5f;
                float wn = s->ws[c] - ((s->dt / r) * (p[c] - p[c - 1])) * s->dzi;
                s->w[c] = wn;
                if (wn * s->dt > s->cflx) flags++;   /* 0.25*dx, 3dvof.py:301 */
            }
    s->courant_flags = flags;
}

/* 3dvof.py:366-427 (axis 0), 430-492 (axis 1), 495-541 (axis 2).  a / cf hold the antidiffusive flux and
 * the limiter of the faces along the sweep axis (ax|ay|az, cx|cy|cz of the reference, one pair per axis);
 * face 1 of cf is never written and stays 0. */
static void fct_sweep(OVof3 *s, int axis) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const size_t st = axis == 0 ? s->si : (axis == 1 ? s->sj : 1);
    const float *vel = axis == 0 ? s->u : (axis == 1 ? s->v : s->w);
    float *F = s->F, *Ftd = s->Ftd, *a = s->a3[axis], *cf = s->cf3[axis], *rp = s->rp, *rm = s->rm;
    const float dt = s->dt, dx = s->dx, dy = s->dy, dz = s->dz, vol = s->vol;
    const float dtd = axis == 0 ? s->dt_yz : (axis == 1 ? s->dt_xz : s->dt_xy);
    const float qs = axis == 2 ? dz : dx;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            for (int k = 1; k <= nz; ++k) {
                size_t c = ID(i, j, k), m = c - st, pl = c + st;
                float vc = vel[c], vp = vel[pl];
                float dv = vol - dtd * (vp - vc);
                float lo = vc >= 0 ? (vc * dt) * F[m] : (vc * dt) * F[c];
                float hi = vp >= 0 ? (vp * dt) * F[c] : (vp * dt) * F[pl];
                float t;
                if (axis == 0) t = ((((F[c] + (((lo - hi) * dy) * dz) / vol) * dx) * dy) * dz) / dv;
                else if (axis == 1) t = ((((F[c] + ((((0.0f - 0.0f) + lo) - hi) * dy) / s->dxdy) * dx) * dy) * dz) / dv;
                else t = ((((F[c] + (((lo - hi) * dy) * dx) / vol) * dx) * dy) * dz) / dv;
                if (t > 1.0f || t < 0) t = var3(0.0f, 1.0f, t);
                Ftd[c] = t;
            }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            for (int k = 1; k <= nz; ++k) {
                size_t c = ID(i, j, k), m = c - st, pl = c + st;
                float fmax = MAXF(MAXF(Ftd[c], Ftd[m]), Ftd[pl]);
                float fmin = MINF(MINF(Ftd[c], Ftd[m]), Ftd[pl]);
                float vc = vel[c], vp = vel[pl];
                float lo_L = vc >= 0 ? (vc * dt) * F[m] : (vc * dt) * F[c];
                float hi_L = vp >= 0 ? (vp * dt) * F[c] : (vp * dt) * F[pl];
                float lo_H = vc <= 0 ? (vc * dt) * F[m] : (vc * dt) * F[c];
                float hi_H = vp <= 0 ? (vp * dt) * F[c] : (vp * dt) * F[pl];
                float ac = lo_H - lo_L, apl = hi_H - hi_L;
                a[pl] = apl; a[c] = ac;       /* same-value double write on shared faces, as in the reference */
                float pp = MAXF(0.0f, ac) - MINF(0.0f, apl);
                float pm = MAXF(0.0f, apl) - MINF(0.0f, ac);
                float qp = (fmax - Ftd[c]) * qs, qm = (Ftd[c] - fmin) * qs;
                rp[c] = pp > 0 ? MINF(1.0f, qp / pp) : 0.0f;
                rm[c] = pm > 0 ? MINF(1.0f, qm / pm) : 0.0f;
            }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            for (int k = 1; k <= nz; ++k) {
                size_t c = ID(i, j, k), pl = c + st;
                cf[pl] = a[pl] >= 0 ? MINF(rp[pl], rm[c]) : MINF(rp[c], rm[pl]);
            }
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j)
            for (int k = 1; k <= nz; ++k) {
                size_t c = ID(i, j, k), pl = c + st;
                float dv = vol - dtd * (vel[pl] - vel[c]);
                float t = a[pl] * cf[pl] - a[c] * cf[c];
                float fn = Ftd[c] - ((((t / (axis == 2 ? dz : dy)) * dx) * dy) * dz) / dv;
                F[c] = var3(0.0f, 1.0f, fn);
            }
}

void ovof3d_fct_x_sweep(OVof3 *s) { fct_sweep(s, 0); }
void ovof3d_fct_y_sweep(OVof3 *s) { fct_sweep(s, 1); }
void ovof3d_fct_z_sweep(OVof3 *s) { fct_sweep(s, 2); }

/* 3dvof.py:351-363 */
void ovof3d_solve_VOF_rudman(OVof3 *s) {
    static const int order[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}};
    const int r = s->istep % 3;
    for (int q = 0; q < 3; ++q) fct_sweep(s, order[r][q]);
}

/* 3dvof.py:544-547 */
void ovof3d_post_process_f(OVof3 *s) {
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < s->n; ++k) s->F[k] = var3(s->F[k], 0.0f, 1.0f);
}

/* 3dvof.py:598-623 */
void ovof3d_step(OVof3 *s) {
    s->istep += 1;
    ovof3d_cal_nu_rho(s);
    ovof3d_advect_upwind(s);
    ovof3d_set_BC(s);
    for (int k = 0; k < s->P.n_jacobi; ++k) ovof3d_solve_p_jacobi(s);
    ovof3d_update_uv(s);
    ovof3d_set_BC(s);
    ovof3d_solve_VOF_rudman(s);
    ovof3d_post_process_f(s);
    ovof3d_set_BC(s);
}

void ovof3d_run(OVof3 *s, int nsteps) {
    for (int k = 0; k < nsteps; ++k) ovof3d_step(s);
}

double ovof3d_mass(const OVof3 *s) {
    double m = 0.0;
    for (int i = 1; i <= s->nx; ++i)
        for (int j = 1; j <= s->ny; ++j)
            for (int k = 1; k <= s->nz; ++k) m += (double)s->F[ID(i, j, k)];
    return m;
}
