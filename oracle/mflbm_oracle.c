/*
 * mflbm_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; NOT PRODUCT CODE)
 *
 * Plain C + OpenMP restatement of the MF-LBM time-step hot path and of the
 * init-time routines that produce its inputs.  See mflbm_oracle.h for the
 * usage rules ("parity unpinned": the reference has no golden vectors).
 * Build (parity): gcc -O2 -fopenmp -ffp-contract=off   (source order == evaluation order)
 * Build (timing): gcc -O3 -march=native -fopenmp
 *
 * Reference citations: MP/ = multiphase_3D/0.src/, SP/ = singlephase_3D/0.src/
 */
#include "mflbm_oracle.h"
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- lattice tables, MP/Module.F90:111-131 ---- */
static const int EX[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
static const int EY[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int EZ[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
static const int OPC[19] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
#define W0 (1.0 / 3.0)
#define W1 (1.0 / 18.0)
#define W2 (1.0 / 36.0)
static const double W_EQU[19] = {W0, W1, W1, W1, W1, W1, W1, W2, W2, W2, W2, W2, W2, W2, W2, W2, W2, W2, W2};
static const double PI_ = 3.14159265358979323846; /* MP/Module.F90:7 */
static const double EPS_MP = 1.110223025e-16;     /* MP/Module.F90:8 */
static const double EPS_SP __attribute__((unused)) = 1e-14; /* SP/Module.F90:6 */
#define MRT_COEF1 (1.0 / 19.0)
#define MRT_COEF2 (1.0 / 2394.0)
#define MRT_COEF3 (1.0 / 252.0)
#define MRT_COEF4 (1.0 / 72.0)
#define MRT_E2_COEF1 0.0
#define MRT_E2_COEF2 (-475.0 / 63.0)
#define MRT_OMEGA_XX 0.0
#define ISO4_1 (1.0 / 6.0)
#define ISO4_2 (1.0 / 12.0)

struct orc_state {
    orc_params p;
    int nx, ny, nz;
    /* derived parameters, MP/Init_multiphase.F90:116-150 */
    double la_nui1, la_nui2, theta, phi_inlet, force_Z, rho_out, rho_in, relaxation;
    double uin_avg, uin_avg_0, flowrate, la_x, la_y, la_z, A_xy, A_xy_effective, rk_weight2;
    double s_e, s_e2, s_q, s_nu, s_pi, s_t; /* singlephase constant rates */
    long long pore_sum, pore_sum_effective;
    int ntime_max;
    /* geometry */
    int8_t *walls_global; /* (1:nxG,1:nyG,1:nzG) */
    int8_t *walls;        /* (-1:n+2)^3 */
    int *pore_profile_z;  /* local slab profile (1:nz) */
    orc_solid_node *solid;
    int num_solid, num_solid_global;
    orc_fluid_node *fluid;
    int num_fluid, num_fluid_global;
    /* fields */
    double *f[19], *g[19];
    double *phi, *phi_old, *cn_x, *cn_y, *cn_z, *c_norm, *curv, *u, *v, *w, *rho;
    double *w_in, *f_convec_bc, *g_convec_bc, *phi_convec_bc;
    double *fl1, *fl2, *vol1, *vol2, *mass1, *mass2, *pre;
};

/* index helpers; Fortran (i,j,k) -> linear */
#define SX1 ((size_t)(s->nx + 2))
#define SY1 ((size_t)(s->ny + 2))
#define SX2 ((size_t)(s->nx + 4))
#define SY2 ((size_t)(s->ny + 4))
#define SX4 ((size_t)(s->nx + 8))
#define SY4 ((size_t)(s->ny + 8))
#define I1(i, j, k) ((size_t)(i) + SX1 * ((size_t)(j) + SY1 * (size_t)(k)))
#define I2(i, j, k) ((size_t)((i) + 1) + SX2 * ((size_t)((j) + 1) + SY2 * (size_t)((k) + 1)))
#define I4(i, j, k) ((size_t)((i) + 3) + SX4 * ((size_t)((j) + 3) + SY4 * (size_t)((k) + 3)))
#define IG(i, j, k) ((size_t)((i)-1) + (size_t)s->p.nxG * ((size_t)((j)-1) + (size_t)s->p.nyG * (size_t)((k)-1)))
#define IP2(i, j) ((size_t)(i) + SX1 * (size_t)(j))
#define IPC(i, j, q) ((size_t)(i) + SX1 * ((size_t)(j) + SY1 * (size_t)(q)))

static size_t n1(const orc_state *s) { return (size_t)(s->nx + 2) * (s->ny + 2) * (s->nz + 2); }
static size_t n2(const orc_state *s) { return (size_t)(s->nx + 4) * (s->ny + 4) * (s->nz + 4); }
static size_t n4(const orc_state *s) { return (size_t)(s->nx + 8) * (s->ny + 8) * (s->nz + 8); }

static double *dalloc(size_t n) {
    double *p = (double *)calloc(n, sizeof(double));
    if (!p) {
        fprintf(stderr, "oracle: out of memory (%zu doubles)\n", n);
        abort();
    }
    return p;
}

orc_state *orc_create(const orc_params *p) {
    orc_state *s = (orc_state *)calloc(1, sizeof(orc_state));
    s->p = *p;
    s->nx = p->nxG;
    s->ny = p->nyG;
    if (p->npz < 1 || p->nzG % p->npz != 0) {
        fprintf(stderr, "oracle: nzG must be divisible by npz\n");
        free(s);
        return NULL;
    }
    s->nz = p->nzG / p->npz;
    s->walls_global = (int8_t *)calloc((size_t)p->nxG * p->nyG * p->nzG, 1);
    s->walls = (int8_t *)calloc(n2(s), 1);
    s->pore_profile_z = (int *)calloc(s->nz, sizeof(int));
    int nq = 19;
    for (int q = 0; q < nq; q++) {
        s->f[q] = dalloc(n1(s));
        if (p->multiphase) s->g[q] = dalloc(n1(s));
    }
    s->u = dalloc(n1(s));
    s->v = dalloc(n1(s));
    s->w = dalloc(n1(s));
    s->rho = dalloc(n1(s));
    s->w_in = dalloc(SX1 * SY1);
    s->f_convec_bc = dalloc(SX1 * SY1 * 19);
    if (p->multiphase) {
        s->curv = dalloc(n1(s));
        s->cn_x = dalloc(n2(s));
        s->cn_y = dalloc(n2(s));
        s->cn_z = dalloc(n2(s));
        s->c_norm = dalloc(n2(s));
        s->phi = dalloc(n4(s));
        s->phi_old = dalloc(n4(s));
        s->g_convec_bc = dalloc(SX1 * SY1 * 19);
        s->phi_convec_bc = dalloc(SX1 * SY1);
    }
    s->fl1 = dalloc(s->nz);
    s->fl2 = dalloc(s->nz);
    s->vol1 = dalloc(s->nz);
    s->vol2 = dalloc(s->nz);
    s->mass1 = dalloc(s->nz);
    s->mass2 = dalloc(s->nz);
    s->pre = dalloc(s->nz);
    s->relaxation = 1.0;                      /* MP/Main_multiphase.F90:86 */
    s->rk_weight2 = 1.0 / sqrt(2.0) / 36.0;   /* MP/Module.F90:225 */
    return s;
}

void orc_destroy(orc_state *s) {
    if (!s) return;
    for (int q = 0; q < 19; q++) {
        free(s->f[q]);
        free(s->g[q]);
    }
    free(s->walls_global); free(s->walls); free(s->pore_profile_z); free(s->solid); free(s->fluid);
    free(s->phi); free(s->phi_old); free(s->cn_x); free(s->cn_y); free(s->cn_z); free(s->c_norm); free(s->curv);
    free(s->u); free(s->v); free(s->w); free(s->rho); free(s->w_in);
    free(s->f_convec_bc); free(s->g_convec_bc); free(s->phi_convec_bc);
    free(s->fl1); free(s->fl2); free(s->vol1); free(s->vol2); free(s->mass1); free(s->mass2); free(s->pre);
    free(s);
}

/* =====================================================================================
 * geometry
 * ===================================================================================== */
int8_t *orc_walls_global(orc_state *s) { return s->walls_global; }

/* MP/Misc.F90:213-244 modify_geometry: tube + centred sphere, buffer=10 */
static void modify_geometry(orc_state *s) {
    const int nxG = s->p.nxG, nyG = s->p.nyG, nzG = s->p.nzG;
    double xc = 0.5 * (double)(nxG + 1);
    double yc = 0.5 * (double)(nyG + 1);
    double zc = 0.5 * (double)(nzG + 1);
    double r1 = 0.25 * nyG;
    double r2 = nyG * 0.5;
    int buffer = 10;
    for (int k = 1; k <= nzG; k++)
        for (int j = 1; j <= nyG; j++)
            for (int i = 1; i <= nxG; i++) {
                double dx = i - xc, dy = j - yc, dz = k - zc;
                if (dx * dx + dy * dy + dz * dz < r1 * r1) s->walls_global[IG(i, j, k)] = 1;
                if (dx * dx + dy * dy > r2 * r2 && k > buffer && k < nzG - buffer + 1) s->walls_global[IG(i, j, k)] = 1;
            }
}

/* MP/Misc.F90:6-210 set_walls (from "modify geometry" on; walls_global already holds the
 * file contents or zeros), pore_profile MP/Misc.F90:298-365, {z,y,x}transport_walls
 * MP/Mpi_misc.F90:337-502 restated for an undivided x,y and a z slab idz of npz. */
void orc_set_walls(orc_state *s) {
    const orc_params *p = &s->p;
    const int nxG = p->nxG, nyG = p->nyG, nzG = p->nzG;
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    memset(s->walls, 0, n2(s));
    if (p->modify_geometry_cmd == 1) modify_geometry(s);
    /* channel walls on the global array, MP/Misc.F90:91-104 */
    for (int k = 1; k <= nzG; k++)
        for (int j = 1; j <= nyG; j++)
            for (int i = 1; i <= nxG; i++) {
                if (p->wsz0 == 1) s->walls_global[IG(i, j, 1)] = 1;
                if (p->wsz1 == 1) s->walls_global[IG(i, j, nzG)] = 1;
                if (p->wsx0 == 1) s->walls_global[IG(1, j, k)] = 1;
                if (p->wsx1 == 1) s->walls_global[IG(nxG, j, k)] = 1;
                if (p->wsy0 == 1) s->walls_global[IG(i, 1, k)] = 1;
                if (p->wsy1 == 1) s->walls_global[IG(i, nyG, k)] = 1;
            }
    /* local copy, MP/Misc.F90:130-140 */
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) s->walls[I2(i, j, k)] = s->walls_global[IG(i, j, p->idz * nz + k)];
    /* channel walls on local array incl. ghost layers, MP/Misc.F90:147-191 */
    for (int k = -1; k <= nz + 2; k++)
        for (int j = -1; j <= ny + 2; j++)
            for (int i = -1; i <= nx + 2; i++) {
                if (p->idz == 0 && k <= 1 && p->wsz0 == 1) s->walls[I2(i, j, k)] = 1;
                if (p->idz == p->npz - 1 && k >= nz && p->wsz1 == 1) s->walls[I2(i, j, k)] = 1;
                if (i <= 1 && p->wsx0 == 1) s->walls[I2(i, j, k)] = 1;
                if (i >= nx && p->wsx1 == 1) s->walls[I2(i, j, k)] = 1;
                if (j <= 1 && p->wsy0 == 1) s->walls[I2(i, j, k)] = 1;
                if (j >= ny && p->wsy1 == 1) s->walls[I2(i, j, k)] = 1;
            }
    /* open area at inlet, MP/Misc.F90:193-201 */
    int icount = 0;
    for (int j = 1; j <= nyG; j++)
        for (int i = 1; i <= nxG; i++)
            if (s->walls_global[IG(i, j, 1)] <= 0) icount++;
    s->A_xy_effective = icount;
    /* pore_profile, MP/Misc.F90:298-365 : totals over the GLOBAL domain (sum over ranks) */
    for (int k = 1; k <= nz; k++) {
        int c = 0;
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++)
                if (s->walls[I2(i, j, k)] <= 0) c++;
        s->pore_profile_z[k - 1] = c;
    }
    long long ps = 0, pse = 0;
    for (int k = 1; k <= nzG; k++) {
        long long c = 0;
        for (int j = 1; j <= nyG; j++)
            for (int i = 1; i <= nxG; i++)
                if (s->walls_global[IG(i, j, k)] <= 0) c++;
        ps += c;
        if (k >= 1 + p->n_exclude_inlet && k <= nzG - p->n_exclude_outlet) pse += c;
    }
    s->pore_sum = ps;
    s->pore_sum_effective = pse;
    /* ztransport_walls(0,0,2): neighbours' interior planes -> ghost planes (i,j interior only) */
    const int lz = 2;
    for (int k = 1; k <= lz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                if (p->kper == 1 || p->idz != 0) { /* recv1 = zM neighbour's walls(nz+k-lz) */
                    int kg = p->idz * nz + (k - lz); /* global plane index in 1-based coords, may wrap */
                    int kk = ((kg - 1) % nzG + nzG) % nzG + 1;
                    s->walls[I2(i, j, k - lz)] = s->walls_global[IG(i, j, kk)];
                }
                if (p->kper == 1 || p->idz != p->npz - 1) { /* recv2 = zP neighbour's walls(k) */
                    int kg = p->idz * nz + nz + k;
                    int kk = ((kg - 1) % nzG + nzG) % nzG + 1;
                    s->walls[I2(i, j, k + nz)] = s->walls_global[IG(i, j, kk)];
                }
            }
    /* ytransport_walls(0,2,2): y is undivided -> self exchange when periodic */
    if (p->jper == 1) {
        const int ly = 2;
        for (int k = 1 - lz; k <= nz + lz; k++)
            for (int j = 1; j <= ly; j++)
                for (int i = 1; i <= nx; i++) {
                    int8_t lo = s->walls[I2(i, j, k)], hi = s->walls[I2(i, ny + j - ly, k)];
                    s->walls[I2(i, j - ly, k)] = hi;
                    s->walls[I2(i, j + ny, k)] = lo;
                }
    }
    /* xtransport_walls(2,2,2): x periodic is rejected by the reference (MP/IO_multiphase.F90:469-472) */
}

/* ISO8 offset tables for the wall normal, MP/Geometry_preprocessing.F90:234-377.
 * Each entry (a,b,c) stands for  ws(i+a,j+b,k+c) - ws(i-a,j-b,k-c), accumulated left to right. */
typedef struct { signed char a, b, c; } off3;
static const int ISO8_CNT[7] = {1, 4, 4, 1, 8, 12, 4};
static const off3 ISO8_X[34] = {
    {1,0,0},
    {1,1,0},{1,-1,0},{1,0,1},{1,0,-1},
    {1,1,1},{1,1,-1},{1,-1,1},{1,-1,-1},
    {2,0,0},
    {2,1,0},{2,-1,0},{2,0,1},{2,0,-1},{1,2,0},{1,-2,0},{1,0,2},{1,0,-2},
    {2,1,1},{2,1,-1},{2,-1,1},{2,-1,-1},{1,2,1},{1,2,-1},{1,-2,1},{1,-2,-1},{1,1,2},{1,1,-2},{1,-1,2},{1,-1,-2},
    {2,2,0},{2,-2,0},{2,0,2},{2,0,-2}};
static const off3 ISO8_Y[34] = {
    {0,1,0},
    {1,1,0},{-1,1,0},{0,1,1},{0,1,-1},
    {1,1,1},{1,1,-1},{-1,1,-1},{-1,1,1},
    {0,2,0},
    {2,1,0},{-2,1,0},{0,2,1},{0,2,-1},{1,2,0},{-1,2,0},{0,1,2},{0,1,-2},
    {2,1,1},{2,1,-1},{-2,1,1},{-2,1,-1},{1,2,1},{1,2,-1},{-1,2,1},{-1,2,-1},{1,1,2},{1,1,-2},{-1,1,2},{-1,1,-2},
    {2,2,0},{-2,2,0},{0,2,2},{0,2,-2}};
static const off3 ISO8_Z[34] = {
    {0,0,1},
    {0,1,1},{0,-1,1},{1,0,1},{-1,0,1},
    {1,1,1},{1,-1,1},{-1,1,1},{-1,-1,1},
    {0,0,2},
    {0,1,2},{0,-1,2},{2,0,1},{-2,0,1},{0,2,1},{0,-2,1},{1,0,2},{-1,0,2},
    {2,1,1},{2,-1,1},{-2,1,1},{-2,-1,1},{1,2,1},{1,-2,1},{-1,2,1},{-1,-2,1},{1,1,2},{1,-1,2},{-1,1,2},{-1,-1,2},
    {0,2,2},{0,-2,2},{2,0,2},{-2,0,2}};
static const double ISO8_W[7] = {4.0 / 45.0, 1.0 / 21.0, 2.0 / 105.0, 5.0 / 504.0, 1.0 / 315.0, 1.0 / 630.0, 1.0 / 5040.0};

static double iso8_component(const double *ws, size_t c, ptrdiff_t sx, ptrdiff_t sy, const off3 *tab) {
    double res = 0.0;
    int t = 0;
    for (int grp = 0; grp < 7; grp++) {
        double acc = 0.0;
        for (int m = 0; m < ISO8_CNT[grp]; m++, t++) {
            ptrdiff_t o = tab[t].a + sx * tab[t].b + sy * tab[t].c;
            if (m == 0) acc = ws[c + o] - ws[c - o];
            else { acc = acc + ws[c + o]; acc = acc - ws[c - o]; }
        }
        if (grp == 0) res = ISO8_W[0] * acc;
        else res = res + ISO8_W[grp] * acc;
    }
    return res;
}

/* MP/Geometry_preprocessing.F90:9-512 geometry_preprocessing_new */
void orc_geometry_preprocess(orc_state *s) {
    const orc_params *p = &s->p;
    const int nxG = p->nxG, nyG = p->nyG, nzG = p->nzG;
    const int gl = 6 + 4; /* ghost_layers, :41-42 */
    const int ophi = 4;
    const ptrdiff_t ex_ = nxG + 2 * gl, ey_ = nyG + 2 * gl, ez_ = nzG + 2 * gl;
    const size_t ntot = (size_t)ex_ * ey_ * ez_;
#define IE(i, j, k) ((size_t)((i) + gl - 1) + (size_t)ex_ * ((size_t)((j) + gl - 1) + (size_t)ey_ * (size_t)((k) + gl - 1)))
    int8_t *wt = (int8_t *)calloc(ntot, 1);
    double *ws1 = (double *)malloc(ntot * sizeof(double));
    double *ws2 = (double *)malloc(ntot * sizeof(double));
    for (int k = 1; k <= nzG; k++)
        for (int j = 1; j <= nyG; j++)
            for (int i = 1; i <= nxG; i++) wt[IE(i, j, k)] = s->walls_global[IG(i, j, k)];
    /* z extension :56-72 */
    for (int j = 1; j <= nyG; j++)
        for (int i = 1; i <= nxG; i++)
            for (int g = 1; g <= gl; g++) {
                if (p->kper == 0) {
                    wt[IE(i, j, 1 - g)] = wt[IE(i, j, 1)];
                    wt[IE(i, j, nzG + g)] = wt[IE(i, j, nzG)];
                } else {
                    wt[IE(i, j, 1 - g)] = wt[IE(i, j, nzG + 1 - g)];
                    wt[IE(i, j, nzG + g)] = wt[IE(i, j, g)];
                }
            }
    /* y extension :74-90 */
    for (int k = 1 - gl; k <= nzG + gl; k++)
        for (int i = 1; i <= nxG; i++)
            for (int g = 1; g <= gl; g++) {
                if (p->jper == 0) {
                    wt[IE(i, 1 - g, k)] = wt[IE(i, 1, k)];
                    wt[IE(i, nyG + g, k)] = wt[IE(i, nyG, k)];
                } else {
                    wt[IE(i, 1 - g, k)] = wt[IE(i, nyG + 1 - g, k)];
                    wt[IE(i, nyG + g, k)] = wt[IE(i, g, k)];
                }
            }
    /* x extension :92-108 */
    for (int k = 1 - gl; k <= nzG + gl; k++)
        for (int j = 1 - gl; j <= nyG + gl; j++)
            for (int g = 1; g <= gl; g++) {
                if (p->iper == 0) {
                    wt[IE(1 - g, j, k)] = wt[IE(1, j, k)];
                    wt[IE(nxG + g, j, k)] = wt[IE(nxG, j, k)];
                } else {
                    wt[IE(1 - g, j, k)] = wt[IE(nxG + 1 - g, j, k)];
                    wt[IE(nxG + g, j, k)] = wt[IE(g, j, k)];
                }
            }
    for (size_t n = 0; n < ntot; n++) ws1[n] = ws2[n] = (double)wt[n]; /* :110-118 */
    /* classify :121-143 (in place; order independent) */
    for (int k = 2 - gl; k <= nzG + gl - 1; k++)
        for (int j = 2 - gl; j <= nyG + gl - 1; j++)
            for (int i = 2 - gl; i <= nxG + gl - 1; i++) {
                size_t c = IE(i, j, k);
                if (wt[c] == 1) {
                    for (int n = 1; n <= 18; n++)
                        if (wt[IE(i + EX[n], j + EY[n], k + EZ[n])] <= 0) { wt[c] = 2; break; }
                }
                if (wt[c] == 0) {
                    for (int n = 1; n <= 18; n++)
                        if (wt[IE(i + EX[n], j + EY[n], k + EZ[n])] >= 1) { wt[c] = -1; break; }
                }
            }
    /* 4x 27-point smoothing :145-169 */
    static const int iex[27] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1};
    static const int iey[27] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, -1, 1, -1, 1};
    static const int iez[27] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1, -1, 1, 1, -1, -1, 1, 1, -1};
    const double we[4] = {8.0 / 27.0, 2.0 / 27.0, 1.0 / 54.0, 1.0 / 216.0};
    ptrdiff_t soff[27];
    double sw[27];
    for (int n = 0; n < 27; n++) {
        soff[n] = iex[n] + ex_ * (iey[n] + ey_ * (ptrdiff_t)iez[n]);
        sw[n] = we[iex[n] * iex[n] + iey[n] * iey[n] + iez[n] * iez[n]];
    }
    for (int it = 0; it < 4; it++) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int k = 2 - gl; k <= nzG + gl - 1; k++)
            for (int j = 2 - gl; j <= nyG + gl - 1; j++)
                for (int i = 2 - gl; i <= nxG + gl - 1; i++) {
                    size_t c = IE(i, j, k);
                    double acc = 0.0;
                    for (int n = 0; n < 27; n++) acc = acc + ws1[c + soff[n]] * sw[n];
                    ws2[c] = acc;
                }
#pragma omp parallel for collapse(2) schedule(static)
        for (int k = 2 - gl; k <= nzG + gl - 1; k++)
            for (int j = 2 - gl; j <= nyG + gl - 1; j++)
                for (int i = 2 - gl; i <= nxG + gl - 1; i++) ws1[IE(i, j, k)] = ws2[IE(i, j, k)];
    }
    /* count :171-185 */
    int ns = 0, nf = 0;
    for (int k = 1 - ophi; k <= nzG + ophi; k++)
        for (int j = 1 - ophi; j <= nyG + ophi; j++)
            for (int i = 1 - ophi; i <= nxG + ophi; i++) {
                int8_t t = wt[IE(i, j, k)];
                if (t == 2) ns++;
                if (t == -1) nf++;
            }
    s->num_solid_global = ns;
    s->num_fluid_global = nf;
    orc_solid_node *sg = (orc_solid_node *)calloc(ns > 0 ? ns : 1, sizeof(orc_solid_node));
    orc_fluid_node *fg = (orc_fluid_node *)calloc(nf > 0 ? nf : 1, sizeof(orc_fluid_node));
    /* fill :195-225 */
    int c1 = 0, c2 = 0;
    for (int k = 1 - ophi; k <= nzG + ophi; k++)
        for (int j = 1 - ophi; j <= nyG + ophi; j++)
            for (int i = 1 - ophi; i <= nxG + ophi; i++) {
                int8_t t = wt[IE(i, j, k)];
                if (t == 2) {
                    orc_solid_node *sn = &sg[c1++];
                    sn->ix = i; sn->iy = j; sn->iz = k;
                    int ncount = 0;
                    sn->la_weight = 0.0;
                    for (int n = 1; n <= 18; n++)
                        if (wt[IE(i + EX[n], j + EY[n], k + EZ[n])] <= 0) {
                            sn->la_weight = sn->la_weight + W_EQU[n];
                            sn->neighbor_list[ncount++] = n;
                        }
                    sn->i_fluid_num = ncount;
                }
                if (t == -1) {
                    orc_fluid_node *fn = &fg[c2++];
                    fn->ix = i; fn->iy = j; fn->iz = k;
                }
            }
    /* normals :227-383 */
#pragma omp parallel for schedule(static)
    for (int num = 0; num < nf; num++) {
        size_t c = IE(fg[num].ix, fg[num].iy, fg[num].iz);
        double nwx = iso8_component(ws2, c, ex_, ex_ * ey_, ISO8_X);
        double nwy = iso8_component(ws2, c, ex_, ex_ * ey_, ISO8_Y);
        double nwz = iso8_component(ws2, c, ex_, ex_ * ey_, ISO8_Z);
        double tmp = 1.0 / (sqrt(nwx * nwx + nwy * nwy + nwz * nwz) + EPS_MP);
        fg[num].nwx = nwx * tmp;
        fg[num].nwy = nwy * tmp;
        fg[num].nwz = nwz * tmp;
    }
    free(wt); free(ws1); free(ws2);
#undef IE
    /* local lists :424-507 (x,y undivided: out1=out2=0, out3=idz) */
    const int nx = s->nx, ny = s->ny, nz = s->nz, o3 = p->idz;
    free(s->solid); free(s->fluid);
    int cs = 0, cf = 0;
    for (int n = 0; n < ns; n++) {
        int i = sg[n].ix, j = sg[n].iy, k = sg[n].iz - o3 * nz;
        if (i >= 1 - 3 && i <= nx + 3 && j >= 1 - 3 && j <= ny + 3 && k >= 1 - 3 && k <= nz + 3) cs++;
    }
    for (int n = 0; n < nf; n++) {
        int i = fg[n].ix, j = fg[n].iy, k = fg[n].iz - o3 * nz;
        if (i >= 1 - 2 && i <= nx + 2 && j >= 1 - 2 && j <= ny + 2 && k >= 1 - 2 && k <= nz + 2) cf++;
    }
    s->solid = (orc_solid_node *)calloc(cs > 0 ? cs : 1, sizeof(orc_solid_node));
    s->fluid = (orc_fluid_node *)calloc(cf > 0 ? cf : 1, sizeof(orc_fluid_node));
    s->num_solid = cs;
    s->num_fluid = cf;
    cs = cf = 0;
    /* theta transform MP/IO_multiphase.F90:467-468 */
    s->theta = (180.0 - p->theta_deg) * PI_ / 180.0;
    for (int n = 0; n < ns; n++) {
        int i = sg[n].ix, j = sg[n].iy, k = sg[n].iz - o3 * nz;
        if (i >= 1 - 3 && i <= nx + 3 && j >= 1 - 3 && j <= ny + 3 && k >= 1 - 3 && k <= nz + 3) {
            s->solid[cs] = sg[n];
            s->solid[cs].iz = k;
            cs++;
        }
    }
    for (int n = 0; n < nf; n++) {
        int i = fg[n].ix, j = fg[n].iy, k = fg[n].iz - o3 * nz;
        if (i >= 1 - 2 && i <= nx + 2 && j >= 1 - 2 && j <= ny + 2 && k >= 1 - 2 && k <= nz + 2) {
            s->fluid[cf] = fg[n];
            s->fluid[cf].iz = k;
            s->fluid[cf].theta = s->theta;
            cf++;
        }
    }
    free(sg); free(fg);
}

/* =====================================================================================
 * initialisation
 * ===================================================================================== */
/* gfortran evaluates integer n**5 in 32-bit arithmetic: wraps for n >= 75 (reference quirk) */
static int32_t ipow_wrap(int32_t n, int e) {
    uint32_t r = 1, b = (uint32_t)n;
    for (int i = 0; i < e; i++) r *= b;
    return (int32_t)r;
}

/* MP/Misc.F90:625-665 inlet_vel_profile_rectangular */
static void inlet_vel_profile_rectangular(orc_state *s, double vel_avg, int num_terms) {
    double a = 0.5 * s->la_x, b = 0.5 * s->la_y;
    double tmp1 = 0.0;
    for (int n = 1; n <= num_terms; n += 2) tmp1 = tmp1 + (tanh(0.5 * (double)n * PI_ * b / a)) / ipow_wrap(n, 5);
    const double pi2 = PI_ * PI_, pi5 = pi2 * pi2 * PI_, pim3 = 1.0 / (pi2 * PI_); /* gfortran powi expansion */
    double tmp2 = 1.0 - 192.0 / pi5 * (a / b) * tmp1;
    tmp2 = -3.0 * vel_avg / (tmp2 * (a * a));
    for (int j = 1; j <= s->ny; j++)
        for (int i = 1; i <= s->nx; i++) {
            int x = i, y = j;
            if (x > 1 && x < s->p.nxG && y > 1 && y < s->p.nyG) {
                double xx = x - 1.5 - a;
                double yy = y - 1.5 - b;
                double tmp3 = 0.0;
                for (int n = 1; n <= num_terms; n += 2) {
                    double sgn = pow(-1.0, 0.5 * (double)(n - 1));
                    tmp3 = tmp3 + sgn * cos(0.5 * n * PI_ * xx / a) / ipow_wrap(n, 3) *
                                      (1.0 - (exp(0.5 * n * PI_ * (yy - b) / a) + exp(0.5 * n * PI_ * (-yy - b) / a)) /
                                                 (1.0 + exp(0.5 * n * PI_ * (-b - b) / a)));
                }
                s->w_in[IP2(i, j)] = tmp3 * (-16.0 * tmp2 * (a * a) * pim3);
            }
        }
}

/* MP/Init_multiphase.F90:68-150,194-236 ; SP/Initialization.F90:76-130,157-197 */
void orc_init_basic(orc_state *s) {
    const orc_params *p = &s->p;
    s->la_z = p->nzG - 1;
    s->la_y = p->nyG - 1 - 0.5f - 0.5f;
    s->la_x = p->nxG - 1 - 0.5f - 0.5f;
    s->A_xy = s->la_x * s->la_y;
    s->la_nui1 = 1.0 / p->la_nu1;
    s->la_nui2 = p->multiphase ? 1.0 / p->la_nu2 : 0.0;
    s->theta = (180.0 - p->theta_deg) * PI_ / 180.0;
    s->phi_inlet = 2.0 * p->sa_inject - 1.0;
    s->force_Z = p->force_z0;
    s->rho_out = 1.0;
    s->rho_in = 1.0;
    s->ntime_max = 0;
    if (!p->multiphase) {
        /* SP/Initialization.F90:87-112, including the duplicated preset==1 test (preset 2 == SRT) */
        double omega = 1.0 / (3.0 * p->la_nu1 + 0.5);
        s->s_nu = omega;
        if (p->mrt_para_preset == 1) {
            s->s_e = omega; s->s_e2 = omega; s->s_pi = omega;
            s->s_q = 8.0 * (2.0 - omega) / (8.0 - omega);
            s->s_t = s->s_q;
        } else {
            s->s_e = omega; s->s_e2 = omega; s->s_pi = omega; s->s_q = omega; s->s_t = omega;
        }
    }
    int open_z = (p->kper == 0 && p->wsz0 == 0 && p->wsz1 == 0);
    if (open_z) {
        if (p->inlet_BC == 1) {
            if (p->multiphase) {
                s->force_Z = 0.0; /* MP/Init_multiphase.F90:143 */
                s->uin_avg_0 = p->ca_0 * p->gamma / p->la_nu1;
            } else {
                s->uin_avg_0 = p->Re * p->la_nu1 / p->char_length; /* SP/Initialization.F90:164 */
            }
            s->uin_avg = s->uin_avg_0;
            s->flowrate = s->uin_avg_0 * s->A_xy;
            for (int j = 1; j <= s->ny; j++)
                for (int i = 1; i <= s->nx; i++) {
                    s->w_in[IP2(i, j)] = 0.0;
                    if (i > 1 && i < p->nxG && j > 1 && j < p->nyG) s->w_in[IP2(i, j)] = s->uin_avg;
                }
            inlet_vel_profile_rectangular(s, s->uin_avg_0, 1000);
            if (p->target_inject_pore_volume > 0) {
                s->ntime_max = (int)((double)(p->target_inject_pore_volume * s->pore_sum) / s->flowrate);
                if (s->ntime_max % 2 == 1) s->ntime_max += 1;
            }
        } else if (p->inlet_BC == 2) {
            if (p->multiphase) {
                double p_gradient = -p->force_z0 / 3.0; /* MP/Init_multiphase.F90:146-147 */
                s->rho_in = s->rho_out - p_gradient * p->nzG;
            } else {
                s->rho_in = s->rho_out + p->rho_drop; /* SP/Initialization.F90:124 */
            }
        }
    }
}

/* MP/Init_multiphase.F90:276-337 (option 6 = irreproducible random field: caller must fill phi) */
void orc_init_phi(orc_state *s) {
    const orc_params *p = &s->p;
    if (!p->multiphase) return;
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const int open_z = (p->kper == 0 && p->wsz0 == 0 && p->wsz1 == 0);
    const double z0 = p->interface_z0;
    for (int k = -3; k <= nz + 4; k++)
        for (int j = -3; j <= ny + 4; j++)
            for (int i = -3; i <= nx + 4; i++) {
                double x = i, y = j, z = p->idz * nz + k;
                double v;
                switch (p->initial_fluid_distribution_option) {
                case 1: v = -1.0; if (z <= z0) v = 1.0; break;
                case 2: v = 1.0; if (z <= z0) v = -1.0; break;
                case 3: case 4: {
                    double dx = x - (p->nxG + 1) * 0.0, dz = z - (p->nzG + 1) * 0.5, dy = y - (p->nyG + 1) * 0.5;
                    int in = dx * dx + dz * dz + dy * dy <= z0 * z0;
                    v = (p->initial_fluid_distribution_option == 3) ? (in ? 1.0 : -1.0) : (in ? -1.0 : 1.0);
                } break;
                case 5: {
                    double dx = x - (p->nxG + 1) * 0.5, dz = z - (p->nzG + 1) * 0.5, dy = y - (p->nyG + 1) * 0.5;
                    v = (dx * dx + dz * dz + dy * dy <= z0 * z0) ? 1.0 : -1.0;
                } break;
                default: v = s->phi[I4(i, j, k)]; break; /* keep caller-provided field */
                }
                if (open_z && z <= 0) v = s->phi_inlet;
                s->phi[I4(i, j, k)] = v;
            }
}

/* MP/Init_multiphase.F90:254-273 + 357-470 ; SP/Initialization.F90:217-309 */
void orc_init_pdf(orc_state *s) {
    const orc_params *p = &s->p;
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    for (size_t n = 0; n < n1(s); n++) { s->u[n] = 0; s->v[n] = 0; s->w[n] = 0; s->rho[n] = 1.0; }
    if (p->multiphase) {
        memset(s->curv, 0, n1(s) * sizeof(double));
        /* reference zeroes only 0:n+1 of the (-1:n+2) arrays; the rest is the allocation's content (calloc: 0) */
        memset(s->cn_x, 0, n2(s) * sizeof(double)); memset(s->cn_y, 0, n2(s) * sizeof(double));
        memset(s->cn_z, 0, n2(s) * sizeof(double)); memset(s->c_norm, 0, n2(s) * sizeof(double));
    }
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= nz + 1; k++)
        for (int j = 0; j <= ny + 1; j++)
            for (int i = 0; i <= nx + 1; i++) {
                size_t c = I1(i, j, k);
                double uu = s->u[c], vv = s->v[c], ww = s->w[c];
                double usqrt = uu * uu + vv * vv + ww * ww;
                double rr[2];
                if (p->multiphase) {
                    double ph = s->phi[I4(i, j, k)];
                    rr[0] = s->rho[c] * (1.0 + ph) * 0.5;
                    rr[1] = s->rho[c] * (1.0 - ph) * 0.5;
                } else {
                    rr[0] = s->rho[c];
                    rr[1] = 0;
                }
                for (int fl = 0; fl < (p->multiphase ? 2 : 1); fl++) {
                    double r = rr[fl];
                    double **F = fl == 0 ? s->f : s->g;
                    F[0][c] = r * W0 + r * W0 * (-1.5 * usqrt);
                    for (int q = 1; q < 19; q++) {
                        double wq = q <= 6 ? W1 : W2;
                        double eu = EX[q] * uu + EY[q] * vv + EZ[q] * ww; /* u=0 at init: exact for any form */
                        F[q][c] = r * wq + r * wq * (3.0 * eu + 4.5 * eu * eu - 1.5 * usqrt);
                    }
                }
            }
    if (p->outlet_BC == 1 && p->idz == p->npz - 1) {
        for (int j = 0; j <= ny + 1; j++)
            for (int i = 0; i <= nx + 1; i++) {
                for (int q = 0; q < 19; q++) {
                    s->f_convec_bc[IPC(i, j, q)] = s->f[q][I1(i, j, nz)];
                    if (p->multiphase) s->g_convec_bc[IPC(i, j, q)] = s->g[q][I1(i, j, nz)];
                }
                if (p->multiphase) s->phi_convec_bc[IP2(i, j)] = s->phi[I4(i, j, nz)];
            }
    }
    if (p->multiphase && p->steady_state_option == 2) memcpy(s->phi_old, s->phi, n4(s) * sizeof(double));
}

/* =====================================================================================
 * collision (shared by odd/even), MP/Kernel_multiphase.F90:86-315 == :450-678
 * a[] = fluid-1 incoming PDFs (g1t*), b[] = fluid-2 incoming (g2t*); outputs overwrite a,b.
 * ===================================================================================== */
typedef struct {
    double s_e, s_e2, s_q, s_nu, s_pi, s_t;
} rates_t;

static inline void mrt_core(double *ft, double den, double fx, double fy, double fz, const rates_t *r) {
    const double s_e = r->s_e, s_e2 = r->s_e2, s_q = r->s_q, s_nu = r->s_nu, s_pi = r->s_pi, s_t = r->s_t;
    double ft0 = ft[0], ft1 = ft[1], ft2 = ft[2], ft3 = ft[3], ft4 = ft[4], ft5 = ft[5], ft6 = ft[6], ft7 = ft[7], ft8 = ft[8],
           ft9 = ft[9], ft10 = ft[10], ft11 = ft[11], ft12 = ft[12], ft13 = ft[13], ft14 = ft[14], ft15 = ft[15], ft16 = ft[16],
           ft17 = ft[17], ft18 = ft[18];
    double ux1 = ft1 - ft2 + ft7 - ft8 + ft9 - ft10 + ft11 - ft12 + ft13 - ft14 + 0.5 * fx;
    double uy1 = ft3 - ft4 + ft7 + ft8 - ft9 - ft10 + ft15 - ft16 + ft17 - ft18 + 0.5 * fy;
    double uz1 = ft5 - ft6 + ft11 + ft12 - ft13 - ft14 + ft15 + ft16 - ft17 - ft18 + 0.5 * fz;
    double u2 = ux1 * ux1 + uy1 * uy1 + uz1 * uz1;
    double sum1 = ft1 + ft2 + ft3 + ft4 + ft5 + ft6;
    double sum2 = ft7 + ft8 + ft9 + ft10 + ft11 + ft12 + ft13 + ft14 + ft15 + ft16 + ft17 + ft18;
    double sum3 = ft7 - ft8 + ft9 - ft10 + ft11 - ft12 + ft13 - ft14;
    double sum4 = ft7 + ft8 - ft9 - ft10 + ft15 - ft16 + ft17 - ft18;
    double sum5 = ft11 + ft12 - ft13 - ft14 + ft15 + ft16 - ft17 - ft18;
    double sum6 = 2.0 * (ft1 + ft2) - ft3 - ft4 - ft5 - ft6;
    double sum7 = ft7 + ft8 + ft9 + ft10 + ft11 + ft12 + ft13 + ft14 - 2.0 * (ft15 + ft16 + ft17 + ft18);
    double sum8 = ft3 + ft4 - ft5 - ft6;
    double sum9 = ft7 + ft8 + ft9 + ft10 - ft11 - ft12 - ft13 - ft14;
    double m_rho = den;
    double m_e = -30.0 * ft0 - 11.0 * sum1 + 8.0 * sum2;
    double m_e2 = 12.0 * ft0 - 4.0 * sum1 + sum2;
    double m_jx = ft1 - ft2 + sum3;
    double m_qx = -4.0 * (ft1 - ft2) + sum3;
    double m_jy = ft3 - ft4 + sum4;
    double m_qy = -4.0 * (ft3 - ft4) + sum4;
    double m_jz = ft5 - ft6 + sum5;
    double m_qz = -4.0 * (ft5 - ft6) + sum5;
    double m_3pxx = sum6 + sum7;
    double m_3pixx = -2.0 * sum6 + sum7;
    double m_pww = sum8 + sum9;
    double m_piww = -2.0 * sum8 + sum9;
    double m_pxy = ft7 - ft8 - ft9 + ft10;
    double m_pyz = ft15 - ft16 - ft17 + ft18;
    double m_pzx = ft11 - ft12 - ft13 + ft14;
    double m_tx = ft7 - ft8 + ft9 - ft10 - ft11 + ft12 - ft13 + ft14;
    double m_ty = -ft7 - ft8 + ft9 + ft10 + ft15 - ft16 + ft17 - ft18;
    double m_tz = ft11 + ft12 - ft13 - ft14 - ft15 - ft16 + ft17 + ft18;
    /* relaxation in moment space :198-216 */
    m_e = m_e - s_e * (m_e - (-11.0 * den + 19.0 * u2)) + (38.0 - 19.0 * s_e) * (fx * ux1 + fy * uy1 + fz * uz1);
    m_e2 = m_e2 - s_e2 * (m_e2 - (MRT_E2_COEF1 * den + MRT_E2_COEF2 * u2)) + (-11.0 + 5.5 * s_e2) * (fx * ux1 + fy * uy1 + fz * uz1);
    m_jx = m_jx + fx;
    m_qx = m_qx - s_q * (m_qx - (-0.666666666666666667 * ux1)) + (-0.666666666666666667 + 0.333333333333333333 * s_q) * fx;
    m_jy = m_jy + fy;
    m_qy = m_qy - s_q * (m_qy - (-0.666666666666666667 * uy1)) + (-0.666666666666666667 + 0.333333333333333333 * s_q) * fy;
    m_jz = m_jz + fz;
    m_qz = m_qz - s_q * (m_qz - (-0.666666666666666667 * uz1)) + (-0.666666666666666667 + 0.333333333333333333 * s_q) * fz;
    m_3pxx = m_3pxx - s_nu * (m_3pxx - (3.0 * ux1 * ux1 - u2)) + (2.0 - s_nu) * (2.0 * fx * ux1 - fy * uy1 - fz * uz1);
    m_3pixx = m_3pixx - s_pi * (m_3pixx - MRT_OMEGA_XX * (3.0 * ux1 * ux1 - u2)) + (1.0 - 0.5 * s_pi) * (-2.0 * fx * ux1 + fy * uy1 + fz * uz1);
    m_pww = m_pww - s_nu * (m_pww - (uy1 * uy1 - uz1 * uz1)) + (2.0 - s_nu) * (fy * uy1 - fz * uz1);
    m_piww = m_piww - s_pi * (m_piww - MRT_OMEGA_XX * (uy1 * uy1 - uz1 * uz1)) + (1.0 - 0.5 * s_pi) * (-fy * uy1 + fz * uz1);
    m_pxy = m_pxy - s_nu * (m_pxy - (ux1 * uy1)) + (1.0 - 0.5 * s_nu) * (fx * uy1 + fy * ux1);
    m_pyz = m_pyz - s_nu * (m_pyz - (uy1 * uz1)) + (1.0 - 0.5 * s_nu) * (fy * uz1 + fz * uy1);
    m_pzx = m_pzx - s_nu * (m_pzx - (ux1 * uz1)) + (1.0 - 0.5 * s_nu) * (fx * uz1 + fz * ux1);
    m_tx = m_tx - s_t * (m_tx);
    m_ty = m_ty - s_t * (m_ty);
    m_tz = m_tz - s_t * (m_tz);
    /* back transform :220-267 */
    m_rho = MRT_COEF1 * m_rho;
    m_e = MRT_COEF2 * m_e;
    m_e2 = MRT_COEF3 * m_e2;
    m_jx = 0.1 * m_jx;
    m_qx = 0.025 * m_qx;
    m_jy = 0.1 * m_jy;
    m_qy = 0.025 * m_qy;
    m_jz = 0.1 * m_jz;
    m_qz = 0.025 * m_qz;
    m_3pxx = 2.0 * MRT_COEF4 * m_3pxx;
    m_3pixx = MRT_COEF4 * m_3pixx;
    m_pww = 6.0 * MRT_COEF4 * m_pww;
    m_piww = 3.0 * MRT_COEF4 * m_piww;
    m_pxy = 0.25 * m_pxy;
    m_pyz = 0.25 * m_pyz;
    m_pzx = 0.25 * m_pzx;
    m_tx = 0.125 * m_tx;
    m_ty = 0.125 * m_ty;
    m_tz = 0.125 * m_tz;
    sum1 = m_rho - 11.0 * m_e - 4.0 * m_e2;
    sum2 = 2.0 * m_3pxx - 4.0 * m_3pixx;
    sum3 = m_pww - 2.0 * m_piww;
    sum4 = m_rho + 8.0 * m_e + m_e2;
    sum5 = m_jx + m_qx;
    sum6 = m_jy + m_qy;
    sum7 = m_jz + m_qz;
    sum8 = m_3pxx + m_3pixx;
    sum9 = m_pww + m_piww;
    ft[0] = m_rho - 30.0 * m_e + 12.0 * m_e2;
    ft[1] = sum1 + m_jx - 4.0 * m_qx + sum2;
    ft[2] = sum1 - m_jx + 4.0 * m_qx + sum2;
    ft[3] = sum1 + m_jy - 4.0 * m_qy - 0.5 * sum2 + sum3;
    ft[4] = sum1 - m_jy + 4.0 * m_qy - 0.5 * sum2 + sum3;
    ft[5] = sum1 + m_jz - 4.0 * m_qz - 0.5 * sum2 - sum3;
    ft[6] = sum1 - m_jz + 4.0 * m_qz - 0.5 * sum2 - sum3;
    ft[7] = sum4 + sum5 + sum6 + sum8 + sum9 + m_pxy + m_tx - m_ty;
    ft[8] = sum4 - sum5 + sum6 + sum8 + sum9 - m_pxy - m_tx - m_ty;
    ft[9] = sum4 + sum5 - sum6 + sum8 + sum9 - m_pxy + m_tx + m_ty;
    ft[10] = sum4 - sum5 - sum6 + sum8 + sum9 + m_pxy - m_tx + m_ty;
    ft[11] = sum4 + sum5 + sum7 + sum8 - sum9 + m_pzx - m_tx + m_tz;
    ft[12] = sum4 - sum5 + sum7 + sum8 - sum9 - m_pzx + m_tx + m_tz;
    ft[13] = sum4 + sum5 - sum7 + sum8 - sum9 - m_pzx - m_tx - m_tz;
    ft[14] = sum4 - sum5 - sum7 + sum8 - sum9 + m_pzx + m_tx - m_tz;
    ft[15] = sum4 + sum6 + sum7 - sum8 * 2.0 + m_pyz + m_ty - m_tz;
    ft[16] = sum4 - sum6 + sum7 - sum8 * 2.0 - m_pyz - m_ty - m_tz;
    ft[17] = sum4 + sum6 - sum7 - sum8 * 2.0 - m_pyz + m_ty + m_tz;
    ft[18] = sum4 - sum6 - sum7 - sum8 * 2.0 + m_pyz - m_ty + m_tz;
}

/* returns phi; a,b are overwritten with the post-collision, recoloured PDFs */
static inline double collide_mp(const orc_state *s, double *a, double *b, double cnx, double cny, double cnz, double curv,
                                double c_norm) {
    double ft[19];
    for (int q = 0; q < 19; q++) ft[q] = a[q] + b[q];
    double rho1 = a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] + a[8] + a[9] + a[10] + a[11] + a[12] + a[13] + a[14] +
                  a[15] + a[16] + a[17] + a[18];
    double rho2 = b[0] + b[1] + b[2] + b[3] + b[4] + b[5] + b[6] + b[7] + b[8] + b[9] + b[10] + b[11] + b[12] + b[13] + b[14] +
                  b[15] + b[16] + b[17] + b[18];
    double phi = (rho1 - rho2) / (rho1 + rho2);
    double tmp = 0.5 * s->p.gamma * curv * c_norm;
    double fx = tmp * cnx, fy = tmp * cny, fz = tmp * cnz + s->force_Z;
    double omega = 1.0 / (6.0 / ((1.0 + phi) * s->la_nui1 + (1.0 - phi) * s->la_nui2) + 0.5);
    rates_t r;
    r.s_nu = omega;
    switch (s->p.mrt) { /* MP/Kernel_multiphase.F90:127-156 */
    case 1: r.s_e = omega; r.s_e2 = omega; r.s_pi = omega; r.s_q = 8.0 * (2.0 - omega) / (8.0 - omega); r.s_t = r.s_q; break;
    case 3: r.s_e = omega; r.s_e2 = omega; r.s_pi = omega; r.s_q = omega; r.s_t = omega; break;
    case 4: r.s_e = omega; r.s_e2 = omega; r.s_pi = omega; r.s_q = (6.0 - 3.0 * omega) / (3.0 - omega); r.s_t = omega; break;
    default: r.s_e = 1.19; r.s_e2 = 1.4; r.s_pi = 1.4; r.s_q = 1.2; r.s_t = 1.98; break;
    }
    double den = rho1 + rho2;
    mrt_core(ft, den, fx, fy, fz, &r);
    /* recolouring :272-315 */
    double tmp1 = rho1 / den;
    a[0] = tmp1 * ft[0];
    b[0] = ft[0] * (1.0 - tmp1);
    tmp = rho1 * rho2 * s->p.beta / den;
    const double rk = s->rk_weight2;
    a[1] = tmp1 * ft[1] + W1 * tmp * (cnx);
    a[2] = tmp1 * ft[2] + W1 * tmp * (-cnx);
    a[3] = tmp1 * ft[3] + W1 * tmp * (cny);
    a[4] = tmp1 * ft[4] + W1 * tmp * (-cny);
    a[5] = tmp1 * ft[5] + W1 * tmp * (cnz);
    a[6] = tmp1 * ft[6] + W1 * tmp * (-cnz);
    a[7] = tmp1 * ft[7] + rk * tmp * (cnx + cny);
    a[8] = tmp1 * ft[8] + rk * tmp * (-cnx + cny);
    a[9] = tmp1 * ft[9] + rk * tmp * (cnx - cny);
    a[10] = tmp1 * ft[10] + rk * tmp * (-cnx - cny);
    a[11] = tmp1 * ft[11] + rk * tmp * (cnx + cnz);
    a[12] = tmp1 * ft[12] + rk * tmp * (-cnx + cnz);
    a[13] = tmp1 * ft[13] + rk * tmp * (cnx - cnz);
    a[14] = tmp1 * ft[14] + rk * tmp * (-cnx - cnz);
    a[15] = tmp1 * ft[15] + rk * tmp * (cny + cnz);
    a[16] = tmp1 * ft[16] + rk * tmp * (-cny + cnz);
    a[17] = tmp1 * ft[17] + rk * tmp * (cny - cnz);
    a[18] = tmp1 * ft[18] + rk * tmp * (-cny - cnz);
    for (int q = 1; q < 19; q++) b[q] = ft[q] - a[q];
    return phi;
}

/* SP/Kernel.F90:54-178 */
static inline void collide_sp(const orc_state *s, double *ft) {
    double den = ft[0] + ft[1] + ft[2] + ft[3] + ft[4] + ft[5] + ft[6] + ft[7] + ft[8] + ft[9] + ft[10] + ft[11] + ft[12] + ft[13] +
                 ft[14] + ft[15] + ft[16] + ft[17] + ft[18];
    rates_t r = {s->s_e, s->s_e2, s->s_q, s->s_nu, s->s_pi, s->s_t};
    mrt_core(ft, den, 0.0, 0.0, s->force_Z, &r);
}

/* MP/Kernel_multiphase.F90:6-362 kernel_odd_color ; SP/Kernel.F90:5-200 kernel_odd */
void orc_kernel_odd(orc_state *s, int ix0, int ix1, int iy0, int iy1, int iz0, int iz1) {
    const int mp = s->p.multiphase;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = iz0; k <= iz1; k++)
        for (int j = iy0; j <= iy1; j++)
            for (int i = ix0; i <= ix1; i++) {
                if (s->walls[I2(i, j, k)] != 0) continue;
                double a[19], b[19];
                for (int q = 0; q < 19; q++) {
                    size_t c = I1(i - EX[q], j - EY[q], k - EZ[q]);
                    a[q] = s->f[q][c];
                    if (mp) b[q] = s->g[q][c];
                }
                if (mp) {
                    size_t c2 = I2(i, j, k);
                    double phi = collide_mp(s, a, b, s->cn_x[c2], s->cn_y[c2], s->cn_z[c2], s->curv[I1(i, j, k)], s->c_norm[c2]);
                    s->phi[I4(i, j, k)] = phi;
                } else {
                    collide_sp(s, a);
                }
                s->f[0][I1(i, j, k)] = a[0];
                if (mp) s->g[0][I1(i, j, k)] = b[0];
                for (int q = 1; q < 19; q++) {
                    size_t c = I1(i + EX[q], j + EY[q], k + EZ[q]);
                    s->f[OPC[q]][c] = a[q];
                    if (mp) s->g[OPC[q]][c] = b[q];
                }
            }
}

/* MP/Kernel_multiphase.F90:371-725 kernel_even_color ; SP/Kernel.F90:206-400 kernel_even */
void orc_kernel_even(orc_state *s, int ix0, int ix1, int iy0, int iy1, int iz0, int iz1) {
    const int mp = s->p.multiphase;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = iz0; k <= iz1; k++)
        for (int j = iy0; j <= iy1; j++)
            for (int i = ix0; i <= ix1; i++) {
                if (s->walls[I2(i, j, k)] != 0) continue;
                const size_t c = I1(i, j, k);
                double a[19], b[19];
                for (int q = 0; q < 19; q++) {
                    a[q] = s->f[OPC[q]][c];
                    if (mp) b[q] = s->g[OPC[q]][c];
                }
                if (mp) {
                    size_t c2 = I2(i, j, k);
                    double phi = collide_mp(s, a, b, s->cn_x[c2], s->cn_y[c2], s->cn_z[c2], s->curv[c], s->c_norm[c2]);
                    s->phi[I4(i, j, k)] = phi;
                } else {
                    collide_sp(s, a);
                }
                for (int q = 0; q < 19; q++) {
                    s->f[q][c] = a[q];
                    if (mp) s->g[q][c] = b[q];
                }
            }
}

/* =====================================================================================
 * colour gradient, MP/Phase_gradient.F90:5-265
 * ===================================================================================== */
#define PH(a, b, c) s->phi[I4(i + (a), j + (b), k + (c))]
#define CX(a, b, c) s->cn_x[I2(i + (a), j + (b), k + (c))]
#define CY(a, b, c) s->cn_y[I2(i + (a), j + (b), k + (c))]
#define CZ(a, b, c) s->cn_z[I2(i + (a), j + (b), k + (c))]
/* the three ISO4 derivative shapes, written in the reference's term order */
#define DX(F) (ISO4_1 * (F(1, 0, 0) - F(-1, 0, 0)) + ISO4_2 * (F(1, 1, 0) - F(-1, -1, 0) + F(1, -1, 0) - F(-1, 1, 0) + F(1, 0, 1) - F(-1, 0, -1) + F(1, 0, -1) - F(-1, 0, 1)))
#define DY(F) (ISO4_1 * (F(0, 1, 0) - F(0, -1, 0)) + ISO4_2 * (F(1, 1, 0) - F(-1, -1, 0) + F(-1, 1, 0) - F(1, -1, 0) + F(0, 1, 1) - F(0, -1, -1) + F(0, 1, -1) - F(0, -1, 1)))
#define DZ(F) (ISO4_1 * (F(0, 0, 1) - F(0, 0, -1)) + ISO4_2 * (F(1, 0, 1) - F(-1, 0, -1) + F(-1, 0, 1) - F(1, 0, -1) + F(0, 1, 1) - F(0, -1, -1) + F(0, -1, 1) - F(0, 1, -1)))

void orc_color_gradient(orc_state *s) {
    if (!s->p.multiphase) return;
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    /* K3: phi -> solid boundary nodes :16-29 */
#pragma omp parallel for schedule(static)
    for (int num = 0; num < s->num_solid; num++) {
        const orc_solid_node *sn = &s->solid[num];
        int i = sn->ix, j = sn->iy, k = sn->iz;
        double acc = 0.0;
        for (int n = 0; n < sn->i_fluid_num; n++) {
            int ie = sn->neighbor_list[n];
            acc = acc + s->phi[I4(i + EX[ie], j + EY[ie], k + EZ[ie])] * W_EQU[ie];
        }
        s->phi[I4(i, j, k)] = acc / sn->la_weight;
    }
    /* K4: gradient + normalise :36-78 */
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = -1; k <= nz + 2; k++)
        for (int j = -1; j <= ny + 2; j++)
            for (int i = -1; i <= nx + 2; i++) {
                double gx = DX(PH), gy = DY(PH), gz = DZ(PH);
                double cn = sqrt(gx * gx + gy * gy + gz * gz);
                size_t c = I2(i, j, k);
                if (cn < 1e-6 || s->walls[c] == 1) {
                    s->cn_x[c] = 0; s->cn_y[c] = 0; s->cn_z[c] = 0; s->c_norm[c] = 0;
                } else {
                    s->cn_x[c] = gx / cn; s->cn_y[c] = gy / cn; s->cn_z[c] = gz / cn; s->c_norm[c] = cn;
                }
            }
    /* K5: alter_color_gradient_solid_surface :210-265 */
#pragma omp parallel for schedule(static)
    for (int num = 0; num < s->num_fluid; num++) {
        const orc_fluid_node *fn = &s->fluid[num];
        size_t c = I2(fn->ix, fn->iy, fn->iz);
        if (s->c_norm[c] > 1e-6) {
            double th = fn->theta, nwx = fn->nwx, nwy = fn->nwy, nwz = fn->nwz;
            double cnx0 = s->cn_x[c], cny0 = s->cn_y[c], cnz0 = s->cn_z[c];
            double tmpCos = cos(th);
            double tmp1 = nwx * cnx0 + nwy * cny0 + nwz * cnz0;
            double tmp2 = 1.0 / sqrt(1 - tmp1 * tmp1);
            double coe1 = sin(th) * tmp1 * tmp2;
            double coe2 = sin(th) * tmp2;
            double cnxp = (tmpCos - coe1) * nwx + coe2 * cnx0;
            double cnyp = (tmpCos - coe1) * nwy + coe2 * cny0;
            double cnzp = (tmpCos - coe1) * nwz + coe2 * cnz0;
            double cnxm = (tmpCos + coe1) * nwx - coe2 * cnx0;
            double cnym = (tmpCos + coe1) * nwy - coe2 * cny0;
            double cnzm = (tmpCos + coe1) * nwz - coe2 * cnz0;
            double distP = (cnxp - cnx0) * (cnxp - cnx0) + (cnyp - cny0) * (cnyp - cny0) + (cnzp - cnz0) * (cnzp - cnz0);
            double distM = (cnxm - cnx0) * (cnxm - cnx0) + (cnym - cny0) * (cnym - cny0) + (cnzm - cnz0) * (cnzm - cnz0);
            if (distP <= distM) { s->cn_x[c] = cnxp; s->cn_y[c] = cnyp; s->cn_z[c] = cnzp; }
            else { s->cn_x[c] = cnxm; s->cn_y[c] = cnym; s->cn_z[c] = cnzm; }
        }
    }
    /* K6: n -> solid boundary nodes :88-109 */
#pragma omp parallel for schedule(static)
    for (int num = 0; num < s->num_solid; num++) {
        const orc_solid_node *sn = &s->solid[num];
        int i = sn->ix, j = sn->iy, k = sn->iz;
        if (i >= 0 && i <= nx + 1 && j >= 0 && j <= ny + 1 && k >= 0 && k <= nz + 1) {
            double ax = 0, ay = 0, az = 0;
            for (int n = 0; n < sn->i_fluid_num; n++) {
                int ie = sn->neighbor_list[n];
                size_t c = I2(i + EX[ie], j + EY[ie], k + EZ[ie]);
                ax = ax + s->cn_x[c] * W_EQU[ie];
                ay = ay + s->cn_y[c] * W_EQU[ie];
                az = az + s->cn_z[c] * W_EQU[ie];
            }
            size_t c = I2(i, j, k);
            s->cn_x[c] = ax / sn->la_weight;
            s->cn_y[c] = ay / sn->la_weight;
            s->cn_z[c] = az / sn->la_weight;
        }
    }
    /* K7: curvature :116-200 */
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                double kxx = DX(CX), kyy = DY(CY), kzz = DZ(CZ);
                double kxy = DY(CX), kxz = DZ(CX);
                double kyx = DX(CY), kyz = DZ(CY);
                double kzx = DX(CZ), kzy = DY(CZ);
                double nx_ = CX(0, 0, 0), ny_ = CY(0, 0, 0), nz_ = CZ(0, 0, 0);
                s->curv[I1(i, j, k)] = (nx_ * nx_ - 1.0) * kxx + (ny_ * ny_ - 1.0) * kyy + (nz_ * nz_ - 1.0) * kzz +
                                       nx_ * ny_ * (kxy + kyx) + nx_ * nz_ * (kxz + kzx) + ny_ * nz_ * (kzy + kyz);
            }
}

/* =====================================================================================
 * boundary conditions
 * ===================================================================================== */
#define F_(q, i, j, k) F[q][I1(i, j, k)]
/* MP/Boundary_multiphase_inlet.F90:6-102 ; SP/Boundary.F90:5-73 */
static void inlet_velocity(orc_state *s, int after) {
    if (s->p.idz != 0) return;
    const int nx = s->nx, ny = s->ny, mp = s->p.multiphase;
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
            int wi = s->walls[I2(i, j, 1)];
            if (mp) {
                double *ph = s->phi;
                ph[I4(i, j, 0)] = s->phi_inlet * (1 - wi) + ph[I4(i, j, 0)] * wi;
                ph[I4(i, j, -1)] = ph[I4(i, j, 0)];
                ph[I4(i, j, -2)] = ph[I4(i, j, 0)];
                ph[I4(i, j, -3)] = ph[I4(i, j, 0)];
            }
            double tmp2 = s->w_in[IP2(i, j)] * s->relaxation;
            double tmp1 = mp ? tmp2 * s->p.sa_inject : tmp2;
            tmp2 = tmp2 - tmp1;
            for (int fl = 0; fl < (mp ? 2 : 1); fl++) {
                double **F = fl == 0 ? s->f : s->g;
                double t = fl == 0 ? tmp1 : tmp2;
                if (!after) {
                    F_(5, i, j, 0) = (F_(6, i, j, 1) + 6.0 * W1 * t) * (1 - wi) + F_(5, i, j, 0) * wi;
                    F_(11, i - 1, j, 0) = (F_(14, i, j, 1) + 6.0 * W2 * t) * (1 - wi) + F_(11, i - 1, j, 0) * wi;
                    F_(12, i + 1, j, 0) = (F_(13, i, j, 1) + 6.0 * W2 * t) * (1 - wi) + F_(12, i + 1, j, 0) * wi;
                    F_(15, i, j - 1, 0) = (F_(18, i, j, 1) + 6.0 * W2 * t) * (1 - wi) + F_(15, i, j - 1, 0) * wi;
                    F_(16, i, j + 1, 0) = (F_(17, i, j, 1) + 6.0 * W2 * t) * (1 - wi) + F_(16, i, j + 1, 0) * wi;
                } else {
                    F_(6, i, j, 1) = (F_(5, i, j, 0) + 6.0 * W1 * t) * (1 - wi) + F_(6, i, j, 1) * wi;
                    F_(13, i, j, 1) = (F_(12, i + 1, j, 0) + 6.0 * W2 * t) * (1 - wi) + F_(13, i, j, 1) * wi;
                    F_(14, i, j, 1) = (F_(11, i - 1, j, 0) + 6.0 * W2 * t) * (1 - wi) + F_(14, i, j, 1) * wi;
                    F_(17, i, j, 1) = (F_(16, i, j + 1, 0) + 6.0 * W2 * t) * (1 - wi) + F_(17, i, j, 1) * wi;
                    F_(18, i, j, 1) = (F_(15, i, j - 1, 0) + 6.0 * W2 * t) * (1 - wi) + F_(18, i, j, 1) * wi;
                }
            }
        }
}

/* MP/Boundary_multiphase_inlet.F90:115-294 ; SP/Boundary.F90:80-187 */
static void inlet_pressure(orc_state *s, int after) {
    if (s->p.idz != 0) return;
    const int nx = s->nx, ny = s->ny, mp = s->p.multiphase;
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
            int wi = s->walls[I2(i, j, 1)];
            if (mp) {
                double *ph = s->phi;
                ph[I4(i, j, 0)] = s->phi_inlet * (1 - wi) + ph[I4(i, j, 0)] * wi;
                ph[I4(i, j, -1)] = ph[I4(i, j, 0)];
                ph[I4(i, j, -2)] = ph[I4(i, j, 0)];
                ph[I4(i, j, -3)] = ph[I4(i, j, 0)];
            }
            double tmpRho2 = s->rho_in;
            double tmpRho1 = mp ? s->rho_in * s->p.sa_inject : s->rho_in;
            tmpRho2 = tmpRho2 - tmpRho1;
            for (int fl = 0; fl < (mp ? 2 : 1); fl++) {
                double **F = fl == 0 ? s->f : s->g;
                double rin = fl == 0 ? tmpRho1 : tmpRho2;
                if (!after) {
                    double t = (rin - (F_(0, i, j, 1) + F_(1, i - 1, j, 1) + F_(2, i + 1, j, 1) + F_(3, i, j - 1, 1) + F_(4, i, j + 1, 1) +
                                       F_(7, i - 1, j - 1, 1) + F_(8, i + 1, j - 1, 1) + F_(9, i - 1, j + 1, 1) + F_(10, i + 1, j + 1, 1) +
                                       2.0 * (F_(6, i, j, 2) + F_(14, i + 1, j, 2) + F_(13, i - 1, j, 2) + F_(18, i, j + 1, 2) + F_(17, i, j - 1, 2)))) *
                               s->relaxation;
                    double tnx = 0.5 * (F_(1, i - 1, j, 1) + F_(7, i - 1, j - 1, 1) + F_(9, i - 1, j + 1, 1) -
                                        (F_(2, i + 1, j, 1) + F_(8, i + 1, j - 1, 1) + F_(10, i + 1, j + 1, 1)));
                    double tny = 0.5 * (F_(3, i, j - 1, 1) + F_(7, i - 1, j - 1, 1) + F_(8, i + 1, j - 1, 1) -
                                        (F_(4, i, j + 1, 1) + F_(10, i + 1, j + 1, 1) + F_(9, i - 1, j + 1, 1)));
                    F_(5, i, j, 0) = (F_(6, i, j, 2) + 0.333333333333333333 * t) * (1 - wi) + F_(5, i, j, 0) * wi;
                    F_(11, i - 1, j, 0) = (F_(14, i + 1, j, 2) + 0.166666666666666667 * t - tnx) * (1 - wi) + F_(11, i - 1, j, 0) * wi;
                    F_(12, i + 1, j, 0) = (F_(13, i - 1, j, 2) + 0.166666666666666667 * t + tnx) * (1 - wi) + F_(12, i + 1, j, 0) * wi;
                    F_(15, i, j - 1, 0) = (F_(18, i, j + 1, 2) + 0.166666666666666667 * t - tny) * (1 - wi) + F_(15, i, j - 1, 0) * wi;
                    F_(16, i, j + 1, 0) = (F_(17, i, j - 1, 2) + 0.166666666666666667 * t + tny) * (1 - wi) + F_(16, i, j + 1, 0) * wi;
                } else {
                    double t = (rin - (F_(0, i, j, 1) + F_(2, i, j, 1) + F_(1, i, j, 1) + F_(4, i, j, 1) + F_(3, i, j, 1) + F_(8, i, j, 1) +
                                       F_(7, i, j, 1) + F_(10, i, j, 1) + F_(9, i, j, 1) +
                                       2.0 * (F_(5, i, j, 1) + F_(11, i, j, 1) + F_(12, i, j, 1) + F_(15, i, j, 1) + F_(16, i, j, 1)))) *
                               s->relaxation;
                    double tnx = 0.5 * (F_(2, i, j, 1) + F_(8, i, j, 1) + F_(10, i, j, 1) - (F_(1, i, j, 1) + F_(7, i, j, 1) + F_(9, i, j, 1)));
                    double tny = 0.5 * (F_(4, i, j, 1) + F_(9, i, j, 1) + F_(10, i, j, 1) - (F_(3, i, j, 1) + F_(8, i, j, 1) + F_(7, i, j, 1)));
                    F_(6, i, j, 1) = (F_(5, i, j, 1) + 0.333333333333333333 * t) * (1 - wi) + F_(6, i, j, 1) * wi;
                    F_(13, i, j, 1) = (F_(12, i, j, 1) + 0.166666666666666667 * t + tnx) * (1 - wi) + F_(13, i, j, 1) * wi;
                    F_(14, i, j, 1) = (F_(11, i, j, 1) + 0.166666666666666667 * t - tnx) * (1 - wi) + F_(14, i, j, 1) * wi;
                    F_(17, i, j, 1) = (F_(16, i, j, 1) + 0.166666666666666667 * t + tny) * (1 - wi) + F_(17, i, j, 1) * wi;
                    F_(18, i, j, 1) = (F_(15, i, j, 1) + 0.166666666666666667 * t - tny) * (1 - wi) + F_(18, i, j, 1) * wi;
                }
            }
        }
}

/* MP/Boundary_multiphase_outlet.F90:7-127 ; SP/Boundary.F90:196-276 */
static void outlet_convective(orc_state *s, int after) {
    if (s->p.idz != s->p.npz - 1) return;
    const int nx = s->nx, ny = s->ny, nz = s->nz, mp = s->p.multiphase;
    const double uc = s->uin_avg;
    const double temp = 1.0 / (1.0 + uc);
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
            int wi = s->walls[I2(i, j, nz)];
            if (mp) {
                double *ph = s->phi;
                ph[I4(i, j, nz + 1)] = ((s->phi_convec_bc[IP2(i, j)] + uc * ph[I4(i, j, nz)]) * temp) * (1 - wi) + ph[I4(i, j, nz + 1)] * wi;
                s->phi_convec_bc[IP2(i, j)] = ph[I4(i, j, nz + 1)];
                ph[I4(i, j, nz + 2)] = ph[I4(i, j, nz + 1)];
                ph[I4(i, j, nz + 3)] = ph[I4(i, j, nz + 1)];
                ph[I4(i, j, nz + 4)] = ph[I4(i, j, nz + 1)];
            }
            for (int fl = 0; fl < (mp ? 2 : 1); fl++) {
                double **F = fl == 0 ? s->f : s->g;
                double *cb = fl == 0 ? s->f_convec_bc : s->g_convec_bc;
#define CB(q) cb[IPC(i, j, q)]
                if (!after) {
                    F_(6, i, j, nz + 1) = ((CB(6) + uc * F_(6, i, j, nz)) * temp) * (1 - wi) + F_(6, i, j, nz + 1) * wi;
                    F_(13, i - 1, j, nz + 1) = ((CB(13) + uc * F_(13, i - 1, j, nz)) * temp) * (1 - wi) + F_(13, i - 1, j, nz + 1) * wi;
                    F_(14, i + 1, j, nz + 1) = ((CB(14) + uc * F_(14, i + 1, j, nz)) * temp) * (1 - wi) + F_(14, i + 1, j, nz + 1) * wi;
                    F_(17, i, j - 1, nz + 1) = ((CB(17) + uc * F_(17, i, j - 1, nz)) * temp) * (1 - wi) + F_(17, i, j - 1, nz + 1) * wi;
                    F_(18, i, j + 1, nz + 1) = ((CB(18) + uc * F_(18, i, j + 1, nz)) * temp) * (1 - wi) + F_(18, i, j + 1, nz + 1) * wi;
                    CB(6) = F_(6, i, j, nz + 1);
                    CB(13) = F_(13, i - 1, j, nz + 1);
                    CB(14) = F_(14, i + 1, j, nz + 1);
                    CB(17) = F_(17, i, j - 1, nz + 1);
                    CB(18) = F_(18, i, j + 1, nz + 1);
                } else {
                    F_(5, i, j, nz) = ((CB(6) + uc * F_(5, i, j, nz - 1)) * temp) * (1 - wi) + F_(5, i, j, nz) * wi;
                    F_(11, i, j, nz) = ((CB(14) + uc * F_(11, i, j, nz - 1)) * temp) * (1 - wi) + F_(11, i, j, nz) * wi;
                    F_(12, i, j, nz) = ((CB(13) + uc * F_(12, i, j, nz - 1)) * temp) * (1 - wi) + F_(12, i, j, nz) * wi;
                    F_(15, i, j, nz) = ((CB(18) + uc * F_(15, i, j, nz - 1)) * temp) * (1 - wi) + F_(15, i, j, nz) * wi;
                    F_(16, i, j, nz) = ((CB(17) + uc * F_(16, i, j, nz - 1)) * temp) * (1 - wi) + F_(16, i, j, nz) * wi;
                    CB(6) = F_(5, i, j, nz);
                    CB(14) = F_(11, i, j, nz);
                    CB(13) = F_(12, i, j, nz);
                    CB(18) = F_(15, i, j, nz);
                    CB(17) = F_(16, i, j, nz);
                }
#undef CB
            }
        }
}

/* MP/Boundary_multiphase_outlet.F90:139-311 ; SP/Boundary.F90:284-390 */
static void outlet_pressure(orc_state *s, int after) {
    if (s->p.idz != s->p.npz - 1) return;
    const int nx = s->nx, ny = s->ny, nz = s->nz, mp = s->p.multiphase;
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
            int wi = s->walls[I2(i, j, nz)];
            double dwi = 1.0 - wi;
            double tmp1, tmp2 = 0.0;
            if (mp) {
                double *ph = s->phi;
                ph[I4(i, j, nz + 1)] = ph[I4(i, j, nz)];
                ph[I4(i, j, nz + 2)] = ph[I4(i, j, nz)];
                ph[I4(i, j, nz + 3)] = ph[I4(i, j, nz)];
                ph[I4(i, j, nz + 4)] = ph[I4(i, j, nz)];
            }
            double **F = s->f, **G = s->g;
#define G_(q, i, j, k) G[q][I1(i, j, k)]
            if (!after) {
                if (mp)
                    tmp1 = (F_(0, i, j, nz) + F_(1, i - 1, j, nz) + F_(2, i + 1, j, nz) + F_(3, i, j - 1, nz) + F_(4, i, j + 1, nz) +
                            F_(7, i - 1, j - 1, nz) + F_(8, i + 1, j - 1, nz) + F_(9, i - 1, j + 1, nz) + F_(10, i + 1, j + 1, nz) +
                            2.0 * (F_(5, i, j, nz - 1) + F_(11, i - 1, j, nz - 1) + F_(12, i + 1, j, nz - 1) + F_(15, i, j - 1, nz - 1) + F_(16, i, j + 1, nz - 1)) +
                            G_(0, i, j, nz) + G_(1, i - 1, j, nz) + G_(2, i + 1, j, nz) + G_(3, i, j - 1, nz) + G_(4, i, j + 1, nz) +
                            G_(7, i - 1, j - 1, nz) + G_(8, i + 1, j - 1, nz) + G_(9, i - 1, j + 1, nz) + G_(10, i + 1, j + 1, nz) +
                            2.0 * (G_(5, i, j, nz - 1) + G_(11, i - 1, j, nz - 1) + G_(12, i + 1, j, nz - 1) + G_(15, i, j - 1, nz - 1) + G_(16, i, j + 1, nz - 1))) -
                           s->rho_out;
                else
                    tmp1 = (F_(0, i, j, nz) + F_(1, i - 1, j, nz) + F_(2, i + 1, j, nz) + F_(3, i, j - 1, nz) + F_(4, i, j + 1, nz) +
                            F_(7, i - 1, j - 1, nz) + F_(8, i + 1, j - 1, nz) + F_(9, i - 1, j + 1, nz) + F_(10, i + 1, j + 1, nz) +
                            2.0 * (F_(5, i, j, nz - 1) + F_(11, i - 1, j, nz - 1) + F_(12, i + 1, j, nz - 1) + F_(15, i, j - 1, nz - 1) + F_(16, i, j + 1, nz - 1))) -
                           s->rho_out;
                if (mp) {
                    tmp2 = tmp1 * 0.5 * (1.0 - s->phi[I4(i, j, nz)]);
                    tmp1 = tmp1 - tmp2;
                }
                for (int fl = 0; fl < (mp ? 2 : 1); fl++) {
                    double **F = fl == 0 ? s->f : s->g;
                    double t = fl == 0 ? tmp1 : tmp2;
                    double tnx = 0.5 * (F_(1, i - 1, j, nz) + F_(7, i - 1, j - 1, nz) + F_(9, i - 1, j + 1, nz) -
                                        (F_(2, i + 1, j, nz) + F_(8, i + 1, j - 1, nz) + F_(10, i + 1, j + 1, nz)));
                    double tny = 0.5 * (F_(3, i, j - 1, nz) + F_(7, i - 1, j - 1, nz) + F_(8, i + 1, j - 1, nz) -
                                        (F_(4, i, j + 1, nz) + F_(10, i + 1, j + 1, nz) + F_(9, i - 1, j + 1, nz)));
                    F_(6, i, j, nz + 1) = (F_(5, i, j, nz - 1) - 0.333333333333333333 * t) * dwi + F_(6, i, j, nz + 1) * wi;
                    F_(13, i - 1, j, nz + 1) = (F_(12, i + 1, j, nz - 1) - 0.166666666666666667 * t - tnx) * dwi + F_(13, i - 1, j, nz + 1) * wi;
                    F_(14, i + 1, j, nz + 1) = (F_(11, i - 1, j, nz - 1) - 0.166666666666666667 * t + tnx) * dwi + F_(14, i + 1, j, nz + 1) * wi;
                    F_(17, i, j - 1, nz + 1) = (F_(16, i, j + 1, nz - 1) - 0.166666666666666667 * t - tny) * dwi + F_(17, i, j - 1, nz + 1) * wi;
                    F_(18, i, j + 1, nz + 1) = (F_(15, i, j - 1, nz - 1) - 0.166666666666666667 * t + tny) * dwi + F_(18, i, j + 1, nz + 1) * wi;
                }
            } else {
                if (mp)
                    tmp1 = (F_(0, i, j, nz) + F_(2, i, j, nz) + F_(1, i, j, nz) + F_(4, i, j, nz) + F_(3, i, j, nz) + F_(8, i, j, nz) + F_(7, i, j, nz) +
                            F_(10, i, j, nz) + F_(9, i, j, nz) + 2.0 * (F_(6, i, j, nz) + F_(14, i, j, nz) + F_(13, i, j, nz) + F_(18, i, j, nz) + F_(17, i, j, nz)) +
                            G_(0, i, j, nz) + G_(2, i, j, nz) + G_(1, i, j, nz) + G_(4, i, j, nz) + G_(3, i, j, nz) + G_(8, i, j, nz) + G_(7, i, j, nz) +
                            G_(10, i, j, nz) + G_(9, i, j, nz) + 2.0 * (G_(6, i, j, nz) + G_(14, i, j, nz) + G_(13, i, j, nz) + G_(18, i, j, nz) + G_(17, i, j, nz))) -
                           s->rho_out;
                else
                    tmp1 = (F_(0, i, j, nz) + F_(2, i, j, nz) + F_(1, i, j, nz) + F_(4, i, j, nz) + F_(3, i, j, nz) + F_(8, i, j, nz) + F_(7, i, j, nz) +
                            F_(10, i, j, nz) + F_(9, i, j, nz) + 2.0 * (F_(6, i, j, nz) + F_(14, i, j, nz) + F_(13, i, j, nz) + F_(18, i, j, nz) + F_(17, i, j, nz))) -
                           s->rho_out;
                if (mp) {
                    tmp2 = tmp1 * 0.5 * (1.0 - s->phi[I4(i, j, nz)]);
                    tmp1 = tmp1 - tmp2;
                }
                for (int fl = 0; fl < (mp ? 2 : 1); fl++) {
                    double **F = fl == 0 ? s->f : s->g;
                    double t = fl == 0 ? tmp1 : tmp2;
                    double tnx = 0.5 * (F_(2, i, j, nz) + F_(8, i, j, nz) + F_(10, i, j, nz) - (F_(1, i, j, nz) + F_(7, i, j, nz) + F_(9, i, j, nz)));
                    double tny = 0.5 * (F_(4, i, j, nz) + F_(10, i, j, nz) + F_(9, i, j, nz) - (F_(3, i, j, nz) + F_(7, i, j, nz) + F_(8, i, j, nz)));
                    F_(5, i, j, nz) = (F_(6, i, j, nz) - 0.333333333333333333 * t) * dwi + F_(5, i, j, nz) * wi;
                    F_(11, i, j, nz) = (F_(14, i, j, nz) - 0.166666666666666667 * t + tnx) * dwi + F_(11, i, j, nz) * wi;
                    F_(12, i, j, nz) = (F_(13, i, j, nz) - 0.166666666666666667 * t - tnx) * dwi + F_(12, i, j, nz) * wi;
                    F_(15, i, j, nz) = (F_(18, i, j, nz) - 0.166666666666666667 * t + tny) * dwi + F_(15, i, j, nz) * wi;
                    F_(16, i, j, nz) = (F_(17, i, j, nz) - 0.166666666666666667 * t - tny) * dwi + F_(16, i, j, nz) * wi;
                }
            }
#undef G_
        }
}

/* MP/Boundary_multiphase_other.F90:8-184 porous_plate_BC_{before,after}_odd */
static void porous_plate(orc_state *s, int after) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    int zmin = s->p.idz * nz + 1, zmax = s->p.idz * nz + nz;
    if (!(s->p.Z_porous_plate >= zmin && s->p.Z_porous_plate <= zmax)) return;
    int zp = s->p.Z_porous_plate - s->p.idz * nz;
    if (s->p.porous_plate_cmd != 1 && s->p.porous_plate_cmd != 2) return;
    double **B = s->p.porous_plate_cmd == 1 ? s->f : s->g; /* blocked fluid: bounce */
    double **T = s->p.porous_plate_cmd == 1 ? s->g : s->f; /* passing fluid: copy through */
#define B_(q, i, j, k) B[q][I1(i, j, k)]
#define T_(q, i, j, k) T[q][I1(i, j, k)]
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
            if (!after) {
                B_(6, i, j, zp) = B_(5, i, j, zp - 1);
                B_(13, i - 1, j, zp) = B_(12, i, j, zp - 1);
                B_(14, i + 1, j, zp) = B_(11, i, j, zp - 1);
                B_(17, i, j - 1, zp) = B_(16, i, j, zp - 1);
                B_(18, i, j + 1, zp) = B_(15, i, j, zp - 1);
                B_(5, i, j, zp) = B_(6, i, j, zp + 1);
                B_(12, i + 1, j, zp) = B_(13, i, j, zp + 1);
                B_(11, i - 1, j, zp) = B_(14, i, j, zp + 1);
                B_(16, i, j + 1, zp) = B_(17, i, j, zp + 1);
                B_(15, i, j - 1, zp) = B_(18, i, j, zp + 1);
                T_(6, i, j, zp) = T_(6, i, j, zp + 1);
                T_(13, i, j, zp) = T_(13, i, j, zp + 1);
                T_(14, i, j, zp) = T_(14, i, j, zp + 1);
                T_(17, i, j, zp) = T_(17, i, j, zp + 1);
                T_(18, i, j, zp) = T_(18, i, j, zp + 1);
                T_(5, i, j, zp) = T_(5, i, j, zp - 1);
                T_(12, i, j, zp) = T_(12, i, j, zp - 1);
                T_(11, i, j, zp) = T_(11, i, j, zp - 1);
                T_(16, i, j, zp) = T_(16, i, j, zp - 1);
                T_(15, i, j, zp) = T_(15, i, j, zp - 1);
            } else {
                B_(5, i, j, zp - 1) = B_(6, i, j, zp);
                B_(11, i, j, zp - 1) = B_(14, i + 1, j, zp);
                B_(12, i, j, zp - 1) = B_(13, i - 1, j, zp);
                B_(15, i, j, zp - 1) = B_(18, i, j + 1, zp);
                B_(16, i, j, zp - 1) = B_(17, i, j - 1, zp);
                B_(6, i, j, zp + 1) = B_(5, i, j, zp);
                B_(14, i, j, zp + 1) = B_(11, i - 1, j, zp);
                B_(13, i, j, zp + 1) = B_(12, i + 1, j, zp);
                B_(18, i, j, zp + 1) = B_(15, i, j - 1, zp);
                B_(17, i, j, zp + 1) = B_(16, i, j + 1, zp);
                T_(5, i, j, zp - 1) = T_(5, i, j, zp);
                T_(11, i, j, zp - 1) = T_(11, i, j, zp);
                T_(12, i, j, zp - 1) = T_(12, i, j, zp);
                T_(15, i, j, zp - 1) = T_(15, i, j, zp);
                T_(16, i, j, zp - 1) = T_(16, i, j, zp);
                T_(6, i, j, zp + 1) = T_(6, i, j, zp);
                T_(14, i, j, zp + 1) = T_(14, i, j, zp);
                T_(13, i, j, zp + 1) = T_(13, i, j, zp);
                T_(18, i, j, zp + 1) = T_(18, i, j, zp);
                T_(17, i, j, zp + 1) = T_(17, i, j, zp);
            }
        }
#undef B_
#undef T_
}

/* =====================================================================================
 * halo exchange for np = 1 (self send/recv on the periodic Cartesian communicator)
 * MP/Mpi.F90:101-598 (z faces only) and :608-867 (phi)
 * ===================================================================================== */
static void periodic_z_pdf(orc_state *s, int push) {
    const int nx = s->nx, ny = s->ny, nz = s->nz, mp = s->p.multiphase;
    static const int qM[5] = {6, 14, 13, 18, 17}; /* e_z = -1 */
    static const int qP[5] = {5, 11, 12, 15, 16}; /* e_z = +1 */
    for (int fl = 0; fl < (mp ? 2 : 1); fl++) {
        double **F = fl == 0 ? s->f : s->g;
        for (int m = 0; m < 5; m++)
            for (int j = 1; j <= ny; j++)
                for (int i = 1; i <= nx; i++) {
                    if (!push) { /* pull, :121-143 + :244-266: own k=1 / k=nz planes -> zM/zP neighbour ghosts */
                        F_(qM[m], i, j, nz + 1) = F_(qM[m], i, j, 1);
                        F_(qP[m], i, j, 0) = F_(qP[m], i, j, nz);
                    } else { /* push, :374-396 + :497-519: own ghost planes -> neighbour interior planes */
                        F_(qP[m], i, j, nz) = F_(qP[m], i, j, 0);
                        F_(qM[m], i, j, 1) = F_(qM[m], i, j, nz + 1);
                    }
                }
    }
}

static void periodic_z_phi(orc_state *s) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    /* MP/Mpi.F90:624-631 pack, :702-727 update; both ends active since kper==1 */
    for (int k = 1; k <= 4; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                double lo = s->phi[I4(i, j, k)];           /* send_phi_zM */
                double hi = s->phi[I4(i, j, nz + k - 4)];  /* send_phi_zP */
                s->phi[I4(i, j, k - 4)] = hi;              /* recv_phi_zM */
                s->phi[I4(i, j, k + nz)] = lo;             /* recv_phi_zP */
            }
}

/* y faces, self exchange of a y-periodic lattice (npy == 1): MP/Mpi.F90:147-180, :270-305 (pull), :398-430, :521-553 (push).
 * Rows i = 1..nx, k = 1..nz only, exactly like the reference's buffers. */
static void periodic_y_pdf(orc_state *s, int push) {
    const int nx = s->nx, ny = s->ny, nz = s->nz, mp = s->p.multiphase;
    static const int qM[5] = {4, 10, 9, 16, 18}; /* e_y = -1 */
    static const int qP[5] = {3, 7, 8, 15, 17};  /* e_y = +1 */
    for (int fl = 0; fl < (mp ? 2 : 1); fl++) {
        double **F = fl == 0 ? s->f : s->g;
        for (int m = 0; m < 5; m++)
            for (int k = 1; k <= nz; k++)
                for (int i = 1; i <= nx; i++) {
                    if (!push) { /* own j=1 / j=ny rows -> yM / yP neighbour ghosts */
                        F_(qM[m], i, ny + 1, k) = F_(qM[m], i, 1, k);
                        F_(qP[m], i, 0, k) = F_(qP[m], i, ny, k);
                    } else { /* own ghost rows -> neighbour interior rows */
                        F_(qP[m], i, ny, k) = F_(qP[m], i, 0, k);
                        F_(qM[m], i, 1, k) = F_(qM[m], i, ny + 1, k);
                    }
                }
    }
}

/* edges along x when y AND z are exchanged (MP/Mpi.F90:184-207, :316-341 pull; :434-456, :566-590 push); applied after
 * the faces ("face communication will contaminate edge communication") */
static void periodic_yz_edges_pdf(orc_state *s, int push) {
    const int nx = s->nx, ny = s->ny, nz = s->nz, mp = s->p.multiphase;
    for (int fl = 0; fl < (mp ? 2 : 1); fl++) {
        double **F = fl == 0 ? s->f : s->g;
        for (int i = 1; i <= nx; i++) {
            if (!push) {
                F_(18, i, ny + 1, nz + 1) = F_(18, i, 1, 1);   /* send yMzM -> recv yPzP */
                F_(16, i, ny + 1, 0) = F_(16, i, 1, nz);       /* send yMzP -> recv yPzM */
                F_(17, i, 0, nz + 1) = F_(17, i, ny, 1);       /* send yPzM -> recv yMzP */
                F_(15, i, 0, 0) = F_(15, i, ny, nz);           /* send yPzP -> recv yMzM */
            } else {
                F_(15, i, ny, nz) = F_(15, i, 0, 0);           /* send yMzM (ghost) -> recv yPzP */
                F_(17, i, ny, 1) = F_(17, i, 0, nz + 1);       /* send yMzP -> recv yPzM */
                F_(16, i, 1, nz) = F_(16, i, ny + 1, 0);       /* send yPzM -> recv yMzP */
                F_(18, i, 1, 1) = F_(18, i, ny + 1, nz + 1);   /* send yPzP -> recv yMzM */
            }
        }
    }
}

/* phi: y faces (k = 1..nz) and, with z periodic too, the four x edges (MP/Mpi.F90:633-655 pack, :729-790 update) */
static void periodic_y_phi(orc_state *s, int with_z_edges) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= 4; j++)
            for (int i = 1; i <= nx; i++) {
                double lo = s->phi[I4(i, j, k)], hi = s->phi[I4(i, ny + j - 4, k)];
                s->phi[I4(i, j - 4, k)] = hi;
                s->phi[I4(i, j + ny, k)] = lo;
            }
    if (!with_z_edges) return;
    for (int k = 1; k <= 4; k++)
        for (int j = 1; j <= 4; j++)
            for (int i = 1; i <= nx; i++) {
                double mm = s->phi[I4(i, j, k)], pp = s->phi[I4(i, ny + j - 4, nz + k - 4)];
                double mp_ = s->phi[I4(i, j, nz + k - 4)], pm = s->phi[I4(i, ny + j - 4, k)];
                s->phi[I4(i, j - 4, k - 4)] = pp;    /* recv yMzM <- send yPzP */
                s->phi[I4(i, j + ny, k - 4)] = mp_;  /* recv yPzM <- send yMzP */
                s->phi[I4(i, j + ny, k + nz)] = mm;  /* recv yPzP <- send yMzM */
                s->phi[I4(i, j - 4, k + nz)] = pm;   /* recv yMzP <- send yPzM */
            }
}

static void periodic_exchange(orc_state *s, int push) {
    const orc_params *p = &s->p;
    if (p->kper == 1) periodic_z_pdf(s, push);
    if (p->jper == 1) periodic_y_pdf(s, push);
    if (p->kper == 1 && p->jper == 1) periodic_yz_edges_pdf(s, push);
    if (p->multiphase) {
        if (p->kper == 1) periodic_z_phi(s);
        if (p->jper == 1) periodic_y_phi(s, p->kper == 1);
    }
}

/* =====================================================================================
 * main_iteration_kernel, MP/Main_multiphase.F90:341-486 ; SP/Main.F90:291-422  (np == 1)
 * ===================================================================================== */
void orc_step(orc_state *s, int ntime) {
    const orc_params *p = &s->p;
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const int open_z = (p->kper == 0 && p->wsz0 == 0 && p->wsz1 == 0);
    if (p->npz != 1) {
        fprintf(stderr, "oracle: orc_step supports np=1 only\n");
        abort();
    }
    if (ntime % 2 == 0) {
        orc_kernel_even(s, 1, nx, 1, ny, 1, nz);
        periodic_exchange(s, 0);
        if (open_z) {
            if (p->inlet_BC == 1) inlet_velocity(s, 0);
            else if (p->inlet_BC == 2) inlet_pressure(s, 0);
            if (p->outlet_BC == 1) outlet_convective(s, 0);
            else if (p->outlet_BC == 2) outlet_pressure(s, 0);
        }
        if (p->multiphase && p->porous_plate_cmd != 0) porous_plate(s, 0);
    } else {
        orc_kernel_odd(s, 1, nx, 1, ny, 1, nz);
        periodic_exchange(s, 1);
        if (open_z) {
            if (p->inlet_BC == 1) inlet_velocity(s, 1);
            else if (p->inlet_BC == 2) inlet_pressure(s, 1);
            if (p->outlet_BC == 1) outlet_convective(s, 1);
            else if (p->outlet_BC == 2) outlet_pressure(s, 1);
        }
        if (p->multiphase && p->porous_plate_cmd != 0) porous_plate(s, 1);
    }
    if (p->multiphase) orc_color_gradient(s);
}

/* =====================================================================================
 * macroscopic variables and monitors
 * ===================================================================================== */
/* MP/Misc.F90:372-430 ; SP/Misc.F90:368-423 */
void orc_compute_macro_vars(orc_state *s) {
    const int nx = s->nx, ny = s->ny, nz = s->nz, mp = s->p.multiphase;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                size_t c = I1(i, j, k);
                int wi = s->walls[I2(i, j, k)];
                double ft[19];
                for (int q = 0; q < 19; q++) ft[q] = mp ? s->f[q][c] + s->g[q][c] : s->f[q][c];
                s->rho[c] = (ft[0] + ft[1] + ft[2] + ft[3] + ft[4] + ft[5] + ft[6] + ft[7] + ft[8] + ft[9] + ft[10] + ft[11] + ft[12] +
                             ft[13] + ft[14] + ft[15] + ft[16] + ft[17] + ft[18]) * (1 - wi);
                double fx = 0, fy = 0, fz = s->force_Z;
                if (mp) {
                    size_t c2 = I2(i, j, k);
                    double tmp = 0.5 * s->p.gamma * s->curv[c] * s->c_norm[c2];
                    fx = tmp * s->cn_x[c2];
                    fy = tmp * s->cn_y[c2];
                    fz = tmp * s->cn_z[c2] + s->force_Z;
                }
                s->u[c] = (ft[1] - ft[2] + ft[7] - ft[8] + ft[9] - ft[10] + ft[11] - ft[12] + ft[13] - ft[14] - 0.5 * fx) * (1 - wi);
                s->v[c] = (ft[3] - ft[4] + ft[7] + ft[8] - ft[9] - ft[10] + ft[15] - ft[16] + ft[17] - ft[18] - 0.5 * fy) * (1 - wi);
                s->w[c] = (ft[5] - ft[6] + ft[11] + ft[12] - ft[13] - ft[14] + ft[15] + ft[16] - ft[17] - ft[18] - 0.5 * fz) * (1 - wi);
                if (mp) s->phi[I4(i, j, k)] = 0.0 * wi + s->phi[I4(i, j, k)] * (1 - wi);
            }
}

/* MP/Monitor.F90:5-277 (np==1: the rank-0 branch) ; SP/Monitor.F90:4-172.
 * Sums run in plain i,j,k order (one OpenMP thread per z-slice in the reference). */
void orc_monitor(orc_state *s, orc_monitor_out *o) {
    const orc_params *p = &s->p;
    const int nx = s->nx, ny = s->ny, nz = s->nz, mp = p->multiphase;
    memset(o, 0, sizeof(*o));
    orc_compute_macro_vars(s);
    double umax = 0, usq1 = 0, usq2 = 0;
    for (int k = 1; k <= nz; k++) {
        double t1 = 0, t2 = 0, prek = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0;
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                size_t c = I1(i, j, k);
                int wi = s->walls[I2(i, j, k)];
                if (mp) {
                    double ph = s->phi[I4(i, j, k)];
                    double temp = (s->u[c] * s->u[c] + s->v[c] * s->v[c] + s->w[c] * s->w[c]) * (1 - wi);
                    if (umax < temp) umax = temp;
                    if (ph > 0.999) usq1 = usq1 + temp;
                    else if (ph < -0.999) usq2 = usq2 + temp;
                    t3 = t3 + 0.5 * (1.0 + ph) * (1 - wi);
                    t4 = t4 + 0.5 * (1.0 - ph) * (1 - wi);
                    t5 = t5 + s->rho[c] * 0.5 * (1.0 + ph) * (1 - wi);
                    t6 = t6 + s->rho[c] * 0.5 * (1.0 - ph) * (1 - wi);
                    t1 = t1 + s->w[c] * 0.5 * (1.0 + ph) * (1 - wi);
                    t2 = t2 + s->w[c] * 0.5 * (1.0 - ph) * (1 - wi);
                    prek = prek + s->rho[c] * (1 - wi);
                } else {
                    double temp = s->u[c] * s->u[c] + s->v[c] * s->v[c] + s->w[c] * s->w[c];
                    if (umax < temp) umax = temp;
                    t1 = t1 + s->w[c];
                    prek = prek + s->rho[c];
                }
            }
        s->pre[k - 1] = prek; s->fl1[k - 1] = t1; s->fl2[k - 1] = t2;
        s->vol1[k - 1] = t3; s->vol2[k - 1] = t4; s->mass1[k - 1] = t5; s->mass2[k - 1] = t6;
    }
    o->umax = umax; o->usq1 = usq1; o->usq2 = usq2;
    o->umax_global = sqrt(umax);
    if (p->npz != 1) return; /* slab: caller gathers the per-slice arrays */
    const int nzG = p->nzG;
    const int k0 = p->n_exclude_inlet + 1, k1 = nzG - p->n_exclude_outlet;
    if (mp) {
        double m1 = 0, m2 = 0, v1 = 0, v2 = 0;
        for (int k = k0; k <= k1; k++) { m1 += s->mass1[k - 1]; m2 += s->mass2[k - 1]; v1 += s->vol1[k - 1]; v2 += s->vol2[k - 1]; }
        o->mass1_sum = m1; o->mass2_sum = m2; o->vol1_sum = v1; o->vol2_sum = v2;
        o->saturation = v1 / (v1 + v2);
        double t3 = 0, t4 = 0;
        for (int k = 1; k <= nzG; k++) { t3 += s->vol1[k - 1]; t4 += s->vol2[k - 1]; }
        o->saturation_full_domain = t3 / (t3 + t4);
        double f1w = 0, f2w = 0, f1 = 0, f2 = 0;
        for (int k = 1; k <= nzG; k++) { f1w += s->fl1[k - 1]; f2w += s->fl2[k - 1]; }
        f1w = f1w / (double)nzG; f2w = f2w / (double)nzG;
        for (int k = k0; k <= k1; k++) { f1 += s->fl1[k - 1]; f2 += s->fl2[k - 1]; }
        f1 = f1 / (double)(nzG - p->n_exclude_outlet - p->n_exclude_inlet);
        f2 = f2 / (double)(nzG - p->n_exclude_outlet - p->n_exclude_inlet);
        o->fl1_avg_whole = f1w; o->fl2_avg_whole = f2w; o->fl_avg_whole = f1w + f2w;
        o->fl1_avg = f1; o->fl2_avg = f2; o->fl_avg = f1 + f2;
        o->ca = (o->fl_avg / s->A_xy) * p->la_nu1 / p->gamma;
    } else {
        double fw = 0;
        for (int k = 1; k <= nzG; k++) fw += s->fl1[k - 1];
        o->fl_avg_whole = fw / (double)nzG;
    }
    if (p->kper == 0 && p->wsz0 == 0 && p->wsz1 == 0) {
        o->pre_in = s->pre[k0 - 1] / s->pore_profile_z[k0 - 1];
        o->pre_out = s->pre[k1 - 1] / s->pore_profile_z[k1 - 1];
    }
}

/* MP/Monitor.F90:512-550 */
void orc_cal_saturation(orc_state *s, orc_monitor_out *o) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    double v1 = 0, v2 = 0;
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                int wi = s->walls[I2(i, j, k)];
                v1 = v1 + 0.5 * (1.0 + s->phi[I4(i, j, k)]) * (1 - wi);
                v2 = v2 + 0.5 * (1.0 - s->phi[I4(i, j, k)]) * (1 - wi);
            }
    o->vol1_sum = v1;
    o->vol2_sum = v2;
    o->saturation_full_domain = v1 / (v1 + v2 + EPS_MP);
}

/* MP/Monitor.F90:472-507 */
void orc_monitor_breakthrough(orc_state *s, orc_monitor_out *o) {
    int itemp = 0;
    if (s->p.idz == s->p.npz - 1) {
        int obs_z = s->nz - 1;
        for (int j = 1; j <= s->ny; j++)
            for (int i = 1; i <= s->nx; i++)
                if (s->walls[I2(i, j, obs_z)] == 0 && s->phi[I4(i, j, obs_z)] > 0.0) itemp++;
    }
    o->outlet_phase1_sum = itemp;
}

/* MP/Monitor.F90:287-365 */
void orc_monitor_steady_phasefield(orc_state *s, orc_monitor_out *o) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    double umax = 0, dmax = 0;
    orc_compute_macro_vars(s);
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                size_t c = I1(i, j, k);
                int wi = s->walls[I2(i, j, k)];
                double t1 = (s->u[c] * s->u[c] + s->v[c] * s->v[c] + s->w[c] * s->w[c]) * (1 - wi);
                if (umax < t1) umax = t1;
                double t2 = fabs(s->phi[I4(i, j, k)] - s->phi_old[I4(i, j, k)]) * (1 - wi);
                if (dmax < t2) dmax = t2;
            }
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) s->phi_old[I4(i, j, k)] = s->phi[I4(i, j, k)];
    o->umax = umax;
    o->umax_global = sqrt(umax);
    o->d_phi_max = dmax;
}

/* MP/Monitor.F90:370-462 */
void orc_monitor_steady_capillarypressure(orc_state *s, orc_monitor_out *o) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    double umax = 0, pw = 0, pnw = 0;
    int iw = 0, inw = 0;
    orc_compute_macro_vars(s);
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                size_t c = I1(i, j, k);
                int wi = s->walls[I2(i, j, k)];
                double t = (s->u[c] * s->u[c] + s->v[c] * s->v[c] + s->w[c] * s->w[c]) * (1 - wi);
                if (umax < t) umax = t;
                if (wi == 0) {
                    double ph = s->phi[I4(i, j, k)];
                    if (ph < -0.99) { pw = pw + s->rho[c]; iw++; }
                    if (ph > 0.99) { pnw = pnw + s->rho[c]; inw++; }
                }
            }
    o->umax = umax;
    o->umax_global = sqrt(umax);
    o->pre_w = pw; o->pre_nw = pnw; o->i_w = iw; o->i_nw = inw;
}

/* =====================================================================================
 * accessors
 * ===================================================================================== */
double *orc_f(orc_state *s, int q) { return s->f[q]; }
double *orc_g(orc_state *s, int q) { return s->g[q]; }
int8_t *orc_walls(orc_state *s) { return s->walls; }
orc_solid_node *orc_solid_nodes(orc_state *s, int *n) { *n = s->num_solid; return s->solid; }
orc_fluid_node *orc_fluid_nodes(orc_state *s, int *n) { *n = s->num_fluid; return s->fluid; }

double *orc_field(orc_state *s, const char *name) {
#define FLD(x) if (!strcmp(name, #x)) return s->x;
    FLD(phi) FLD(phi_old) FLD(cn_x) FLD(cn_y) FLD(cn_z) FLD(c_norm) FLD(curv) FLD(u) FLD(v) FLD(w) FLD(rho) FLD(w_in)
    FLD(f_convec_bc) FLD(g_convec_bc) FLD(phi_convec_bc) FLD(fl1) FLD(fl2) FLD(vol1) FLD(vol2) FLD(mass1) FLD(mass2) FLD(pre)
#undef FLD
    return NULL;
}

int orc_get_int(orc_state *s, const char *name) {
#define GI(x) if (!strcmp(name, #x)) return (int)s->x;
    GI(nx) GI(ny) GI(nz) GI(pore_sum) GI(pore_sum_effective) GI(num_solid) GI(num_fluid) GI(num_solid_global) GI(num_fluid_global)
    GI(ntime_max)
#undef GI
    if (!strcmp(name, "pore_profile_z_ptr")) return 0;
    return -1;
}
long long orc_get_i64(orc_state *s, const char *name) {
    if (!strcmp(name, "pore_sum")) return s->pore_sum;
    if (!strcmp(name, "pore_sum_effective")) return s->pore_sum_effective;
    return -1;
}
double orc_get_double(orc_state *s, const char *name) {
#define GD(x) if (!strcmp(name, #x)) return s->x;
    GD(la_nui1) GD(la_nui2) GD(theta) GD(phi_inlet) GD(force_Z) GD(rho_out) GD(rho_in) GD(relaxation) GD(uin_avg) GD(uin_avg_0)
    GD(flowrate) GD(la_x) GD(la_y) GD(la_z) GD(A_xy) GD(A_xy_effective) GD(rk_weight2) GD(s_e) GD(s_e2) GD(s_q) GD(s_nu) GD(s_pi) GD(s_t)
#undef GD
    return NAN;
}
void orc_set_double(orc_state *s, const char *name, double v) {
#define SD(x) if (!strcmp(name, #x)) { s->x = v; return; }
    SD(force_Z) SD(rho_in) SD(rho_out) SD(uin_avg) SD(phi_inlet) SD(relaxation) SD(theta)
#undef SD
}
void orc_set_int(orc_state *s, const char *name, int v) {
    if (!strcmp(name, "mrt")) s->p.mrt = v;
    if (!strcmp(name, "porous_plate_cmd")) s->p.porous_plate_cmd = v;
    if (!strcmp(name, "Z_porous_plate")) s->p.Z_porous_plate = v;
}
