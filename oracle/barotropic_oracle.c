/* barotropic_oracle.c -- CPU restatement of XLab-FFTBarotropic's pseudospectral RK4 step.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this; the product path (xlab_fftbarotropic_b200/csrc) never
 * links or calls it and fails loudly when its CUDA library is missing.
 *
 * Every function cites the reference file:line it follows (paths under /root/reference).
 * Arithmetic is float32 in the reference's own evaluation order so that, linked against the
 * same FFT (oracle/fftw3_shim), the results are BIT-IDENTICAL to the unmodified reference
 * main.cpp / invert_pres.cpp built by oracle/build_oracle.py into oracle/_ref/
 * (tests/test_oracle.py::test_restatement_matches_reference_binary pins that).
 * Compile WITHOUT -mfma / -ffast-math: the reference is built with plain `g++ -O3`
 * (Makefile:1-2), so no multiply-add is ever contracted.
 *
 * Parity pin: the reference holds no golden vectors.  This oracle is
 * pinned against the reference itself, run here: tests/golden/ holds outputs of the unmodified
 * reference binaries (generator script: tests/golden/make_golden.py).
 * The three diagnostics (filamentation time, effective diffusivity, deformation factor) have
 * no reference code at all (README.md:5-7 only): for those this file is "parity unpinned",
 * a restatement of the cited papers' formulas.
 */
#include "fftw3.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n;          /* XPTS == YPTS == n */
    int hy;         /* HALF_YPTS = n/2+1 */
    size_t grids, hgrids;
    float lx, ly, nu;
    float *gradx_coe, *grady_coe, *lap, *lapinv, *mask;
    /* model state, main.cpp:24,31 */
    float *vort, *u, *v, *dvortdx, *dvortdy, *dvortdt, *workspace, *vort_src;
    fftwf_complex *vort_c0, *vort_c, *lvort_c, *dvortdt_c, *tmp_c, *psi_c, *rk1_c, *rk2_c, *rk3_c, *rk4_c;
    fftwf_plan p_fwd, p_bwd;
    /* passive tracer (no reference code, see orc_set_tracer) */
    int has_tracer;
    float kappa;
    float *trc, *dtrcdx, *dtrcdy;
    fftwf_complex *trc_c0, *trc_c, *ltrc_c, *dtrcdt_c, *trk1_c, *trk2_c, *trk3_c, *trk4_c;
} orc_t;

#define HIDX(o, i, j) ((size_t)(o)->hy * (size_t)(i) + (size_t)(j))

/* fftwfop.cpp:5-79 (constructor): wavenumber, Laplacian and dealiasing tables */
static void build_tables(orc_t *o)
{
    const int n = o->n, hx = n / 2 + 1, hy = o->hy;
    const float TWOPI = acosf(-1.0f) * 2.0f;                    /* fftwfop.hpp:7 */
    const int dkx = (int)ceil(((float)n) / 3.0);                /* fftwfop.cpp:11 */
    const int dky = (int)ceil(((float)n) / 3.0);                /* fftwfop.cpp:12 */
    for (int i = 0; i < hx; ++i) o->gradx_coe[i] = TWOPI * ((float)i) / o->lx;      /* :15-17 */
    for (int i = hx; i < n; ++i) o->gradx_coe[i] = -o->gradx_coe[n - i];            /* :18-20 */
    for (int j = 0; j < hy; ++j) o->grady_coe[j] = TWOPI * ((float)j) / o->ly;      /* :22-24 */
    for (int i = 0; i < hx; ++i)                                                     /* :40-47 */
        for (int j = 0; j < hy; ++j) {
            double kx = o->gradx_coe[i], ky = o->grady_coe[j];
            float l = (float)(-(kx * kx + ky * ky));
            o->lapinv[HIDX(o, i, j)] = (i == 0 && j == 0) ? 1.0f : l;
            o->lap[HIDX(o, i, j)] = l;
        }
    for (int i = hx; i < n; ++i)                                                     /* :49-54 */
        for (int j = 0; j < hy; ++j) {
            o->lapinv[HIDX(o, i, j)] = o->lapinv[HIDX(o, n - i, j)];
            o->lap[HIDX(o, i, j)] = o->lap[HIDX(o, n - i, j)];
        }
    const float gws = (float)((double)dkx * dkx + (double)dky * dky);               /* :57 */
    for (int i = 0; i < hx; ++i)                                                     /* :59-63 */
        for (int j = 0; j < hy; ++j)
            o->mask[HIDX(o, i, j)] = ((double)i * i + (double)j * j >= (double)gws) ? 0.0f : 1.0f;
    for (int i = hx; i < n; ++i)                                                     /* :64-68 */
        for (int j = 0; j < hy; ++j) o->mask[HIDX(o, i, j)] = o->mask[HIDX(o, n - i, j)];
}

orc_t *orc_create(int n, float lx, float ly, float nu)
{
    orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
    o->n = n; o->hy = n / 2 + 1;
    o->grids = (size_t)n * n; o->hgrids = (size_t)n * o->hy;
    o->lx = lx; o->ly = ly; o->nu = nu;
    o->gradx_coe = (float *)fftwf_malloc(sizeof(float) * n);
    o->grady_coe = (float *)fftwf_malloc(sizeof(float) * o->hy);
    o->lap = (float *)fftwf_malloc(sizeof(float) * o->hgrids);
    o->lapinv = (float *)fftwf_malloc(sizeof(float) * o->hgrids);
    o->mask = (float *)fftwf_malloc(sizeof(float) * o->hgrids);
    build_tables(o);
    float **r[] = {&o->vort, &o->u, &o->v, &o->dvortdx, &o->dvortdy, &o->dvortdt, &o->workspace, &o->vort_src};
    for (size_t k = 0; k < sizeof(r) / sizeof(r[0]); ++k) {
        *r[k] = (float *)fftwf_malloc(sizeof(float) * o->grids);
        memset(*r[k], 0, sizeof(float) * o->grids);   /* main.cpp:110 leaves vort_src to fresh zero pages */
    }
    fftwf_complex **c[] = {&o->vort_c0, &o->vort_c, &o->lvort_c, &o->dvortdt_c, &o->tmp_c, &o->psi_c,
                           &o->rk1_c, &o->rk2_c, &o->rk3_c, &o->rk4_c};
    for (size_t k = 0; k < sizeof(c) / sizeof(c[0]); ++k) {
        *c[k] = (fftwf_complex *)fftwf_malloc(sizeof(fftwf_complex) * o->hgrids);
        memset(*c[k], 0, sizeof(fftwf_complex) * o->hgrids);
    }
    o->p_fwd = fftwf_plan_dft_r2c_2d(n, n, o->vort, o->vort_c, FFTW_ESTIMATE);       /* main.cpp:126 */
    o->p_bwd = fftwf_plan_dft_c2r_2d(n, n, o->vort_c, o->vort, FFTW_ESTIMATE);       /* main.cpp:129 */
    return o;
}

void orc_destroy(orc_t *o)
{
    if (!o) return;
    float *r[] = {o->gradx_coe, o->grady_coe, o->lap, o->lapinv, o->mask, o->vort, o->u, o->v,
                  o->dvortdx, o->dvortdy, o->dvortdt, o->workspace, o->vort_src};
    for (size_t k = 0; k < sizeof(r) / sizeof(r[0]); ++k) fftwf_free(r[k]);
    fftwf_complex *c[] = {o->vort_c0, o->vort_c, o->lvort_c, o->dvortdt_c, o->tmp_c, o->psi_c,
                          o->rk1_c, o->rk2_c, o->rk3_c, o->rk4_c};
    for (size_t k = 0; k < sizeof(c) / sizeof(c[0]); ++k) fftwf_free(c[k]);
    if (o->has_tracer) {
        float *tr[] = {o->trc, o->dtrcdx, o->dtrcdy};
        for (size_t k = 0; k < sizeof(tr) / sizeof(tr[0]); ++k) fftwf_free(tr[k]);
        fftwf_complex *tc[] = {o->trc_c0, o->trc_c, o->ltrc_c, o->dtrcdt_c, o->trk1_c, o->trk2_c, o->trk3_c, o->trk4_c};
        for (size_t k = 0; k < sizeof(tc) / sizeof(tc[0]); ++k) fftwf_free(tc[k]);
    }
    fftwf_destroy_plan(o->p_fwd); fftwf_destroy_plan(o->p_bwd);
    free(o);
}

/* table accessors: which = 0 gradx_coe[n], 1 grady_coe[n/2+1], 2 laplacian_coe[H],
 * 3 laplacian_coe_inverse[H], 4 dealiasing_mask[H] */
const float *orc_table(orc_t *o, int which)
{
    switch (which) {
    case 0: return o->gradx_coe;
    case 1: return o->grady_coe;
    case 2: return o->lap;
    case 3: return o->lapinv;
    default: return o->mask;
    }
}

/* fftwfop.cpp:87-94 */
void orc_gradx(orc_t *o, const fftwf_complex *in, fftwf_complex *out)
{
    for (int i = 0; i < o->n; ++i)
        for (int j = 0; j < o->hy; ++j) {
            out[HIDX(o, i, j)][0] = -in[HIDX(o, i, j)][1] * o->gradx_coe[i];
            out[HIDX(o, i, j)][1] = in[HIDX(o, i, j)][0] * o->gradx_coe[i];
        }
}

/* fftwfop.cpp:96-103 */
void orc_grady(orc_t *o, const fftwf_complex *in, fftwf_complex *out)
{
    for (int i = 0; i < o->n; ++i)
        for (int j = 0; j < o->hy; ++j) {
            out[HIDX(o, i, j)][0] = -in[HIDX(o, i, j)][1] * o->grady_coe[j];
            out[HIDX(o, i, j)][1] = in[HIDX(o, i, j)][0] * o->grady_coe[j];
        }
}

/* fftwfop.cpp:105-110 */
void orc_laplacian(orc_t *o, const fftwf_complex *in, fftwf_complex *out)
{
    for (size_t i = 0; i < o->hgrids; ++i) {
        out[i][0] = in[i][0] * o->lap[i];
        out[i][1] = in[i][1] * o->lap[i];
    }
}

/* fftwfop.cpp:112-117 (divides by the table whose (0,0) entry is 1) */
void orc_invert_laplacian(orc_t *o, const fftwf_complex *in, fftwf_complex *out)
{
    for (size_t i = 0; i < o->hgrids; ++i) {
        out[i][0] = in[i][0] / o->lapinv[i];
        out[i][1] = in[i][1] / o->lapinv[i];
    }
}

/* fftwfop.cpp:119-124 */
void orc_dealias(orc_t *o, const fftwf_complex *in, fftwf_complex *out)
{
    for (size_t i = 0; i < o->hgrids; ++i) {
        out[i][0] = in[i][0] * o->mask[i];
        out[i][1] = in[i][1] * o->mask[i];
    }
}

/* main.cpp:126-135 + fftwf_execute: unnormalised 2-D transforms */
void orc_r2c(orc_t *o, float *in, fftwf_complex *out) { fftwf_execute_dft_r2c(o->p_fwd, in, out); }
void orc_c2r(orc_t *o, fftwf_complex *in, float *out) { fftwf_execute_dft_c2r(o->p_bwd, in, out); }

/* main.cpp:37-41 */
static void backward_normalize(orc_t *o, float *data)
{
    const int grids = (int)o->grids;
    for (int i = 0; i < grids; ++i) data[i] /= grids;
}

/* main.cpp:146-244 (getDvortdt); psi/u/v stay in o->workspace/u/v for the record path */
static void get_dvortdt(orc_t *o, int want_psi)
{
    const int grids = (int)o->grids, hgrids = (int)o->hgrids;
    orc_laplacian(o, o->vort_c, o->lvort_c);                                     /* :148 */
    orc_gradx(o, o->vort_c, o->tmp_c);                                           /* :151 */
    orc_c2r(o, o->tmp_c, o->dvortdx); backward_normalize(o, o->dvortdx);         /* :154 */
    orc_grady(o, o->vort_c, o->tmp_c);                                           /* :165 */
    orc_c2r(o, o->tmp_c, o->dvortdy); backward_normalize(o, o->dvortdy);         /* :168 */
    orc_invert_laplacian(o, o->vort_c, o->psi_c);                                /* :179 */
    if (want_psi) { orc_c2r(o, o->psi_c, o->workspace); backward_normalize(o, o->workspace); } /* :183-192 */
    orc_grady(o, o->psi_c, o->tmp_c);                                            /* :198 */
    orc_c2r(o, o->tmp_c, o->u); backward_normalize(o, o->u);                     /* :200 */
    for (int i = 0; i < grids; ++i) o->u[i] = -o->u[i];                          /* :201 */
    orc_gradx(o, o->psi_c, o->tmp_c);                                            /* :212 */
    orc_c2r(o, o->tmp_c, o->v); backward_normalize(o, o->v);                     /* :214 */
    for (int i = 0; i < grids; ++i)                                              /* :225-227 */
        o->dvortdt[i] = -o->u[i] * o->dvortdx[i] - o->v[i] * o->dvortdy[i] + o->vort_src[i];
    orc_r2c(o, o->dvortdt, o->dvortdt_c);                                        /* :237 */
    for (int i = 0; i < hgrids; ++i) {                                           /* :240-243 */
        o->dvortdt_c[i][0] += o->lvort_c[i][0] * o->nu;
        o->dvortdt_c[i][1] += o->lvort_c[i][1] * o->nu;
    }
}

/* Passive tracer c advected by the flow of the SAME Runge-Kutta stage (for effective-
 * diffusivity diagnostics on a tracer other than the vorticity, Hendricks & Schubert 2009):
 *     dc/dt = -u c_x - v c_y + kappa lap(c)
 * The reference has no tracer: PARITY UNPINNED.  The restatement mirrors the vorticity tendency operation for
 * operation (main.cpp:151-168 for the gradients, :225-227 for the product, :237-243 for the transform and the
 * diffusion term) so that c == vort with kappa == nu and no source reproduces the vorticity exactly.
 * Must run right after get_dvortdt: o->u, o->v hold the stage's velocity. */
static void get_dtrcdt(orc_t *o)
{
    const int grids = (int)o->grids, hgrids = (int)o->hgrids;
    orc_laplacian(o, o->trc_c, o->ltrc_c);
    orc_gradx(o, o->trc_c, o->tmp_c);
    orc_c2r(o, o->tmp_c, o->dtrcdx); backward_normalize(o, o->dtrcdx);
    orc_grady(o, o->trc_c, o->tmp_c);
    orc_c2r(o, o->tmp_c, o->dtrcdy); backward_normalize(o, o->dtrcdy);
    for (int i = 0; i < grids; ++i) o->dvortdt[i] = -o->u[i] * o->dtrcdx[i] - o->v[i] * o->dtrcdy[i];
    orc_r2c(o, o->dvortdt, o->dtrcdt_c);
    for (int i = 0; i < hgrids; ++i) {
        o->dtrcdt_c[i][0] += o->ltrc_c[i][0] * o->kappa;
        o->dtrcdt_c[i][1] += o->ltrc_c[i][1] * o->kappa;
    }
}

static void evolve_tracer(orc_t *o, fftwf_complex *rk, float dt)
{
    const int hgrids = (int)o->hgrids;
    for (int i = 0; i < hgrids; ++i) {
        o->trc_c[i][0] = o->trc_c0[i][0] + rk[i][0] * dt;
        o->trc_c[i][1] = o->trc_c0[i][1] + rk[i][1] * dt;
    }
}

/* physical tracer field + diffusivity; allocates the tracer state on first use */
void orc_set_tracer(orc_t *o, const float *c, float kappa)
{
    if (!o->has_tracer) {
        float **r[] = {&o->trc, &o->dtrcdx, &o->dtrcdy};
        for (size_t k = 0; k < sizeof(r) / sizeof(r[0]); ++k) *r[k] = (float *)fftwf_malloc(sizeof(float) * o->grids);
        fftwf_complex **cc[] = {&o->trc_c0, &o->trc_c, &o->ltrc_c, &o->dtrcdt_c, &o->trk1_c, &o->trk2_c, &o->trk3_c, &o->trk4_c};
        for (size_t k = 0; k < sizeof(cc) / sizeof(cc[0]); ++k) {
            *cc[k] = (fftwf_complex *)fftwf_malloc(sizeof(fftwf_complex) * o->hgrids);
            memset(*cc[k], 0, sizeof(fftwf_complex) * o->hgrids);
        }
        o->has_tracer = 1;
    }
    o->kappa = kappa;
    memcpy(o->trc, c, sizeof(float) * o->grids);
    orc_r2c(o, o->trc, o->trc_c);
}

/* current tracer in physical space */
void orc_get_tracer(orc_t *o, float *out)
{
    memcpy(o->tmp_c, o->trc_c, sizeof(fftwf_complex) * o->hgrids);      /* c2r destroys its input */
    orc_c2r(o, o->tmp_c, o->trc); backward_normalize(o, o->trc);
    memcpy(out, o->trc, sizeof(float) * o->grids);
}

/* main.cpp:246-251 */
static void evolve(orc_t *o, fftwf_complex *rk, float dt)
{
    const int hgrids = (int)o->hgrids;
    for (int i = 0; i < hgrids; ++i) {
        o->vort_c[i][0] = o->vort_c0[i][0] + rk[i][0] * dt;
        o->vort_c[i][1] = o->vort_c0[i][1] + rk[i][1] * dt;
    }
}

/* main.cpp:143-144,256: physical initial field -> spectral state */
void orc_set_vorticity(orc_t *o, const float *vort)
{
    memcpy(o->vort, vort, sizeof(float) * o->grids);
    orc_r2c(o, o->vort, o->vort_c);
}

void orc_set_spectrum(orc_t *o, const fftwf_complex *z) { memcpy(o->vort_c, z, sizeof(fftwf_complex) * o->hgrids); }
void orc_get_spectrum(orc_t *o, fftwf_complex *z) { memcpy(z, o->vort_c, sizeof(fftwf_complex) * o->hgrids); }

/* main-shallow-water.cpp:304 / vorticity_source.cpp:112-133: piecewise-constant forcing */
void orc_set_source(orc_t *o, const float *src)
{
    if (src) memcpy(o->vort_src, src, sizeof(float) * o->grids);
    else memset(o->vort_src, 0, sizeof(float) * o->grids);
}

/* main.cpp:286-317: nsteps RK4 steps, no record I/O */
void orc_step(orc_t *o, int nsteps, float dt)
{
    const int hgrids = (int)o->hgrids;
    for (int s = 0; s < nsteps; ++s) {
        memcpy(o->vort_c0, o->vort_c, sizeof(fftwf_complex) * o->hgrids);          /* :286 */
        if (o->has_tracer) memcpy(o->trc_c0, o->trc_c, sizeof(fftwf_complex) * o->hgrids);
        for (int k = 0; k < 4; ++k) {
            get_dvortdt(o, 0);                                                     /* :290 */
            if (o->has_tracer) {
                /* same stage, same velocity; same dealias / evolve / final-combine expressions as below */
                get_dtrcdt(o);
                switch (k) {
                case 0: orc_dealias(o, o->dtrcdt_c, o->trk1_c); evolve_tracer(o, o->trk1_c, dt / 2.0f); break;
                case 1: orc_dealias(o, o->dtrcdt_c, o->trk2_c); evolve_tracer(o, o->trk2_c, dt / 2.0f); break;
                case 2: orc_dealias(o, o->dtrcdt_c, o->trk3_c); evolve_tracer(o, o->trk3_c, dt); break;
                case 3:
                    orc_dealias(o, o->dtrcdt_c, o->trk4_c);
                    for (int i = 0; i < hgrids; ++i) {
                        o->trc_c[i][0] = o->trc_c0[i][0] + (o->trk1_c[i][0] + 2.0f * o->trk2_c[i][0] + 2.0f * o->trk3_c[i][0] + o->trk4_c[i][0]) * dt / 6.0f;
                        o->trc_c[i][1] = o->trc_c0[i][1] + (o->trk1_c[i][1] + 2.0f * o->trk2_c[i][1] + 2.0f * o->trk3_c[i][1] + o->trk4_c[i][1]) * dt / 6.0f;
                    }
                    break;
                }
            }
            switch (k) {
            case 0: orc_dealias(o, o->dvortdt_c, o->rk1_c); evolve(o, o->rk1_c, dt / 2.0f); break; /* :296 */
            case 1: orc_dealias(o, o->dvortdt_c, o->rk2_c); evolve(o, o->rk2_c, dt / 2.0f); break; /* :299 */
            case 2: orc_dealias(o, o->dvortdt_c, o->rk3_c); evolve(o, o->rk3_c, dt); break;        /* :302 */
            case 3:
                orc_dealias(o, o->dvortdt_c, o->rk4_c);                            /* :306 */
                for (int i = 0; i < hgrids; ++i) {                                 /* :309-312 */
                    o->vort_c[i][0] = o->vort_c0[i][0] + (o->rk1_c[i][0] + 2.0f * o->rk2_c[i][0] + 2.0f * o->rk3_c[i][0] + o->rk4_c[i][0]) * dt / 6.0f;
                    o->vort_c[i][1] = o->vort_c0[i][1] + (o->rk1_c[i][1] + 2.0f * o->rk2_c[i][1] + 2.0f * o->rk3_c[i][1] + o->rk4_c[i][1]) * dt / 6.0f;
                }
                break;
            }
        }
    }
}

/* Record-step fields of the CURRENT state, as main.cpp writes them at a record step:
 * which = 0 vort (:273-281), 1 psi (:183-192), 2 u (:198-209), 3 v (:212-222), 4 source (:268-270) */
void orc_get_field(orc_t *o, int which, float *out)
{
    if (which == 0) {
        orc_c2r(o, o->vort_c, o->vort); backward_normalize(o, o->vort);
        memcpy(out, o->vort, sizeof(float) * o->grids);
        return;
    }
    if (which == 4) { memcpy(out, o->vort_src, sizeof(float) * o->grids); return; }
    const int grids = (int)o->grids;
    orc_invert_laplacian(o, o->vort_c, o->psi_c);
    if (which == 1) {
        orc_c2r(o, o->psi_c, o->workspace); backward_normalize(o, o->workspace);
        memcpy(out, o->workspace, sizeof(float) * o->grids);
    } else if (which == 2) {
        orc_grady(o, o->psi_c, o->tmp_c);
        orc_c2r(o, o->tmp_c, o->u); backward_normalize(o, o->u);
        for (int i = 0; i < grids; ++i) o->u[i] = -o->u[i];
        memcpy(out, o->u, sizeof(float) * o->grids);
    } else {
        orc_gradx(o, o->psi_c, o->tmp_c);
        orc_c2r(o, o->tmp_c, o->v); backward_normalize(o, o->v);
        memcpy(out, o->v, sizeof(float) * o->grids);
    }
}

/* invert_pres.cpp:135-187: psi -> pressure anomaly.  rho, f: configuration.hpp:10-11 */
void orc_invert_pres(orc_t *o, const float *psi_in, float *pres, size_t ref_x, size_t ref_y, float rho, float f)
{
    const int grids = (int)o->grids, hgrids = (int)o->hgrids;
    float *psi = o->workspace, *dx2 = o->dvortdx, *dy2 = o->dvortdy, *dxdy = o->u, *gaus = o->dvortdt;
    fftwf_complex *psi_c = o->psi_c, *tmp_c = o->tmp_c, *dx2_c = o->rk1_c, *dy2_c = o->rk2_c, *dxdy_c = o->rk3_c,
                  *lap_pres_c = o->rk4_c;
    memcpy(psi, psi_in, sizeof(float) * o->grids);
    orc_r2c(o, psi, psi_c);                                                       /* :135 */
    orc_gradx(o, psi_c, tmp_c); orc_gradx(o, tmp_c, dx2_c);                       /* :139-140 */
    orc_grady(o, psi_c, tmp_c); orc_grady(o, tmp_c, dy2_c);                       /* :142-143 */
    orc_gradx(o, tmp_c, dxdy_c);                                                  /* :145 */
    orc_dealias(o, dx2_c, dx2_c); orc_dealias(o, dy2_c, dy2_c); orc_dealias(o, dxdy_c, dxdy_c); /* :148-150 */
    orc_c2r(o, dx2_c, dx2); backward_normalize(o, dx2);                           /* :153 */
    orc_c2r(o, dy2_c, dy2); backward_normalize(o, dy2);                           /* :154 */
    orc_c2r(o, dxdy_c, dxdy); backward_normalize(o, dxdy);                        /* :155 */
    /* :159  `pow(float, 2.0f)` there is the C library's ::pow(double,double) (no `using namespace std` in
     * invert_pres.cpp): the square and the subtraction are done in double, the float product is promoted */
    for (int i = 0; i < grids; ++i) gaus[i] = (float)((double)(dx2[i] * dy2[i]) - pow((double)dxdy[i], 2.0));
    orc_r2c(o, gaus, lap_pres_c);                                                 /* :161 */
    orc_laplacian(o, psi_c, tmp_c);                                               /* :164 */
    for (int i = 0; i < hgrids; ++i) {                                            /* :166-169, double intermediates */
        lap_pres_c[i][0] = rho * (f * tmp_c[i][0] + 2.0 * lap_pres_c[i][0]);
        lap_pres_c[i][1] = rho * (f * tmp_c[i][1] + 2.0 * lap_pres_c[i][1]);
    }
    orc_invert_laplacian(o, lap_pres_c, tmp_c);                                   /* :171 */
    orc_c2r(o, tmp_c, pres); backward_normalize(o, pres);                         /* :172 */
    float ref_val = pres[ref_x + (size_t)o->n * ref_y];                           /* :182 */
    for (int i = 0; i < grids; ++i) pres[i] -= ref_val;                           /* :183-185 */
}

/* ---------------------------------------------------------------- diagnostics (parity unpinned)
 * README.md:5-7 names them; no reference code exists.  Definitions:
 *   S1 = u_x - v_y = -2 psi_xy,  S2 = v_x + u_y = psi_xx - psi_yy,  zeta = psi_xx + psi_yy
 *   Q  = S1^2 + S2^2 - zeta^2
 *   filamentation time (Rozoff et al. 2006, eq. 5):  tau = 2 / sqrt(Q) where Q > 0, else 0 (sentinel)
 *   deformation factor (builder-defined):            D = Q / (S1^2 + S2^2 + zeta^2), 0 where the denominator is 0
 * All derivatives are spectral, normalised like every other inverse transform (/GRIDS).
 * out: tfil[G], deform[G]; optional s1[G], s2[G] (may be NULL).
 */
void orc_diagnostics(orc_t *o, float *tfil, float *deform, float *s1_out, float *s2_out)
{
    const int grids = (int)o->grids;
    float *pxy = o->dvortdx, *pxx = o->dvortdy, *pyy = o->dvortdt;
    orc_invert_laplacian(o, o->vort_c, o->psi_c);
    orc_gradx(o, o->psi_c, o->tmp_c); orc_grady(o, o->tmp_c, o->rk1_c);
    orc_c2r(o, o->rk1_c, pxy); backward_normalize(o, pxy);
    orc_gradx(o, o->psi_c, o->tmp_c); orc_gradx(o, o->tmp_c, o->rk1_c);
    orc_c2r(o, o->rk1_c, pxx); backward_normalize(o, pxx);
    orc_grady(o, o->psi_c, o->tmp_c); orc_grady(o, o->tmp_c, o->rk1_c);
    orc_c2r(o, o->rk1_c, pyy); backward_normalize(o, pyy);
    for (int i = 0; i < grids; ++i) {
        float s1 = -2.0f * pxy[i], s2 = pxx[i] - pyy[i], z = pxx[i] + pyy[i];
        float ss = s1 * s1 + s2 * s2, zz = z * z;
        float q = ss - zz, den = ss + zz;
        tfil[i] = (q > 0.0f) ? 2.0f / sqrtf(q) : 0.0f;
        deform[i] = (den > 0.0f) ? q / den : 0.0f;
        if (s1_out) s1_out[i] = s1;
        if (s2_out) s2_out[i] = s2;
    }
}

/* Effective-diffusivity histograms (Hendricks & Schubert 2009 after Nakamura 1996), tracer c = zeta:
 *   bin b = clamp(floor((c - cmin) / (cmax - cmin) * nbins), 0, nbins-1)
 *   area[b] += dx dy ;  grad2[b] += |grad c|^2 dx dy       (float64 accumulators)
 * |grad c|^2 from the spectral derivatives the tendency already forms (main.cpp:151-168).
 */
static void keff_hist_of(orc_t *o, fftwf_complex *c_c, int nbins, float cmin, float cmax, double *area, double *grad2)
{
    const int grids = (int)o->grids;
    const double da = ((double)o->lx / o->n) * ((double)o->ly / o->n);
    memcpy(o->psi_c, c_c, sizeof(fftwf_complex) * o->hgrids);           /* c2r destroys its input */
    orc_c2r(o, o->psi_c, o->vort); backward_normalize(o, o->vort);
    orc_gradx(o, c_c, o->tmp_c);
    orc_c2r(o, o->tmp_c, o->dvortdx); backward_normalize(o, o->dvortdx);
    orc_grady(o, c_c, o->tmp_c);
    orc_c2r(o, o->tmp_c, o->dvortdy); backward_normalize(o, o->dvortdy);
    for (int b = 0; b < nbins; ++b) { area[b] = 0; grad2[b] = 0; }
    const float scale = (float)nbins / (cmax - cmin);
    for (int i = 0; i < grids; ++i) {
        int b = (int)floorf((o->vort[i] - cmin) * scale);
        if (b < 0) b = 0;
        if (b >= nbins) b = nbins - 1;
        float g2 = o->dvortdx[i] * o->dvortdx[i] + o->dvortdy[i] * o->dvortdy[i];
        area[b] += da;
        grad2[b] += (double)g2 * da;
    }
}

void orc_keff_hist(orc_t *o, int nbins, float cmin, float cmax, double *area, double *grad2)
{
    keff_hist_of(o, o->vort_c, nbins, cmin, cmax, area, grad2);
}

/* the same histograms over the passive tracer */
void orc_tracer_keff_hist(orc_t *o, int nbins, float cmin, float cmax, double *area, double *grad2)
{
    keff_hist_of(o, o->trc_c, nbins, cmin, cmax, area, grad2);
}

/* kappa_eff(C_b) on the bin edges C_b = cmin + b (cmax-cmin)/nbins, b = 1..nbins-1:
 *   A(C) = area{c >= C}, G(C) = int_{c>=C} |grad c|^2 dA (cumulative sums from the top bin)
 *   Le^2 = (dG/dA) / (dC/dA)^2 by centred differences over neighbouring edges
 *   kappa_eff = kappa * Le^2 / (2 pi r_e)^2,  r_e = sqrt(A/pi)  ->  (2 pi r_e)^2 = 4 pi A
 * Edges where A does not change get kappa_eff = 0.  a_out/k_out have nbins+1 entries (edge 0 and nbins = 0). */
void orc_keff_from_hist(int nbins, float cmin, float cmax, double kappa, const double *area, const double *grad2,
                        double *a_out, double *k_out)
{
    double *A = (double *)calloc((size_t)nbins + 2, sizeof(double));
    double *G = (double *)calloc((size_t)nbins + 2, sizeof(double));
    for (int b = nbins - 1; b >= 0; --b) { A[b] = A[b + 1] + area[b]; G[b] = G[b + 1] + grad2[b]; }
    const double dc = ((double)cmax - (double)cmin) / nbins;
    for (int b = 0; b <= nbins; ++b) { a_out[b] = A[b]; k_out[b] = 0.0; }
    for (int b = 1; b < nbins; ++b) {
        double dA = A[b + 1] - A[b - 1], dG = G[b + 1] - G[b - 1];
        if (dA == 0.0 || A[b] <= 0.0) continue;
        double dCdA = 2.0 * dc / dA, dGdA = dG / dA;
        double le2 = dGdA / (dCdA * dCdA);
        k_out[b] = kappa * le2 / (4.0 * M_PI * A[b]);
    }
    free(A); free(G);
}
