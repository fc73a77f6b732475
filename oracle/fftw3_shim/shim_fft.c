/* shim_fft.c -- CPU implementation of the seven FFTW3f names the reference uses.
 *
 * TEST INFRASTRUCTURE (oracle/): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load or link this.  The product library
 * (xlab_fftbarotropic_b200/csrc) never does.
 *
 * What it restates: FFTW 3.3.x single precision, the un-vendored dependency named at
 * /root/reference/README.md:15-18 and Makefile:10 ("-lfftw3f", no pinned version).
 * FFTW's source is absent from /root/reference, so this file restates the PUBLISHED
 * definition of the two transforms the reference plans
 * (/root/reference/src/main.cpp:126-135, src/invert_pres.cpp:100-107):
 *   fftwf_plan_dft_r2c_2d / fftwf_plan_dft_c2r_2d, FFTW_ESTIMATE, executed by fftwf_execute.
 * Algorithm: Stockham autosort mixed radix (4,2,3,5, generic odd prime), float32 data,
 * twiddles generated in float64 and rounded once.  A real transform of even length n is a
 * complex transform of length n/2 plus the standard split/merge step.  The 2-D r2c does the
 * real transform along the last (contiguous) dimension, then complex transforms along the
 * first; c2r does them in the opposite order, which is the order FFTW's rdft2 solver uses
 * and the reason imaginary parts at k1 = 0 and k1 = n1/2 are dropped AFTER the first-dimension
 * pass (SURVEY.md section 8a "quirks").
 * Lines are processed VL at a time in a transposed [n][VL] layout so that gcc vectorises the
 * butterflies across lines; OpenMP spreads blocks of lines over threads
 * (XFB_SHIM_THREADS, default 1 = the reference never enables FFTW threads).
 */
#include "fftw3.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VL 16
#define MAXST 40

typedef struct {
    int n;
    int nst;
    int radix[MAXST];
    int ns[MAXST];
    float *twr[MAXST]; /* [k][r-1], k in [0,Ns), r in [1,R): exp(-2 pi i r k /(Ns R)) */
    float *twi[MAXST];
    double *gr, *gi;   /* exp(-2 pi i q / R) for the generic radix, R = largest prime factor */
    int gR;
} plan1d;

struct xfb_shim_plan_s {
    int kind; /* 0 = r2c, 1 = c2r */
    int n0, n1;
    float *rbuf;
    fftwf_complex *cbuf;
    plan1d *p0;   /* length n0, complex */
    plan1d *ph;   /* length n1/2 complex (n1 even) or n1 (n1 odd) */
    float *hwr, *hwi; /* exp(-2 pi i k / n1), k in [0, n1/2] */
    fftwf_complex *scratch; /* c2r works on a copy: the caller's spectrum is left intact */
};

static int g_threads = -1;

void xfb_shim_set_threads(int n) { g_threads = n; }

static int shim_threads(void)
{
    if (g_threads < 0) {
        const char *e = getenv("XFB_SHIM_THREADS");
        g_threads = e ? atoi(e) : 1;
    }
#ifdef _OPENMP
    if (g_threads == 0) return omp_get_max_threads();
#endif
    return g_threads > 0 ? g_threads : 1;
}

void *fftwf_malloc(size_t n)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, n ? n : 64) != 0) return NULL;
    return p;
}

void fftwf_free(void *p) { free(p); }

/* ------------------------------------------------------------------ 1-D plan */

static plan1d *plan1d_new(int n)
{
    plan1d *p = (plan1d *)calloc(1, sizeof(plan1d));
    p->n = n;
    int m = n, ns = 1;
    while (m > 1) {
        int r;
        if (m % 4 == 0) r = 4;
        else if (m % 2 == 0) r = 2;
        else if (m % 3 == 0) r = 3;
        else if (m % 5 == 0) r = 5;
        else { r = 7; while (m % r) r += 2; }
        int s = p->nst++;
        p->radix[s] = r;
        p->ns[s] = ns;
        p->twr[s] = (float *)malloc(sizeof(float) * (size_t)ns * (r - 1));
        p->twi[s] = (float *)malloc(sizeof(float) * (size_t)ns * (r - 1));
        for (int k = 0; k < ns; ++k)
            for (int q = 1; q < r; ++q) {
                double a = -2.0 * M_PI * (double)q * (double)k / ((double)ns * r);
                p->twr[s][k * (r - 1) + q - 1] = (float)cos(a);
                p->twi[s][k * (r - 1) + q - 1] = (float)sin(a);
            }
        if (r > 5 && r > p->gR) p->gR = r;
        ns *= r;
        m /= r;
    }
    if (p->gR) {
        /* one table per distinct generic radix would be needed in general; sizes with two
         * different prime factors > 5 are refused by the planner below */
        p->gr = (double *)malloc(sizeof(double) * p->gR);
        p->gi = (double *)malloc(sizeof(double) * p->gR);
        for (int q = 0; q < p->gR; ++q) {
            p->gr[q] = cos(-2.0 * M_PI * q / p->gR);
            p->gi[q] = sin(-2.0 * M_PI * q / p->gR);
        }
        for (int s = 0; s < p->nst; ++s)
            if (p->radix[s] > 5 && p->radix[s] != p->gR) {
                fprintf(stderr, "fftw3 shim: unsupported size %d\n", n);
                abort();
            }
    }
    return p;
}

static void plan1d_free(plan1d *p)
{
    if (!p) return;
    for (int s = 0; s < p->nst; ++s) { free(p->twr[s]); free(p->twi[s]); }
    free(p->gr); free(p->gi);
    free(p);
}

/* One Stockham pass over VL interleaved lines.  sg = -1 forward, +1 backward: the tables
 * hold the forward twiddles, the backward pass conjugates them. */
static void pass(const plan1d *p, int s, float sg, const float *restrict ir, const float *restrict ii,
                 float *restrict outr, float *restrict outi)
{
    const int n = p->n, R = p->radix[s], Ns = p->ns[s], m = n / R;
    const float *twr = p->twr[s], *twi = p->twi[s];
    for (int j = 0; j < m; ++j) {
        const int k = j % Ns;
        const int base = (j / Ns) * Ns * R + k;
        const float *tr = twr + (size_t)k * (R - 1), *ti = twi + (size_t)k * (R - 1);
        if (R == 4) {
            const float *a0r = ir + (size_t)j * VL, *a0i = ii + (size_t)j * VL;
            const float *a1r = a0r + (size_t)m * VL, *a1i = a0i + (size_t)m * VL;
            const float *a2r = a1r + (size_t)m * VL, *a2i = a1i + (size_t)m * VL;
            const float *a3r = a2r + (size_t)m * VL, *a3i = a2i + (size_t)m * VL;
            float *o0r = outr + (size_t)base * VL, *o0i = outi + (size_t)base * VL;
            float *o1r = o0r + (size_t)Ns * VL, *o1i = o0i + (size_t)Ns * VL;
            float *o2r = o1r + (size_t)Ns * VL, *o2i = o1i + (size_t)Ns * VL;
            float *o3r = o2r + (size_t)Ns * VL, *o3i = o2i + (size_t)Ns * VL;
            const float w1r = tr[0], w1i = sg < 0 ? ti[0] : -ti[0];
            const float w2r = tr[1], w2i = sg < 0 ? ti[1] : -ti[1];
            const float w3r = tr[2], w3i = sg < 0 ? ti[2] : -ti[2];
            for (int v = 0; v < VL; ++v) {
                float x0r = a0r[v], x0i = a0i[v];
                float x1r = a1r[v] * w1r - a1i[v] * w1i, x1i = a1r[v] * w1i + a1i[v] * w1r;
                float x2r = a2r[v] * w2r - a2i[v] * w2i, x2i = a2r[v] * w2i + a2i[v] * w2r;
                float x3r = a3r[v] * w3r - a3i[v] * w3i, x3i = a3r[v] * w3i + a3i[v] * w3r;
                float t0r = x0r + x2r, t0i = x0i + x2i, t1r = x0r - x2r, t1i = x0i - x2i;
                float t2r = x1r + x3r, t2i = x1i + x3i;
                /* t3 = sg * i * (x1 - x3) */
                float dr = x1r - x3r, di = x1i - x3i;
                float t3r = -sg * di, t3i = sg * dr;
                o0r[v] = t0r + t2r; o0i[v] = t0i + t2i;
                o2r[v] = t0r - t2r; o2i[v] = t0i - t2i;
                o1r[v] = t1r + t3r; o1i[v] = t1i + t3i;
                o3r[v] = t1r - t3r; o3i[v] = t1i - t3i;
            }
        } else if (R == 2) {
            const float *a0r = ir + (size_t)j * VL, *a0i = ii + (size_t)j * VL;
            const float *a1r = a0r + (size_t)m * VL, *a1i = a0i + (size_t)m * VL;
            float *o0r = outr + (size_t)base * VL, *o0i = outi + (size_t)base * VL;
            float *o1r = o0r + (size_t)Ns * VL, *o1i = o0i + (size_t)Ns * VL;
            const float w1r = tr[0], w1i = sg < 0 ? ti[0] : -ti[0];
            for (int v = 0; v < VL; ++v) {
                float x1r = a1r[v] * w1r - a1i[v] * w1i, x1i = a1r[v] * w1i + a1i[v] * w1r;
                float x0r = a0r[v], x0i = a0i[v];
                o0r[v] = x0r + x1r; o0i[v] = x0i + x1i;
                o1r[v] = x0r - x1r; o1i[v] = x0i - x1i;
            }
        } else if (R == 3) {
            const float *a0r = ir + (size_t)j * VL, *a0i = ii + (size_t)j * VL;
            const float *a1r = a0r + (size_t)m * VL, *a1i = a0i + (size_t)m * VL;
            const float *a2r = a1r + (size_t)m * VL, *a2i = a1i + (size_t)m * VL;
            float *o0r = outr + (size_t)base * VL, *o0i = outi + (size_t)base * VL;
            float *o1r = o0r + (size_t)Ns * VL, *o1i = o0i + (size_t)Ns * VL;
            float *o2r = o1r + (size_t)Ns * VL, *o2i = o1i + (size_t)Ns * VL;
            const float w1r = tr[0], w1i = sg < 0 ? ti[0] : -ti[0];
            const float w2r = tr[1], w2i = sg < 0 ? ti[1] : -ti[1];
            const float c3 = -0.5f, s3 = sg * 0.86602540378443864676f;
            for (int v = 0; v < VL; ++v) {
                float x0r = a0r[v], x0i = a0i[v];
                float x1r = a1r[v] * w1r - a1i[v] * w1i, x1i = a1r[v] * w1i + a1i[v] * w1r;
                float x2r = a2r[v] * w2r - a2i[v] * w2i, x2i = a2r[v] * w2i + a2i[v] * w2r;
                float sr = x1r + x2r, si = x1i + x2i, dr = x1r - x2r, di = x1i - x2i;
                float mr = x0r + c3 * sr, mi = x0i + c3 * si;
                /* + i*s3*d */
                float er = -s3 * di, ei = s3 * dr;
                o0r[v] = x0r + sr; o0i[v] = x0i + si;
                o1r[v] = mr + er;  o1i[v] = mi + ei;
                o2r[v] = mr - er;  o2i[v] = mi - ei;
            }
        } else if (R == 5) {
            const float *ar[5], *ai[5];
            float *orr[5], *oii[5];
            for (int q = 0; q < 5; ++q) {
                ar[q] = ir + ((size_t)j + (size_t)q * m) * VL;
                ai[q] = ii + ((size_t)j + (size_t)q * m) * VL;
                orr[q] = outr + ((size_t)base + (size_t)q * Ns) * VL;
                oii[q] = outi + ((size_t)base + (size_t)q * Ns) * VL;
            }
            const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
            const float s1 = sg * 0.95105651629515357212f, s2 = sg * 0.58778525229247312917f;
            for (int v = 0; v < VL; ++v) {
                float xr[5], xi[5];
                xr[0] = ar[0][v]; xi[0] = ai[0][v];
                for (int q = 1; q < 5; ++q) {
                    float wr = tr[q - 1], wi = sg < 0 ? ti[q - 1] : -ti[q - 1];
                    xr[q] = ar[q][v] * wr - ai[q][v] * wi;
                    xi[q] = ar[q][v] * wi + ai[q][v] * wr;
                }
                float s14r = xr[1] + xr[4], s14i = xi[1] + xi[4], d14r = xr[1] - xr[4], d14i = xi[1] - xi[4];
                float s23r = xr[2] + xr[3], s23i = xi[2] + xi[3], d23r = xr[2] - xr[3], d23i = xi[2] - xi[3];
                float m1r = xr[0] + c1 * s14r + c2 * s23r, m1i = xi[0] + c1 * s14i + c2 * s23i;
                float m2r = xr[0] + c2 * s14r + c1 * s23r, m2i = xi[0] + c2 * s14i + c1 * s23i;
                /* i*(s1 d14 + s2 d23), i*(s2 d14 - s1 d23) */
                float e1r = -(s1 * d14i + s2 * d23i), e1i = (s1 * d14r + s2 * d23r);
                float e2r = -(s2 * d14i - s1 * d23i), e2i = (s2 * d14r - s1 * d23r);
                orr[0][v] = xr[0] + s14r + s23r; oii[0][v] = xi[0] + s14i + s23i;
                orr[1][v] = m1r + e1r; oii[1][v] = m1i + e1i;
                orr[4][v] = m1r - e1r; oii[4][v] = m1i - e1i;
                orr[2][v] = m2r + e2r; oii[2][v] = m2i + e2i;
                orr[3][v] = m2r - e2r; oii[3][v] = m2i - e2i;
            }
        } else {
            /* generic odd prime radix: direct O(R^2) DFT in float64 accumulators */
            for (int v = 0; v < VL; ++v) {
                double xr[64], xi[64];
                xr[0] = ir[(size_t)j * VL + v]; xi[0] = ii[(size_t)j * VL + v];
                for (int q = 1; q < R; ++q) {
                    float wr = tr[q - 1], wi = sg < 0 ? ti[q - 1] : -ti[q - 1];
                    float a = ir[((size_t)j + (size_t)q * m) * VL + v], b = ii[((size_t)j + (size_t)q * m) * VL + v];
                    xr[q] = a * wr - b * wi; xi[q] = a * wi + b * wr;
                }
                for (int o = 0; o < R; ++o) {
                    double accr = 0, acci = 0;
                    for (int q = 0; q < R; ++q) {
                        int e = (int)(((long)o * q) % R);
                        double wr = p->gr[e], wi = sg < 0 ? p->gi[e] : -p->gi[e];
                        accr += xr[q] * wr - xi[q] * wi; acci += xr[q] * wi + xi[q] * wr;
                    }
                    outr[((size_t)base + (size_t)o * Ns) * VL + v] = (float)accr;
                    outi[((size_t)base + (size_t)o * Ns) * VL + v] = (float)acci;
                }
            }
        }
    }
}

/* VL complex lines of length p->n, layout [n][VL] split re/im.  Result lands in (ar,ai);
 * (br,bi) is scratch of the same size. */
static void fft_lines(const plan1d *p, float sg, float *ar, float *ai, float *br, float *bi)
{
    float *sr = ar, *si = ai, *dr = br, *di = bi;
    for (int s = 0; s < p->nst; ++s) {
        pass(p, s, sg, sr, si, dr, di);
        float *t;
        t = sr; sr = dr; dr = t;
        t = si; si = di; di = t;
    }
    if (sr != ar) {
        memcpy(ar, sr, sizeof(float) * (size_t)p->n * VL);
        memcpy(ai, si, sizeof(float) * (size_t)p->n * VL);
    }
}

/* ------------------------------------------------------------------ 2-D plans */

static struct xfb_shim_plan_s *plan_new(int kind, int n0, int n1, float *r, fftwf_complex *c)
{
    struct xfb_shim_plan_s *p = (struct xfb_shim_plan_s *)calloc(1, sizeof(*p));
    p->kind = kind; p->n0 = n0; p->n1 = n1; p->rbuf = r; p->cbuf = c;
    p->p0 = plan1d_new(n0);
    p->ph = plan1d_new((n1 % 2 == 0) ? n1 / 2 : n1);
    int h = n1 / 2;
    p->hwr = (float *)malloc(sizeof(float) * (h + 1));
    p->hwi = (float *)malloc(sizeof(float) * (h + 1));
    for (int k = 0; k <= h; ++k) {
        double a = -2.0 * M_PI * (double)k / (double)n1;
        p->hwr[k] = (float)cos(a); p->hwi[k] = (float)sin(a);
    }
    if (kind == 1) p->scratch = (fftwf_complex *)fftwf_malloc(sizeof(fftwf_complex) * (size_t)n0 * (h + 1));
    return p;
}

fftwf_plan fftwf_plan_dft_r2c_2d(int n0, int n1, float *in, fftwf_complex *out, unsigned flags)
{
    (void)flags;
    return plan_new(0, n0, n1, in, out);
}

fftwf_plan fftwf_plan_dft_c2r_2d(int n0, int n1, fftwf_complex *in, float *out, unsigned flags)
{
    (void)flags;
    return plan_new(1, n0, n1, out, in);
}

void fftwf_destroy_plan(fftwf_plan p)
{
    if (!p) return;
    plan1d_free(p->p0); plan1d_free(p->ph);
    free(p->hwr); free(p->hwi);
    fftwf_free(p->scratch);
    free(p);
}

/* complex transforms along dimension 0 for all columns of a [n0][nc] complex array, in place */
static void columns(const struct xfb_shim_plan_s *p, float sg, fftwf_complex *a, int nc)
{
    const int n0 = p->n0;
    const int nblk = (nc + VL - 1) / VL;
    const int nt = shim_threads();
#pragma omp parallel num_threads(nt)
    {
        float *w = (float *)fftwf_malloc(sizeof(float) * 4 * (size_t)n0 * VL);
        float *ar = w, *ai = w + (size_t)n0 * VL, *br = ai + (size_t)n0 * VL, *bi = br + (size_t)n0 * VL;
#pragma omp for schedule(static)
        for (int b = 0; b < nblk; ++b) {
            const int j0 = b * VL, nv = (nc - j0 < VL) ? nc - j0 : VL;
            for (int i = 0; i < n0; ++i) {
                const fftwf_complex *row = a + (size_t)i * nc + j0;
                for (int v = 0; v < nv; ++v) { ar[(size_t)i * VL + v] = row[v][0]; ai[(size_t)i * VL + v] = row[v][1]; }
                for (int v = nv; v < VL; ++v) { ar[(size_t)i * VL + v] = 0.f; ai[(size_t)i * VL + v] = 0.f; }
            }
            fft_lines(p->p0, sg, ar, ai, br, bi);
            for (int i = 0; i < n0; ++i) {
                fftwf_complex *row = a + (size_t)i * nc + j0;
                for (int v = 0; v < nv; ++v) { row[v][0] = ar[(size_t)i * VL + v]; row[v][1] = ai[(size_t)i * VL + v]; }
            }
        }
        fftwf_free(w);
    }
}

void fftwf_execute_dft_r2c(const fftwf_plan p, float *in, fftwf_complex *out)
{
    const int n0 = p->n0, n1 = p->n1, h = n1 / 2, nc = h + 1;
    const int even = (n1 % 2 == 0);
    const int L = even ? h : n1;
    const int nblk = (n0 + VL - 1) / VL;
    const int nt = shim_threads();
#pragma omp parallel num_threads(nt)
    {
        float *w = (float *)fftwf_malloc(sizeof(float) * 4 * (size_t)L * VL);
        float *ar = w, *ai = w + (size_t)L * VL, *br = ai + (size_t)L * VL, *bi = br + (size_t)L * VL;
#pragma omp for schedule(static)
        for (int b = 0; b < nblk; ++b) {
            const int i0 = b * VL, nv = (n0 - i0 < VL) ? n0 - i0 : VL;
            if (even) {
                for (int v = 0; v < VL; ++v) {
                    if (v < nv) {
                        const float *x = in + (size_t)(i0 + v) * n1;
                        for (int m = 0; m < h; ++m) { ar[(size_t)m * VL + v] = x[2 * m]; ai[(size_t)m * VL + v] = x[2 * m + 1]; }
                    } else {
                        for (int m = 0; m < h; ++m) { ar[(size_t)m * VL + v] = 0.f; ai[(size_t)m * VL + v] = 0.f; }
                    }
                }
                fft_lines(p->ph, -1.f, ar, ai, br, bi);
                /* X[k] = E[k] + w^k O[k], E = (Z[k] + conj Z[h-k])/2, O = -i (Z[k] - conj Z[h-k])/2 */
                for (int k = 0; k <= h; ++k) {
                    const int k1 = (k == h) ? 0 : k, k2 = (k == 0) ? 0 : h - k;
                    const float wr = p->hwr[k], wi = p->hwi[k];
                    for (int v = 0; v < nv; ++v) {
                        float zr = ar[(size_t)k1 * VL + v], zi = ai[(size_t)k1 * VL + v];
                        float yr = ar[(size_t)k2 * VL + v], yi = -ai[(size_t)k2 * VL + v];
                        float er = 0.5f * (zr + yr), ei = 0.5f * (zi + yi);
                        float dr = 0.5f * (zr - yr), di = 0.5f * (zi - yi);
                        /* O = -i d = (di, -dr) */
                        float orr = di, oi = -dr;
                        fftwf_complex *o = out + (size_t)(i0 + v) * nc + k;
                        (*o)[0] = er + (orr * wr - oi * wi);
                        (*o)[1] = ei + (orr * wi + oi * wr);
                    }
                }
            } else {
                for (int v = 0; v < VL; ++v)
                    for (int m = 0; m < n1; ++m) {
                        ar[(size_t)m * VL + v] = (v < nv) ? in[(size_t)(i0 + v) * n1 + m] : 0.f;
                        ai[(size_t)m * VL + v] = 0.f;
                    }
                fft_lines(p->ph, -1.f, ar, ai, br, bi);
                for (int k = 0; k <= h; ++k)
                    for (int v = 0; v < nv; ++v) {
                        fftwf_complex *o = out + (size_t)(i0 + v) * nc + k;
                        (*o)[0] = ar[(size_t)k * VL + v]; (*o)[1] = ai[(size_t)k * VL + v];
                    }
            }
        }
        fftwf_free(w);
    }
    columns(p, -1.f, out, nc);
}

void fftwf_execute_dft_c2r(const fftwf_plan p, fftwf_complex *in, float *out)
{
    const int n0 = p->n0, n1 = p->n1, h = n1 / 2, nc = h + 1;
    const int even = (n1 % 2 == 0);
    const int L = even ? h : n1;
    fftwf_complex *a = p->scratch;
    memcpy(a, in, sizeof(fftwf_complex) * (size_t)n0 * nc);
    columns(p, +1.f, a, nc);
    const int nblk = (n0 + VL - 1) / VL;
    const int nt = shim_threads();
#pragma omp parallel num_threads(nt)
    {
        float *w = (float *)fftwf_malloc(sizeof(float) * 4 * (size_t)L * VL);
        float *ar = w, *ai = w + (size_t)L * VL, *br = ai + (size_t)L * VL, *bi = br + (size_t)L * VL;
#pragma omp for schedule(static)
        for (int b = 0; b < nblk; ++b) {
            const int i0 = b * VL, nv = (n0 - i0 < VL) ? n0 - i0 : VL;
            if (even) {
                /* Z[k] = (X[k] + conj X[h-k]) + i (X[k] - conj X[h-k]) conj(w^k); Im X[0], Im X[h] dropped */
                for (int k = 0; k < h; ++k) {
                    const float wr = p->hwr[k], wi = -p->hwi[k];
                    for (int v = 0; v < VL; ++v) {
                        if (v >= nv) { ar[(size_t)k * VL + v] = 0.f; ai[(size_t)k * VL + v] = 0.f; continue; }
                        const fftwf_complex *x = a + (size_t)(i0 + v) * nc;
                        float xr = x[k][0], xi = (k == 0) ? 0.f : x[k][1];
                        float yr = x[h - k][0], yi = (k == 0) ? 0.f : -x[h - k][1];
                        float er = xr + yr, ei = xi + yi, dr = xr - yr, di = xi - yi;
                        float orr = dr * wr - di * wi, oi = dr * wi + di * wr;
                        ar[(size_t)k * VL + v] = er - oi;
                        ai[(size_t)k * VL + v] = ei + orr;
                    }
                }
                fft_lines(p->ph, +1.f, ar, ai, br, bi);
                for (int v = 0; v < nv; ++v) {
                    float *x = out + (size_t)(i0 + v) * n1;
                    for (int m = 0; m < h; ++m) { x[2 * m] = ar[(size_t)m * VL + v]; x[2 * m + 1] = ai[(size_t)m * VL + v]; }
                }
            } else {
                for (int v = 0; v < VL; ++v) {
                    const fftwf_complex *x = a + (size_t)(i0 + (v < nv ? v : 0)) * nc;
                    ar[v] = (v < nv) ? x[0][0] : 0.f; ai[v] = 0.f;
                    for (int k = 1; k <= h; ++k) {
                        float xr = (v < nv) ? x[k][0] : 0.f, xi = (v < nv) ? x[k][1] : 0.f;
                        ar[(size_t)k * VL + v] = xr; ai[(size_t)k * VL + v] = xi;
                        ar[(size_t)(n1 - k) * VL + v] = xr; ai[(size_t)(n1 - k) * VL + v] = -xi;
                    }
                }
                fft_lines(p->ph, +1.f, ar, ai, br, bi);
                for (int v = 0; v < nv; ++v)
                    for (int m = 0; m < n1; ++m) out[(size_t)(i0 + v) * n1 + m] = ar[(size_t)m * VL + v];
            }
        }
        fftwf_free(w);
    }
}

void fftwf_execute(const fftwf_plan p)
{
    if (p->kind == 0) fftwf_execute_dft_r2c(p, p->rbuf, p->cbuf);
    else fftwf_execute_dft_c2r(p, p->cbuf, p->rbuf);
}
