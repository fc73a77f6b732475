/* fftw3.h (compat) -- the eight FFTW3f names the reference's FFT path uses, on the xfb C ABI.
 *
 * /root/reference/src/main.cpp:103-135,154,168,186,200,214,237,256,275 and src/invert_pres.cpp:84-107,135,153-172 use
 * exactly: fftwf_complex, fftwf_plan, FFTW_ESTIMATE, fftwf_malloc, fftwf_free, fftwf_plan_dft_r2c_2d,
 * fftwf_plan_dft_c2r_2d, fftwf_execute.  With this directory ahead of the system include path
 * (`g++ -Iinclude/compat -Iinclude ... -lxfb`) those sources compile UNCHANGED and every transform runs on the GPU:
 *
 *   fftwf_plan_dft_r2c_2d(n0, n1, in, out, flags)  -> a plan bound to (in, out); one xfb handle per (n0, n1) per process
 *   fftwf_execute(plan)                             -> xfb_r2c / xfb_c2r on the plan's buffers (host or device pointers)
 *   fftwf_malloc / fftwf_free                       -> pinned host memory (xfb_host_alloc), so the copies run at PCIe speed
 *
 * Conventions are FFTW's (manual 4.3 / 4.8): r2c_2d is the unnormalised forward transform with the last dimension
 * halved, c2r_2d the unnormalised inverse that ignores the imaginary parts of the k1 = 0 and k1 = n1/2 bins.  One
 * difference, on the safe side: FFTW's multi-dimensional c2r may destroy its input (the reference keeps backups for
 * that, main.cpp:185,191,273,281); xfb_c2r never does.
 * The companion file fftwfop.cpp in this directory makes `#include "fftwfop.cpp"` (main.cpp:19, invert_pres.cpp:28)
 * resolve to include/fftwfop.hpp.  Header-only, C++11.
 */
#ifndef XFB_COMPAT_FFTW3_H
#define XFB_COMPAT_FFTW3_H

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>

#include "../xfb.h"

typedef float fftwf_complex[2];

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

struct xfb_compat_plan_s {
    xfb_handle h;
    int kind;              /* 0: r2c, 1: c2r */
    void *in, *out;
};
typedef struct xfb_compat_plan_s *fftwf_plan;

/* one handle per grid size (the reference creates up to eight plans of the same size, main.cpp:126-135) */
static inline xfb_handle xfb_compat_handle(int n0, int n1)
{
    static struct { int n0, n1; xfb_handle h; } cache[8];
    static int used = 0;
    for (int i = 0; i < used; ++i)
        if (cache[i].n0 == n0 && cache[i].n1 == n1) return cache[i].h;
    xfb_handle h = NULL;
    const char *dev = getenv("XFB_DEVICE");
    /* Lx, Ly and nu do not enter the plain transforms */
    if (xfb_create(&h, n0, n1, 1.0f, 1.0f, 0.0f, 1, dev ? atoi(dev) : 0) != 0) {
        fprintf(stderr, "fftw3.h (xfb compat): %s\n", xfb_last_error());
        exit(EXIT_FAILURE);            /* FFTW's planner has no error path the reference checks */
    }
    if (used < 8) { cache[used].n0 = n0; cache[used].n1 = n1; cache[used].h = h; ++used; }
    return h;
}

static inline void *fftwf_malloc(size_t n)
{
    float *p = NULL;
    if (xfb_host_alloc(&p, (n + sizeof(float) - 1) / sizeof(float)) != 0) return malloc(n);
    return p;
}

static inline void fftwf_free(void *p)
{
    if (p && xfb_host_free((float *)p) != 0) free(p);
}

static inline fftwf_plan fftwf_plan_dft_r2c_2d(int n0, int n1, float *in, fftwf_complex *out, unsigned flags)
{
    (void)flags;
    fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
    p->h = xfb_compat_handle(n0, n1); p->kind = 0; p->in = in; p->out = out;
    return p;
}

static inline fftwf_plan fftwf_plan_dft_c2r_2d(int n0, int n1, fftwf_complex *in, float *out, unsigned flags)
{
    (void)flags;
    fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
    p->h = xfb_compat_handle(n0, n1); p->kind = 1; p->in = in; p->out = out;
    return p;
}

static inline void fftwf_execute(const fftwf_plan p)
{
    const int rc = p->kind == 0 ? xfb_r2c(p->h, (const float *)p->in, (float *)p->out)
                                : xfb_c2r(p->h, (const float *)p->in, (float *)p->out);
    if (rc != 0) {
        fprintf(stderr, "fftwf_execute (xfb compat): %s\n", xfb_last_error());
        exit(EXIT_FAILURE);
    }
}

static inline void fftwf_destroy_plan(fftwf_plan p) { free(p); }

#endif
