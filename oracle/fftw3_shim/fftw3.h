/* fftw3.h -- declare-only stand-in for FFTW3 single precision.
 *
 * TEST INFRASTRUCTURE (oracle/): never linked into the product library.
 *
 * FFTW3 (>= 3.3.4, /root/reference/README.md:15-18, Makefile:10) is an un-vendored,
 * un-pinned system dependency of the reference and is not installed in this image.
 * The reference's FFT path uses exactly seven names from it
 * (/root/reference/src/main.cpp:103-135,154,...; src/invert_pres.cpp:84-107):
 *   fftwf_complex, fftwf_plan, FFTW_ESTIMATE, fftwf_malloc, fftwf_free,
 *   fftwf_plan_dft_r2c_2d, fftwf_plan_dft_c2r_2d, fftwf_execute.
 * This header declares them; shim_fft.c implements them with a plain mixed-radix
 * CPU FFT following FFTW's published conventions (FFTW manual 4.3.x "Real-data DFTs",
 * 4.8 "What FFTW Really Computes"):
 *   r2c_2d(n0,n1): Y[k0][k1] = sum x[j0][j1] exp(-2 pi i (j0 k0/n0 + j1 k1/n1)),
 *                  k1 = 0..n1/2, row-major, last dimension halved;
 *   c2r_2d(n0,n1): the unnormalised inverse (sign +), input is the half spectrum,
 *                  imaginary parts of the k1 = 0 and k1 = n1/2 bins of the last-dimension
 *                  transform are ignored, and the input array may be overwritten.
 */
#ifndef XFB_ORACLE_FFTW3_SHIM_H
#define XFB_ORACLE_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct xfb_shim_plan_s *fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

void *fftwf_malloc(size_t n);
void fftwf_free(void *p);

fftwf_plan fftwf_plan_dft_r2c_2d(int n0, int n1, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_2d(int n0, int n1, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);

/* new-array execute (FFTW's guru "execute_dft_r2c/c2r"): used only by the oracle
 * restatement, which owns no plans bound to fixed buffers */
void fftwf_execute_dft_r2c(const fftwf_plan p, float *in, fftwf_complex *out);
void fftwf_execute_dft_c2r(const fftwf_plan p, fftwf_complex *in, float *out);

/* shim-only: number of OpenMP threads the transforms use (0 = OpenMP default) */
void xfb_shim_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
