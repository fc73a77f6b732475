// fftwfop.hpp -- header-compatible replacement of the reference's operator class
// (/root/reference/src/fftwfop.hpp:9-29, implementation src/fftwfop.cpp:5-124) on top of the xfb C ABI.
//
// Same class name, template arguments, constructor and method names; the five methods run the CUDA
// pointwise kernels of libxfb on caller-owned HOST buffers of HALF_GRIDS fftwf_complex (they may also be
// device pointers).  Host code written against the reference header compiles unchanged after replacing
//     #include "fftwfop.cpp"      by      #include <fftwfop.hpp>
// (see INTEGRATION.md).  The 2-D transforms the reference does through raw FFTW plans are available as
// r2c()/c2r() on the same object, or through include/compat/fftw3.h.
#ifndef XFB_FFTWF_OPERATION_HPP
#define XFB_FFTWF_OPERATION_HPP

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "xfb.h"

#ifndef XFB_FFTWF_COMPLEX_DEFINED
#define XFB_FFTWF_COMPLEX_DEFINED
#ifndef FFTW_ESTIMATE   /* a real or compat fftw3.h already defines fftwf_complex */
typedef float fftwf_complex[2];
#endif
#endif

const float TWOPI = (acos(-1.0f) * 2.0f);

template <int XPTS, int YPTS>
class fftwf_operation {
private:
    xfb_handle h_;
    const int HALF_XPTS = (int)(XPTS / 2) + 1, HALF_YPTS = (int)(YPTS / 2) + 1, HALF_GRIDS = XPTS * HALF_YPTS;

    void check(int rc, const char *what)
    {
        if (rc != 0) {
            std::fprintf(stderr, "fftwf_operation::%s: %s\n", what, xfb_last_error());
            std::exit(EXIT_FAILURE);   // the reference's methods are void and cannot report errors
        }
    }

public:
    // nu and device are extensions with defaults; the reference constructor is (Lx, Ly)
    fftwf_operation(float Lx, float Ly, float nu = 0.0f, int device = 0) : h_(nullptr)
    {
        check(xfb_create(&h_, XPTS, YPTS, Lx, Ly, nu, 1, device), "fftwf_operation");
    }
    ~fftwf_operation() { xfb_destroy(h_); }
    fftwf_operation(const fftwf_operation &) = delete;
    fftwf_operation &operator=(const fftwf_operation &) = delete;

    void gradx(fftwf_complex *in, fftwf_complex *out) { check(xfb_gradx(h_, (const float *)in, (float *)out), "gradx"); }
    void grady(fftwf_complex *in, fftwf_complex *out) { check(xfb_grady(h_, (const float *)in, (float *)out), "grady"); }
    void laplacian(fftwf_complex *in, fftwf_complex *out) { check(xfb_laplacian(h_, (const float *)in, (float *)out), "laplacian"); }
    void invertLaplacian(fftwf_complex *in, fftwf_complex *out)
    {
        check(xfb_invert_laplacian(h_, (const float *)in, (float *)out), "invertLaplacian");
    }
    void dealiase(fftwf_complex *in, fftwf_complex *out) { check(xfb_dealias(h_, (const float *)in, (float *)out), "dealiase"); }

    // the transforms main.cpp:126-135 performs with raw FFTW plans (unnormalised, non-destructive)
    void r2c(float *in, fftwf_complex *out) { check(xfb_r2c(h_, in, (float *)out), "r2c"); }
    void c2r(fftwf_complex *in, float *out) { check(xfb_c2r(h_, (const float *)in, out), "c2r"); }

    inline int reflectedXWavenumberIndex(int i)
    {
        assert(i >= 1 && "Input of ReflectedXWavenumberIndex must >= 1");
        return XPTS - i;
    }
    inline int HIDX(int i, int j) { return HALF_YPTS * i + j; }
    inline int R_HIDX(int i, int j) { return HIDX(this->reflectedXWavenumberIndex(i), j); }

    xfb_handle handle() { return h_; }
};

#endif
