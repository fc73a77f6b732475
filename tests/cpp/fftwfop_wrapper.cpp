// fftwfop_wrapper.cpp -- exercises include/fftwfop.hpp exactly as the reference's host code uses the class
// (/root/reference/src/fftwfop.hpp:9-29): a global object constructed before main (main.cpp:33), the five void methods
// on caller-owned fftwf_complex buffers, and the index helpers HIDX / R_HIDX / reflectedXWavenumberIndex (:26-28).
// Reads a half spectrum from argv[1] (N x (N/2+1) complex64), writes gradx, grady, laplacian, invertLaplacian, dealiase
// (in place, as invert_pres.cpp:148-150 calls it) to argv[2].{gradx,...}; prints the helper values for the test to check.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fftw3.h>       // include/compat/fftw3.h
#include "fftwfop.cpp"   // include/compat/fftwfop.cpp -> include/fftwfop.hpp

#ifndef NPTS
#define NPTS 256
#endif
const int XPTS = NPTS, YPTS = NPTS, HALF_YPTS = YPTS / 2 + 1, HALF_GRIDS = XPTS * HALF_YPTS;
const float LX = 600000.0f, LY = 600000.0f;

fftwf_operation<XPTS, YPTS> fop(LX, LY);          // constructed before main, like the reference's

static int dump(const std::string &path, const fftwf_complex *a)
{
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return 1;
    std::fwrite(a, sizeof(fftwf_complex), HALF_GRIDS, f);
    std::fclose(f);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    fftwf_complex *in = (fftwf_complex *)fftwf_malloc(sizeof(fftwf_complex) * HALF_GRIDS);
    fftwf_complex *out = (fftwf_complex *)fftwf_malloc(sizeof(fftwf_complex) * HALF_GRIDS);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(in, sizeof(fftwf_complex), HALF_GRIDS, f) != (size_t)HALF_GRIDS) return 3;
    std::fclose(f);
    const std::string stem = argv[2];
    fop.gradx(in, out);            if (dump(stem + ".gradx", out)) return 4;
    fop.grady(in, out);            if (dump(stem + ".grady", out)) return 4;
    fop.laplacian(in, out);        if (dump(stem + ".laplacian", out)) return 4;
    fop.invertLaplacian(in, out);  if (dump(stem + ".invertLaplacian", out)) return 4;
    std::memcpy(out, in, sizeof(fftwf_complex) * HALF_GRIDS);
    fop.dealiase(out, out);        if (dump(stem + ".dealiase", out)) return 4;      // in == out
    std::printf("HIDX(3,5)=%d R_HIDX(3,5)=%d reflected(1)=%d reflected(%d)=%d\n", fop.HIDX(3, 5), fop.R_HIDX(3, 5),
                fop.reflectedXWavenumberIndex(1), XPTS / 2, fop.reflectedXWavenumberIndex(XPTS / 2));
    fftwf_free(in);
    fftwf_free(out);
    return 0;
}
