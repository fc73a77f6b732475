// makefield.cpp -- the four initial-vortex generators, host-side (north_star: generators stay on the host).
// Built four times with -DXFB_GEN=<kind> into makefield-{elliptic-vortex,const-vortex,gaussian,Kuo2004}.out;
// each writes <input>/<init_file> like the reference programs.  NPTS is a run-time option (-n, default 768).
// The float/double promotion of every expression follows the reference sources so that the 768^2 files are
// byte-identical (md5 pinned in tests/test_host_programs.py):
//   elliptic  /root/reference/src/makefield-elliptic-vortex.cpp:14-50
//   const     src/makefield-const-vortex.cpp:14-35
//   gaussian  src/makefield-gaussian.cpp:14-31
//   Kuo2004   src/makefield-Kuo2004.cpp:30-44 + src/field_generator.cpp:5-28
#include <getopt.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "fieldio.hpp"

#define GEN_ELLIPTIC 1
#define GEN_CONST 2
#define GEN_GAUSSIAN 3
#define GEN_KUO2004 4
#ifndef XFB_GEN
#define XFB_GEN GEN_ELLIPTIC
#endif

static float LX = 600000.0f, LY = 600000.0f;
static int XPTS = 768, YPTS = 768;
static inline size_t IDX(int i, int j) { return (size_t)YPTS * i + j; }

static float radius(float x, float y, float cx, float cy) { return sqrtf(pow(x - cx, 2) + pow(y - cy, 2)); }

#if XFB_GEN == GEN_KUO2004
static void addCake(float *data, float cx, float cy, float zeta_0, float scale_r)
{
    const float DX = LX / XPTS, DY = LY / YPTS;
    for (size_t j = 0; j < (size_t)XPTS; ++j) {
        const float y = j * DY;
        for (size_t i = 0; i < (size_t)YPTS; ++i) {
            const float x = i * DX;
            const float r = sqrtf(pow(x - cx, 2.0) + pow(y - cy, 2.0)) / scale_r;
            if (r < 1) data[IDX((int)i, (int)j)] += zeta_0 * (1 - exp(-30.0 / r * exp(1.0 / (r - 1.0))));
        }
    }
}
#endif

int main(int argc, char *args[])
{
    std::string input = "input", init_file = "initial_vorticity.bin";
    int opt;
    while ((opt = getopt(argc, args, "I:i:n:L:")) != EOF) {
        switch (opt) {
        case 'I': input = optarg; break;
        case 'i': init_file = optarg; break;
        case 'n': XPTS = YPTS = std::atoi(optarg); break;
        case 'L': LX = LY = (float)std::atof(optarg); break;
        }
    }
    const size_t GRIDS = (size_t)XPTS * YPTS;
    std::vector<float> vort(GRIDS, 0.0f);
    const float centerx = LX / 2.0, centery = LY / 2.0;
    const float dx = LX / XPTS, dy = LY / YPTS;
    (void)dx; (void)dy; (void)centerx; (void)centery;

#if XFB_GEN == GEN_KUO2004
    addCake(vort.data(), LX / 2.0, LY / 2.0, 1.5e-2, 10000.0);
    addCake(vort.data(), LX / 2.0 + 50000.0, LY / 2.0, 3e-3, 30000.0);
#else
    for (int i = 0; i < XPTS; ++i) {
        const float x = i * dx;
        for (int j = 0; j < YPTS; ++j) {
            const float y = j * dy;
            const float r = radius(x, y, centerx, centery);
#if XFB_GEN == GEN_ELLIPTIC
            const float epsilon = 0.7, lambda = 2.0, zeta0 = .005f, r_i = 30000.0, r_o = 60000.0;
            float c;
            if (r == 0.0f) c = 0; else c = (y - centery) / r;
            const float alpha = sqrtf((1.0 - pow(epsilon, 2)) / (1.0 - pow(epsilon * c, 2)));
            const float r_i_alpha = r_i * alpha, r_o_alpha = r_o * alpha;
            if (r <= r_i_alpha) {
                vort[IDX(i, j)] = zeta0;
            } else if (r <= r_o_alpha) {
                const float r_prime = (r - r_i_alpha) / (r_o_alpha - r_i_alpha);
                vort[IDX(i, j)] = zeta0 * (1.0 - exp(-lambda / r_prime * exp(1.0 / (r_prime - 1))));
            } else {
                vort[IDX(i, j)] = 0;
            }
#elif XFB_GEN == GEN_CONST
            const float r_bound = 6000.0, zeta0 = 2e-5;
            vort[IDX(i, j)] = (r <= r_bound) ? zeta0 : 0;
#elif XFB_GEN == GEN_GAUSSIAN
            const float zeta0 = 1e-3;
            vort[IDX(i, j)] = zeta0 * exp(-pow(r / 60000.0, 2.0));
#endif
        }
    }
#endif
    const std::string file = input + "/" + init_file;
    return writeFieldChecked(file.c_str(), vort.data(), GRIDS) == 0 ? 0 : 1;
}
