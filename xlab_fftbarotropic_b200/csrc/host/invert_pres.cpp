// invert_pres.cpp -- invert_pres.out on the xfb backend: psi files -> pressure anomaly files.
// Protocol and flags of /root/reference/src/invert_pres.cpp:65-192: stdin lines `from=>to`,
// -x <ref_x> -y <ref_y>; -n <NPTS> -L <length> -g <device> are run-time versions of configuration.hpp.
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../../include/xfb.h"
#include "fieldio.hpp"

int main(int argc, char *args[])
{
    size_t ref_x = 0, ref_y = 0;
    int npts = 768, device = 0;
    float L = 600000.0f;
    const float rho = 1.0f, f = 1e-5;                  // configuration.hpp:10-11
    int opt;
    while ((opt = getopt(argc, args, "x:y:n:L:g:")) != EOF) {
        switch (opt) {
        case 'x': ref_x = std::atoi(optarg); break;
        case 'y': ref_y = std::atoi(optarg); break;
        case 'n': npts = std::atoi(optarg); break;
        case 'L': L = (float)std::atof(optarg); break;
        case 'g': device = std::atoi(optarg); break;
        }
    }
    xfb_handle h = nullptr;
    if (xfb_create(&h, npts, npts, L, L, 0.0f, 1, device) != 0) {
        std::fprintf(stderr, "invert_pres.out: %s\n", xfb_last_error());
        return 1;
    }
    const size_t GRIDS = (size_t)npts * npts;
    std::vector<float> psi(GRIDS), pres(GRIDS);
    char line[1024], from_file[1024], to_file[1024];
    const char sep[] = "=>";
    while (std::fgets(line, sizeof(line), stdin) != NULL) {
        line[std::strcspn(line, "\r\n")] = '\0';
        char *sep_beg = std::strstr(line, sep);
        if (sep_beg == NULL) {
            std::printf("Error reading input: %s. Continue next line...\n", line);     // invert_pres.cpp:128
            continue;
        }
        const size_t l = (size_t)(sep_beg - line);
        std::memcpy(from_file, line, l);
        from_file[l] = '\0';
        std::strcpy(to_file, sep_beg + std::strlen(sep));
        if (readFieldChecked(from_file, psi.data(), GRIDS) != 0) continue;
        if (xfb_invert_pres(h, psi.data(), pres.data(), ref_x, ref_y, rho, f) != 0) {
            std::fprintf(stderr, "invert_pres.out: %s\n", xfb_last_error());
            return 1;
        }
        writeField(to_file, pres.data(), GRIDS);
    }
    xfb_destroy(h);
    std::printf("Program ends. Congrats!\n");
    return 0;
}
