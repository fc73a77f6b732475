// main.cpp -- main.out: the reference's RK4 driver process on the xfb backend.
//
// Command line, stdout, `log` file and output file names follow /root/reference/src/main.cpp:65-328 and
// src/main-shallow-water.cpp:75-94 (the -s/-f forcing variant):
//   -I <input dir> -O <output dir> -i <initial file>      (main.cpp:68-80; -o is accepted as an alias of -O,
//                                                          README.md:36 documents lower-case flags)
//   -f <fifo> | -s <script>                                (main-shallow-water.cpp:86-93)
// Compile-time constants of src/configuration.hpp:13-36 are run-time options with the same defaults:
//   -n <NPTS=768> -d <dt=3> -t <total_steps=1200> -r <record_step=100> -L <600000> -N <NU=6.5>
// Additions: -g <cuda device>, -D (also write filamentation time / deformation factor at record steps),
//   -q (no per-step line).
// All arithmetic runs on the GPU through the C ABI (include/xfb.h); this file is I/O and control only.
#include <getopt.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../../include/xfb.h"
#include "fieldio.hpp"
#include "vorticity_source.hpp"

using namespace VORT_SRC_READER;

#define CHECK(call)                                                        \
    do {                                                                   \
        if ((call) != 0) {                                                 \
            std::fprintf(stderr, "main.out: %s\n", xfb_last_error());      \
            return 1;                                                      \
        }                                                                  \
    } while (0)

int main(int argc, char *args[])
{
    std::string input = "input", output = "output", init_file = "initial_vorticity.bin";   // configuration.hpp:39-41
    std::string vort_src_filename;
    RECIPE_TYPE recipe_type = EMPTY;
    int npts = 768, record_step = 100, total_steps = -1, device = 0;
    float L = 600000.0f, NU = 6.5f, dt = 3.0f;
    bool diagnostics = false, quiet = false;

    int opt;
    while ((opt = getopt(argc, args, "I:O:o:i:s:f:n:d:t:r:L:N:g:Dq")) != EOF) {
        switch (opt) {
        case 'I': input = optarg; break;
        case 'O': case 'o': output = optarg; break;
        case 'i': init_file = optarg; break;
        case 's': vort_src_filename = optarg; recipe_type = SCRIPT; break;
        case 'f': vort_src_filename = optarg; recipe_type = FIFO; break;
        case 'n': npts = std::atoi(optarg); break;
        case 'd': dt = (float)std::atof(optarg); break;
        case 't': total_steps = std::atoi(optarg); break;
        case 'r': record_step = std::atoi(optarg); break;
        case 'L': L = (float)std::atof(optarg); break;
        case 'N': NU = (float)std::atof(optarg); break;
        case 'g': device = std::atoi(optarg); break;
        case 'D': diagnostics = true; break;
        case 'q': quiet = true; break;
        }
    }
    if (total_steps < 0) total_steps = (int)(60 * 60 / dt);          // configuration.hpp:36
    if (record_step < 1) record_step = 1;
    const size_t GRIDS = (size_t)npts * npts;
    const float dx = L / npts, dy = L / npts;

    std::printf("##### Model setting #####\n");
    std::printf("Initial file          : %s \n", init_file.c_str());
    std::printf("Input folder          : %s \n", input.c_str());
    std::printf("Output folder         : %s \n", output.c_str());
    std::printf("Length X              : %.3f [m]\n", L);
    std::printf("Length Y              : %.3f [m]\n", L);
    std::printf("Spatial Resolution dx : %.3f [m]\n", dx);      // the reference prints 0 here (main.cpp:89-90 before :140)
    std::printf("Spatial Resolution dy : %.3f [m]\n", dy);
    std::printf("Time Resolution dt    : %.3f [s]\n", dt);
    std::printf("#########################\n\n\n");
    std::printf("Start project.\n");

    FILE *log_fd = std::fopen("log", "w");                        // main.cpp:97, in the CWD
    if (log_fd == NULL) perror("Open log file");

    xfb_handle h = nullptr;
    CHECK(xfb_create(&h, npts, npts, L, L, NU, 1, device));

    std::vector<float> field(GRIDS), src(GRIDS, 0.0f);
    char filename[1024];
    std::snprintf(filename, sizeof(filename), "%s/%s", input.c_str(), init_file.c_str());
    if (readFieldChecked(filename, field.data(), GRIDS) != 0) {
        std::fprintf(stderr, "main.out: cannot read the initial field %s\n", filename);
        return 1;
    }
    VortSrcRecipeReader vs_reader;
    if (vs_reader.init(recipe_type, vort_src_filename, GRIDS) != 0) return 1;

    std::printf("Initialization complete.\n");
    CHECK(xfb_set_vorticity(h, 0, field.data()));                  // step 01, main.cpp:256

    auto record = [&](const char *stem, int which, int step) -> int {
        if (xfb_get_field(h, 0, which, field.data()) != 0) return 1;
        std::snprintf(filename, sizeof(filename), "%s/%s_step_%d.bin", output.c_str(), stem, step);
        writeField(filename, field.data(), GRIDS);
        if (log_fd) { std::fprintf(log_fd, "%s\n", filename); std::fflush(log_fd); }
        return 0;
    };

    int step = 0;
    while (step < total_steps) {
        const bool record_flag = (step % record_step) == 0;
        if (!quiet) {
            std::printf("# Step %d, time = %.2f", step, step * dt);
            if (record_flag) std::printf(", record now!");
            std::printf("\n");
        }
        if (record_flag) {
            // same order as the reference's log: source, vort (main.cpp:268-278), then psi, u, v (:183-222)
            if (record("vort_src_input", XFB_SRC, step) || record("vort", XFB_VORT, step) ||
                record("psi", XFB_PSI, step) || record("u", XFB_U, step) || record("v", XFB_V, step)) {
                std::fprintf(stderr, "main.out: %s\n", xfb_last_error());
                return 1;
            }
            if (diagnostics) {
                // filamentation time and deformation factor from one set of second derivatives of psi
                std::vector<float> deform(GRIDS);
                if (xfb_get_diagnostics(h, 0, field.data(), deform.data()) != 0) {
                    std::fprintf(stderr, "main.out: %s\n", xfb_last_error());
                    return 1;
                }
                const char *stems[2] = {"tfil", "deform"};
                float *bufs[2] = {field.data(), deform.data()};
                for (int o = 0; o < 2; ++o) {
                    std::snprintf(filename, sizeof(filename), "%s/%s_step_%d.bin", output.c_str(), stems[o], step);
                    writeField(filename, bufs[o], GRIDS);
                    if (log_fd) { std::fprintf(log_fd, "%s\n", filename); std::fflush(log_fd); }
                }
            }
        }
        int chunk = 1;
        if (recipe_type == EMPTY) {
            // no per-step host work until the next record step: run the whole stretch on the device
            const int next_record = (step / record_step + 1) * record_step;
            chunk = (next_record < total_steps ? next_record : total_steps) - step;
            if (!quiet)
                for (int s = step + 1; s < step + chunk; ++s) std::printf("# Step %d, time = %.2f\n", s, s * dt);
        } else {
            const int got = vs_reader.read(step * dt, src.data());       // main-shallow-water.cpp:304
            if (got == 1) CHECK(xfb_set_source(h, 0, src.data()));
        }
        CHECK(xfb_step(h, chunk, dt));
        step += chunk;
    }
    CHECK(xfb_sync(h));
    if (log_fd) std::fclose(log_fd);
    xfb_destroy(h);
    std::printf("Program ends. Congrats!\n");
    return 0;
}
