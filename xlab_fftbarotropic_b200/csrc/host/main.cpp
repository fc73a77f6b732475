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
//   -q (no per-step line), -c <tracer file> [-k <kappa>] (advect a passive tracer read from <input dir>, written as
//   tracer_step_N.bin after the reference's five files).
// Record output is asynchronous: at a record step the five fields are formed on the GPU and copied
// to pinned host buffers on a second stream (xfb_get_field_async) while the next stretch of steps already runs; a
// writer thread waits for each buffer, calls writeField and appends the `log` line -- same files, same order as the
// reference (main.cpp:266-282,183-222).
// All arithmetic runs on the GPU through the C ABI (include/xfb.h); this file is I/O and control only.
#include <getopt.h>
#include <unistd.h>

#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
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
    std::string vort_src_filename, tracer_file;
    float kappa = -1.0f;
    RECIPE_TYPE recipe_type = EMPTY;
    int npts = 768, record_step = 100, total_steps = -1, device = 0;
    float L = 600000.0f, NU = 6.5f, dt = 3.0f;
    bool diagnostics = false, quiet = false;

    int opt;
    while ((opt = getopt(argc, args, "I:O:o:i:s:f:n:d:t:r:L:N:g:c:k:Dq")) != EOF) {
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
        case 'c': tracer_file = optarg; break;        // passive tracer: initial field <input>/<file>, tracer_step_N.bin
        case 'k': kappa = (float)std::atof(optarg); break;   // its diffusivity (default: NU)
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

    const bool tracer = !tracer_file.empty();
    if (tracer) {
        std::snprintf(filename, sizeof(filename), "%s/%s", input.c_str(), tracer_file.c_str());
        if (readFieldChecked(filename, field.data(), GRIDS) != 0) {
            std::fprintf(stderr, "main.out: cannot read the tracer field %s\n", filename);
            return 1;
        }
        CHECK(xfb_set_tracer(h, 0, field.data(), kappa >= 0.0f ? kappa : NU));
    }

    // ---- asynchronous record output: one pinned buffer per field kind, a writer thread, jobs in log order
    struct Job { int ticket; float *buf; std::string file; };
    std::deque<Job> jobs;
    std::mutex mu;
    std::condition_variable cv_jobs, cv_idle;
    bool closing = false, writer_failed = false;
    int pending = 0;
    std::thread writer([&]() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_jobs.wait(lk, [&] { return closing || !jobs.empty(); });
                if (jobs.empty()) return;
                j = jobs.front();
                jobs.pop_front();
            }
            if (xfb_wait_field(h, j.ticket) != 0) writer_failed = true;
            writeField(j.file.c_str(), j.buf, GRIDS);
            if (log_fd) { std::fprintf(log_fd, "%s\n", j.file.c_str()); std::fflush(log_fd); }
            {
                std::lock_guard<std::mutex> lk(mu);
                --pending;
            }
            cv_idle.notify_all();
        }
    });
    auto wait_writer = [&]() {
        std::unique_lock<std::mutex> lk(mu);
        cv_idle.wait(lk, [&] { return pending == 0; });
    };
    const int NKIND = 8;
    float *pinned[NKIND] = {nullptr};
    auto record = [&](int kind, const char *stem, int which, int step) -> int {
        if (!pinned[kind] && xfb_host_alloc(&pinned[kind], GRIDS) != 0) return 1;
        int ticket = -1;
        if (xfb_get_field_async(h, 0, which, pinned[kind], &ticket) != 0) return 1;
        std::snprintf(filename, sizeof(filename), "%s/%s_step_%d.bin", output.c_str(), stem, step);
        {
            std::lock_guard<std::mutex> lk(mu);
            jobs.push_back(Job{ticket, pinned[kind], filename});
            ++pending;
        }
        cv_jobs.notify_one();
        return 0;
    };
    auto shutdown_writer = [&]() {
        {
            std::lock_guard<std::mutex> lk(mu);
            closing = true;
        }
        cv_jobs.notify_all();
        writer.join();
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
            wait_writer();                     // the pinned buffers of the previous record step are on disk
            // same order as the reference's log: source, vort (main.cpp:268-278), then psi, u, v (:183-222)
            if (record(0, "vort_src_input", XFB_SRC, step) || record(1, "vort", XFB_VORT, step) ||
                record(2, "psi", XFB_PSI, step) || record(3, "u", XFB_U, step) || record(4, "v", XFB_V, step) ||
                (diagnostics && (record(5, "tfil", XFB_TFIL, step) || record(6, "deform", XFB_DEFORM, step))) ||
                (tracer && record(7, "tracer", XFB_TRACER, step))) {
                std::fprintf(stderr, "main.out: %s\n", xfb_last_error());
                shutdown_writer();
                return 1;
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
            if (got == 1 && xfb_set_source(h, 0, src.data()) != 0) {
                std::fprintf(stderr, "main.out: %s\n", xfb_last_error());
                shutdown_writer();
                return 1;
            }
        }
        if (xfb_step(h, chunk, dt) != 0) {
            std::fprintf(stderr, "main.out: %s\n", xfb_last_error());
            shutdown_writer();
            return 1;
        }
        step += chunk;
    }
    const int sync_rc = xfb_sync(h);
    wait_writer();
    shutdown_writer();
    if (sync_rc != 0 || writer_failed) {
        std::fprintf(stderr, "main.out: %s\n", xfb_last_error());
        return 1;
    }
    for (int k = 0; k < NKIND; ++k) xfb_host_free(pinned[k]);
    if (log_fd) std::fclose(log_fd);
    xfb_destroy(h);
    std::printf("Program ends. Congrats!\n");
    return 0;
}
