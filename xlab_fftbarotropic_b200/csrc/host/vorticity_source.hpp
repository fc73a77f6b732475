// vorticity_source.hpp -- host-side forcing reader with the reference's recipe types
// (/root/reference/src/vorticity_source.cpp:11 RECIPE_TYPE {SCRIPT, FIFO, EMPTY}).
//   FIFO   : per step one flag byte; if it is 1, `grids` float32 follow       (vorticity_source.cpp:112-133)
//   SCRIPT : text file, one `[time] [binary filename]` per line, '#' comments (format documented at
//            vorticity_source.cpp:13-19; the reference's readScript is an empty stub, :100-110 --
//            implemented here: the field of the last line whose time <= t is active)
//   EMPTY  : no forcing
// read(t, buf) returns 1 when `buf` received a new field (the caller uploads it with xfb_set_source),
// 0 when nothing changed, <0 on a short read.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "fieldio.hpp"

namespace VORT_SRC_READER {

enum RECIPE_TYPE { SCRIPT, FIFO, EMPTY };

class VortSrcRecipeReader {
    RECIPE_TYPE type_ = EMPTY;
    size_t grids_ = 0;
    FILE *fifo_ = nullptr;
    struct Recipe { float time; std::string filename; };
    std::vector<Recipe> recipes_;
    int active_ = -1;

public:
    ~VortSrcRecipeReader() { if (fifo_) std::fclose(fifo_); }

    int init(RECIPE_TYPE type, const std::string &filename, size_t grids)
    {
        type_ = type;
        grids_ = grids;
        if (type == FIFO) {
            fifo_ = std::fopen(filename.c_str(), "rb");
            if (!fifo_) { std::printf("ERROR: cannot open file [%s].\n", filename.c_str()); return -1; }
        } else if (type == SCRIPT) {
            FILE *fd = std::fopen(filename.c_str(), "r");
            if (!fd) { std::printf("ERROR: cannot open file [%s].\n", filename.c_str()); return -1; }
            char buf[1024];
            while (std::fgets(buf, sizeof(buf), fd)) {
                if (char *hash = std::strchr(buf, '#')) *hash = '\0';      // stripComment, vorticity_source.cpp:29-36
                float tm; char name[512];
                if (std::sscanf(buf, "%f %511s", &tm, name) == 2) recipes_.push_back({tm, name});
            }
            std::fclose(fd);
        }
        return 0;
    }

    int read(float time, float *buf)
    {
        if (type_ == EMPTY) return 0;
        if (type_ == FIFO) {
            char flag;
            if (std::fread(&flag, 1, 1, fifo_) != 1) {
                std::fprintf(stderr, "No flag was detected, assume flag = 0\n");
                return 0;
            }
            if ((unsigned int)flag == 1) {
                if (std::fread(buf, sizeof(float), grids_, fifo_) != grids_) {
                    std::fprintf(stderr, "ERROR: Cannot read vorticity source input.\n");
                    return -2;
                }
                std::fprintf(stderr, "New vorticity source was given.\n");
                return 1;
            }
            std::fprintf(stderr, "No new vorticity source input was given.\n");
            return 0;
        }
        int want = -1;
        for (size_t k = 0; k < recipes_.size(); ++k)
            if (recipes_[k].time <= time) want = (int)k;
        if (want == active_ || want < 0) return 0;
        active_ = want;
        return readFieldChecked(recipes_[want].filename.c_str(), buf, grids_) == 0 ? 1 : -2;
    }
};

}  // namespace VORT_SRC_READER
