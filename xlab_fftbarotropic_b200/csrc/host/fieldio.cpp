#include "fieldio.hpp"

#include <cerrno>
#include <cstdio>
#include <cstring>

int writeFieldChecked(const char *filename, const float *data, size_t len)
{
    FILE *f = std::fopen(filename, "wb");
    if (!f) {
        std::fprintf(stderr, "Write field: cannot open %s: %s\n", filename, std::strerror(errno));
        return -1;
    }
    const size_t n = std::fwrite(data, sizeof(float), len, f);
    std::fclose(f);
    if (n != len) {
        std::fprintf(stderr, "Write field: short write on %s (%zu of %zu)\n", filename, n, len);
        return -2;
    }
    std::fprintf(stderr, "Output %s\n", filename);          // same chatter as src/fieldio.cpp:18
    return 0;
}

int readFieldChecked(const char *filename, float *data, size_t len)
{
    FILE *f = std::fopen(filename, "rb");
    if (!f) {
        std::fprintf(stderr, "Read field: cannot open %s: %s\n", filename, std::strerror(errno));
        return -1;
    }
    const size_t n = std::fread(data, sizeof(float), len, f);
    std::fclose(f);
    std::fprintf(stderr, "%zu bytes read: %s\n", n, filename);   // element count, worded as in src/fieldio.cpp:32
    return n == len ? 0 : -2;
}

void writeField(const char *filename, float *data, size_t len) { (void)writeFieldChecked(filename, data, len); }
void readField(const char *filename, float *data, size_t len) { (void)readFieldChecked(filename, data, len); }
