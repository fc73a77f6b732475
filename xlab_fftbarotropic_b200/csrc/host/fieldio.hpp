// fieldio.hpp -- headerless little-endian float32 field files.
// Same two signatures as the reference (/root/reference/src/fieldio.hpp:5-6) so host code that calls
// writeField/readField links unchanged; unlike the reference (src/fieldio.cpp:9-18,23-32, which
// dereferences a NULL FILE* on a missing file) failures are reported through the return value of the
// *_checked variants and never crash.
#pragma once
#include <cstddef>

void writeField(const char *filename, float *data, size_t len);
void readField(const char *filename, float *data, size_t len);

// 0 on success, -1 on open failure, -2 on short read/write
int writeFieldChecked(const char *filename, const float *data, size_t len);
int readFieldChecked(const char *filename, float *data, size_t len);
