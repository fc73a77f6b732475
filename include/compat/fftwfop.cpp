// fftwfop.cpp (compat) -- the reference includes its template implementation by file name
// (/root/reference/src/main.cpp:19, src/invert_pres.cpp:28: `#include "fftwfop.cpp"`); with include/compat on the quote
// include path that line resolves here and picks up the GPU-backed class of the same name, constructor and methods.
#include "../fftwfop.hpp"
