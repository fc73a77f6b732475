// fftwfop.hpp (compat) -- same name as /root/reference/src/fftwfop.hpp; forwards to the GPU-backed class.
#include "../fftwfop.hpp"
