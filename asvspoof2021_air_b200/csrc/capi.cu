// Miscellaneous C-ABI entry points.
#include "common.cuh"
extern "C" int air_version() { return 100; }
