// Kernel instantiations: 128 threads per frame, 9 bins per thread, shared-memory rows of 1032 floats,
// 1 merge chain(s) per thread.
#include "sot_launch.cuh"
SOT_DEFINE_CONFIG(128, 9, 1032, 1)
