// Kernel instantiations for 128 threads per frame, 33 bins per thread, 1 frame(s) per CTA.
#include "sot_launch.cuh"
SOT_DEFINE_CONFIG(128, 33, 1)
