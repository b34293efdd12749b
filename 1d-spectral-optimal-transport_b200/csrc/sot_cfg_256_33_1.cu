// Kernel instantiations for 256 threads per frame, 33 bins per thread, 1 frame(s) per CTA.
#include "sot_launch.cuh"
SOT_DEFINE_CONFIG(256, 33, 1)
