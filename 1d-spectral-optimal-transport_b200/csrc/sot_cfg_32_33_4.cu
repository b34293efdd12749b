// Kernel instantiations for 32 threads per frame, 33 bins per thread, 4 frame(s) per CTA.
#include "sot_launch.cuh"
SOT_DEFINE_CONFIG(32, 33, 4)
