// Kernel instantiations for 128 threads per frame, 17 bins per thread, 4 frame(s) per CTA.
#include "sot_launch.cuh"
SOT_DEFINE_CONFIG(128, 17, 4)
