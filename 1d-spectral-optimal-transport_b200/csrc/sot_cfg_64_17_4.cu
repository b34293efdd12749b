// Kernel instantiations for 64 threads per frame, 17 bins per thread, 4 frame(s) per CTA.
#include "sot_launch.cuh"
SOT_DEFINE_CONFIG(64, 17, 4)
