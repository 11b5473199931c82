#define VCRT_TU_TRAV 2
#define VCRT_TU_NAME launch_render_brute
#include "vcrt_kernels.inl"
