#define VCRT_TU_TRAV 1
#define VCRT_TU_NAME launch_render_fast
#include "vcrt_kernels.inl"
