#define VCRT_TU_TRAV 0
#define VCRT_TU_NAME launch_render_reference
#include "vcrt_kernels.inl"
