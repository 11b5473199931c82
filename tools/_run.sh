mkdir -p gpurun_out/v15
bash tools/ab_libs.sh "lib_w1 lib_w1p9 lib_w1p8 lib_w1r lib_w1r12" --spp 32 --trace 2>&1 | tee gpurun_out/v15/ab.log
VCRT_LIB=$PWD/vulkan_compute_ray_tracing_b200/lib_w1/libvcrt.so python tools/sweep.py --spp 32 --trace leaf_threshold=4,6,8,12 shade_threshold=4,8,12 continue_threshold=20,26 2>&1 | tee gpurun_out/v15/sweep.log
