mkdir -p gpurun_out/v19
bash tools/ab_libs.sh "lib lib_sh6 lib_sh8 lib_sh4g20 lib_sh5g20 lib_sh6g24" --spp 32 --trace 2>&1 | tee gpurun_out/v19/ab.log
