mkdir -p gpurun_out/v28
bash tools/ab_libs.sh "lib lib_spf lib_spf2 lib lib_spf lib_spf2" --spp 64 --trace 2>&1 | tee gpurun_out/v28/ab.log
