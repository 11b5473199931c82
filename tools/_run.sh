mkdir -p gpurun_out/v22
(./tools/ubench/gather_tex 20 64; ./tools/ubench/gather_tex 16 64; ./tools/ubench/gather_tex 12 64) 2>&1 | tee gpurun_out/v22/gather_tex.log
