mkdir -p gpurun_out/v29
python tools/sweep.py --triangles 10000000 --spp 8 --trace --count fast_nodes=q15x4,q15 2>&1 | tee gpurun_out/v29/c4_sweep.log
python tools/sweep.py --triangles 10000000 --spp 8 --trace 2>&1 | tee -a gpurun_out/v29/c4_sweep.log
ncu --set full --clock-control none -k regex:wf_trace_kernel --launch-skip 8 -c 2 -o gpurun_out/v29/trace_c4 -f python tools/sweep.py --triangles 10000000 --spp 8 --reps 1 > gpurun_out/v29/ncu_c4.log 2>&1
tail -2 gpurun_out/v29/ncu_c4.log
