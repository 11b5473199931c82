#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, A/B of library builds, ncu launch list + full captures of the trace kernel.
# usage (from the repo root, on the GPU box): bash tools/gpu_check.sh <tag> [quick] ; A/B libs: AB_LIBS="lib lib_x" (built with make OUT=../lib_x EXTRA=-D...)
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu.txt 2>&1
python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
python bench.py > $out/bench.json 2> $out/bench.err
python tools/latency.py > $out/latency.log 2>&1
python tools/latency.py --bundled >> $out/latency.log 2>&1
for lib in ${AB_LIBS:-lib}; do
  echo "== $lib C3" >> $out/ab.log
  VCRT_LIB=$PWD/vulkan_compute_ray_tracing_b200/$lib/libvcrt.so timeout 300 python tools/sweep.py --spp 16 --trace $AB_ARGS 2>&1 | grep Mrays >> $out/ab.log
  echo "== $lib C4" >> $out/ab.log
  VCRT_LIB=$PWD/vulkan_compute_ray_tracing_b200/$lib/libvcrt.so timeout 400 python tools/sweep.py --triangles 10000000 --spp 8 --trace $AB_ARGS 2>&1 | grep Mrays >> $out/ab.log
done
if [ "$2" != "quick" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-c4 > $out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_trace_kernel --launch-skip 8 -c 3 -o $out/trace_full -f python tools/sweep.py --spp 8 --reps 1 > $out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_trace_kernel --launch-skip 9 -c 2 -o $out/trace_c4 -f python tools/sweep.py --triangles 10000000 --spp 8 --reps 1 > $out/ncu_c4.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wf_trace_kernel --csv --log-file $out/traffic.csv python tools/sweep.py --spp 64 --reps 0 wf_streams=1 > $out/traffic.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wf_trace_kernel --csv --log-file $out/traffic_c4.csv python tools/sweep.py --triangles 10000000 --spp 8 --reps 0 wf_streams=1 > $out/traffic_c4.log 2>&1
fi
ls -la $out
