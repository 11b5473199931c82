#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list + one full capture of the trace kernel.
# usage (from the repo root, on the GPU box): bash tools/gpu_check.sh <tag> [quick]
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
python bench.py > $out/bench.json 2> $out/bench.err
timeout 300 python tools/sweep.py --spp 16 --count --trace >> $out/ab.log 2>&1
if [ "$2" != "quick" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_trace_kernel --launch-skip 8 -c 3 -o $out/trace_full -f python tools/sweep.py --spp 8 --reps 1 > $out/ncu_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wf_trace_kernel --csv --log-file $out/traffic.csv python tools/sweep.py --spp 64 --reps 0 > $out/traffic.log 2>&1
fi
ls -la $out
