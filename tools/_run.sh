mkdir -p gpurun_out/v24
python bench.py --config c2 > gpurun_out/v24/bench_c2.json 2> gpurun_out/v24/bench_c2.err; tail -c 1500 gpurun_out/v24/bench_c2.json; tail -3 gpurun_out/v24/bench_c2.err
python bench.py --no-cpu-baseline > gpurun_out/v24/bench_c3.json 2> gpurun_out/v24/bench_c3.err; tail -c 1200 gpurun_out/v24/bench_c3.json; tail -3 gpurun_out/v24/bench_c3.err
