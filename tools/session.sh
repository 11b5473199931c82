out=gpurun_out/r02_v9; mkdir -p $out
python bench.py --config c4 --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_c4_n1.json 2> $out/bench_c4_n1.err
ls -la $out
