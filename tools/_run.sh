mkdir -p gpurun_out/v23
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/v23/gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/v23/bench_c3_n8.json 2> gpurun_out/v23/bench_c3_n8.err
tail -c 900 gpurun_out/v23/bench_c3_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 --config c5 --no-e2e > gpurun_out/v23/bench_c5_n8.json 2> gpurun_out/v23/bench_c5_n8.err
tail -c 600 gpurun_out/v23/bench_c5_n8.json
