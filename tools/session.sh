out=gpurun_out/r02_v5_n2; mkdir -p $out
nvidia-smi -L > $out/gpus.txt 2>&1
python -m pytest tests -m gpu -q -k "two_gpus or group_api" > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $out/bench_ref_n2.json 2> $out/bench_ref_n2.err ) 2> $out/ref_time.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config c4 --steps 3 --warmup 3 --no-e2e > $out/bench_c4_n2.json 2> $out/bench_c4_n2.err
ls -la $out
