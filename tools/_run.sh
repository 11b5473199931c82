mkdir -p gpurun_out/v26
python tools/sweep.py --spp 64 --reps 3 --trace wf_batch_paths=33554432,67108864,140000000 2>&1 | tee gpurun_out/v26/batch.log
