mkdir -p gpurun_out/v11
python -m pytest tests -m gpu -x -q > gpurun_out/v11/pytest.log 2>&1; tail -3 gpurun_out/v11/pytest.log
bash tools/ab_libs.sh "lib lib_nohint" --spp 16 --trace 2>&1 | tee gpurun_out/v11/ab.log
bash tools/ab_libs.sh "lib lib_nohint" --spp 64 --trace 2>&1 | tee -a gpurun_out/v11/ab.log
