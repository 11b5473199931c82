out=gpurun_out/r02_v4; mkdir -p $out
python -m pytest tests -m gpu -q -k "device_record or c4_size or c3_full or node_formats or edge_cases or wavefront_pipelines or portable_trig" > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
for lib in lib lib_s64 lib lib_s64; do
  echo "== $lib C3" >> $out/ab.log
  VCRT_LIB=$PWD/vulkan_compute_ray_tracing_b200/$lib/libvcrt.so timeout 300 python tools/sweep.py --spp 32 --trace 2>&1 | grep Mrays >> $out/ab.log
done
echo "== C4 device build" >> $out/ab.log
timeout 400 python tools/sweep.py --triangles 10000000 --spp 8 --trace fast_build=host,device >> $out/ab.log 2>&1
python bench.py --no-cpu-baseline --no-c4 > $out/bench.json 2> $out/bench.err
ls -la $out
