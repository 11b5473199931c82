out=gpurun_out/r02_v6; mkdir -p $out
python tools/latency.py --bounces 1,2,3,4,8 > $out/latency_c3.log 2>&1
python tools/latency.py --bundled --bounces 1,2,4,8 > $out/latency_bundled.log 2>&1
python -m pytest tests -m gpu -q -k "device_record or glass_metal_deep or edge_cases or primary_hits or portable_trig" > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
ls -la $out
