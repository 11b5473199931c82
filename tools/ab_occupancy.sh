for lib in lib lib_b128_m8 lib_b128_m10 lib_b256_m4 lib_b64_m16 lib_b64_m12; do
  echo "== $lib"; VCRT_LIB=$PWD/vulkan_compute_ray_tracing_b200/$lib/libvcrt.so timeout 200 python tools/sweep.py 2>&1 | tail -1
done
