#!/bin/bash
# A/B of library builds (make OUT=../lib_x EXTRA=-D...) on the bench workload; usage: tools/ab_libs.sh "lib lib_x ..." [sweep args]
libs=$1; shift
for lib in $libs; do
  echo "== $lib"; VCRT_LIB=$PWD/vulkan_compute_ray_tracing_b200/$lib/libvcrt.so timeout 300 python tools/sweep.py "$@" 2>&1 | grep Mrays
done
