#!/bin/bash
# Scratch entry for `gpurun -- 'bash tools/_run.sh'` during development (A/B runs of library builds, one-off ncu passes).
# The round's standard pass is tools/gpu_check.sh <tag> [quick]; A/B of builds: tools/ab_libs.sh "lib lib_x" --spp 32 --trace
bash tools/gpu_check.sh dev quick
