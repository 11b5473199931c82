#!/bin/bash
# Scratch entry for `gpurun -- 'bash tools/_run.sh'` during development (A/B runs of library builds, one-off ncu passes).
# The round's standard pass is tools/gpu_check.sh <tag> [quick]; A/B of builds: tools/ab_libs.sh "lib lib_x" --spp 32 --trace
out=gpurun_out/r02_v56; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log; tail -3 $out/pytest.log
python bench.py > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_v56/bench.json").read().strip().splitlines()[0])
r=d["roofline"]
print(d["value"], d["ms_per_step"], "frac", r["frac"], "bounce", d["bounce_mrays"], "e2e", d["e2e"]["value"], "frame", d["e2e"]["frame_1spp"]["ms_per_frame"])
print(r.get("timed_as")); print(r.get("timed_region")); print("share", r.get("kernel_share_of_step"))
print("c4", d.get("c4",{}).get("value"), d.get("c4",{}).get("roofline",{}).get("frac"), d.get("c4",{}).get("error"))
print("upload", d["e2e"].get("with_scene_upload",{}).get("ms_per_step"))
PY
