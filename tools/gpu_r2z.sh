#!/bin/bash
# r2z: predictor variant of the tiled update (own values from the staged tile) vs the plain one (knob 5=9); GPU suite
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for k in "5=0" "5=9" "5=0" "5=9"; do
  timeout 300 python bench.py --workload c5 --knob $k --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_v.json 2> gpurun_out/${TAG}_v.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_v.json") if l.startswith("{")][0]
    print("c5 knob $k", round(d["ms_per_step"],4), round(d["best_ms_per_step"],4), d["state_sha256"][:16], {k[6:]:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
except Exception as e:
    print("c5 knob $k no line", e)
PY
  tail -2 gpurun_out/${TAG}_v.err
done
timeout 200 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_c4.json 2>/dev/null
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_c4.json") if l.startswith("{")][0]
print("c4", round(d["ms_per_step"],4), round(d["best_ms_per_step"],4))
PY
