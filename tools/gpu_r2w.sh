#!/bin/bash
TAG=${1:-r2w}
mkdir -p gpurun_out
for k in 0 1 2 3 0; do
  timeout 300 python bench.py --workload c5 --knob 17=$k --steps 20 --warmup 5 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_v.json 2> gpurun_out/${TAG}_v.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_v.json") if l.startswith("{")][0]
    print("c5 knob17=$k", round(d["ms_per_step"],4), round(d["best_ms_per_step"],4), {k[6:]:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
except Exception as e:
    print("c5 knob17=$k no line", e)
PY
  tail -2 gpurun_out/${TAG}_v.err
done
