#!/bin/bash
# r2t: aflux fused into the filter (knob 16=1) and the tile form of the hydro kernel (knob 7=2) against the defaults
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "parity or bands or extras" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
i=0
for wl in c5 c5b8 c3 c4; do
  for k in "" "--knob 16=1" "--knob 7=2" "--knob 7=2 --knob 15=30" "--knob 16=1 --knob 7=2"; do
    if [ "$wl" = "c4" ] && [ "$k" != "" ] && [ "$k" != "--knob 16=1" ]; then continue; fi
    timeout 300 python bench.py --workload $wl $k --steps 20 --warmup 5 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_v${i}.json 2> gpurun_out/${TAG}_v${i}.err
    python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_v${i}.json") if l.startswith("{")][0]
    print("$wl [$k]", round(d["ms_per_step"],4), round(d["best_ms_per_step"],4), {k[6:]:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
except Exception as e:
    print("$wl [$k] no line", e)
PY
    tail -2 gpurun_out/${TAG}_v${i}.err
    i=$((i+1))
  done
done
