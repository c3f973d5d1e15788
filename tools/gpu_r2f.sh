#!/bin/bash
# r2f: persistent pipelined filter kernel (bulk-copy prefetch) A/B against the one-unit-per-CTA kernel (knob 14=1)
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "parity or bands" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
i=0
for args in "--workload c5" "--workload c5 --knob 14=1" "--workload c5b8" "--workload c5b8 --knob 14=1" "--workload c3" "--workload c3 --knob 14=1" "--workload c2" "--workload c2 --knob 14=1" "--workload c4" "--workload c4 --knob 14=1"; do
  timeout 300 python bench.py $args --steps 20 --warmup 5 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_v${i}.json 2> gpurun_out/${TAG}_v${i}.err
  echo "v$i [$args] exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_v${i}.json"))
    print(round(d["ms_per_step"],4), round(d["best_ms_per_step"],4), {k[6:]:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
except Exception as e:
    print("no line", e)
PY
  tail -2 gpurun_out/${TAG}_v${i}.err
  i=$((i+1))
done
