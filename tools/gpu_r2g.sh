#!/bin/bash
# r2g: register budgets (knob 15): filter / hydro / tiled update at more resident CTAs per SM
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "benchmark_grids or fused_filter" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
i=0
for k in 0 5 6 8 50 60 80 500 600 56 66 566 656; do
  for wl in c5 c5b8; do
    if [ "$wl" = "c5b8" ] && [ $k -ne 0 ] && [ $k -ne 6 ] && [ $k -ne 66 ] && [ $k -ne 566 ]; then continue; fi
    timeout 300 python bench.py --workload $wl --knob 15=$k --steps 20 --warmup 5 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_v${i}.json 2> gpurun_out/${TAG}_v${i}.err
    echo "v$i [$wl knob15=$k] exit $?"; python - <<PY
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
done
