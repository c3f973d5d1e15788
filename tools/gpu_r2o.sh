#!/bin/bash
TAG=${1:-r2o}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "step_host" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
for args in "" "--e2e-serial" "--knob 6=16" "--knob 6=32" "--knob 6=12"; do
  timeout 300 python bench.py --workload c5 $args --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_e2e.json 2> gpurun_out/${TAG}_e2e.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_e2e.json") if l.startswith("{")][0]
    print("[$args] e2e %.4g" % d["e2e"]["value"], "-> GB/s per direction %.1f" % (d["e2e"]["value"]/9331200*306892800/1e9), "| resident ms/step", round(d["ms_per_step"],4))
except Exception as e:
    print("[$args] no line", e)
PY
  tail -3 gpurun_out/${TAG}_e2e.err
done
