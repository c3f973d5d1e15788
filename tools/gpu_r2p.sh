#!/bin/bash
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "ensemble or step_host" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for args in "" "--knob 4=6" "--knob 15=400"; do
  timeout 300 python bench.py --workload c4 $args --steps 20 --warmup 5 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_c4.json") if l.startswith("{")][0]
    print("c4 [$args]", round(d["ms_per_step"],4), round(d["best_ms_per_step"],4), {k[6:]:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
except Exception as e:
    print("c4 [$args] no line", e)
PY
  tail -2 gpurun_out/${TAG}_c4.err
done
for args in "--knob 6=4" "--knob 6=6" "--knob 6=8" "--knob 6=10"; do
  timeout 300 python bench.py --workload c5 $args --steps 20 --warmup 3 --repeats 1 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_e2e.json 2> gpurun_out/${TAG}_e2e.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_e2e.json") if l.startswith("{")][0]
    print("[$args] e2e %.4g" % d["e2e"]["value"], "-> GB/s per direction %.1f" % (d["e2e"]["value"]/9331200*306892800/1e9))
except Exception as e:
    print("[$args] no line", e)
PY
done
