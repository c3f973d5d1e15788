#!/bin/bash
TAG=${1:-r2n}
mkdir -p gpurun_out
python tools/pcie_probe.py > gpurun_out/${TAG}_pcie.log 2>&1; cat gpurun_out/${TAG}_pcie.log
for k in 8 4 16 32; do
  timeout 300 python bench.py --workload c5 --knob 6=$k --steps 10 --warmup 3 --repeats 1 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_blocks${k}.json 2>/dev/null
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_blocks${k}.json") if l.startswith("{")][0]
print("blocks $k e2e %.4g" % d["e2e"]["value"], "-> GB/s per direction %.1f" % (d["e2e"]["value"]/9331200*306892800/1e9))
PY
done
for wl in c1 c1big p2d; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_${wl}.json 2>/dev/null
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_${wl}.json") if l.startswith("{")][0]
print("$wl", round(d["ms_per_step"],5), "%.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"])
PY
done
for k in "2=2" "2=4" "2=0"; do
  timeout 300 python bench.py --workload c5b8 --knob $k --steps 20 --warmup 5 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_rg_${k}.json 2>/dev/null
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_rg_${k}.json") if l.startswith("{")][0]
print("c5b8 knob $k", round(d["ms_per_step"],4), {k[6:]:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
PY
done
