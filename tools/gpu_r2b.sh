#!/bin/bash
# r2b: TMA update kernel: parity on the GPU, A/B against the LDGSTS kernel (knob 4=4)
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "parity or bands or extras" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
i=0
for args in "--workload c5" "--workload c5 --knob 4=4" "--workload c3" "--workload c3 --knob 4=4" "--workload c5b8" "--workload c5b8 --knob 4=4"; do
  timeout 300 python bench.py $args --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_v${i}.json 2> gpurun_out/${TAG}_v${i}.err
  echo "v$i [$args] exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_v${i}.json"))
    print(d["ms_per_step"], d["best_ms_per_step"], d["state_sha256"][:12], d["roofline"]["kernels_ms_per_step"])
except Exception as e:
    print("no line", e)
PY
  tail -2 gpurun_out/${TAG}_v${i}.err
  i=$((i+1))
done
