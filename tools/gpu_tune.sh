#!/bin/bash
# Run on the GPU box: GPU parity tests, then bench variants (one JSON line each), then optional ncu capture.
# Usage: tools/gpu_tune.sh TAG "variant args 1" "variant args 2" ...   (NCU_KERNEL=regex to add a full capture)
TAG=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
n=0
for v in "$@"; do
  python bench.py --no-cpu-baseline $v > gpurun_out/${TAG}_v${n}.json 2> gpurun_out/${TAG}_v${n}.err
  echo "variant $n [$v] exit $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_v${n}.json"))
    r = d.get("roofline") or {}
    print("  value %.4g  ms/step %.4f  e2e %.4g  whole_frac %.4f  finite %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], (r.get("whole_step") or {}).get("frac", 0), d["finite"]))
    print("  kernels ms/step:", {k: round(v, 4) for k, v in (r.get("kernels_ms_per_step") or {}).items()})
except Exception as e:
    print("  (no json)", e)
PY
  tail -2 gpurun_out/${TAG}_v${n}.err
  n=$((n+1))
done
if [ -n "$NCU_KERNEL" ]; then
  CMD="python bench.py --workload ${NCU_WL:-c5} --steps 2 --warmup 3 --no-cpu-baseline $NCU_ARGS"
  $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$NCU_KERNEL -s ${NCU_SKIP:-20} -c ${NCU_COUNT:-4} -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
  echo "ncu full exit $?"
fi
