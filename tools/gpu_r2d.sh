#!/bin/bash
# r2d: warp-specialised TMA update kernel: parity, A/B of its variants against the LDGSTS kernel, ncu of both (small reports)
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "parity or bands or extras" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
i=0
for args in "" "--knob 4=4" "--knob 10=4" "--knob 13=1" "--knob 13=1 --knob 10=4" "--knob 11=8" "--knob 12=1" "--workload c3" "--workload c3 --knob 4=4" "--workload c5b8" "--workload c5b8 --knob 4=4"; do
  case "$args" in *workload*) wl="";; *) wl="--workload c5";; esac
  timeout 300 python bench.py $wl $args --steps 20 --warmup 5 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_v${i}.json 2> gpurun_out/${TAG}_v${i}.err
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
CMD="python bench.py --workload c5 --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-hash"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pe25f_update -s 10 -c 2 -o gpurun_out/${TAG}_prof_tma $CMD > gpurun_out/${TAG}_ncu_tma.log 2>&1
echo "ncu tma exit $?"
CMD="python bench.py --workload c5 --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-hash --knob 4=4"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pe25f_update -s 10 -c 2 -o gpurun_out/${TAG}_prof_ldgsts $CMD > gpurun_out/${TAG}_ncu_ldgsts.log 2>&1
echo "ncu ldgsts exit $?"
ls -la gpurun_out | tail -30; du -sh gpurun_out
