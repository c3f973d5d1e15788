#!/bin/bash
# r2m: final single-GPU evidence: full GPU suite, smoke, bench lines of every workload, stream-priority A/B,
# ncu launch list + full capture (reports summarised on the box, only CSVs come back)
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
for wl in c5 c3 c2 c4 c1 c1dt700 c1big p2d; do
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
  echo "bench $wl exit $?"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_${wl}.json") if l.startswith("{")][0]
    print("$wl", round(d["ms_per_step"],5), round(d["best_ms_per_step"],5), "%.4g"%d["value"], "frac %.4f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], "cpu %.4g"%d["cpu_baseline"]["value"])
except Exception as e:
    print("no line", e)
PY
done
timeout 300 python bench.py --workload c2 --options limit_q,limit_t,viscosity=1e5 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c2_opts.json 2>/dev/null; echo "bench c2 opts exit $?"
timeout 300 python bench.py --workload c5 --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_c5_reference.json 2>/dev/null; echo "reference arm exit $?"
for k in "3=2" "3=0"; do
  timeout 300 python bench.py --workload c5 --knob $k --steps 20 --warmup 5 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_prio_${k}.json 2>/dev/null
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_prio_${k}.json") if l.startswith("{")][0]
print("knob $k", round(d["ms_per_step"],4), round(d["best_ms_per_step"],4), {k[6:]:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
PY
done
CMD="python bench.py --workload c5 --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-hash"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pe25f -s 40 -c 10 -o /tmp/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"
python tools/ncu_summary.py /tmp/${TAG}_prof.ncu-rep gpurun_out/${TAG}_ncu_full_summary.csv > /dev/null 2>&1; echo "summary exit $?"
ls -la gpurun_out | wc -l; du -sh gpurun_out
