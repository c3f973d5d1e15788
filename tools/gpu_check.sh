#!/bin/bash
# Run on the GPU box (via gpurun): parity tests, smoke, bench on every workload, ncu launch list + full capture.
# Usage: tools/gpu_check.sh [tag] [skip_ncu]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
for wl in c5 c3 c2 c4; do
  python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
  echo "bench $wl exit $?"; cat gpurun_out/${TAG}_bench_${wl}.json; tail -3 gpurun_out/${TAG}_bench_${wl}.err
done
if [ -z "$2" ]; then
  CMD="python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
  echo "ncu launches exit $?"
  $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:pe25f -s 40 -c 10 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
  echo "ncu full exit $?"
fi
