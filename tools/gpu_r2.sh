#!/bin/bash
# Round-2 GPU check (via gpurun): parity tests, smoke, bench lines (2.5-D and 2-D workloads), optional ncu passes.
# Usage: tools/gpu_r2.sh TAG [tests|notests] [ncu|noncu] [extra bench args for the c5 line]
TAG=${1:-r2a}; TESTS=${2:-tests}; NCU=${3:-noncu}; shift 3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ "$TESTS" = "tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
fi
timeout 600 python bench.py --workload c5 --steps 20 --warmup 5 "$@" > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
echo "bench c5 exit $?"; cat gpurun_out/${TAG}_bench_c5.json; tail -3 gpurun_out/${TAG}_bench_c5.err
for wl in c3 c2 c4 c1 c1big p2d; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
  echo "bench $wl exit $?"; cut -c1-600 gpurun_out/${TAG}_bench_${wl}.json; tail -3 gpurun_out/${TAG}_bench_${wl}.err
done
if [ "$NCU" = "ncu" ]; then
  CMD="python bench.py --workload c5 --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-hash"
  $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
  echo "ncu launches exit $?"
  $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pe25f -s 40 -c 10 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
  echo "ncu full exit $?"
fi
