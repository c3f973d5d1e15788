#!/bin/bash
# Round-end evidence on the GPU box: tools/gpu_check.sh (tests, smoke, bench on every workload, ncu launch list + full
# capture of the half-step kernels) + the opt-in terms (bench lines, full capture of pe25x_extras_kernel) + reference arm.
# Usage: tools/gpu_final.sh TAG
TAG=${1:-r03}
bash tools/gpu_check.sh $TAG
OPT="--options limit_q,limit_t,viscosity=1e5"
python bench.py --workload c2 --steps 20 --warmup 3 $OPT > gpurun_out/${TAG}_bench_c2_opts.json 2> gpurun_out/${TAG}_bench_c2_opts.err
echo "bench c2 + options exit $?"; cut -c1-400 gpurun_out/${TAG}_bench_c2_opts.json
OPT5="--options coriolis,limit_q,limit_t,viscosity=10"
python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu-baseline $OPT5 > gpurun_out/${TAG}_bench_c5_opts.json 2> gpurun_out/${TAG}_bench_c5_opts.err
echo "bench c5 + options exit $?"; cut -c1-400 gpurun_out/${TAG}_bench_c5_opts.json
CMD="python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline $OPT5"
ncu --set full --clock-control none --import-source on -k regex:pe25x -s 6 -c 2 -o gpurun_out/${TAG}_prof_extras $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu extras exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
echo "reference arm exit $?"; cut -c1-300 gpurun_out/${TAG}_ref.json
