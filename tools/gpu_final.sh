#!/bin/bash
# Round-end evidence on the GPU box: tools/gpu_check.sh (tests, smoke, bench on every workload, ncu launch list + full
# capture of the half-step kernels) + the opt-in terms (bench lines, full capture of pe25x_extras_kernel) + reference arm.
# Usage: tools/gpu_final.sh TAG
TAG=${1:-r03}
bash tools/gpu_check.sh $TAG
OPT="--options limit_q,limit_t,viscosity=1e5"
python bench.py --workload c2 --steps 20 --warmup 3 $OPT > gpurun_out/${TAG}_bench_c2_opts.json 2> gpurun_out/${TAG}_bench_c2_opts.err
echo "bench c2 + options exit $?"; cut -c1-400 gpurun_out/${TAG}_bench_c2_opts.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
echo "reference arm exit $?"; cut -c1-300 gpurun_out/${TAG}_ref.json
