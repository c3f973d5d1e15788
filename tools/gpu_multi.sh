#!/bin/bash
# On an N-GPU box (gpurun --gpus N): bitwise check of the band stepper against the whole-grid step, timing of the band
# schedules, and the bench line launched the way the driver launches it.  Usage: tools/gpu_multi.sh TAG N
TAG=$1; N=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $RUN tools/multi_gpu_check.py --time > gpurun_out/${TAG}_mgpu${N}.log 2>&1; echo "multi_gpu_check exit $?"
grep -E "bitwise|time " gpurun_out/${TAG}_mgpu${N}.log
timeout 300 $RUN bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_g${N}.json 2> gpurun_out/${TAG}_bench_g${N}.err; echo "bench exit $?"
cut -c1-330 gpurun_out/${TAG}_bench_g${N}.json
python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_g1.json 2> gpurun_out/${TAG}_bench_g1.err; echo "bench 1 exit $?"
cut -c1-330 gpurun_out/${TAG}_bench_g1.json
