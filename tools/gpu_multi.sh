#!/bin/bash
# On an N-GPU box (gpurun --gpus N): bitwise check of the band stepper against the whole-grid step (peer mailboxes and
# NCCL), timing of the band schedules, and the bench lines launched the way the driver launches them.
# Usage: tools/gpu_multi.sh TAG N [extra single-GPU variants: 1]
TAG=$1; N=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN tools/multi_gpu_check.py --time > gpurun_out/${TAG}_mgpu${N}.log 2>&1; echo "multi_gpu_check exit $?"
grep -E "bitwise|time |Error|error" gpurun_out/${TAG}_mgpu${N}.log | head -40
timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_g${N}.json 2> gpurun_out/${TAG}_bench_g${N}.err; echo "bench exit $?"
cut -c1-700 gpurun_out/${TAG}_bench_g${N}.json; tail -3 gpurun_out/${TAG}_bench_g${N}.err
GCM_BAND_PEER=0 timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_g${N}_nccl.json 2> gpurun_out/${TAG}_bench_g${N}_nccl.err; echo "bench nccl exit $?"
cut -c1-400 gpurun_out/${TAG}_bench_g${N}_nccl.json
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_g1.json 2> gpurun_out/${TAG}_bench_g1.err; echo "bench 1 exit $?"
cut -c1-400 gpurun_out/${TAG}_bench_g1.json
if [ -n "$3" ]; then
  i=0
  for args in "--knob 4=5" "--knob 4=5 --knob 10=3" "--workload c5b8" "--workload c5b8 --knob 4=5"; do
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
    i=$((i+1))
  done
fi
