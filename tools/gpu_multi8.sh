#!/bin/bash
# N-GPU validation, lean: bitwise check of the band schedules (peer mailboxes and NCCL) + the driver-style bench lines
TAG=$1; N=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $RUN tools/multi_gpu_check.py > gpurun_out/${TAG}_mgpu${N}.log 2>&1; echo "multi_gpu_check exit $?"
grep -E "bitwise|Error|error" gpurun_out/${TAG}_mgpu${N}.log | head -20
timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_g${N}.json 2> gpurun_out/${TAG}_bench_g${N}.err; echo "bench exit $?"
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_g${N}.json") if l.startswith("{")][0]
print("peer", d["n_gpus"], d["ms_per_step"], d["best_ms_per_step"], d["state_sha256"][:16], d["halo_transport"], d["peer_timeouts"], d["roofline"]["kernels_ms_per_step"], d["roofline"].get("halo_exchange_ms_per_step"))
PY
tail -3 gpurun_out/${TAG}_bench_g${N}.err
GCM_BAND_PEER=0 timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_g${N}_nccl.json 2> gpurun_out/${TAG}_bench_g${N}_nccl.err; echo "bench nccl exit $?"
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_g${N}_nccl.json") if l.startswith("{")][0]
print("nccl", d["n_gpus"], d["ms_per_step"], d["best_ms_per_step"], d["state_sha256"][:16], d["halo_transport"], d["roofline"].get("halo_exchange_ms_per_step"))
PY
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_g1.json 2> gpurun_out/${TAG}_bench_g1.err; echo "bench 1 exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_g1.json"))
print("n1", d["ms_per_step"], d["best_ms_per_step"], d["state_sha256"][:16])
PY
