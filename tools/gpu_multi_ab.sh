#!/bin/bash
# N-GPU A/B of the band schedule: exchange overlapped with the interior predictor (default) vs exchange first
TAG=$1; N=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $RUN tools/multi_gpu_check.py > gpurun_out/${TAG}_mgpu${N}.log 2>&1; echo "multi_gpu_check exit $?"
grep -E "bitwise|Error|error" gpurun_out/${TAG}_mgpu${N}.log | head -20
for ov in 1 0; do
  GCM_BAND_OVERLAP=$ov timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_g${N}_ov${ov}.json 2> gpurun_out/${TAG}_bench_g${N}_ov${ov}.err; echo "bench overlap=$ov exit $?"
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_g${N}_ov${ov}.json") if l.startswith("{")][0]
print("overlap=$ov", d["n_gpus"], round(d["ms_per_step"],4), round(d["best_ms_per_step"],4), d["state_sha256"][:16], d["halo_transport"], d["peer_timeouts"], {k[6:]:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
PY
done
