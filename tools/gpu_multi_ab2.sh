#!/bin/bash
# 2-GPU proxy of the 8-GPU band size (workload c5q = 1440x180x9): A/B of the overlap modes of the band loop
TAG=$1; N=${2:-2}; WL=${3:-c5q}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $RUN tools/multi_gpu_check.py > gpurun_out/${TAG}_mgpu${N}.log 2>&1; echo "multi_gpu_check exit $?"
grep -E "bitwise|Error|error" gpurun_out/${TAG}_mgpu${N}.log | head -20
for ov in 0 1 2 0 1; do
  GCM_BAND_OVERLAP=$ov timeout 300 $RUN bench.py --gpus $N --workload $WL --steps 20 --warmup 5 --no-hash > gpurun_out/${TAG}_${WL}_g${N}_ov${ov}.json 2> gpurun_out/${TAG}_${WL}_g${N}_ov${ov}.err; echo "bench overlap=$ov exit $?"
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_${WL}_g${N}_ov${ov}.json") if l.startswith("{")][0]
print("overlap=$ov", d["n_gpus"], round(d["ms_per_step"],4), round(d["best_ms_per_step"],4), d["halo_transport"], d["peer_timeouts"], {k[6:]:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
PY
done
python bench.py --workload $WL --steps 20 --warmup 5 --no-cpu-baseline --no-hash > gpurun_out/${TAG}_${WL}_g1.json 2>/dev/null
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_${WL}_g1.json") if l.startswith("{")][0]
print("n1", round(d["ms_per_step"],4), round(d["best_ms_per_step"],4))
PY
