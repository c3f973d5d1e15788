#!/bin/bash
# N-GPU bench line only (driver-style launch), final code
TAG=$1; N=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_g${N}.json 2> gpurun_out/${TAG}_bench_g${N}.err; echo "bench exit $?"
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_g${N}.json") if l.startswith("{")][0]
print("peer", d["n_gpus"], round(d["ms_per_step"],4), round(d["best_ms_per_step"],4), d["state_sha256"][:16], d["halo_transport"], d["peer_timeouts"], {k[6:]:round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
PY
