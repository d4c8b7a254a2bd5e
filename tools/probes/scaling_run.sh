#!/bin/bash
# usage: scaling_run.sh N  -- the weak bench and the strong configs[2] run at N GPUs of one box
N=$1
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
[ "$N" = "1" ] && L="python"
timeout 900 $L bench.py --gpus $N --steps 3 --warmup 3 --skip-cpu-baseline --e2e-steps 3 > gpurun_out/scale_weak_n$N.json 2> gpurun_out/scale_weak_n$N.err; echo "weak rc=$?"
timeout 900 $L bench.py --gpus $N --steps 3 --warmup 3 --skip-cpu-baseline --skip-e2e --strong > gpurun_out/scale_strong_n$N.json 2> gpurun_out/scale_strong_n$N.err; echo "strong rc=$?"
python - <<PY
import json
for kind in ("weak","strong"):
    try:
        j=json.loads(open("gpurun_out/scale_%s_n$N.json" % kind).read().strip().splitlines()[-1])
    except Exception as e:
        print(kind, "no line", e); continue
    print(kind, "N", j["n_gpus"], "value", round(j["value"],1), "ms", round(j["ms_per_step"],2), "e2e", (j.get("e2e") or {}).get("value"))
    for c in j.get("configs",[]): print("   ", c["name"][:50], round(c["value"],1), "GB/s", round(c["ms_per_step"],2), "ms", c["k1_kernel"], [round(x,1) for x in c["ms_per_step_per_rank"]])
PY
