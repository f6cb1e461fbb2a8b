#!/bin/bash
# Round-2 scaling lines beyond the driver's weak-scaling run: BASELINE config 4 (32768^2, strong scaling) and
# config 5 (3-D 512^3 slabs), for one GPU count.  Usage (on a GPU box): bash profiles/run_scaling.sh N TAG
N=${1:-1}; TAG=${2:-r2}
run() {  # name, bench args...
  local name=$1; shift
  if [ "$N" -eq 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_n${N}.json 2> gpurun_out/${TAG}_${name}_n${N}.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" \
        > gpurun_out/${TAG}_${name}_n${N}.json 2> gpurun_out/${TAG}_${name}_n${N}.err; fi
  tail -c 300 gpurun_out/${TAG}_${name}_n${N}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${name}_n${N}.json").read().strip().splitlines()[-1])
    print("${name} N=${N}:", round(d["ms_per_step"], 4), "ms/step", round(d["value"], 1), d["unit"], "parity", (d.get("parity") or {}).get("identical"), "mass", d.get("mass"))
except Exception as e:
    print("${name} N=${N}: no line", e)
PY
}
run strong32768 --nx-global 32768 --preroll 200 --steps 20 --warmup 3 --no-e2e --no-cpu --no-general
run dim3_512 --dim 3 --size 512 --steps 20 --warmup 3 --no-e2e --no-cpu
if [ "$N" -gt 1 ]; then run weak8192 --steps 20 --warmup 3 --no-cpu; fi
