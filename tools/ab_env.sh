#!/bin/bash
# A/B of an environment switch on ONE box, runs interleaved (A B A B ...) so that clock / thermal drift hits both arms:
#   tools/ab_env.sh TAG "ENV_A=.." "ENV_B=.." [pairs]      -> gpurun_out/TAG_{a,b}{i}.json, one summary line per run
cd ${GRAFT_REPO_ROOT:-.}
tag=$1; a=$2; b=$3; pairs=${4:-2}
run() {
  env $2 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-fp32-frames --no-store-e2e --e2e-steps 2 --profile-passes 1 --no-graph-profile > gpurun_out/${tag}_$1.json 2> gpurun_out/${tag}_$1.err
  python - <<PY
import json
for line in open('gpurun_out/${tag}_$1.json'):
    if line.startswith('{'):
        d=json.loads(line); print('$1', '$2', d['value'], d['ms_per_step'], d['gpu_launches'], d['clocks']['sm_mhz'])
PY
}
for i in $(seq 1 $pairs); do
  run a$i "$a"
  run b$i "$b"
done
