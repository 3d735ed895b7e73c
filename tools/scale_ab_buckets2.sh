cd $GRAFT_REPO_ROOT
run() {
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 4 --steps 100 --warmup 5 --no-cpu-baseline --no-fp32-frames --no-store-e2e --e2e-steps 2 --profile-passes 1 --no-graph-profile > gpurun_out/r02ac_n4_$1.json 2> gpurun_out/r02ac_n4_$1.err
  python - <<PY
import json
for line in open('gpurun_out/r02ac_n4_$1.json'):
    if line.startswith('{'):
        d=json.loads(line); print('$1', d['value'], d['ms_per_step'])
PY
}
run mb256 "HULC2_BUCKET_MB=256" 29541
run mb150 "HULC2_BUCKET_MB=150" 29542
run mb130 "HULC2_BUCKET_MB=130" 29543
run mb256b "HULC2_BUCKET_MB=256" 29544
