cd $GRAFT_REPO_ROOT
run() { # name, extra env
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 4 --steps 100 --warmup 5 --no-cpu-baseline --no-fp32-frames --no-store-e2e --e2e-steps 2 --profile-passes 1 --no-graph-profile > gpurun_out/r02w_n4_$1.json 2> gpurun_out/r02w_n4_$1.err
  python - <<PY
import json
for line in open('gpurun_out/r02w_n4_$1.json'):
    if line.startswith('{'):
        d=json.loads(line); print('$1', d['value'], d['ms_per_step'])
PY
}
run default "X=1" 29541
run ctas16 "NCCL_MAX_CTAS=16" 29542
run ctas8 "NCCL_MAX_CTAS=8" 29543
run ctas4 "NCCL_MAX_CTAS=4" 29544
