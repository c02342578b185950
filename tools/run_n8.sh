#!/bin/bash
# one 8-GPU call: BASELINE configs[1] strong + weak, configs[2], configs[3], and the A/B switches at N = 8
mkdir -p gpurun_out
N=${1:-8}
TAG=${2:-c11}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"
run() { name=$1; shift; timeout 300 $TR bench.py --gpus $N "$@" > gpurun_out/${TAG}_n${N}_$name.json 2> gpurun_out/${TAG}_n${N}_$name.err; echo "$name exit $?"; }
run cfg2 --steps 50 --warmup 5
run cfg2_nccl --steps 50 --warmup 5 --no-peer
run cfg2_weak --steps 50 --warmup 5 --scaling weak
run cfg3 --steps 30 --warmup 5 --config 3
run cfg4 --steps 20 --warmup 3 --config 4
for f in cfg2 cfg2_nccl cfg2_weak cfg3 cfg4; do python - <<PY
import json
try:
    j = json.load(open('gpurun_out/${TAG}_n${N}_$f.json'))
    p = j['parity_check']
    print('$f', round(j['ms_per_step'], 4), round(j['value']), 'e2e', round(j['e2e']['value']), 'parity', p['ok'], p['loss_rel_err'], p['dx_cos_min'], p.get('update_cos_min'), p['sampled_index_sets_equal'], j['config']['exchange'][:12], j['config']['launch'][:10], j['gpu_launches'] // j['steps'], j['clocks']['sm_mhz'], j['clocks']['reasons'])
except Exception as e:
    print('$f', 'ERR', e)
PY
grep -v "Warning\|warn\|rel = \|NCCL version\|^$" gpurun_out/${TAG}_n${N}_$f.err | tail -2 | cut -c1-300; done
