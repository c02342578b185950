mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q -x > gpurun_out/c9_pytest_dist.log 2>&1; tail -5 gpurun_out/c9_pytest_dist.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/c9_bench_n2.json 2> gpurun_out/c9_bench_n2.err; echo "n2 exit $?"
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-peer > gpurun_out/c9_bench_n2_nccl.json 2> gpurun_out/c9_bench_n2_nccl.err; echo "n2 nccl exit $?"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --config 3 > gpurun_out/c9_bench_n2_cfg3.json 2> gpurun_out/c9_bench_n2_cfg3.err; echo "n2 cfg3 exit $?"
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --scaling weak > gpurun_out/c9_bench_n2_weak.json 2> gpurun_out/c9_bench_n2_weak.err; echo "n2 weak exit $?"
for f in n2 n2_nccl n2_cfg3 n2_weak; do python -c "
import json; j=json.load(open('gpurun_out/c9_bench_$f.json')); print('$f', j['ms_per_step'], j['value'], j['e2e']['value'], j['parity_check'], j['config']['exchange'], j['config']['launch'])"; tail -2 gpurun_out/c9_bench_$f.err | cut -c1-300; done
