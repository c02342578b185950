N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_r01c_n${N}.json 2> gpurun_out/bench_r01c_n${N}.err; tail -2 gpurun_out/bench_r01c_n${N}.err; cat gpurun_out/bench_r01c_n${N}.json
