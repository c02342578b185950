set -x
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15
for flag in "" "--no-peer"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 $flag > gpurun_out/bench_n2_peer$flag.json 2> gpurun_out/bench_n2_peer$flag.err; tail -3 gpurun_out/bench_n2_peer$flag.err; cat gpurun_out/bench_n2_peer$flag.json
done
