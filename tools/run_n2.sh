set -x
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -5
timeout 400 python -m pytest tests/test_gpu_head.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2_peer.json 2> gpurun_out/bench_n2_peer.err; tail -2 gpurun_out/bench_n2_peer.err; cat gpurun_out/bench_n2_peer.json
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_tmp.json 2> gpurun_out/err.log; cat gpurun_out/bench_n1_tmp.json
