set -x
N=${1:-8}
for flag in "" "--no-peer"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 30 --warmup 5 $flag > gpurun_out/bench_n${N}_peer$flag.json 2> gpurun_out/bench_n${N}_peer$flag.err; tail -2 gpurun_out/bench_n${N}_peer$flag.err; cat gpurun_out/bench_n${N}_peer$flag.json
done
