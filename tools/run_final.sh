set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|rror"
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r01d_n1.json 2> gpurun_out/err.log; cat gpurun_out/bench_r01d_n1.json
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r01d_ref.json 2>> gpurun_out/err.log; cat gpurun_out/bench_r01d_ref.json
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/prof_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_gemm|dw_sgd_rows" -s 24 -c 4 -o gpurun_out/prof_r01e -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_f.log 2>&1
tail -1 gpurun_out/ncu_f.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
