set -x
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/prof_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_gemm|dw_sgd_rows" -s 24 -c 4 -o gpurun_out/prof_r01b -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm_modes.py -m gpu -x -q > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer_memcheck.log
