mkdir -p gpurun_out
timeout 300 python -m pytest "tests/test_gpu_dropin.py" -q -x -k "True" 2>&1 | grep -E "^E |assert|Error" | head -30
for o in 1 0; do timeout 200 python bench.py --steps 40 --warmup 5 --dw-first $o --no-cpu-baseline > gpurun_out/c14_bench_dwfirst$o.json 2> gpurun_out/c14_bench_dwfirst$o.err; python -c "
import json; j=json.load(open('gpurun_out/c14_bench_dwfirst$o.json')); print('dw_first=$o', j['ms_per_step'], j['value'], j['e2e']['value'], j['step_roofline']['frac'])"; done
ncu --set full --clock-control none -k regex:'roc_kernel|kfold_kernel|pair_score_kernel|acc_kernel|cross_score' -c 8 -o gpurun_out/c14_eval -f python -m pytest tests/test_gpu_kernels.py -q -k "cfg5" > gpurun_out/c14_ncu.log 2>&1
ncu -i gpurun_out/c14_eval.ncu-rep --page raw --csv > gpurun_out/c14_eval_raw.csv 2>/dev/null; rm -f gpurun_out/c14_eval.ncu-rep
python tools/ncu_summary.py gpurun_out/c14_eval_raw.csv --all > gpurun_out/c14_ncu_eval_summary.txt 2>&1
grep -E "^----|gpu__time_duration" gpurun_out/c14_ncu_eval_summary.txt | cut -c1-110
