#!/bin/bash
# ncu evidence for the bench step and the kernels around it (run on the GPU box, AFTER the same commands have exited 0
# without ncu):      bash tools/prof.sh <tag> [bench args...]
# writes gpurun_out/<tag>_launches.csv (every launch of three eager steps, gpu__time_duration), gpurun_out/<tag>_full.ncu-rep
# (--set full, one launch of each step kernel) and gpurun_out/<tag>_aux.ncu-rep (sampler, row moves, normalise, scorer from
# tools/bench_aux.py), plus text summaries of both.  Eager launches (--no-graph): ncu serialises kernels anyway.
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --no-graph --no-parity --no-cpu-baseline --steps 3 --warmup 3 "$@" > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 60 \
    -k regex:'umma_gemm|dw_sgd|row_stats|prepare|finalize|l2norm' -c 12 -o gpurun_out/${tag}_full -f \
    python bench.py --no-graph --no-parity --no-cpu-baseline --steps 3 --warmup 3 "$@" > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full.ncu-rep --page details --csv > gpurun_out/${tag}_full_details.csv 2>/dev/null
rm -f gpurun_out/${tag}_full.ncu-rep        # gpurun_out/ travels back only below 64 MiB: keep the CSV exports
python tools/ncu_summary.py gpurun_out/${tag}_full_raw.csv > gpurun_out/${tag}_ncu_full_summary.txt 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'hist_kernel|count_kernel|compact_kernel|mark_positive|roc_kernel|kfold_kernel|pair_score_kernel|acc_kernel|move_rows|l2norm_rows_kernel|dw_sgd_rows' \
    -c 60 -o gpurun_out/${tag}_aux -f python tools/bench_aux.py --iters 1 --out gpurun_out/${tag}_aux_under_ncu.json > gpurun_out/${tag}_ncu3.log 2>&1
ncu -i gpurun_out/${tag}_aux.ncu-rep --page raw --csv > gpurun_out/${tag}_aux_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_aux.ncu-rep
python tools/ncu_summary.py gpurun_out/${tag}_aux_raw.csv --all > gpurun_out/${tag}_ncu_aux_summary.txt 2>&1
tail -n 3 gpurun_out/${tag}_ncu2.log; tail -n 3 gpurun_out/${tag}_ncu3.log; du -sh gpurun_out
