#!/bin/bash
# ncu evidence for one bench mode (run on the GPU box, AFTER the same command has exited 0 without ncu):
#   bash tools/prof.sh <tag> <bench args...>
# writes gpurun_out/<tag>_launches.csv (every launch, gpu__time_duration) and gpurun_out/<tag>_full.ncu-rep (--set full
# of the step kernels, one launch each), plus text summaries.  Eager launches (--no-graph): ncu serialises kernels, and the
# lazy update must be ordered before the forward + dX kernel that waits for it.
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --no-graph --no-parity --no-cpu-baseline --steps 3 --warmup 3 "$@" > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 60 \
    -k regex:'fx_kernel|umma_gemm|dw_sgd|row_stats|prepare|finalize|l2norm' -c 14 -o gpurun_out/${tag}_full -f \
    python bench.py --no-graph --no-parity --no-cpu-baseline --steps 3 --warmup 3 "$@" > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${tag}_full_raw.csv > gpurun_out/${tag}_ncu_full_summary.txt 2>&1
tail -5 gpurun_out/${tag}_ncu2.log
