#!/bin/bash
# ncu evidence of the final build's no-autograd step (the kernel sequence GraphedHeadStep replays) on one GPU, after the
# same commands have exited 0 without ncu:      bash tools/prof_final.sh <tag>
#   <tag>_ncu_launches.csv        every launch of three eager cfg-2 steps (gpu__time_duration)
#   <tag>_ncu_full_summary.txt    --set full, one launch of each cfg-2 step kernel (incl. row_stats_loss_prepare)
#   <tag>_ncu_cfg3_summary.txt    --set full of the sampled step at the cfg-3 shape on one GPU: sample_cluster_kernel,
#                                 l2norm_rows / dw_sgd_rows through the index list, the GEMMs at the sampled shape
tag=$1; shift
mkdir -p gpurun_out
B="python bench.py --no-graph --no-parity --no-cpu-baseline --steps 3 --warmup 3"
$B > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain cfg-2 run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
$B --config 3 > gpurun_out/${tag}_plain3.log 2>&1 || { echo "plain cfg-3 run failed"; tail -5 gpurun_out/${tag}_plain3.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_ncu_launches.csv \
    $B > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 40 \
    -k regex:'umma_gemm|dw_sgd|row_stats|finalize|l2norm' -c 8 -o gpurun_out/${tag}_full -f $B > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_full.ncu-rep
python tools/ncu_summary.py gpurun_out/${tag}_full_raw.csv > gpurun_out/${tag}_ncu_full_summary.txt 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 40 \
    -k regex:'sample_|umma_gemm|dw_sgd|row_stats|finalize|l2norm' -c 9 -o gpurun_out/${tag}_cfg3 -f $B --config 3 > gpurun_out/${tag}_ncu3.log 2>&1
ncu -i gpurun_out/${tag}_cfg3.ncu-rep --page raw --csv > gpurun_out/${tag}_cfg3_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_cfg3.ncu-rep
python tools/ncu_summary.py gpurun_out/${tag}_cfg3_raw.csv > gpurun_out/${tag}_ncu_cfg3_summary.txt 2>&1
rm -f gpurun_out/${tag}_full_raw.csv gpurun_out/${tag}_cfg3_raw.csv
tail -n 2 gpurun_out/${tag}_ncu2.log; tail -n 2 gpurun_out/${tag}_ncu3.log
grep -c "^----" gpurun_out/${tag}_ncu_full_summary.txt gpurun_out/${tag}_ncu_cfg3_summary.txt; du -sh gpurun_out
