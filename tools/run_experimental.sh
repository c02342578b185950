# FIRST GPU CALL of the next round: verify and measure everything that was built after this round's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash tools/run_experimental.sh 1 > gpurun_out/experimental_n1.log 2>&1; tail -60 gpurun_out/experimental_n1.log'
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/run_experimental.sh 2 > gpurun_out/experimental_n2.log 2>&1; tail -60 gpurun_out/experimental_n2.log'
# exp_step fields: pdl-mode : pdl-mask : fork(-1 auto) : direct(no autograd) : prio : early_dx : l2_grad
N=${1:-1}
set -x
if [ "$N" = "1" ]; then
  # one process per test function: a trapping kernel (sticky CUDA error) must not take the other checks down
  for t in test_graph_without_autograd_is_bit_identical test_fused_step_eager_matches_autograd \
           test_high_priority_side_stream_changes_nothing test_adamw_sampled_fused_matches_unfused_and_reference \
           test_head_with_interclass_filter_matches_reference test_early_dx_matches_the_serial_order \
           test_early_dx_graph_replay test_l2_resident_gradient_is_bit_identical test_fused_sampler_matches_the_oracle; do
    PFC_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_experimental.py -q -k $t 2>&1 | grep -E "passed|failed|rror" | tail -2 | sed "s/^/$t: /"
  done
  timeout 200 python -m pytest tests/test_gpu_z_cfg1.py -q 2>&1 | tail -2
  timeout 300 python tools/exp_step.py --configs 0:0:-1:0:0:0:0,0:0:-1:1:0:0:0,0:0:-1:0:0:1:0,0:0:-1:1:0:1:0,0:0:-1:0:0:0:1,0:0:-1:1:0:1:1,0:0:1:1:0:1:1,0:0:-1:0:0:0:0 2>&1 | grep -E "^mode|rror"
  timeout 120 python tools/check_pick.py 2>&1 | grep -E "PARALLEL|pfc_sample"     # incl. the one-launch sampler's time
else
  PFC_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_dist.py -q -k experimental 2>&1 | tail -8
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 \
    tools/exp_step.py --configs 0:0:-1:0:0:0:0,0:0:-1:1:0:0:0,0:0:-1:0:1:0:0,0:0:-1:0:0:1:0,0:0:-1:1:0:1:0,0:0:-1:1:1:1:0,0:0:-1:1:1:1:1,0:0:-1:0:0:0:0 \
    2>&1 | grep -E "^mode|rror"
fi
