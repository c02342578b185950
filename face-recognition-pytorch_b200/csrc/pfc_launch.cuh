// Launch helper shared by the row / peer kernels.
//
// Round 1 carried programmatic dependent launch (griddepcontrol.wait / launch_dependents + the stream-serialisation
// attribute) through every step kernel.  Measured on B200 (profiles/r01d_exp_pdl.txt): 0.402 ms against 0.380 ms per
// graph-replayed step on one GPU, 0.282 against 0.274 ms on two -- every programmatic edge cost 2-5 us more than the plain
// graph edge it replaced -- so it was removed.  The `id` argument survives as documentation of the step position.
#pragma once
#include <cuda_runtime.h>

namespace pfc {

enum StepKernelId {
    PDL_NORMALISE = 0,   // l2norm_rows, peer_l2norm_gather
    PDL_FORWARD = 1,     // forward GEMM
    PDL_STATS = 2,       // row_stats(_loss), loss, peer_row_stats, peer_loss
    PDL_PREPARE = 3,     // backward_prepare
    PDL_DW = 4,          // dW GEMM
    PDL_DX = 5,          // dX GEMM
    PDL_DX_FINAL = 6,    // dx_finalize, peer_dx_scatter, peer_dx_finalize
    PDL_UPDATE = 7,      // dw_sgd_rows, dw_finalize (normalise-backward / SGD / AdamW rows)
    PDL_LABELS = 8,      // localize_labels, peer_localize_labels, peer_barrier
};

template <class... KArgs, class... Args>
static inline cudaError_t launch_step_kernel(int /*id*/, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                             cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.numAttrs = 0;
    cfg.attrs = nullptr;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace pfc
