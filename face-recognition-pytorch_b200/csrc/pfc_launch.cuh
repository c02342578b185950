// Programmatic dependent launch (PDL) for the kernels of one head step.
//
// A step is a chain of 8 (one GPU) to 12 (peer exchanges) dependent launches whose small members run 5-10 us, so the
// gap between "last CTA of kernel k exits" and "first CTA of kernel k+1 does useful work" -- launch latency plus the
// prologue of k+1 (tensor-map prefetch, mbarrier init, TMEM allocation, cluster sync) -- is a visible share of the step,
// above all at 4-8 GPUs where the GEMMs themselves shrink to 10-20 us.  With the stream-serialisation attribute the
// next kernel's CTAs are scheduled while the current one still runs and park in `griddepcontrol.wait`, which returns
// once every prerequisite grid has completed and flushed its memory.
//
// Rules every step kernel follows (so that "predecessor complete" stays transitive along the chain):
//   * pdl_wait() is executed by every thread BEFORE the first global-memory access (reads of a predecessor's output,
//     and writes, which could otherwise race with a predecessor still reading the buffer) and before any early return;
//     only CTA-local setup (shared-memory barriers, TMEM allocation, descriptor prefetch) may precede it;
//   * pdl_launch_dependents() follows immediately: the dependent grid may be scheduled as soon as all CTAs of this grid
//     are resident (it then blocks in its own pdl_wait), it can never start its body before this grid has finished.
// Both instructions are no-ops when the kernel was launched without the attribute (pfc_set_pdl(0), or an ordinary
// <<<>>> launch), so the same binaries serve both modes.
//
// Deferred wait (mode 2, GEMM kernels only).  The caller may declare that the next GEMM does not depend on the kernel
// launched just before it (pfc_pdl_independent_next(): dX after dW on one GPU, dW after the dX scatter on several).
// That GEMM then triggers at once and starts working as soon as its CTAs get an SM -- it fills the tail of the
// preceding kernel instead of idling behind it -- and executes griddepcontrol.wait as its LAST instruction, so that
// "this grid complete => predecessor complete" still holds for everything launched after it.  Sound because the
// predecessor itself passed its wait before it triggered: whatever the GEMM reads was complete before it was scheduled.
#pragma once
#include <cuda_runtime.h>

namespace pfc {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// entry sequence of every step kernel
__device__ __forceinline__ void pdl_entry() {
    pdl_wait();
    pdl_launch_dependents();
}

// Which step kernels take the attribute (bit = 1 << id; pfc_debug_pdl_mask / PFC_PDL_MASK, default: see pfc_api.cu).
enum PdlId {
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
bool pdl_enabled();            // pfc_api.cu (pfc_set_pdl / PFC_PDL): mode >= 1
bool pdl_enabled_for(int id);  // mode >= 1 and bit `id` of the mask set
bool pdl_take_independent();   // one-shot flag set by pfc_pdl_independent_next(); true only in mode 2.  Clears it.

// Fills `at` with the stream-serialisation attribute when PDL is on; returns the number of attributes written.
static inline int pdl_attr(cudaLaunchAttribute* at, int id) {
    if (!pdl_enabled_for(id)) return 0;
    at->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at->val.programmaticStreamSerializationAllowed = 1;
    return 1;
}

// <<<grid, block, smem, stream>>> with the PDL attribute.  Only for kernels that start with pdl_entry().
template <class... KArgs, class... Args>
static inline cudaError_t launch_step_kernel(int id, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                             cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    cfg.numAttrs = pdl_attr(at, id);
    cfg.attrs = at;
    (void)pdl_take_independent();   // the hint is for the launch that follows it, and only GEMMs honour it
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace pfc
