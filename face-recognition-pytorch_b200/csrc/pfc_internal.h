// Internal constants shared by the .cu translation units (error codes mirror include/pfc.h).
#pragma once
#include <stdint.h>

#define PFC_OK 0
#define PFC_ERR_CUDA (-1)         /* a CUDA runtime call failed */
#define PFC_ERR_LAUNCH (-2)       /* kernel launch failed */
#define PFC_ERR_SHAPE (-3)        /* invalid / unsupported shape argument */
#define PFC_ERR_ALIGNMENT (-4)    /* pointer or stride not 16-byte aligned */
#define PFC_ERR_DRIVER (-5)       /* cuTensorMapEncodeTiled not available */
#define PFC_ERR_TENSORMAP (-6)    /* tensor-map encoding rejected */
#define PFC_ERR_SCALE_RANGE (-7)  /* logit scale s too large for the fixed-shift exponent range */
#define PFC_ERR_WORKSPACE (-8)    /* workspace too small */

// e = 2^(log2e*(z - s) + PFC_EXP_TOP): largest possible term is 2^PFC_EXP_TOP
#define PFC_EXP_TOP 64

#ifdef __cplusplus
// 1: keep the bf16 gradient of the dW GEMM in L2 for the update kernel (PFC_L2_GRAD / pfc_debug_l2_grad, pfc_gemm.cu)
extern "C" int pfc_l2_grad_enabled(void);
#endif
