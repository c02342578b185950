// Library identity and error strings (include/pfc.h).
#include <stdlib.h>
#include "pfc_internal.h"

extern "C" {

int pfc_version(void) { return 200; }   // 0.2.0

const char* pfc_error_string(int code) {
    switch (code) {
        case PFC_OK: return "ok";
        case PFC_ERR_CUDA: return "CUDA runtime call failed";
        case PFC_ERR_LAUNCH: return "kernel launch failed";
        case PFC_ERR_SHAPE: return "invalid or unsupported shape";
        case PFC_ERR_ALIGNMENT: return "pointer or stride not 16-byte aligned";
        case PFC_ERR_DRIVER: return "cuTensorMapEncodeTiled not available";
        case PFC_ERR_TENSORMAP: return "tensor map encoding rejected";
        case PFC_ERR_SCALE_RANGE: return "logit scale outside the fixed-shift exponent range";
        case PFC_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown error";
    }
}

}  // extern "C"
