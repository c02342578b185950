// Library identity and error strings (include/pfc.h).
#include <stdlib.h>
#include "pfc_internal.h"
#include "pfc_launch.cuh"

namespace pfc {

// Programmatic dependent launch of the step kernels (pfc_launch.cuh): -1 = not decided yet (PFC_PDL in the
// environment, default PFC_PDL_DEFAULT), 0 = off, 1 = on, 2 = on + deferred waits for GEMMs the caller declares
// independent of their predecessor.  Read on every launch; plain ints because all launches of a process come from the
// one thread that owns the stream (threading).
#ifndef PFC_PDL_DEFAULT
#define PFC_PDL_DEFAULT 0
#endif
static int g_pdl = -1;
static int g_independent_next = 0;

static int pdl_mode() {
    if (g_pdl < 0) {
        const char* e = getenv("PFC_PDL");
        g_pdl = e ? atoi(e) : PFC_PDL_DEFAULT;
        if (g_pdl < 0) g_pdl = 0;
        if (g_pdl > 2) g_pdl = 2;
    }
    return g_pdl;
}

bool pdl_enabled() { return pdl_mode() >= 1; }

#ifndef PFC_PDL_MASK_DEFAULT
#define PFC_PDL_MASK_DEFAULT 0x1ff
#endif
static int g_pdl_mask = -1;
bool pdl_enabled_for(int id) {
    if (g_pdl_mask < 0) {
        const char* e = getenv("PFC_PDL_MASK");
        g_pdl_mask = e ? (int)strtol(e, nullptr, 0) & 0xffff : PFC_PDL_MASK_DEFAULT;
    }
    return pdl_mode() >= 1 && ((g_pdl_mask >> id) & 1);
}

bool pdl_take_independent() {
    const bool v = g_independent_next != 0 && pdl_mode() >= 2;
    g_independent_next = 0;
    return v;
}

}  // namespace pfc

extern "C" {

void pfc_set_pdl(int mode) { pfc::g_pdl = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }
int pfc_get_pdl(void) { return pfc::pdl_mode(); }
void pfc_pdl_independent_next(void) { pfc::g_independent_next = 1; }
// not part of the public header: which step kernels take the PDL attribute (bit = 1 << PdlId, pfc_launch.cuh)
void pfc_debug_pdl_mask(unsigned mask) { pfc::g_pdl_mask = (int)(mask & 0xffff); }

int pfc_version(void) { return 100; }   // 0.1.0

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
