"""ctypes binding of libpfc_b200.so -- the only way the Python host code reaches the CUDA kernels.

There is no fallback: if the library is missing and cannot be built, importing this module raises.
Signatures mirror include/pfc.h one to one.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))


class PfcError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        super().__init__(f"{where}: libpfc_b200 error {code} ({error_string(code)})")


def _load():
    path = _build.LIB
    if os.environ.get("PFC_REBUILD") == "1" or _build.needs_build():
        # missing (fresh checkout) or older than a source under csrc/: compile in-tree.  build_locked() serialises the ranks
        # of one node behind a file lock and re-checks staleness once it holds it, so only the first one compiles.
        # Without nvcc a missing library is fatal by design; a merely stale one is loaded with a warning.
        try:
            _build.build_locked(force=os.environ.get("PFC_REBUILD") == "1")
        except Exception:
            if not os.path.exists(path):
                raise
            import warnings
            warnings.warn("libpfc_b200.so is older than its sources and could not be rebuilt here; loading it as it is")
    try:
        return ctypes.CDLL(path)
    except OSError as e:   # pragma: no cover
        raise ImportError(f"cannot load {path}: {e}. The CUDA extension is mandatory (no CPU fallback).") from e


lib = _load()

p = c_void_p
_SIGS = {
    "pfc_version": (c_int, []),
    "pfc_error_string": (c_char_p, [c_int]),
    "pfc_exp_top": (c_int, []),
    "pfc_padded_classes": (c_int, [c_int]),
    "pfc_padded_batch": (c_int, [c_int]),
    "pfc_num_class_tiles": (c_int, [c_int]),
    "pfc_part_sum_cols": (c_int, []),
    "pfc_row_stats_loss": (c_int, [p, c_int, c_int, p, p, p, p, p, p, p]),
    "pfc_row_stats_loss_prepare": (c_int, [p, c_int, c_int, p, p, p, p, p, p, p, c_float, c_int, p, c_int, c_float, p, p, p, p, c_int, c_int, p]),
    "pfc_l2norm_rows_localize": (c_int, [p, c_int, c_int, p, p, p, c_int64, c_int, p, c_int, p]),
    "pfc_dx_splits": (c_int, [c_int, c_int, c_int]),
    "pfc_dx_max_splits": (c_int, [c_int, c_int]),
    "pfc_cast_f16_to_bf16": (c_int, [p, p, c_size_t, p]),
    "pfc_l2norm_rows": (c_int, [p, p, c_int, c_int, p, p, c_int, p]),
    "pfc_localize_labels": (c_int, [p, c_int, c_int64, c_int, p, p]),
    "pfc_sample_workspace_bytes": (c_size_t, [c_int]),
    "pfc_sample": (c_int, [p, p, c_int, c_int, c_int, p, p, p, p, c_size_t, p]),
    "pfc_host_mt19937_state_bytes": (c_size_t, []),
    "pfc_host_mt19937_uniform": (c_int, [p, c_size_t, p, c_size_t]),
    "pfc_sample_launches": (c_int, [c_int]),
    "pfc_sample_debug_cluster": (c_int, [c_int]),
    "pfc_gather_rows": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_int, p, c_int, c_int, p]),
    "pfc_scatter_rows": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_int, p, c_int, c_int, p]),
    "pfc_forward": (c_int, [p, p, p, c_int, c_int, c_int, c_float, c_int, c_float, c_float, c_float, p, c_int, p, p, p, p, c_int, p]),
    "pfc_margin_apply": (c_int, [p, p, c_int, c_int, c_int, c_float, c_float, c_float, c_float, p, p, p]),
    "pfc_row_stats": (c_int, [p, c_int, c_int, p, p, p, p]),
    "pfc_loss": (c_int, [p, c_int, p, p, p]),
    "pfc_backward_prepare": (c_int, [p, p, p, c_float, c_int, c_int, p, p, c_int, c_float, p, p, p, p, c_int, c_int, p]),
    "pfc_backward_dx": (c_int, [p, c_int, p, c_int, c_int, c_int, p, c_int, p]),
    "pfc_dx_finalize": (c_int, [p, c_int, p, p, p, c_float, c_int, c_int, c_int, p, p]),
    "pfc_backward_dw": (c_int, [p, c_int, p, c_int, c_int, c_int, p, c_int, p]),
    "pfc_dw_finalize": (c_int, [p, p, p, c_int, c_int, c_float, p, p]),
    "pfc_dw_sgd": (c_int, [p, c_int, p, p, p, c_int, c_int, c_float, c_float, c_float, p, p, p, p, c_int, p, p]),
    "pfc_dw_adam": (c_int, [p, p, p, p, p, c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_int, c_int, p, p, p, p, p, c_int, p, p]),
    "pfc_peer_max_ranks": (c_int, []),
    "pfc_peer_set_timeout_ms": (c_int, [c_double]),
    "pfc_peer_barrier": (c_int, [POINTER(c_void_p), p, c_int, c_int, p]),
    "pfc_peer_l2norm_gather": (c_int, [p, p, c_int, c_int, c_int, c_int, POINTER(c_void_p), POINTER(c_void_p), p, c_int, p]),
    "pfc_peer_row_stats": (c_int, [p, c_int, c_int, p, p, c_int, c_int, POINTER(c_void_p), p]),
    "pfc_peer_loss": (c_int, [POINTER(c_void_p), p, c_int, p, c_int, c_int, p, p, p, p]),
    "pfc_peer_loss_prepare": (c_int, [POINTER(c_void_p), p, c_int, p, c_int, c_int, p, p, p, p, p, c_float, c_int, p, p, c_int, c_float, p, p, p, p, c_int, c_int, p]),
    "pfc_peer_localize_labels": (c_int, [POINTER(c_void_p), p, c_int, c_int, p, c_int, ctypes.c_int64, c_int, p, p]),
    "pfc_peer_dx_finalize": (c_int, [POINTER(c_void_p), p, c_int, c_int, p, p, p, c_float, c_int, c_int, p, p]),
    "pfc_peer_dx_scatter": (c_int, [p, c_int, p, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p), p]),
    "pfc_eval_hist_bins": (c_int, []),
    "fr_pair_score": (c_int, [p, p, p, c_int, c_int, p, p, p, p, p]),
    "fr_cross_score": (c_int, [p, p, c_int, c_int, p, p, p, p, p]),
    "fr_roc": (c_int, [p, p, c_int, c_int, p, p]),
    "fr_acc_counts": (c_int, [p, p, c_int, c_double, p, p]),
    "fr_kfold_acc": (c_int, [p, p, c_int, c_int, c_int, c_double, p, p, p, p]),
}
EXPORTS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)      # AttributeError here == header / library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def error_string(code):
    return lib.pfc_error_string(int(code)).decode()


def check(code, where):
    if code != 0:
        raise PfcError(code, where)


class RocOut(ctypes.Structure):
    """fr_roc_out_t of include/pfc.h."""
    _fields_ = [("eer_threshold", c_int32), ("pad", c_int32), ("eer", c_double), ("total_genuine", c_double),
                ("total_imposter", c_double), ("frr_at", c_double * 16), ("th_at", c_int32 * 16)]


def ptr_array(ptrs):
    arr = (c_void_p * len(ptrs))()
    for i, v in enumerate(ptrs):
        arr[i] = v
    return arr
