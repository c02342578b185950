"""The three launch modes of the tcgen05 GEMMs (single CTA, multicast CTA pair, cta_group::2 CTA pair) must give
the same results; auto picks the single-CTA kernel, so the pair kernels are exercised here explicitly."""
import pytest

from tools import gpu_probe

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("case", ["fwd_mid", "dx", "dw"])
def test_gemm_mode(mode, case):
    from face_recognition_pytorch_b200 import _lib
    _lib.lib.pfc_debug_cluster(mode)
    try:
        ok = getattr(gpu_probe, "case_" + case)()
    finally:
        _lib.lib.pfc_debug_cluster(0)
    assert ok


def test_odd_tile_counts_in_pair_mode():
    """B = 300 -> 3 sample tiles (padded to 4 in cta_group::2 mode), n = 1000 -> 8 class tiles of 128 for dW."""
    from face_recognition_pytorch_b200 import _lib
    _lib.lib.pfc_debug_cluster(3)
    try:
        assert gpu_probe._fwd(300, 1000, 512)
        assert gpu_probe._dx(300, 1000, 512, False)
        assert gpu_probe._dw(300, 1100, 512, False)      # 9 class tiles: odd, one all-padding tile
    finally:
        _lib.lib.pfc_debug_cluster(0)
