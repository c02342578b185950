"""The sampling draw of nets/PartialFC.py:110 -- `torch.rand(size=[num_local])` on the CPU generator -- in bulk.

Bit-exact sampling needs exactly the numbers torch's CPU generator (MT19937) would produce, in the same order, and the
generator left in the same state.  torch draws them one generator call at a time, which costs more host time per step than
the whole sampled step takes on the GPU once a rank holds a few hundred thousand classes.  `cpu_rand_` hands the
generator's state blob to `pfc_host_mt19937_uniform` (csrc/pfc_hostrng.cu: the same published algorithm, 624 words at a
time, written straight into the caller's -- possibly pinned -- buffer) and stores the advanced state back.  The first call
checks the result and the final state against torch.rand on a copy of the generator; any difference (another torch build
with another state layout) switches the module back to torch.rand for good.  PFC_HOST_RNG=0 does the same by hand.
Pinned by tests/test_host_rng.py."""
import os
import warnings

import torch

from ._lib import lib

_MIN_BULK = 4096            # below this torch.rand is as fast as the state round trip
_state = {"ok": None}       # None: not verified yet; True / False afterwards


def _bulk(out, gen):
    st = gen.get_state()
    if st.dtype != torch.uint8 or st.numel() != lib.pfc_host_mt19937_state_bytes() or not st.is_contiguous():
        return False
    if lib.pfc_host_mt19937_uniform(st.data_ptr(), st.numel(), out.data_ptr(), out.numel()) != 0:
        return False
    gen.set_state(st)
    return True


def _verify():
    """One-off: 3 draws that cross block boundaries, compared with torch.rand on a generator in the same state."""
    try:
        a, b = torch.Generator(), torch.Generator()
        a.manual_seed(20240607)
        b.manual_seed(20240607)
        for n in (5, 1000, 4099):
            x = torch.empty(n)
            if not _bulk(x, a) or not torch.equal(x, torch.rand(n, generator=b)):
                return False
        return torch.equal(a.get_state(), b.get_state())
    except Exception:       # noqa: BLE001 -- any surprise means "do not use it"
        return False


def enabled():
    if _state["ok"] is None:
        if os.environ.get("PFC_HOST_RNG", "1") == "0":
            _state["ok"] = False
        else:
            _state["ok"] = _verify()
            if not _state["ok"]:
                warnings.warn("bulk MT19937 draw does not reproduce torch.rand on this torch build; using torch.rand")
    return _state["ok"]


def cpu_rand_(out, generator=None):
    """Fill the CPU float32 tensor `out` with what torch.rand(out.shape, generator=generator) would return, advancing the
    generator identically; returns `out`."""
    gen = torch.default_generator if generator is None else generator
    if (out.numel() >= _MIN_BULK and out.device.type == "cpu" and out.dtype == torch.float32 and out.is_contiguous()
            and gen.device.type == "cpu" and enabled() and _bulk(out, gen)):
        return out
    return torch.rand(out.shape, generator=generator, out=out)


def cpu_rand(n, generator=None):
    return cpu_rand_(torch.empty(int(n), dtype=torch.float32), generator)
