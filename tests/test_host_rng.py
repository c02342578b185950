"""The bulk sampling draw (face_recognition_pytorch_b200.hostrng / csrc/pfc_hostrng.cu) against the thing it replaces:
`torch.rand(size=[num_local])` on the CPU generator (nets/PartialFC.py:110).  Bit-exact values, bit-exact generator state
afterwards (so every later torch draw -- the next step's included -- is unchanged), for sizes around the 624-word block
boundaries, from arbitrary generator positions, interleaved with other torch draws.  Host code only: runs without a GPU."""
import ctypes

import pytest
import torch

import face_recognition_pytorch_b200  # noqa: F401  (loads / builds the library)
from face_recognition_pytorch_b200 import hostrng
from face_recognition_pytorch_b200._lib import lib


def _bulk(n, gen):
    out = torch.empty(n)
    assert hostrng._bulk(out, gen)
    return out


def test_self_check_passes_on_this_torch_build():
    assert hostrng.enabled(), "the bulk draw must reproduce torch.rand on the image's torch build"


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 623, 624, 625, 1247, 1248, 1249, 4096, 45029, 360232])
@pytest.mark.parametrize("advance", [0, 1, 376, 623, 624, 1000])
def test_values_and_state_match_torch_rand(n, advance):
    a, b = torch.Generator(), torch.Generator()
    a.manual_seed(1234)
    b.manual_seed(1234)
    if advance:
        torch.rand(advance, generator=a)
        torch.rand(advance, generator=b)
    x = _bulk(n, a)
    y = torch.rand(n, generator=b)
    assert torch.equal(x, y)
    assert torch.equal(a.get_state(), b.get_state())
    # and the generators keep agreeing through draws of other kinds
    assert torch.equal(torch.randn(33, generator=a), torch.randn(33, generator=b))
    assert torch.equal(torch.randint(0, 1 << 30, (9,), generator=a), torch.randint(0, 1 << 30, (9,), generator=b))
    assert torch.equal(_bulk(700, a), torch.rand(700, generator=b))


def test_default_generator_interleaved_with_other_draws():
    outs = []
    for bulk in (True, False):
        torch.manual_seed(77)
        seq = []
        for step in range(4):
            seq.append(torch.randn(5))                                   # e.g. a dropout mask drawn elsewhere
            seq.append(hostrng.cpu_rand(20000) if bulk else torch.rand(size=[20000]))
            seq.append(torch.randperm(11))
        seq.append(torch.get_rng_state())
        outs.append(seq)
    for u, v in zip(*outs):
        assert torch.equal(u, v)


def test_small_and_unsupported_outputs_take_the_torch_path():
    torch.manual_seed(3)
    a = hostrng.cpu_rand(100)                   # below the bulk threshold
    b = hostrng.cpu_rand_(torch.empty(5000, dtype=torch.float64))
    torch.manual_seed(3)
    assert torch.equal(a, torch.rand(100))
    assert torch.equal(b, torch.rand(5000, dtype=torch.float64))


def test_rejects_a_blob_that_is_not_a_cpu_generator_state():
    out = torch.empty(10)
    blob = torch.zeros(5056, dtype=torch.uint8)            # seeded == 0
    assert lib.pfc_host_mt19937_uniform(blob.data_ptr(), blob.numel(), out.data_ptr(), 10) != 0
    short = torch.zeros(100, dtype=torch.uint8)
    assert lib.pfc_host_mt19937_uniform(short.data_ptr(), short.numel(), out.data_ptr(), 10) != 0
    assert lib.pfc_host_mt19937_state_bytes() == torch.Generator().get_state().numel()
    assert isinstance(ctypes.c_size_t(lib.pfc_host_mt19937_state_bytes()).value, int)


def test_env_switch_restores_torch_rand(monkeypatch):
    monkeypatch.setitem(hostrng._state, "ok", False)
    torch.manual_seed(9)
    a = hostrng.cpu_rand(10000)
    torch.manual_seed(9)
    assert torch.equal(a, torch.rand(10000))


def test_published_mt19937_known_answers():
    """Independent of torch: a state blob seeded with the published init_genrand recurrence (seed 5489) must yield the
    published outputs -- the first is 3499211612, the 10000th is 4123659995 (ISO C++ [rand.predef]) -- as floats of their
    low 24 bits.  torch.Generator().manual_seed(5489) must be that same state."""
    import numpy as np
    s = np.zeros(624, dtype=np.uint64)
    s[0] = 5489
    for j in range(1, 624):
        prev = int(s[j - 1])
        s[j] = (1812433253 * (prev ^ (prev >> 30)) + j) & 0xFFFFFFFF
    blob = np.zeros(5056, dtype=np.uint8)
    blob[0:8] = np.frombuffer(np.uint64(5489).tobytes(), dtype=np.uint8)
    blob[8:12] = np.frombuffer(np.int32(1).tobytes(), dtype=np.uint8)        # left
    blob[12:16] = np.frombuffer(np.int32(1).tobytes(), dtype=np.uint8)       # seeded
    blob[24:24 + 624 * 8] = np.frombuffer(s.tobytes(), dtype=np.uint8)
    t = torch.from_numpy(blob)
    out = torch.empty(10000)
    assert lib.pfc_host_mt19937_uniform(t.data_ptr(), t.numel(), out.data_ptr(), out.numel()) == 0
    assert float(out[0]) == (3499211612 & 0xFFFFFF) / 2 ** 24
    assert float(out[9999]) == (4123659995 & 0xFFFFFF) / 2 ** 24
    g = torch.Generator()
    g.manual_seed(5489)
    assert torch.equal(out, torch.rand(10000, generator=g))
