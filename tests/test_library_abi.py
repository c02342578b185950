"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/pfc.h declares; the Python binding table matches the header; shape helpers behave."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pfc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:pfc|fr)_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from face_recognition_pytorch_b200 import _lib
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/pfc.h but not exported"
    assert set(_lib.EXPORTS) == set(names), set(_lib.EXPORTS) ^ set(names)


def test_version_and_error_strings():
    from face_recognition_pytorch_b200 import _lib
    assert _lib.lib.pfc_version() >= 100
    assert _lib.error_string(0) == "ok"
    assert "scale" in _lib.error_string(-7)
    with pytest.raises(_lib.PfcError):
        _lib.check(-3, "unit-test")


def test_shape_helpers():
    from face_recognition_pytorch_b200 import kernels as K
    assert K.padded_classes(93431) == 93440 and K.padded_classes(64) == 64
    assert K.num_class_tiles(93431) == 730 and K.num_class_tiles(256) == 2 and K.num_class_tiles(257) == 4
    assert K.padded_batch(1000) == 1024
    assert K.exp_top() == 64
    for B, n, d in [(1024, 93431, 512), (128, 10000, 512), (4096, 51497, 512), (32, 100, 64)]:
        s = K.dx_splits(B, n, d)
        assert 1 <= s <= K.dx_max_splits(B, d)
        k_total = (n + 63) // 64
        per = (k_total + s - 1) // s
        assert (s - 1) * per < k_total          # no empty split
    # one cluster kernel (slot table only) up to ~376 k classes per rank, the tiled six-launch path beyond
    assert K.sample_workspace_bytes(45029) >= 45029 * 4
    assert K.sample_workspace_bytes(2000000) > 2000000 * 5


def test_kernels_refuse_cpu_tensors():
    import torch
    from face_recognition_pytorch_b200 import kernels as K
    x = torch.zeros(4, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        K.l2norm_rows(x, None, 4, torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(4))


def test_shard_range_matches_oracle():
    from face_recognition_pytorch_b200 import shard_range
    from oracle import head_oracle as ho
    for C in (10, 301, 93431, 360232, 2059906):
        for W in (1, 2, 3, 4, 8):
            for r in range(W):
                assert shard_range(C, r, W) == ho.shard_range(C, r, W)


def test_header_is_plain_c_and_cxx():
    """The drop-in boundary is a C ABI: include/pfc.h must compile on its own as C99 and as C++ (no torch / CUDA types in
    the signatures, every type it uses declared by the headers it includes)."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "pfc.h")
    for cc, lang, std in (("gcc", "c", "-std=c99"), ("g++", "c++", "-std=c++17")):
        exe = shutil.which(cc)
        if exe is None:
            pytest.skip(f"{cc} not installed")
        r = subprocess.run([exe, "-fsyntax-only", std, "-Wall", "-Werror", "-x", lang, hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_plain_c_client(tmp_path):
    """tests/c_client/abi_client.c -- a C99 program compiled against include/pfc.h and linked with libpfc_b200.so, as a
    non-Python binding would be -- loads the library without a GPU and gets the published MT19937 known answers (first
    and 10000th output of seed 5489) out of the host sampling-draw entry point."""
    import shutil
    import subprocess
    import face_recognition_pytorch_b200  # noqa: F401  (builds the library if it is missing)
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not installed")
    libdir = os.path.join(ROOT, "face-recognition-pytorch_b200")
    exe = str(tmp_path / "abi_client")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c_client", "abi_client.c"), "-o", exe, "-L", libdir,
                        "-l:libpfc_b200.so", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "abi_client: ok" in r.stdout, r.stdout + r.stderr
