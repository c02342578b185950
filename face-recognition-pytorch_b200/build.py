"""Builds libpfc_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only container; the resulting .so is git-ignored
but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpfc_b200.so")
SOURCES = ["pfc_api.cu", "pfc_gemm.cu", "pfc_rows.cu", "pfc_sample.cu", "pfc_eval.cu", "pfc_peer.cu", "pfc_hostrng.cu"]
HEADERS = ["pfc_ptx.cuh", "pfc_umma.cuh", "pfc_umma2.cuh", "pfc_launch.cuh", "pfc_internal.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libpfc_b200.so")
    return exe


STAMP = LIB + ".srchash"


def _source_hash():
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def needs_build():
    """True when the library is missing or was built from other sources (content hash, not mtimes: a repo snapshot copied
    to another machine does not keep them)."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into libpfc_b200.so; returns the library path."""
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", f".{os.getpid()}.o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose:
            print(out)
        objs.append(obj)
    tmp = f"{LIB}.{os.getpid()}.tmp"          # per process: concurrent builders never publish a half-linked file
    link = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB)
    with open(STAMP + f".{os.getpid()}", "w") as fh:
        fh.write(_source_hash())
    os.replace(STAMP + f".{os.getpid()}", STAMP)
    for obj in objs:
        try:
            os.remove(obj)
        except OSError:
            pass
    return LIB


def build_locked(force=False, verbose=False):
    """build() behind an exclusive file lock: the ranks of one node (torchrun starts them together) compile once."""
    import fcntl
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            return build(force=force, verbose=verbose)     # re-checks needs_build() now that the lock is held
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
