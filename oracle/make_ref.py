"""Recipe for oracle/_ref: the reference's own head modules, for use as the CPU reference arm.  TEST INFRASTRUCTURE ONLY.

The reference (aanna0701/face-recognition-pytorch) has no packaging and no native code: its hot path is three
pure-Python files over torch / numba.  Where /root/reference exists (the build container) this script copies those
three files, byte for byte, into oracle/_ref/{nets,utils}/ -- a BUILD OUTPUT, git-ignored like a compiled oracle would
be, but not gpurun-ignored, so it travels to the GPU box where /root/reference does not exist.  Nothing is committed and
nothing under the product package imports it; only `bench.py --impl reference` / its cpu_baseline leg load the modules
(kind = "reference") and fall back to the port in oracle/head_oracle.py (kind = "port") when the directory is absent.

    python oracle/make_ref.py            # idempotent; prints what it did
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PFC_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
# the head (nets/), the scorer (utils/eval.py) and the head's CALL SITE: the trainer's Model with the two small utility modules
# it imports -- tests/test_gpu_dropin.py drives Model.training_step with the reference head and with this package's head
FILES = ["nets/ArcFace.py", "nets/PartialFC.py", "utils/eval.py", "model/FR_PartialFC.py", "utils/logger.py",
         "utils/scheduler.py"]


def make_ref(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"make_ref: {REF} not present; keeping {OUT} as it is ({'present' if os.path.isdir(OUT) else 'absent'})")
        return os.path.isdir(OUT)
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            with open(dst, "rb") as fh:
                print(f"make_ref: {rel}  sha256 {hashlib.sha256(fh.read()).hexdigest()[:16]}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
