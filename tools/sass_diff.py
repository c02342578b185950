"""DEV TOOL: prove that an edit left the DEFAULT instantiation of a kernel untouched when no GPU is at hand -- compile
the old revision of a .cu file, dump both objects with cuobjdump and compare the instruction streams of the kernels
whose mangled name contains a pattern (constant-bank offsets, which move when parameters are appended, are masked).

    python tools/sass_diff.py <git-rev> <file.cu under csrc/> <pattern> [<pattern> ...]
    python tools/sass_diff.py 20f42e1 pfc_rows.cu backward_prepare dx_finalize_d512 dx_finalize_kernel
"""
import difflib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "face-recognition-pytorch_b200", "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"]


def functions(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], check=True, capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            out[cur].append(re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[P]", m.group(1)).strip())
    return out


def main():
    rev, name, patterns = sys.argv[1], sys.argv[2], sys.argv[3:]
    with tempfile.TemporaryDirectory() as tmp:
        # old source against the CURRENT headers' directory layout: take the old headers too
        for f in os.listdir(CSRC):
            old = subprocess.run(["git", "-C", ROOT, "show", f"{rev}:face-recognition-pytorch_b200/csrc/{f}"],
                                 capture_output=True, text=True)
            if old.returncode == 0:
                open(os.path.join(tmp, f), "w").write(old.stdout)
        old_obj, new_obj = os.path.join(tmp, "old.o"), os.path.join(tmp, "new.o")
        subprocess.run(["nvcc", *FLAGS, "-I", tmp, "-c", os.path.join(tmp, name), "-o", old_obj], check=True)
        subprocess.run(["nvcc", *FLAGS, "-I", CSRC, "-c", os.path.join(CSRC, name), "-o", new_obj], check=True)
        o, n = functions(old_obj), functions(new_obj)
    bad = 0
    for pat in patterns:
        for ko in [k for k in o if pat in k]:
            # the new kernel of the same name, or its <false> / <..., false> instantiation if it became a template
            cands = [k for k in n if pat in k and (k == ko or "Lb0" in k)]
            best = min(cands, key=lambda k: sum(1 for _ in difflib.unified_diff(o[ko], n[k], lineterm="", n=0)), default=None)
            if best is None:
                print(f"{ko}: no counterpart")
                bad += 1
                continue
            same = o[ko] == n[best]
            bad += not same
            print(f"{'IDENTICAL' if same else 'DIFFERS  '} {len(o[ko]):5d} / {len(n[best]):5d} instr  {ko}  ->  {best}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
