"""Is the device code of the current library the device code of an earlier commit?  (no GPU needed)

    python tools/sass_equivalence.py <git-rev>

Builds the library from <git-rev> in a scratch directory with the flags of moondream_b200/build.py, dumps both libraries
with `cuobjdump -sass` and compares (a) the raw dumps, (b) the dumps with register / predicate numbers and instruction
encodings masked, (c) those as multisets of instructions per kernel.  ptxas is not deterministic from run to run
(register numbering, the order of adjacent independent instructions), so (a) usually differs even for identical sources —
pass the same revision twice (`HEAD`) to see that baseline — while (c) is the equivalence that matters here.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moondream_b200 import build as B  # noqa: E402


def sass(lib):
    return subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout.splitlines()


def masked(lines):
    out = []
    for ln in lines:
        ln = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln)
        ln = re.sub(r"/\*[0-9a-f]{4,}\*/", "", ln)                  # instruction address
        ln = re.sub(r"\bU?R[0-9]+\b", "REG", ln)
        ln = re.sub(r"\bU?P[0-9]+\b", "PRED", ln)
        out.append(" ".join(ln.split()))
    return out


def per_kernel(lines):
    kernels, cur = {}, None
    for ln in lines:
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
        elif cur is not None and ln:
            cur[ln] += 1
    return kernels


def main():
    rev = sys.argv[1] if len(sys.argv) > 1 else "HEAD"
    tmp = tempfile.mkdtemp()
    subprocess.run(f"git archive {rev} moondream_b200/csrc include | tar -x -C {tmp}", shell=True, check=True, cwd=ROOT)
    old = os.path.join(tmp, "old.so")
    subprocess.run(["nvcc"] + B.NVCC_FLAGS + ["-o", old] + B.SOURCES, check=True, cwd=os.path.join(tmp, "moondream_b200", "csrc"))
    B.build()
    a, b = sass(old), sass(B.OUT)
    raw = sum(x != y for x, y in zip(a, b)) + abs(len(a) - len(b))
    ma, mb = masked(a), masked(b)
    msk = sum(x != y for x, y in zip(ma, mb)) + abs(len(ma) - len(mb))
    ka, kb = per_kernel(ma), per_kernel(mb)
    differing = sorted(k for k in set(ka) | set(kb) if ka.get(k) != kb.get(k))
    print(f"{rev} vs working tree: {len(a)} / {len(b)} lines; raw lines differing {raw}; with registers masked {msk}; "
          f"kernels whose instruction multiset differs: {len(differing)} of {len(kb)}")
    for k in differing:
        print("  ", k)
    return 1 if differing else 0


if __name__ == "__main__":
    sys.exit(main())
