"""Static resource table of every kernel in the library: registers, spills, static shared memory, as `ptxas -v` reports
them for sm_100a (no GPU needed).  Dynamic shared memory is chosen at launch and is not in this table.

    python tools/ptxas_resources.py > profiles/rNN_ptxas_resources.json
"""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moondream_b200 import build as B  # noqa: E402


def main():
    out = os.path.join(tempfile.mkdtemp(), "lib.so")
    cmd = ["nvcc"] + B.NVCC_FLAGS + ["-Xptxas", "-v", "-o", out] + B.SOURCES
    log = subprocess.run(cmd, cwd=B.CSRC, capture_output=True, text=True, check=True).stderr
    names = re.findall(r"Compiling entry function '(\S+)'", log)
    demangled = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    rows = []
    blocks = log.split("ptxas info    : Compiling entry function ")[1:]
    for name, blk in zip(demangled, blocks):
        regs = re.search(r"Used (\d+) registers", blk)
        bars = re.search(r"used (\d+) barriers", blk)
        smem = re.search(r"(\d+) bytes smem", blk)
        spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
        rows.append({"kernel": re.sub(r"\(.*", "", name).replace("void ", "").replace("md::", ""), "registers": int(regs.group(1)),
                     "barriers": int(bars.group(1)) if bars else 0, "static_smem_bytes": int(smem.group(1)) if smem else 0,
                     "stack_bytes": int(spill.group(1)), "spill_store_bytes": int(spill.group(2)),
                     "spill_load_bytes": int(spill.group(3))})
    rows.sort(key=lambda r: r["kernel"])
    print(json.dumps({"how": "nvcc " + " ".join(B.NVCC_FLAGS) + " -Xptxas -v (tools/ptxas_resources.py); static, no GPU",
                      "kernels": rows}, indent=1))


if __name__ == "__main__":
    main()
