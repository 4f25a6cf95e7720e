"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).

    python tools/ncu_launch_summary.py gpurun_out/launches.csv [--skip N] > profiles/rNN_launches_summary.md

--skip N drops the first N launches (model upload, warm-up step) so that only the timed step is summarised; without
it every launch of this repo's kernels (namespace md::) is counted.
"""
import argparse
import csv
import re
import sys
from collections import defaultdict

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--skip", type=int, default=0)
ap.add_argument("--only-md", action="store_true", help="count only this repo's kernels (md::)")
ap.add_argument("--second-half", action="store_true",
                help="after --only-md: keep the second half (bench.py --steps 1 --warmup 1 runs two identical steps)")
args = ap.parse_args()

rows = []
with open(args.csv, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e6))       # ns -> ms
rows = rows[args.skip:]
if args.only_md:
    rows = [x for x in rows if "md::" in x[0]]
if args.second_half:
    rows = rows[len(rows) // 2:]


def short(name):
    name = re.sub(r"\(.*$", "", name)                                    # drop the parameter list
    return name.strip()


tot = defaultdict(float)
cnt = defaultdict(int)
for name, ms in rows:
    k = short(name)
    tot[k] += ms
    cnt[k] += 1
total = sum(tot.values())
print(f"launches {len(rows)}, summed device time {total:.2f} ms\n")
print("| kernel | launches | total ms | share | avg us |")
print("|---|---|---|---|---|")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"| `{k}` | {cnt[k]} | {tot[k]:.3f} | {100 * tot[k] / total:.1f}% | {1000 * tot[k] / cnt[k]:.1f} |")
sys.exit(0)
