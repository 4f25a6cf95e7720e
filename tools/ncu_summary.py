"""Summarise `ncu --set full` reports (read here, no GPU needed) into one JSON: per captured launch the duration, DRAM
bytes, achieved GB/s, tensor-pipe %, issue activity, registers, occupancy.

    python tools/ncu_summary.py gpurun_out/r02_*.ncu-rep > profiles/r02_ncu_decode_summary.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "duration_us": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_pct": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_inst_pct": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "registers": "launch__registers_per_thread",
    "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "shared_mem_per_block": "launch__shared_mem_per_block_dynamic",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9,
              "usecond": 1e3, "nsecond": 1.0, "msecond": 1e6, "second": 1e9}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return [{"report": path, "error": "no captured launches"}]
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {"report": path.split("/")[-1], "kernel": r[hdr.index("Kernel Name")][:60]}
        for name, metric in KEYS.items():
            cols = [i for i, h in enumerate(hdr) if h == metric or h.endswith("." + metric)]
            if not cols:
                continue
            v = num(r[cols[0]])
            if v is None:
                continue
            u = units[cols[0]]
            if name == "duration_us":
                v = v * UNIT_SCALE.get(u, 1.0) / 1e3
            elif name.endswith("_bytes"):
                v = v * UNIT_SCALE.get(u, 1.0)
            rec[name] = v
        if "dram_read_bytes" in rec and "duration_us" in rec:
            tot = rec["dram_read_bytes"] + rec.get("dram_write_bytes", 0.0)
            rec["dram_bytes"] = tot
            rec["achieved_gb_s"] = tot / (rec["duration_us"] * 1e-6) / 1e9
        out.append(rec)
    return out


if __name__ == "__main__":
    res = []
    for p in sys.argv[1:]:
        res += summarise(p)
    print(json.dumps(res, indent=1))
