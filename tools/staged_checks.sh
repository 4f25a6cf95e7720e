#!/usr/bin/env bash
# First GPU call after a round that ended with staged, unvalidated work:
#   gpurun --timeout 900 -- 'bash tools/staged_checks.sh'
# 1. the env-gated tests, 2. the M = 64 small-batch stream against the shipped M = 128 one in one box.
set -u
mkdir -p gpurun_out
echo "== staged tests"
MD_EXPERIMENTAL=1 timeout 300 python -m pytest tests -q -m gpu -k "experimental or spatial_refs_match" --tb=short 2>&1 | tail -15
echo "== decode timeline, M = 128 (shipped) / M = 64 (md_debug_gemm bit 6), twice each"
for f in 0 64 0 64; do
  echo "-- gemm-debug $f"
  timeout 200 python tools/decode_timeline.py --brief --gemm-debug $f --out gpurun_out/decode_timeline_dbg$f.json 2>&1 |
    grep -E "Error|error|layer period|^gemm[12] |^attn |^epi "
done
echo "== encode / decode overlap probe (DESIGN.md section 9, 1b)"
timeout 400 python tools/overlap_probe.py --batches 6 2>&1 | tail -8
echo "== compute-sanitizer (SURVEY.md section 5: memcheck / racecheck on the hand-written kernels; bounded subsets)"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gemm_gpu.py -q -m gpu \
  -k "small_batch and (32-6144 or 7-1032)" -p no:cacheprovider 2>&1 | tail -4
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -q -m gpu \
  -k "decode_attention or layernorm" -p no:cacheprovider 2>&1 | tail -4
echo "== ncu --set full of the decode-side kernels (BASELINE.json: each kernel evidenced by a capture); reports in gpurun_out/"
for k in decode_attention_kernel smallbatch_gemm_kernel decode_residual_ln_epilogue_kernel fa_tc_prefill_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 30 -c 2 -f \
    -o gpurun_out/r02_$k python tools/decode_only.py > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/r02_$k.ncu-rep --page raw --csv 2>/dev/null |
    python - "$k" <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    print(sys.argv[1], "no capture"); sys.exit(0)
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]
for r in rows[2:]:
    print(sys.argv[1], {w: r[hdr.index(w)] for w in want if w in hdr})
PY
done
