#!/usr/bin/env bash
# Round 2, call 5: full GPU suite, the stream-tail epilogue (md_debug_gemm bit 4) A/B + parity under it, final bench with
# comparators, launch list and the ncu --set full capture of the dominant GEMM.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu --durations=5 > $O/c5_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c5_pytest.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c5_pytest.log | head -30 | cut -c1-300
echo "== [2] parity suites with the stream tail enabled (MD_DEBUG_GEMM=16)"
MD_DEBUG_GEMM=16 timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_parity_2b_gpu.py tests/test_features_gpu.py -q -m gpu -x > $O/c5_pytest_tail.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c5_pytest_tail.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c5_pytest_tail.log | head -20 | cut -c1-300
echo "== [3] decode timeline: default vs stream tail (bit 4), twice"
for f in 0 16 0 16; do
  echo "-- gemm-debug $f"
  timeout 200 python tools/decode_timeline.py --brief --gemm-debug $f --out $O/c5_decode_timeline_dbg$f.json 2>&1 | grep -E "Error|error|layer period|^gemm[12] |^attn |^epi "
done
for f in 0 16; do MD_DEBUG_GEMM=$f timeout 200 python tools/phase_times.py 2>&1 | grep -E "decode_ms"; done
echo "== [4] bench, default flags"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/c5_bench.json 2> $O/c5_bench.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c5_bench.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"),
          d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("in_process_value"), d.get("clocks"))
    print({k: (v.get("images_per_s"), v.get("unavailable")) for k, v in d.get("comparators", {}).items()})
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 $O/c5_bench.err | cut -c1-300
echo "== [5] reference arm"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['cpu_baseline']['sample'][-170:])"
echo "== [6] ncu launch list of one bench step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 8000 --csv --log-file $O/r02_launches_final.csv \
  python bench.py --steps 1 --warmup 1 --no-parity --no-e2e --comparator none > $O/c5_ncu_bench.log 2>&1
echo "rc=$?"; wc -l $O/r02_launches_final.csv
echo "== [7] ncu --set full of the dominant GEMM (4 ViT launches)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel --launch-skip 8 -c 4 -f -o $O/r02_gemm_full \
  python tools/phase_times.py > $O/c5_ncu_gemm.log 2>&1
echo "rc=$?"; ls -la $O/r02_gemm_full.ncu-rep 2>/dev/null
