#!/usr/bin/env bash
# Round 2, call 11: TS-mode weight stream A/B (timeline + parity), attention default flipped to the single-pass softmax.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] GEMM + kernel tests"
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -q -m gpu > $O/c11_pytest_a.log 2>&1
rc=$?; echo "rc=$rc"; grep -E "passed|failed|error" $O/c11_pytest_a.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert |error" $O/c11_pytest_a.log | head -20 | cut -c1-300
echo "== [2] decode timeline: default vs TS (bit 4), twice"
for f in 0 16 0 16; do
  echo "-- gemm-debug $f"
  timeout 200 python tools/decode_timeline.py --brief --gemm-debug $f --out $O/c11_decode_timeline_dbg$f.json 2>&1 | grep -E "Error|error|layer period|^gemm[12] |^attn |^epi "
done
echo "== [3] parity suites under TS"
MD_DEBUG_GEMM=16 timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_parity_2b_gpu.py -q -m gpu -x > $O/c11_pytest_ts_parity.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c11_pytest_ts_parity.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c11_pytest_ts_parity.log | head -20 | cut -c1-300
echo "== [4] bench default vs TS (no comparators, short)"
timeout 400 python bench.py --steps 10 --warmup 3 --comparator none --no-cpu-subprocess > $O/c11_bench_default.json 2> $O/c11_bench_default.err
MD_DEBUG_GEMM=16 timeout 400 python bench.py --steps 10 --warmup 3 --comparator none --no-cpu-subprocess > $O/c11_bench_ts.json 2> $O/c11_bench_ts.err
python - <<'PY'
import json
for n in ("default", "ts"):
    try:
        d = json.loads(open(f"gpurun_out/c11_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, {k: d.get(k) for k in ("value", "ms_per_step")}, d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), d.get("parity", {}).get("ok"), d.get("clocks", {}).get("sm_mhz"))
    except Exception as e:
        print(n, "bench parse failed", e)
PY
tail -2 $O/c11_bench_default.err | cut -c1-300
