#!/usr/bin/env bash
# Round 2, call 6 (after the container was re-created): full GPU suite, then the two pending A/Bs
#   * stream-tail epilogue (md_debug_gemm bit 4): parity suites under it + decode timeline + phase times
#   * FMA-pipe exp2 flash attention (md_debug_attention_impl 2): kernel timing + parity suites under it + phase times
# and a short bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu --durations=5 > $O/c6_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c6_pytest.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c6_pytest.log | head -30 | cut -c1-300
echo "== [2] parity suites with the stream tail enabled (MD_DEBUG_GEMM=16)"
MD_DEBUG_GEMM=16 timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_parity_2b_gpu.py -q -m gpu -x > $O/c6_pytest_tail.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c6_pytest_tail.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c6_pytest_tail.log | head -20 | cut -c1-300
echo "== [3] parity suites with the FMA-pipe exp2 attention (MD_ATTENTION_IMPL=2)"
MD_ATTENTION_IMPL=2 timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_parity_2b_gpu.py -q -m gpu -x > $O/c6_pytest_poly.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c6_pytest_poly.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c6_pytest_poly.log | head -20 | cut -c1-300
echo "== [4] attention kernels A/B"
timeout 300 python tools/attn_bench.py 2>&1 | tail -10
echo "== [5] decode timeline: default vs stream tail (bit 4), twice"
for f in 0 16 0 16; do
  echo "-- gemm-debug $f"
  timeout 200 python tools/decode_timeline.py --brief --gemm-debug $f --out $O/c6_decode_timeline_dbg$f.json 2>&1 | grep -E "Error|error|layer period|^gemm[12] |^attn |^epi "
done
echo "== [6] phase times: default / tail / poly attention"
timeout 200 python tools/phase_times.py 2>&1 | grep -E "vit_encode|image_prefill_est|decode_ms|generate_total"
MD_DEBUG_GEMM=16 timeout 200 python tools/phase_times.py 2>&1 | grep -E "vit_encode|image_prefill_est|decode_ms|generate_total"
MD_ATTENTION_IMPL=2 timeout 200 python tools/phase_times.py 2>&1 | grep -E "vit_encode|image_prefill_est|decode_ms|generate_total"
echo "== [7] bench (no comparators)"
timeout 600 python bench.py --steps 10 --warmup 3 --comparator none > $O/c6_bench.json 2> $O/c6_bench.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c6_bench.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}).get("value"), d.get("parity", {}).get("teacher_forced"),
          d.get("cpu_baseline", {}).get("value"), d.get("roofline", {}).get("frac"), d.get("clocks"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 $O/c6_bench.err | cut -c1-300
