#!/usr/bin/env bash
# Round 2, call 7: quantised weight stream (int4 g128 / int8) tests, BN = 192 pair tiles, decode K/V early-prefetch A/B.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] quant tests first (new kernel), bounded"
timeout 600 python -m pytest tests/test_quant_gpu.py -q -m gpu -x > $O/c7_pytest_quant.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c7_pytest_quant.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert |error" $O/c7_pytest_quant.log | head -20 | cut -c1-300
echo "== [2] full pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu --durations=5 --deselect tests/test_quant_gpu.py > $O/c7_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c7_pytest.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c7_pytest.log | head -30 | cut -c1-300
echo "== [3] decode timeline: K/V early prefetch pages 0 2 4 6 8 13"
for pg in 0 2 4 6 8 13; do
  f=$((pg * 256))
  echo "-- pages $pg (gemm-debug $f)"
  timeout 200 python tools/decode_timeline.py --brief --gemm-debug $f --out $O/c7_decode_timeline_pf$pg.json 2>&1 | grep -E "Error|error|layer period|^gemm[12] |^attn |^epi "
done
echo "== [4] phase times: default / BN=256 forced (bit 7) / prefetch 4 / prefetch 8"
for f in 0 128 1024 2048; do
  echo "-- MD_DEBUG_GEMM=$f"
  MD_DEBUG_GEMM=$f timeout 200 python tools/phase_times.py 2>&1 | grep -E "vit_encode|image_prefill_est|decode_ms|generate_total"
done
echo "== [5] parity suites with prefetch 4 pages"
MD_DEBUG_GEMM=1024 timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_parity_2b_gpu.py -q -m gpu -x > $O/c7_pytest_pf.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c7_pytest_pf.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c7_pytest_pf.log | head -20 | cut -c1-300
echo "== [6] memcheck of one small quantised stream case"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_quant_gpu.py -q -m gpu -k "5-640-128 or 64-128-False" > $O/c7_memcheck.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/c7_memcheck.log | tail -4
