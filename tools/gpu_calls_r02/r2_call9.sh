#!/usr/bin/env bash
# Round 2, call 9: single-pass flash-attention softmax (impl 2 / 3), im2col-fused patch embedding, early PDL trigger A/B.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] kernel tests + fused patch embed"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_parity_gpu.py -q -m gpu > $O/c9_pytest_a.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c9_pytest_a.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert |error" $O/c9_pytest_a.log | head -30 | cut -c1-300
echo "== [2] attention kernels A/B"
timeout 300 python tools/attn_bench.py 2>&1 | tail -12
echo "== [3] parity suites under impl 2"
MD_ATTENTION_IMPL=2 timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_parity_2b_gpu.py -q -m gpu -x > $O/c9_pytest_fa2.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c9_pytest_fa2.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c9_pytest_fa2.log | head -20 | cut -c1-300
echo "== [4] parity suites under impl 3"
MD_ATTENTION_IMPL=3 timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_parity_2b_gpu.py -q -m gpu -x > $O/c9_pytest_fa3.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c9_pytest_fa3.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c9_pytest_fa3.log | head -20 | cut -c1-300
echo "== [5] phase times"
for cfg in "0 0" "2 0" "3 0" "0 32" "0 128" "2 0" "0 0"; do
  set -- $cfg
  echo "-- MD_ATTENTION_IMPL=$1 MD_DEBUG_GEMM=$2"
  MD_ATTENTION_IMPL=$1 MD_DEBUG_GEMM=$2 timeout 200 python tools/phase_times.py 2>&1 | grep -E "vit_encode|image_prefill_est|decode_ms"
done
