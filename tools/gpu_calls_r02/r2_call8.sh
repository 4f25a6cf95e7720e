#!/usr/bin/env bash
# Round 2, call 8: quantised engine tests, patchify / stitch bit-exact tests, flash attention with two softmax
# warpgroups (impl 2) A/B + parity under it, other BASELINE configs incl. int8 / int4 config 5.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] quant + kernel tests"
timeout 900 python -m pytest tests/test_quant_gpu.py tests/test_kernels_gpu.py -q -m gpu > $O/c8_pytest_a.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c8_pytest_a.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert |error" $O/c8_pytest_a.log | head -30 | cut -c1-300
echo "== [2] attention kernels A/B (impl 0 one softmax warpgroup, 1 mma.sync, 2 two softmax warpgroups)"
timeout 300 python tools/attn_bench.py 2>&1 | tail -10
echo "== [3] parity suites under impl 2"
MD_ATTENTION_IMPL=2 timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_parity_2b_gpu.py -q -m gpu -x > $O/c8_pytest_fa2.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c8_pytest_fa2.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c8_pytest_fa2.log | head -20 | cut -c1-300
echo "== [4] phase times: impl 0 / impl 2, twice"
for i in 0 2 0 2; do
  echo "-- MD_ATTENTION_IMPL=$i"
  MD_ATTENTION_IMPL=$i timeout 200 python tools/phase_times.py 2>&1 | grep -E "vit_encode|image_prefill_est|decode_ms"
done
echo "== [5] full suite (after the clean-up of the rejected experiments)"
timeout 1200 python -m pytest tests -q -m gpu --deselect tests/test_quant_gpu.py --deselect tests/test_kernels_gpu.py > $O/c8_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c8_pytest.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c8_pytest.log | head -30 | cut -c1-300
echo "== [6] other configs (C3, C4', C2, C5 bf16 / int8 / int4)"
timeout 900 python tools/config_runs.py > $O/c8_config_runs.log 2>&1
echo "rc=$?"; tail -8 $O/c8_config_runs.log | cut -c1-600
