#!/usr/bin/env bash
# Round 2, call 12: prefill flash attention with separately released K / V stages; detect / point under the per-object
# CUDA graph; kernel tests.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] kernel tests + model parity (detect graph)"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_parity_gpu.py tests/test_api_gpu.py -q -m gpu > $O/c12_pytest_a.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c12_pytest_a.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert |error" $O/c12_pytest_a.log | head -30 | cut -c1-300
echo "== [2] attention kernels A/B"
timeout 300 python tools/attn_bench.py 2>&1 | tail -12
