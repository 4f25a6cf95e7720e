#!/usr/bin/env bash
# Round 2, call 18: single-pass softmax V3 (shared-memory base agreement + FMA-pipe exp2 for 3/8).
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] kernel tests + model parity + features"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_parity_gpu.py tests/test_features_gpu.py -q -m gpu > $O/c18_pytest_a.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c18_pytest_a.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert |error" $O/c18_pytest_a.log | head -30 | cut -c1-300
echo "== [2] attention kernels A/B"
timeout 300 python tools/attn_bench.py 2>&1 | tail -14
echo "== [3] phases"
timeout 200 python tools/attn_phases.py 2>&1 | tail -16
