#!/usr/bin/env bash
# Round 2, call 10: TS-mode weight stream (activation tile through TMEM), ncu --set full of the flash kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] TS-mode stream tests"
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -k "ts_mode or small_batch" > $O/c10_pytest_ts.log 2>&1
rc=$?; echo "rc=$rc"; grep -E "passed|failed|error" $O/c10_pytest_ts.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert |error" $O/c10_pytest_ts.log | head -20 | cut -c1-300
if [ $rc -eq 0 ]; then
  echo "== [2] decode timeline: default vs TS (bit 4), twice"
  for f in 0 16 0 16; do
    echo "-- gemm-debug $f"
    timeout 200 python tools/decode_timeline.py --brief --gemm-debug $f --out $O/c10_decode_timeline_dbg$f.json 2>&1 | grep -E "Error|error|layer period|^gemm[12] |^attn |^epi "
  done
  echo "== [3] parity suites under TS"
  MD_DEBUG_GEMM=16 timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_parity_2b_gpu.py -q -m gpu -x > $O/c10_pytest_ts_parity.log 2>&1
  echo "rc=$?"; grep -E "passed|failed|error" $O/c10_pytest_ts_parity.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c10_pytest_ts_parity.log | head -20 | cut -c1-300
fi
echo "== [4] ncu --set full: flash attention impl 0 and 2 (third launch of each)"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fa_tc_ --launch-skip 4 --launch-count 2 -f -o $O/r02_fa_impl0 \
  python tools/attn_profile.py 0 > $O/c10_ncu0.log 2>&1
echo "rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fa_tc_ --launch-skip 4 --launch-count 2 -f -o $O/r02_fa_impl2 \
  python tools/attn_profile.py 2 > $O/c10_ncu2.log 2>&1
echo "rc=$?"; ls -la $O/*.ncu-rep
