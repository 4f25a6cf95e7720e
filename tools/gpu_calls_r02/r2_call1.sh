#!/usr/bin/env bash
# Round 2, first GPU call: 2B parity tests + the work staged at the end of round 1, all bounded.
#   gpurun --timeout 1500 -- 'bash tools/gpu_calls_r02/r2_call1.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt 2>&1
python tools/box_info.py > $O/c1_box.txt 2>&1

echo "== [1] pytest -m gpu (everything that is not env-gated, incl. the 2B parity tests)"
timeout 900 python -m pytest tests -q -m gpu --durations=15 -s > $O/c1_pytest.log 2>&1
echo "rc=$?"; tail -30 $O/c1_pytest.log | cut -c1-300

echo "== [2] env-gated experiments (M = 64 small-batch stream) in their own process"
MD_EXPERIMENTAL=1 timeout 300 python -m pytest tests -q -m gpu -k "experimental" --tb=short > $O/c1_pytest_exp.log 2>&1
echo "rc=$?"; tail -8 $O/c1_pytest_exp.log | cut -c1-300

echo "== [3] decode timeline, M = 128 (shipped) / M = 64, twice each"
for f in 0 64 0; do
  echo "-- gemm-debug $f"
  timeout 200 python tools/decode_timeline.py --brief --gemm-debug $f --out $O/c1_decode_timeline_dbg$f.json 2>&1 |
    grep -E "Error|error|layer period|^gemm[12] |^attn |^epi "
done

echo "== [4] phase times"
timeout 200 python tools/phase_times.py 2>&1 | tail -12

echo "== [5] encode / decode overlap probe"
timeout 420 python tools/overlap_probe.py --batches 6 2>&1 | tail -8

echo "== [6] CPU arm probe (CFS throttling?)"
timeout 600 python tools/cpu_arm_probe.py 2>&1 | tail -10 | cut -c1-400

echo "== [7] bench quick"
timeout 300 python bench.py --steps 5 --warmup 3 --comparator none > $O/c1_bench.json 2> $O/c1_bench.err
echo "rc=$?"; cut -c1-600 $O/c1_bench.json

echo "== [7b] torch-CUDA comparator (eager, then the reference's compile() recipe)"
timeout 200 python -m oracle.torch_cuda_comparator --images 8 --tokens 64 > $O/c1_comparator_eager.json 2> $O/c1_comparator_eager.err
echo "rc=$?"; cut -c1-700 $O/c1_comparator_eager.json
timeout 600 python -m oracle.torch_cuda_comparator --images 8 --tokens 64 --compile > $O/c1_comparator_compiled.json 2> $O/c1_comparator_compiled.err
echo "rc=$?"; cut -c1-900 $O/c1_comparator_compiled.json; tail -5 $O/c1_comparator_compiled.err | cut -c1-300

echo "== [8] ncu --set full of the decode-side kernels and the prefill attention"
for k in decode_attention_kernel smallbatch_gemm_kernel decode_residual_ln_epilogue_kernel fa_tc_prefill_kernel; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 30 -c 2 -f \
    -o $O/r02_$k python tools/decode_only.py > $O/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la $O | tail -30
