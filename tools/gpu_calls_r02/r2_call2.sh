#!/usr/bin/env bash
# Round 2, second GPU call: the new rows (sampling kernel, reasoning, text-only, GQA, streaming, seam), M = 64 default,
# bench through ShardedEngine with the in-run comparators, CPU-leg placement.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu --durations=12 -s > $O/c2_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c2_pytest.log | tail -5; grep -E "^(FAILED|ERROR)|Error|assert " $O/c2_pytest.log | head -40 | cut -c1-300
echo "== [2] bench (default flags + steps 10)"
timeout 900 python bench.py --steps 10 --warmup 3 > $O/c2_bench.json 2> $O/c2_bench.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c2_bench.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e"), d.get("parity", {}).get("teacher_forced"),
          d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("placement"))
    print({k: (v.get("images_per_s"), v.get("compile_and_warmup_s"), v.get("unavailable")) for k, v in d.get("comparators", {}).items()})
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 $O/c2_bench.err | cut -c1-300
echo "== [3] reference arm, twice"
for i in 1 2; do
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['cpu_baseline']['placement'], d['cpu_baseline']['sample'][-160:])"
done
echo "== [4] CPU arm probe: pinned vs unpinned"
PROBE_ONLY=4 timeout 500 python tools/cpu_arm_probe.py 2>&1 | tail -5 | cut -c1-420
echo "== [5] decode timeline (default = M 64) and phase times"
for f in 0 8 0 8; do
  echo "-- gemm-debug $f (bit 3 = operand-read-model plan: wide weight tiles + K splits)"
  timeout 200 python tools/decode_timeline.py --brief --gemm-debug $f --out $O/c2_decode_timeline_dbg$f.json 2>&1 | grep -E "Error|error|layer period|^gemm[12] |^attn |^epi "
done
MD_DEBUG_GEMM=8 timeout 200 python tools/phase_times.py 2>&1 | grep decode_ms
timeout 200 python tools/phase_times.py 2>&1 | tail -9
echo "== [6] compute-sanitizer memcheck on the new kernels (bounded)"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_features_gpu.py -q -m gpu \
  -k "kept_set or grouped_query or reasoning_batch" -p no:cacheprovider > $O/c2_sanitizer.log 2>&1
echo "rc=$?"; tail -4 $O/c2_sanitizer.log | cut -c1-300
