#!/usr/bin/env bash
# Round 2, third GPU call: LoRA / device preprocessing / tie-aware sampling tests, FMA-pipe exp2 A/B, wide stream plan
# as the default, CPU-leg follow-up probe.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu --durations=8 > $O/c3_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c3_pytest.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c3_pytest.log | head -40 | cut -c1-300
echo "== [2] attention A/B (tcgen05 / mma.sync / tcgen05 + FMA-pipe exp2)"
timeout 300 python tools/attn_bench.py 2>&1 | tail -10
echo "== [3] phase times: default, attention impl 2"
timeout 200 python tools/phase_times.py 2>&1 | grep -E "vit_encode|image_prefill_est|decode_ms"
MD_ATTENTION_IMPL=2 timeout 200 python tools/phase_times.py 2>&1 | grep -E "vit_encode|image_prefill_est|decode_ms"
echo "== [4] decode timeline: default (wide plan), bit 3 (previous plan)"
for f in 0 8; do
  echo "-- gemm-debug $f"
  timeout 200 python tools/decode_timeline.py --brief --gemm-debug $f --out $O/c3_decode_timeline_dbg$f.json 2>&1 | grep -E "Error|error|layer period|^gemm[12] |^attn |^epi "
done
echo "== [5] CPU leg follow-up"
timeout 400 python tools/cpu_arm_probe2.py 2>&1 | tail -40 | cut -c1-300
echo "== [6] bench (no comparators)"
timeout 600 python bench.py --steps 10 --warmup 3 --comparator none > $O/c3_bench.json 2> $O/c3_bench.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c3_bench.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}).get("value"), d.get("parity", {}).get("teacher_forced"),
          d.get("cpu_baseline", {}).get("value"), d.get("roofline", {}).get("frac"), d.get("clocks"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 $O/c3_bench.err | cut -c1-300
