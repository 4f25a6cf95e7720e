#!/usr/bin/env bash
# Round 2, call 14 (final state): full GPU suite, full bench line (parity gate, cpu_baseline, in-run comparators),
# reference arm, other configs.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== [1] pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu --durations=6 > $O/c14_pytest.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/c14_pytest.log | tail -3; grep -E "^(FAILED|ERROR)|Error|assert " $O/c14_pytest.log | head -30 | cut -c1-300
echo "== [2] bench, default flags, 20 steps"
timeout 1200 python bench.py --steps 20 --warmup 5 > $O/c14_bench.json 2> $O/c14_bench.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c14_bench.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"),
          d.get("cpu_baseline", {}).get("value"), d.get("parity", {}).get("teacher_forced"), d.get("clocks"))
    print({k: (v.get("images_per_s"), v.get("unavailable")) for k, v in d.get("comparators", {}).items()})
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 $O/c14_bench.err | cut -c1-300
echo "== [3] reference arm"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/c14_bench_reference.json 2>/dev/null
python -c "import json; d=json.loads(open('gpurun_out/c14_bench_reference.json').read().strip().splitlines()[-1]); print(d['value'], d['cpu_baseline'].get('cores'), d['cpu_baseline']['sample'][-170:])"
echo "== [4] other configs"
timeout 900 python tools/config_runs.py > $O/c14_config_runs.log 2>&1
echo "rc=$?"; tail -7 $O/c14_config_runs.log | cut -c1-400
