#!/usr/bin/env bash
# Round 2, call 16 (2 GPUs): bench through torchrun as the driver launches it, ShardedModel API check, CPU-leg probe.
set -u
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== [1] bench N=2"
timeout 600 $T bench.py --gpus 2 --steps 5 --warmup 3 --comparator none > $O/n2_bench.json 2> $O/n2_bench.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/n2_bench.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "gpu_launches", "scaling")}, d.get("e2e"), d.get("strong_scaling_point"),
          d.get("parity", {}).get("teacher_forced"), d.get("clocks"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 $O/n2_bench.err | cut -c1-300
echo "== [2] ShardedModel on 2 ranks"
timeout 300 $T tools/sharded_api_check.py 2>&1 | grep -E "rank|Error|error|assert" | head -10
echo "== [3] CPU-leg probe (one process: decode before / after CUDA init)"
timeout 500 python tools/cpu_arm_probe2.py 2>&1 | tail -30 | cut -c1-260
