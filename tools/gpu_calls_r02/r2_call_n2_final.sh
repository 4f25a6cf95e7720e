set -u
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $T bench.py --gpus 2 --steps 3 --warmup 3 --comparator none > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/n2_bench.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "scaling")}, d.get("e2e", {}).get("value"), d.get("strong_scaling_point", {}).get("value"),
          d.get("parity", {}).get("ok"), (d.get("roofline_decode") or {}).get("frac"), d.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/n2_bench.err | cut -c1-300
