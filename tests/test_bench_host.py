"""Host-side pieces of bench.py that need no GPU: the FLOP model behind `roofline` / `step_model_tflops` against the
survey's figures, the nvidia-smi clock parser, and the one-JSON-line stdout guard."""
import json
import os
import subprocess
import sys


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_flop_model_matches_the_survey():
    import bench
    from moondream_b200 import config as C

    cfg2b = C.preset("moondream-2b")
    # per image 3.514 TF with 2 crops, 8.845 TF with 10, 10.84 TF with 13 (2B, 32-token prompt, 64 new)
    for crops, tf in ((2, 3.514), (10, 8.845), (13, 10.84)):
        got = bench.flops_per_image(cfg2b, crops) / 1e12
        assert abs(got - tf) / tf < 0.01, (crops, got, tf)
    got = bench.flops_per_image(C.preset("moondream-0.5b"), 2) / 1e12
    assert abs(got - 1.147) / 1.147 < 0.02, got


def test_clock_sampler_parses_nvidia_smi_lines():
    import bench

    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    s.lines = ["1965, 1965, 400.1, Not Active, Not Active, Not Active, Not Active",
               "1575, 1965, 990.0, Not Active, Not Active, Not Active, Active",
               "1590, 1965, 985.2, Not Active, Not Active, Not Active, Active",
               "210, 1965, 150.0, Not Active, Not Active, Not Active, Not Active",     # idle sample: ignored by the median
               "garbage"]
    out = s.stop()
    assert out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"] and out["samples"] == 4
    assert out["sm_mhz"] == 1590.0


def test_stdout_carries_exactly_one_json_line():
    code = ("import bench, os; bench.protect_stdout(); os.write(1, b'NCCL version banner\\n'); print('library chatter'); "
            "bench.emit_line({'metric': 'x', 'value': 1.0})")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "x", "value": 1.0}
    assert "NCCL version banner" in r.stderr and "library chatter" in r.stderr


def test_reference_arm_exits_quietly_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.returncode, r.stdout[-300:], r.stderr[-300:])
