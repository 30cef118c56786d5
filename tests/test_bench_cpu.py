"""bench.py's reference arm runs on the host alone (the CPU port of the Go engine): its JSON line is checked here without a GPU --
same metric / unit / config object as the repo arm would print for the same flags, the keys the contract asks for, a real timing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_on_the_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--tier", "nano", "--dtype", "q8_0", "--steps", "2", "--warmup", "1",
                        "--cpu-tokens", "2", "--no-config1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tok/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["metric"] == "decode tok/s (Q8_0, bs=1)" and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and abs(line["value"] - 2 * 1000.0 / line["ms_per_step"]) < 1e-6 * line["value"]      # cpu-tokens forwards per measured step
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "extrapolat" in cb["sample"]   # "... nothing extrapolated"
    assert line["e2e"] == {"value": line["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same config object the repo arm prints for these flags (the driver compares the two arms' configs)
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    from nanollama_b200 import tiers as T
    args = argparse.Namespace(tier="nano", dtype="q8_0", tokens_per_step=256)
    assert line["config"] == bench.decode_config(args, T)
