"""CPU: `bench.py --impl reference` (the CPU restatement of the step, oracle/step_port.py) runs end to end on the
tiny workload and prints one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in line


def test_usable_cores_respects_cgroup_quota(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    n = bench.usable_cores()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_eager_arm_plumbing_on_cpu(capsys):
    """`--impl eager` needs a GPU; its code path (state on device, host banks at capacity, timing, JSON line) is
    exercised here on the CPU device with the tiny workload."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    args = argparse.Namespace(workload="tiny", steps=1, warmup=0, gpus=1)
    line = bench.eager_arm(args, dev="cpu")
    assert line["impl"] == "eager" and line["value"] > 0 and line["config"]["global_batch"] == 4
    assert len(line["losses"]) == 3 and line["losses"][2] > 0            # contrastive branch ran against full banks
    assert json.loads(capsys.readouterr().out.strip().splitlines()[-1])["impl"] == "eager"
