"""The reference's OWN scripts running against this repo's modules on a GPU (north_star: "drops into
train.py / simple_inference.py unchanged").  tools/run_reference_script.py puts dropin/ + the repo in front of the unmodified
reference checkout (baseline/_ref, staged by __graft_entry__.build()) and executes simple_inference.py's __main__ and the body
of train.py's training loop; each run is a subprocess (the launcher rewires sys.path / sys.modules)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(*args):
    if not os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "train.py")):
        pytest.skip("baseline/_ref is not staged on this box (python baseline/stage_reference.py in the build container)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), *args], cwd=ROOT,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, f"launcher failed (rc {out.returncode}):\n{out.stdout[-2000:]}\n{out.stderr[-3000:]}"
    return json.loads(lines[-1])


def test_reference_simple_inference_script_runs_on_our_modules(cuda_lib):
    r = _run("simple_inference", "--config", "PlaneRecNet_50_config")
    assert r["ok"] and r["model_class"].startswith("planerecnet_b200")
    assert r["state_dict_keys"] == 520 and r["keys_roundtrip"]
    assert any(o.endswith(".png") for o in r["outputs"])          # the rendered detections + depth map of the --image mode


def test_reference_train_loop_runs_on_our_modules(cuda_lib):
    """train.py's loop body (NetLoss + CustomDataParallel + Adam over five groups, train.py:128-150, 251-256, 344-354) for three
    iterations: finite losses, parameters move, all 816 keys round-trip through save_weights / load_weights (strict)."""
    r = _run("train_loop", "--config", "PlaneRecNet_101_config", "--iters", "3", "--batch", "2")
    assert r["finite"] and r["tensors_updated"] > 150 and r["state_dict_keys"] == 816 and r["save_load_roundtrip"]
    assert set(r["losses"][0].keys()) == {"ins", "cat", "dpt", "pln", "lav"}
