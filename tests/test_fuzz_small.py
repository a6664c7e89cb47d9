"""Short runs of the random-parameter fuzzers in tools/ (random costs, modes, heuristics, truncated queries): the search core
the device runs and the dense model of the DP kernels against the A* oracle. CPU only."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), *args], capture_output=True, text=True, cwd=ROOT, env=e)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_search_core_fuzz():
    assert "product search core == oracle" in _run("fuzz_core_vs_oracle.py", "40", "5")


def test_search_core_fuzz_with_stale_table_records():
    # the visited table pre-filled with records of other epochs (never reset between alignments on the device)
    assert "product search core == oracle" in _run("fuzz_core_vs_oracle.py", "30", "6", env={"RP_HOST_EPOCH": "777"})


def test_dp_model_fuzz():
    assert "DP optimum <= reference score everywhere" in _run("fuzz_dp_model_vs_oracle.py", "40", "9")
