"""The eight example scripts mirror the reference notebooks 1:1; run them end to end in small CPU mode."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = ["part1/00_setup.py", "part1/01_data_prep.py", "part1/02_train_single_node.py",
           "part1/03_train_distributed.py", "part2/01_hpo_single.py", "part2/02_hpo_distributed.py",
           "part2/03_pyfunc_inference.py"]


def test_workshop_runs_end_to_end(tmp_path):
    env = dict(os.environ, WORKSHOP_SMALL="1", B200DDL_HOME=str(tmp_path / "home"), B200DDL_USER="test.user@example.com",
               B200DDL_FORCE_CPU="1", CUDA_VISIBLE_DEVICES="", WORKSHOP_IMAGES="64", NUM_EVALS="3", MAX_EVALS="2")
    env.pop("B200DDL_TRACKING_URI", None)
    for s in SCRIPTS:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "examples", s)], env=env, capture_output=True, text=True,
                           timeout=900, cwd=str(tmp_path))
        assert p.returncode == 0, f"{s} failed:\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}"
    out = p.stdout
    assert "scored" in out and "worker(s)" in out
    # artefacts of the whole flow exist
    home = tmp_path / "home"
    assert (home / "warehouse" / "distributed_dl_workshop_test_user.db" / "silver_train" / "_log").is_dir()
    assert (home / "mlruns" / "models" / "test_user_flower_classifier" / "meta.json").exists()
