"""The TF-free counterparts of the reference's scripts (n1270.py, n882.py) and the example scripts."""
import ast
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = ["n1270.py", "n882.py", "sweep.py", "bench.py", "examples/train_feedback_gnn.py", "examples/train_recipe.py"]


@pytest.mark.parametrize("name", SCRIPTS)
def test_script_is_present_and_parses(name):
    src = open(os.path.join(ROOT, name)).read()
    assert len(src) > 500, f"{name} is empty"
    ast.parse(src)


def test_scripts_keep_the_reference_command_line():
    """n1270.py takes -nG / -p / -id, n882.py fixes nG = 5 (reference n882.py:13) and takes -p / -id."""
    a, b = open(os.path.join(ROOT, "n1270.py")).read(), open(os.path.join(ROOT, "n882.py")).read()
    assert '"-nG"' in a and '"-p"' in a and '"-id"' in a
    assert '"-nG"' not in b and "nG = 5" in b and '"-p"' in b and '"-id"' in b
    assert "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy" in b and "create_cyclic_permuting_matrix(7, [27, 54, 0])" in b


@pytest.mark.gpu
@pytest.mark.parametrize("cmd", [["n1270.py", "-nG", "1", "-p", "0.13", "--batch_size", "1000", "--max_iter", "2"],
                                 ["n882.py", "-p", "0.12", "--batch_size", "1000", "--max_iter", "2"]])
def test_reference_scripts_run(cmd):
    r = subprocess.run([sys.executable] + cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "BLER is" in r.stdout and "rounds of GNN feedback" in r.stdout
