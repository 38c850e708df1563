"""The example scripts build their flows, boundaries and reporters on a CPU context (`--dry`) without stepping."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.parametrize("script,args", [("00_taylor_green.py", ["--resolution", "16"]),
                                         ("01_cylinder_drag.py", ["--diameter", "8", "--bc", "hwbb"])])
def test_example_dry_run(script, args, tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), "--dry", *args],
                       capture_output=True, text=True, timeout=300, cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip()
