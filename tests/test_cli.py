"""tests/test_cli.py of the reference: the console script answers --help."""
import subprocess
import sys


def test_cli_help():
    r = subprocess.run([sys.executable, "-m", "lettuce_b200.cli", "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "benchmark" in r.stdout
