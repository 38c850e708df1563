"""The working tree is what `gpurun` and the driver snapshot onto the GPU box (limit 512 MiB): keep it small, keep
build products out of the package, and keep the shipped library in step with the sources."""
import os

from conftest import ROOT

SKIP_DIRS = {".git", "gpurun_out", ".pytest_cache"}          # .gpurunignore


def tree_bytes(root):
    total = 0
    for base, dirs, files in os.walk(root):
        if base == root:
            dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        for name in files:
            path = os.path.join(base, name)
            if not os.path.islink(path):
                total += os.path.getsize(path)
    return total


def test_snapshot_stays_far_below_the_gpurun_limit():
    mib = tree_bytes(ROOT) / 2 ** 20
    assert mib < 200, f"working tree is {mib:.0f} MiB without .git/ and gpurun_out/ (gpurun refuses above 512 MiB)"


def test_gpurunignore_uses_literal_prefixes_only():
    with open(os.path.join(ROOT, ".gpurunignore")) as fh:
        entries = [line.strip() for line in fh if line.strip() and not line.startswith("#")]
    assert entries, "empty .gpurunignore"
    for e in entries:
        assert not any(c in e for c in "*?["), f"glob in .gpurunignore is not honoured by the driver: {e}"


def test_no_build_trees_or_variant_libraries_inside_the_package():
    pkg = os.path.join(ROOT, "lettuce_b200")
    for name in os.listdir(pkg):
        assert not name.startswith("build") or name == "build.py", f"build directory inside the package: {name}"
        if name.endswith(".so"):
            assert name == "liblbm_b200.so", f"variant library inside the package: {name}"
    for base, _, files in os.walk(ROOT):
        if any(part in SKIP_DIRS for part in os.path.relpath(base, ROOT).split(os.sep)):
            continue
        assert not [f for f in files if f.endswith(".o")], f"object files in {base}"


def test_shipped_library_was_built_from_the_current_sources():
    from lettuce_b200 import build
    assert os.path.exists(build.LIB), "liblbm_b200.so is missing"
    assert os.path.exists(build.STAMP), "liblbm_b200.so has no source stamp"
    with open(build.STAMP) as fh:
        assert fh.read().strip() == build.source_digest(), "liblbm_b200.so is older than csrc/ or include/"


def test_packed_fp32_operators_are_free_of_ptxas_contractions():
    """ptxas fuses `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 (never the scalar forms): the two-nodes-per-thread
    kernels equal the one-node kernels bit for bit only while every fusable multiply-add in csrc/lbm_core.cuh is
    spelled vfma.  The checker compiles every operator for float2 and compares PTX with SASS (no GPU needed)."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_packed_contraction.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_gpu_evidence_was_recorded_on_the_shipped_sources():
    """profiles/r2_gpu_tests_tail.txt (the round's full GPU test run + smoke) names the source digest of the library
    that ran on the B200; it must be the digest of the sources in this tree."""
    import re
    from lettuce_b200 import build
    with open(os.path.join(ROOT, "profiles", "r2_gpu_tests_tail.txt")) as fh:
        text = fh.read()
    m = re.search(r"source digest ([0-9a-f]{64})", text)
    assert m, "no source digest in profiles/r2_gpu_tests_tail.txt"
    assert m.group(1) == build.source_digest(), "the GPU evidence under profiles/ is older than csrc/ or include/"
    assert re.search(r"\b\d+ passed", text) and " failed" not in text
