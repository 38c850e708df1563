"""bench.py's host-side pieces that need no GPU: the reference arm's JSON line (the unmodified reference from
baseline/_ref on the host cores), its behaviour under torchrun's environment, the MEASURED_PEAKS.json reader and the
look-up of the reference's prebuilt generated kernels."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)


def _reference_arm(extra_env, *flags):
    env = dict(os.environ, **extra_env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--size", "64", *flags], capture_output=True, text=True, env=env,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_line_and_all_cores_under_torchrun():
    from baseline import reference
    if not reference.available():
        pytest.skip("baseline/_ref is not installed")
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference must still get every host core
    lines = _reference_arm({"OMP_NUM_THREADS": "1", "RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert len(lines) == 1
    line = lines[0]
    assert line["impl"] == "reference" and line["unit"] == "MLUPS" and line["higher_is_better"] is True
    assert line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["reference_lattice"] == [64] * 3 and line["config"]["same_config"] is False
    # the other ranks exit 0 without work
    assert _reference_arm({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2") == []


def test_measured_peaks_reader(tmp_path, monkeypatch):
    import bench
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.measured_hbm_peak() == (6650.0, "fallback (B200_PROFILING.md)")
    for content, want in (({"hbm_gbs": 6544.3, "bf16_tflops": 1500.0}, 6544.3),
                          ({"hbm": {"burst_gbs": 6700.0, "sustained_gbs": 6500.0}}, 6500.0),
                          ({"peaks": {"hbm_copy_gb_s": 6600.0, "bf16_dense_tflops": 1700.0}}, 6600.0)):
        (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(content))
        value, source = bench.measured_hbm_peak()
        assert value == want and source.startswith("measured")
    (tmp_path / "MEASURED_PEAKS.json").write_text("not json")
    assert bench.measured_hbm_peak()[0] == 6650.0


def test_prebuilt_generated_kernels_are_found_by_the_reference():
    """baseline/build_native.py's packages are looked up by stencil / strategy / operators, not by the reference's
    version-dependent hash, and handed to the reference's own Generator.resolve() (no GPU needed up to the call)."""
    import torch
    from baseline import reference
    if not reference.available() or not reference.native_modules():
        pytest.skip("the reference's generated kernels are not prebuilt (python baseline/build_native.py)")
    lt = reference.load()
    ctx = lt.Context(device="cpu", dtype=torch.float32, use_native=False)
    flow = lt.TaylorGreenVortex(ctx, [8] * 3, 1600.0, 0.05, stencil=lt.D3Q19())
    ctx.use_native = True
    for strategy in (lt.StreamingStrategy.PRE_STREAMING, lt.StreamingStrategy.POST_STREAMING):
        sim = reference.native_simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [], strategy)
        assert sim is not None and sim._collide_and_stream.__module__.startswith("lettuce_")
    # no generated kernel for KBC / D3Q27: the look-up declines instead of starting the reference's setup.py install
    flow27 = lt.TaylorGreenVortex(ctx, [8] * 3, 1600.0, 0.05, stencil=lt.D3Q27())
    assert reference.native_simulation(flow27, lt.BGKCollision(flow27.units.relaxation_parameter_lu), [],
                                       lt.StreamingStrategy.PRE_STREAMING) is None
