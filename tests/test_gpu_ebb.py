"""Link-wise bounce-back boundaries applied after streaming on the B200 engine (`lbm_apply_links`), against the
goldens produced by the reference's example project examples/advanced_projects/efficient_bounce_back_obstacle
(tests/golden/make_golden.py ebb_cases) and against the oracle at another size."""
import numpy as np
import pytest
import torch

from conftest import load_golden, max_rel

pytestmark = pytest.mark.gpu

lt = pytest.importorskip("lettuce_b200")
from oracle import lbm_oracle as lo  # noqa: E402
from test_ebb_oracle import CASES, ebb_setup  # noqa: E402

STENCILS = {"D2Q9": lt.D2Q9, "D3Q19": lt.D3Q19, "D3Q27": lt.D3Q27}


def cuda_ctx(dtype):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return lt.Context("cuda", dtype=dtype)


def make_flow(g, dtype):
    stencil, bc, walls, steps, diameter = g["meta"]
    res = [int(r) for r in g["res"]]
    flow = lt.ObstacleCylinder(cuda_ctx(dtype), res, 100.0, 0.05, char_length_pu=1.0, char_length_lu=float(diameter),
                               bc_type=bc, lateral_walls=walls, calc_force_coefficients=True,
                               stencil=STENCILS[stencil](), u_init=1, perturb_init=len(res) == 2)
    return flow, int(steps)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_ebb_simulation_matches_reference_golden(name, dtype):
    g = load_golden(name)
    flow, steps = make_flow(g, dtype)
    if dtype == torch.float64:
        assert max_rel(flow.f.cpu().numpy(), g["f0"]) < 1e-13
    flow.f = flow.context.convert_to_tensor(g["f0"]).contiguous()
    sim = lt.EbbSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
    assert np.array_equal(sim.no_collision_mask.cpu().numpy(), g["ncm"])
    assert np.array_equal(sim.no_streaming_mask.cpu().numpy(), g["nsm"])
    sim(steps)
    assert flow.i == steps
    assert max_rel(flow.f.cpu().numpy(), g["f"]) < (1e-12 if dtype == torch.float64 else 1e-5)
    force = sim.post_streaming_boundaries[-1].force_sum.cpu().numpy()
    assert force.shape == g["force"].shape
    assert np.max(np.abs(force - g["force"])) < (1e-11 if dtype == torch.float64 else 2e-4) * np.max(np.abs(g["force"]))


@pytest.mark.parametrize("coll", ["trt", "regularized"])
def test_ebb_with_other_collisions_matches_oracle(coll):
    """the link kernels re-evaluate the collide phase of their fluid nodes: check a collision other than BGK, a
    node with links in opposite directions (a one-node gap between two solid blocks) and the drag observable"""
    g = load_golden("ebb2d_ibb1")
    flow, steps = make_flow(g, torch.float64)
    flow.f = flow.context.convert_to_tensor(g["f0"]).contiguous()
    st, units, _, post, post_streaming, _ = ebb_setup(g)
    tau = flow.units.relaxation_parameter_lu
    collision = lt.TRTCollision(tau) if coll == "trt" else lt.RegularizedCollision()
    sim = lt.EbbSimulation(flow, collision, [])
    # add a half-way block pair with a one-node channel upstream of the cylinder (both simulations)
    block = np.zeros(g["obstacle_mask"].shape, dtype=bool)
    block[2:5, 3:5] = True
    block[2:5, 6:8] = True
    extra = lt.HalfwayBounceBackBoundary(flow.context, flow, _sbd(block), periodicity=(False, False), calc_force=True)
    sim.post_streaming_boundaries.insert(0, extra)
    label = int(sim.no_collision_mask.max()) + 1
    sim.no_collision_mask[torch.as_tensor(block, device=flow.f.device)] = label
    sim.no_streaming_mask |= torch.as_tensor(block, device=flow.f.device).to(torch.uint8)
    o_extra = lo.halfway_links(st, block, (False, False))
    o_list = [o_extra] + post_streaming
    ncm, nsm = lo.ebb_masks(st, g["f0"].shape[1:], [], post, post_streaming)
    ncm[block] = label
    nsm |= block.astype(np.uint8)[None]
    assert np.array_equal(ncm, sim.no_collision_mask.cpu().numpy())
    both = np.concatenate([o_extra["q"][:, None], o_extra["nodes"]], axis=1)
    opposite_too = {(int(st["opposite"][q]), x, y) for q, x, y in both} & {(int(q), x, y) for q, x, y in both}
    assert opposite_too, "the test geometry must contain a node with links in opposite directions"
    f = g["f0"].copy()
    cdict = dict(kind=coll, tau=tau)
    for _ in range(6):
        f, forces = lo.ebb_step(st, f, cdict, [], post, o_list, ncm, nsm)
    sim(6)
    assert max_rel(flow.f.cpu().numpy(), f) < 1e-11
    for b, want in zip(sim.post_streaming_boundaries, forces):
        got = b.force_sum.cpu().numpy()
        assert np.max(np.abs(got - want)) < 1e-10 * max(np.max(np.abs(want)), 1e-3)
    drag = lt.DragCoefficient(flow, sim.post_streaming_boundaries[-1], flow.solid_mask, area_pu=1.0)
    rho_mean = lo.rho(f)[~g["obstacle_mask"].astype(bool)].mean()
    want = forces[-1][0] / (0.5 * rho_mean * units.u_lu ** 2 * float(g["meta"][4]))
    assert float(drag().cpu()) == pytest.approx(want, rel=1e-9)


def _sbd(mask):
    sbd = lt.SolidBoundaryData()
    sbd.solid_mask = mask
    return sbd


def test_ebb_rejects_pre_streaming():
    g = load_golden("ebb2d_hwbb")
    flow, _ = make_flow(g, torch.float64)
    sim = lt.EbbSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
    sim.streaming_strategy = lt.StreamingStrategy.PRE_STREAMING
    with pytest.raises(RuntimeError):
        sim(1)


@pytest.mark.skipif(__import__("os").environ.get("LBM_B200_EXPERIMENTAL") != "1",
                    reason="lbm_step_links_n (LBM_B200_EBB_BATCH=1) is written but not yet run on hardware; "
                           "set LBM_B200_EXPERIMENTAL=1 to include it")
def test_ebb_batched_library_call_equals_step_by_step(monkeypatch):
    g = load_golden("ebb3d_hwbb_walls")
    results = []
    for batch in ("0", "1"):
        monkeypatch.setenv("LBM_B200_EBB_BATCH", batch)
        flow, steps = make_flow(g, torch.float64)
        flow.f = flow.context.convert_to_tensor(g["f0"]).contiguous()
        sim = lt.EbbSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
        drag = lt.ObservableReporter(lt.DragCoefficient(flow, sim.post_streaming_boundaries[-1], flow.solid_mask, 1.0),
                                     interval=4, out=None)
        sim.reporter.append(drag)
        sim(steps + 1)                                   # odd number of steps: buffer parity
        results.append((flow.f.clone(), [r[2] for r in drag.out]))
    assert torch.equal(results[0][0], results[1][0])
    assert results[0][1] == results[1][1]
