"""Link-wise bounce-back boundaries applied after streaming (the reference's example project
examples/advanced_projects/efficient_bounce_back_obstacle): the oracle's link search, wall distances, masks, the
C -> S -> B step and the momentum-exchange force against goldens produced by that project's own classes
(tests/golden/make_golden.py ebb_cases)."""
import numpy as np
import pytest

from conftest import load_golden, max_rel
from oracle import lbm_oracle as lo

CASES = ["ebb2d_hwbb", "ebb2d_ibb1", "ebb2d_fwbb", "ebb3d_fwbb_walls", "ebb3d_ibb1", "ebb3d_hwbb_walls"]


def ebb_setup(g):
    """(stencil tables, units, collision, post boundaries, post-streaming boundaries) of a golden case, built
    with the oracle's own link search"""
    stencil, bc, walls, steps, diameter = g["meta"]
    st = lo.stencil(stencil)
    d = st["d"]
    units = lo.Units(100.0, 0.05, characteristic_length_lu=float(diameter))
    assert units.tau == pytest.approx(float(g["tau"]), rel=1e-14)
    direction = [1] + [0] * (d - 1)
    u_inlet = np.asarray(g["u_inlet"], dtype=np.float64)
    if u_inlet.ndim == 1:
        u_inlet = u_inlet.reshape((d,) + (1,) * d)
    post = [lo.equilibrium_pu(g["in_mask"], units.pressure_pu_to_density_lu(np.zeros((1,) * (d + 1))),
                              units.velocity_to_lu(u_inlet)),
            lo.outlet_p(direction, 1.0)]
    obstacle, wall = g["obstacle_mask"].astype(bool), g["wall_mask"].astype(bool)
    periodicity = (False, False) if d == 2 else (False, False, True)
    cx, cy, radius = (float(v) for v in g["center"])
    post_streaming = []
    if walls == "bounceback":
        post_streaming.append(lo.fullway_links(st, wall, periodicity))
    if bc == "fwbb":
        post_streaming.append(lo.fullway_links(st, obstacle, periodicity))
    else:
        post_streaming.append(lo.cylinder_links(st, obstacle, cx, cy, radius,
                                                kind="halfway" if bc == "hwbb" else "interpolated"))
    return st, units, dict(kind="bgk", tau=units.tau), post, post_streaming, int(steps)


def index_rows(b):
    return np.concatenate([b["q"][:, None], b["nodes"]], axis=1)


@pytest.mark.parametrize("name", CASES)
def test_link_lists_and_masks_match_reference(name):
    g = load_golden(name)
    st, units, coll, post, post_streaming, steps = ebb_setup(g)
    bc, walls = g["meta"][1], g["meta"][2]
    obstacle = post_streaming[-1]
    if bc == "fwbb":
        assert np.array_equal(index_rows(obstacle), g["f_index_fwbb"])
    elif bc == "hwbb":
        assert np.array_equal(index_rows(obstacle), g["f_index"])
    else:
        n_lt = len(g["d_lt"])
        assert np.array_equal(index_rows(obstacle)[:n_lt], g["f_index_lt"])
        assert np.array_equal(index_rows(obstacle)[n_lt:], g["f_index_gt"])
        assert np.allclose(obstacle["d"][:n_lt], g["d_lt"], rtol=1e-13, atol=0)
        assert np.allclose(obstacle["d"][n_lt:], g["d_gt"], rtol=1e-13, atol=0)
        assert (obstacle["d"][:n_lt] <= 0.5).all() and (obstacle["d"][n_lt:] > 0.5).all()
    if walls == "bounceback":
        assert np.array_equal(index_rows(post_streaming[0]), g["wall_f_index_fwbb"])
    ncm, nsm = lo.ebb_masks(st, g["f0"].shape[1:], [], post, post_streaming)
    assert np.array_equal(ncm, g["ncm"]) and np.array_equal(nsm, g["nsm"])


@pytest.mark.parametrize("name", CASES)
def test_ebb_steps_and_force_match_reference(name):
    g = load_golden(name)
    st, units, coll, post, post_streaming, steps = ebb_setup(g)
    f, forces = lo.ebb_run(st, g["f0"].copy(), steps, coll, [], post, post_streaming)
    assert max_rel(f, g["f"]) < 1e-12
    # components that vanish by symmetry are sums of cancelling terms: tolerance relative to the force's size
    assert np.max(np.abs(forces[-1] - g["force"])) < 1e-12 * np.max(np.abs(g["force"]))


@pytest.mark.parametrize("name", CASES)
def test_package_host_side_matches_reference(name):
    """lettuce_b200's ObstacleCylinder / EbbSimulation / link boundaries on a CPU context (host logic only, no
    kernels): initial populations, label and no-streaming masks, link lists and wall distances"""
    import torch
    import lettuce_b200 as lt
    g = load_golden(name)
    stencil, bc, walls, steps, diameter = g["meta"]
    res = [int(r) for r in g["res"]]
    flow = lt.ObstacleCylinder(lt.Context("cpu", dtype=torch.float64), res, 100.0, 0.05, char_length_pu=1.0,
                               char_length_lu=float(diameter), bc_type=bc, lateral_walls=walls,
                               calc_force_coefficients=True, u_init=1, perturb_init=len(res) == 2,
                               stencil={"D2Q9": lt.D2Q9, "D3Q19": lt.D3Q19, "D3Q27": lt.D3Q27}[stencil]())
    assert max_rel(flow.f.numpy(), g["f0"]) < 1e-14
    assert flow.units.relaxation_parameter_lu == pytest.approx(float(g["tau"]), rel=1e-14)
    sim = lt.EbbSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
    assert np.array_equal(sim.no_collision_mask.numpy(), g["ncm"])
    assert np.array_equal(sim.no_streaming_mask.numpy(), g["nsm"])
    obstacle = sim.post_streaming_boundaries[-1]
    if bc == "fwbb":
        assert np.array_equal(obstacle.f_index_fwbb.numpy(), g["f_index_fwbb"])
    elif bc == "hwbb":
        assert np.array_equal(obstacle.f_index.numpy(), g["f_index"])
    else:
        assert np.array_equal(obstacle.f_index_lt.numpy(), g["f_index_lt"])
        assert np.array_equal(obstacle.f_index_gt.numpy(), g["f_index_gt"])
        assert np.allclose(obstacle.d_lt.numpy(), g["d_lt"], rtol=1e-13, atol=0)
        assert np.allclose(obstacle.d_gt.numpy(), g["d_gt"], rtol=1e-13, atol=0)
    if walls == "bounceback":
        assert np.array_equal(sim.post_streaming_boundaries[0].f_index_fwbb.numpy(), g["wall_f_index_fwbb"])
    assert [o["kind"] for o in lt.native.describe(sim)["ops"]] == [1, 17, 18]     # BGK, inlet, outlet
    with pytest.raises(AttributeError):
        lt.FullwayBounceBackBoundary(flow.context, flow, flow.wall_mask).force_sum


def test_link_search_border_rules_match_reference_on_random_masks():
    """random solid masks touching the border, every periodicity combination, a second solid body next to the
    first: the oracle's and the package's vectorised link search against the reference's loops
    (fullway_bounce_back_boundary.py:50-123; index -1 wraps, index n is skipped on non-periodic axes)"""
    import torch
    import lettuce_b200 as lt
    g = load_golden("ebb_random_links")
    for tag, stencil, cls in (("2d", "D2Q9", lt.D2Q9), ("3d", "D3Q19", lt.D3Q19), ("3d27", "D3Q27", lt.D3Q27)):
        st = lo.stencil(stencil)
        mask, other = g[f"mask_{tag}"], g[f"other_{tag}"]
        flow = lt.TaylorGreenVortex(lt.Context("cpu", dtype=torch.float64), list(mask.shape), 10, 0.05, stencil=cls())
        k = 0
        while f"fw_{tag}_{k}" in g:
            per = tuple(bool(p) for p in g[f"per_{tag}_{k}"])
            want = g[f"fw_{tag}_{k}"]
            assert len(want) > 0
            ours = lo.fullway_links(st, mask, per, mask | other)
            assert np.array_equal(index_rows(ours), want), (tag, per)
            boundary = lt.FullwayBounceBackBoundary(flow.context, flow, mask, global_solid_mask=mask | other,
                                                    periodicity=per)
            assert np.array_equal(boundary.f_index_fwbb.numpy(), want), (tag, per)
            # the half-way search stores the same links on the fluid side
            hw = lo.halfway_links(st, mask, per, mask | other)
            assert len(hw["q"]) == len(want) and np.array_equal(hw["q"], want[:, 0])
            k += 1
        assert k == 2 ** mask.ndim
