"""CPU oracle for the lattice-Boltzmann stream+collide hot path (TEST INFRASTRUCTURE ONLY).

This module is a NumPy restatement of the algorithm that lettuce's torch path
executes for one time step and for the reporter moment reductions.  It is the
checker that `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline
leg use; nothing under `lettuce_b200/` may import it.  It is written from the
algorithm, not from the reference's source, and every
function cites the reference file:line whose behaviour it restates.

Parity status: PINNED.  `tests/golden/make_golden.py` imported the reference's
torch CPU path in the build container and stored its inputs/outputs under
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this module against
every one of those vectors (fp64, <= 1e-13 relative).

Conventions
-----------
* populations `f` have shape `[q, *res]` (q slowest, last spatial axis fastest),
  exactly the reference's layout (`lettuce/_flow.py:92`).
* all index arithmetic is periodic (`torch.roll`, `lettuce/_simulation.py:241-243`).
* a *transformer list* is `pre_boundaries + [collision] + post_boundaries`
  (`lettuce/_simulation.py:70`); the integer label field `ncm` selects which
  entry acts on a node, `nsm[q, x]` freezes slot `(q, x)` during streaming.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# velocity sets (lettuce/ext/_stencil/d2q9.py:8-10, d3q19.py:8-13, d3q27.py:8-12)
# --------------------------------------------------------------------------


def _axis_pairs(d):
    """+axis / -axis unit vectors in lettuce's ordering for the 3-D sets."""
    out = []
    for a in range(d):
        for s in (1, -1):
            v = [0] * d
            v[a] = s
            out.append(v)
    return out


def stencil(name: str):
    """Return dict(e=int array [q,d], w=float64 [q], opposite=int [q], d, q)."""
    name = name.upper()
    if name == "D2Q9":
        e = [[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1],
             [1, 1], [-1, 1], [-1, -1], [1, -1]]
        w = [4 / 9] + [1 / 9] * 4 + [1 / 36] * 4
    elif name in ("D3Q19", "D3Q27"):
        e = [[0, 0, 0]] + _axis_pairs(3)
        # face diagonals, listed as (v, -v) pairs: yz, yz', xz, xz', xy, xy'
        for v in ([0, 1, 1], [0, 1, -1], [1, 0, 1], [1, 0, -1], [1, 1, 0], [1, -1, 0]):
            e += [v, [-c for c in v]]
        w = [1 / 3] + [1 / 18] * 6 + [1 / 36] * 12
        if name == "D3Q27":
            for v in ([1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1]):
                e += [v, [-c for c in v]]
            w = [8 / 27] + [2 / 27] * 6 + [1 / 54] * 12 + [1 / 216] * 8
    else:
        raise ValueError(f"unknown stencil {name}")
    e = np.asarray(e, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64)
    opposite = np.array([int(np.flatnonzero((e == -e[i]).all(axis=1))[0])
                         for i in range(len(e))])
    return dict(name=name, e=e, w=w, opposite=opposite, d=e.shape[1], q=e.shape[0])


# Python floats (not np.float64 scalars) so that float32 populations stay float32 under NumPy's
# promotion rules, like torch keeps the context dtype when multiplying by Python scalars.
CS = float(1.0 / np.sqrt(3.0))    # lettuce/_stencil.py:19
CS2 = CS ** 2                     # the reference squares the rounded 1/sqrt(3)


# --------------------------------------------------------------------------
# units (lettuce/_unit.py:34-68, 94-101)
# --------------------------------------------------------------------------
class Units:
    """Subset of UnitConversion used by the hot path (lettuce/_unit.py:13-145)."""

    def __init__(self, reynolds_number, mach_number, characteristic_length_lu,
                 characteristic_length_pu=1.0, characteristic_velocity_pu=1.0,
                 characteristic_density_lu=1.0, characteristic_density_pu=1.0):
        self.re = reynolds_number
        self.ma = mach_number
        self.l_lu = characteristic_length_lu
        self.l_pu = characteristic_length_pu
        self.u_pu = characteristic_velocity_pu
        self.rho_lu = characteristic_density_lu
        self.rho_pu = characteristic_density_pu

    @property
    def u_lu(self):                       # _unit.py:34-36
        return CS * self.ma

    @property
    def p_char_pu(self):                  # _unit.py:38-41
        return self.rho_pu * self.u_pu ** 2

    @property
    def p_char_lu(self):                  # _unit.py:43-46
        return self.rho_lu * self.u_lu ** 2

    @property
    def tau(self):                        # _unit.py:48-60
        return (self.l_lu * self.u_lu / self.re) / CS2 + 0.5

    def velocity_to_lu(self, u_pu):       # _unit.py:66-68
        return u_pu / self.u_pu * self.u_lu

    def velocity_to_pu(self, u_lu):       # _unit.py:62-64
        return u_lu / self.u_lu * self.u_pu

    def pressure_pu_to_density_lu(self, p_pu):   # _unit.py:99-101, 119-121
        return (p_pu / self.p_char_pu * self.p_char_lu) / CS2 + self.rho_lu

    def length_to_pu(self, l_lu):         # _unit.py:123-125
        return l_lu * self.l_pu / self.l_lu

    def time_to_pu(self, t_lu):           # _unit.py:84-87
        return t_lu / (self.l_lu / self.u_lu) * (self.l_pu / self.u_pu)

    def incompressible_energy_to_pu(self, e_lu):  # _unit.py:137-140
        return e_lu * self.u_pu ** 2 / self.u_lu ** 2


# --------------------------------------------------------------------------
# moments and equilibrium
# --------------------------------------------------------------------------
def rho(f):
    """Density, sum over q (lettuce/_flow.py:157-159)."""
    return f.sum(axis=0)


def j(st, f):
    """Momentum sum_q e_q f_q (lettuce/_flow.py:173-176)."""
    return np.tensordot(st["e"].T.astype(f.dtype), f, axes=1)


def u(st, f):
    """Velocity j/rho without force correction (lettuce/_flow.py:178-193)."""
    return j(st, f) / rho(f)[None]


def equilibrium(st, rho_, u_):
    """Second-order equilibrium (lettuce/ext/_equilibrium/quadratic_equilibrium.py:11-24).

    `rho_` broadcasts against the spatial shape, `u_` has a leading axis d.
    """
    u_ = np.asarray(u_)
    e = st["e"].astype(u_.dtype)
    eu = np.tensordot(e, u_, axes=1)                       # [q, ...]
    uu = (u_ * u_).sum(axis=0)                              # [...]
    w = st["w"].astype(u_.dtype).reshape((-1,) + (1,) * (eu.ndim - 1))
    return w * (rho_ * ((2 * eu - uu) / (2 * CS2) + 0.5 * (eu / CS2) ** 2 + 1))


# --------------------------------------------------------------------------
# collisions
# --------------------------------------------------------------------------
def collide_none(st, f, **_):
    """NoCollision (lettuce/ext/_collision/no_collision.py:9-11)."""
    return f.copy()


def collide_bgk(st, f, tau, **_):
    """Single relaxation time (lettuce/ext/_collision/bgk_collision.py:17-22), no force."""
    feq = equilibrium(st, rho(f), u(st, f))
    return f - (1.0 / tau) * (f - feq)


def collide_bgk_forced(st, f, tau, acceleration, scheme="guo", tau_force=None, **_):
    """BGK with a body force (lettuce/ext/_collision/bgk_collision.py:17-22).  Guo (ext/_force/guo.py:16-38):
    u_eq = a/(2 rho), source (1 - 1/(2 tau_f)) w [(e - u)/cs^2 + (e.u) e/cs^4] . a ; ShanChen
    (ext/_force/shan_chen.py:13-26): u_eq = tau_f a / rho, no source."""
    tau_force = tau if tau_force is None else tau_force
    d = st["d"]
    a = np.asarray(acceleration, dtype=f.dtype).reshape((d,) + (1,) * d)
    r = rho(f)
    scale = 0.5 if scheme == "guo" else tau_force
    uu = u(st, f) + scale * a / r
    feq = equilibrium(st, r, uu)
    out = f - (1.0 / tau) * (f - feq)
    if scheme == "guo":
        e = st["e"].astype(f.dtype)
        emu = e.reshape((st["q"], d) + (1,) * d) - uu[None]
        eu = np.tensordot(e, uu, axes=1)
        eeu = e.reshape((st["q"], d) + (1,) * d) * eu[:, None]
        term = ((emu / CS2 + eeu / CS2 ** 2) * a[None]).sum(axis=1)
        out = out + (1 - 1 / (2 * tau_force)) * st["w"].astype(f.dtype).reshape((-1,) + (1,) * d) * term
    return out


def collide_trt(st, f, tau, tau_minus=1.0, **_):
    """Two relaxation times (lettuce/ext/_collision/trt_collision.py:16-27)."""
    feq = equilibrium(st, rho(f), u(st, f))
    o = st["opposite"]
    even = ((f + f[o]) - (feq + feq[o])) / (2.0 * tau)
    odd = ((f - f[o]) - (feq - feq[o])) / (2.0 * tau_minus)
    return f - (even + odd)


def _kbc_delta_s(st, fn):
    """Shear part of a non-equilibrium vector from its raw second moments.

    Closed form of `compute_s_seq_from_m_{2d,3d}` applied to f and feq and
    subtracted (lettuce/ext/_collision/kbc_collision.py:44-94, 129-133); the
    reference normalises the moments by rho and multiplies rho back in, which
    cancels.
    """
    e = st["e"].astype(fn.dtype)
    d = st["d"]

    def P(a, b):
        return np.tensordot(e[:, a] * e[:, b], fn, axes=1)

    ds = np.zeros_like(fn)
    if d == 2:
        T = P(0, 0) + P(1, 1)
        N = P(0, 0) - P(1, 1)
        Pxy = P(0, 1)
        ds[0] = -T
        ds[1] = ds[3] = 0.5 * (0.5 * (T + N))
        ds[2] = ds[4] = 0.5 * (0.5 * (T - N))
        ds[5] = ds[7] = 0.25 * Pxy
        ds[6] = ds[8] = -0.25 * Pxy
    else:
        T = P(0, 0) + P(1, 1) + P(2, 2)
        Nxz = P(0, 0) - P(2, 2)
        Nyz = P(1, 1) - P(2, 2)
        Pxy, Pxz, Pyz = P(0, 1), P(0, 2), P(1, 2)
        ds[0] = -T
        ds[1] = ds[2] = (2 * Nxz - Nyz + T) / 6.0
        ds[3] = ds[4] = (2 * Nyz - Nxz + T) / 6.0
        ds[5] = ds[6] = (-Nxz - Nyz + T) / 6.0
        ds[7] = ds[8] = 0.25 * Pyz
        ds[9] = ds[10] = -0.25 * Pyz
        ds[11] = ds[12] = 0.25 * Pxz
        ds[13] = ds[14] = -0.25 * Pxz
        ds[15] = ds[16] = 0.25 * Pxy
        ds[17] = ds[18] = -0.25 * Pxy
    return ds


def collide_kbc(st, f, tau, **_):
    """Entropic KBC (lettuce/ext/_collision/kbc_collision.py:96-160).

    `tau` must be `units.relaxation_parameter_lu`: the reference overwrites the
    constructor argument with it on first call (kbc_collision.py:97-99).
    """
    if st["name"] not in ("D2Q9", "D3Q27"):
        raise NotImplementedError("KBC exists for D2Q9 and D3Q27 only (kbc_collision.py:101,116)")
    beta = 1.0 / (2.0 * tau)
    feq = equilibrium(st, rho(f), u(st, f))
    fn = f - feq
    ds = _kbc_delta_s(st, fn)
    dh = fn - ds
    with np.errstate(divide="ignore", invalid="ignore"):
        sum_s = (ds * dh / feq).sum(axis=0)
        sum_h = (dh * dh / feq).sum(axis=0)
        gamma = 1.0 / beta - (2.0 - 1.0 / beta) * sum_s / sum_h
    gamma = np.where(gamma < 1e-15, 2.0, gamma)              # kbc_collision.py:154
    gamma = np.where(np.isnan(gamma), 2.0, gamma)            # kbc_collision.py:156
    return f - beta * (2.0 * ds + gamma * dh)


def _second_moments(st, g):
    """Pi_ab = sum_q g_q e_qa e_qb (Flow.shear_tensor, lettuce/_flow.py:230-237)"""
    e = st["e"].astype(g.dtype)
    return np.einsum("q...,qa,qb->ab...", g, e, e)


def collide_regularized(st, f, tau, **_):
    """Regularized LBM of Latt & Chopard (lettuce/ext/_collision/regularized_collision.py:17-43):
    f = feq + (1 - 1/tau) w_q Q_q:Pi_neq / (2 cs^4).  `tau` must be units.relaxation_parameter_lu
    (the reference overwrites the constructor argument with it, regularized_collision.py:19)."""
    d = st["d"]
    feq = equilibrium(st, rho(f), u(st, f))
    pi = _second_moments(st, f - feq)
    e = st["e"].astype(f.dtype)
    Q = np.einsum("qa,qb->qab", e, e) - np.eye(d, dtype=f.dtype) * CS2
    w = st["w"].astype(f.dtype).reshape((-1,) + (1,) * d)
    fi1 = w * np.einsum("qab,ab...->q...", Q, pi) / (2 * CS2 ** 2)
    return feq + (1.0 - 1.0 / tau) * fi1


def collide_smagorinsky(st, f, tau, constant=0.17, **_):
    """Smagorinsky LES on top of BGK (lettuce/ext/_collision/smagorinsky_collision.py:22-40), no force:
    two fixed-point iterations for the effective relaxation time, then BGK with it."""
    r = rho(f)
    feq = equilibrium(st, r, u(st, f))
    s_shear = _second_moments(st, f - feq) / (2.0 * r * CS2)
    tau_eff = tau
    nu = (tau - 0.5) / 3.0
    for _ in range(2):
        s = s_shear / tau_eff
        ss = (s * s).sum(axis=(0, 1))
        tau_eff = (nu + constant ** 2 * ss) * 3.0 + 0.5
    return f - 1.0 / tau_eff * (f - feq)


COLLISIONS = {"none": collide_none, "bgk": collide_bgk, "trt": collide_trt, "kbc": collide_kbc,
              "regularized": collide_regularized, "smagorinsky": collide_smagorinsky,
              "bgk_forced": collide_bgk_forced}


# --------------------------------------------------------------------------
# boundaries: each is a dict(kind=..., params)
# --------------------------------------------------------------------------
def bounce_back(mask):
    """BounceBackBoundary(mask) (lettuce/ext/_boundary/bounce_back_boundary.py:10-32)."""
    return dict(kind="bounce_back", mask=np.asarray(mask, dtype=bool))


def equilibrium_pu(mask, rho_lu, u_lu):
    """EquilibriumBoundaryPU already converted to lattice units
    (lettuce/ext/_boundary/equilibrium_boundary_pu.py:79-84).  `rho_lu`
    broadcasts to `[1,*res]`, `u_lu` to `[d,*res]`."""
    return dict(kind="equilibrium_pu", mask=np.asarray(mask, dtype=bool),
                rho=np.asarray(rho_lu), u=np.asarray(u_lu))


def outlet_p(direction, rho_outlet=1.0):
    """EquilibriumOutletP(direction, rho_outlet)
    (lettuce/ext/_boundary/equilibrium_outlet_p.py:16-61)."""
    direction = [int(c) for c in direction]
    assert direction.count(0) == len(direction) - 1 and ((1 in direction) ^ (-1 in direction))
    return dict(kind="outlet_p", direction=direction, rho=float(rho_outlet))


def anti_bounce_back_outlet(direction):
    """AntiBounceBackOutlet(direction) (lettuce/ext/_boundary/anti_bounce_back_outlet.py:22-69)."""
    direction = [int(c) for c in direction]
    assert direction.count(0) == len(direction) - 1 and ((1 in direction) ^ (-1 in direction))
    return dict(kind="anti_bounce_back", direction=direction)


def _plane_index(direction):
    """(here, neighbour) spatial index tuples (equilibrium_outlet_p.py:36-49)."""
    here, other = [], []
    for c in direction:
        if c == 0:
            here.append(slice(None)); other.append(slice(None))
        elif c == 1:
            here.append(-1); other.append(-2)
        else:
            here.append(0); other.append(1)
    return tuple(here), tuple(other)


def _outgoing(st, direction):
    """q with e_q . direction == 1 (equilibrium_outlet_p.py:32-34)."""
    return np.flatnonzero(st["e"] @ np.asarray(direction) > 1 - 1e-6)


def boundary_masks(st, res, b):
    """(no_collision_mask [*res] bool or None, no_streaming_mask [q,*res] bool or None)
    of one boundary (bounce_back_boundary.py:20-26, equilibrium_boundary_pu.py:86-92,
    equilibrium_outlet_p.py:75-85, anti_bounce_back_outlet.py:93-103)."""
    k = b["kind"]
    if k in ("bounce_back", "equilibrium_pu"):
        return np.broadcast_to(b["mask"], res).copy(), None
    here, _ = _plane_index(b["direction"])
    ncm = np.zeros(res, dtype=bool)
    ncm[here] = True
    nsm = np.zeros((st["q"], *res), dtype=bool)
    out = _outgoing(st, b["direction"])
    if k == "outlet_p":
        frozen = np.setdiff1d(np.arange(st["q"]), out)
    else:
        frozen = st["opposite"][out]
    nsm[(frozen,) + here] = True
    return ncm, nsm


def build_masks(st, res, pre, post):
    """Label field and no-stream mask (lettuce/_simulation.py:100-146).

    Returns (ncm uint8 [*res], nsm uint8 [q,*res]) or (None, None) when there is
    no boundary.  NOTE the reference fills *both* with `collision_index`
    (:104-107); this is restated faithfully.
    """
    if len(pre) + len(post) == 0:
        return None, None
    ci = len(pre)
    ncm = np.full(res, ci, dtype=np.uint8)
    nsm = np.full((st["q"], *res), ci, dtype=np.uint8)
    for i, b in list(enumerate(pre)) + list(enumerate(post, start=ci + 1)):
        m, s = boundary_masks(st, res, b)
        if m is not None:
            ncm[m] = i
        if s is not None:
            nsm |= s.astype(np.uint8)
    return ncm, nsm


def apply_boundary(st, f, b):
    """Full-grid result of one boundary operator; may modify `f` in place exactly
    where the reference does (outlet_p: equilibrium_outlet_p.py:63-73)."""
    k = b["kind"]
    if k == "bounce_back":
        return f[st["opposite"]]
    if k == "equilibrium_pu":
        feq = equilibrium(st, np.asarray(b["rho"], dtype=f.dtype), np.asarray(b["u"], dtype=f.dtype))
        return np.broadcast_to(feq, f.shape).astype(f.dtype)
    here, other = _plane_index(b["direction"])
    if k == "outlet_p":
        uu = u(st, f)
        u_w = uu[(slice(None),) + other]
        f[(slice(None),) + here] = equilibrium(st, np.asarray(b["rho"], dtype=f.dtype), u_w)
        return f.copy()
    if k == "anti_bounce_back":
        # anti_bounce_back_outlet.py:71-91
        uu = u(st, f)
        rr = rho(f)
        u_here = uu[(slice(None),) + here]
        u_w = u_here + 0.5 * (u_here - uu[(slice(None),) + other])
        out = _outgoing(st, b["direction"])
        e = st["e"][out].astype(f.dtype)
        w = st["w"][out].astype(f.dtype).reshape((-1,) + (1,) * (u_w.ndim - 1))
        eu = np.tensordot(e, u_w, axes=1)
        unorm2 = (u_w * u_w).sum(axis=0)
        new = (-f[(out,) + here]
               + w * rr[here][None] * (2 + eu ** 2 / CS2 ** 2 - unorm2 / CS2))
        f[(st["opposite"][out],) + here] = new
        return f.copy()
    raise ValueError(k)


# --------------------------------------------------------------------------
# stream / collide / step
# --------------------------------------------------------------------------
def stream(st, f, nsm=None):
    """Periodic streaming (lettuce/_simulation.py:241-256): slot (q,x) receives
    f_q(x - e_q) unless nsm[q,x] == 1."""
    d = st["d"]
    out = f.copy()
    for i in range(1, st["q"]):
        moved = np.roll(f[i], shift=tuple(int(c) for c in st["e"][i]), axis=tuple(range(d)))
        out[i] = moved if nsm is None else np.where(nsm[i] == 1, f[i], moved)
    return out


def collide(st, f, collision, pre=(), post=(), ncm=None):
    """`Simulation._collide` (lettuce/_simulation.py:258-305).  `collision` is
    dict(kind=..., tau=..., tau_minus=...)."""
    f = f.copy()
    coll = lambda g: COLLISIONS[collision["kind"]](st, g, **{k: v for k, v in collision.items() if k != "kind"})
    ci = len(pre)
    if ncm is None:
        for b in pre:
            f = apply_boundary(st, f, b)
        f = coll(f)
        for b in post:
            f = apply_boundary(st, f, b)
        return f
    for i, b in enumerate(pre):
        f = np.where(ncm[None] == i, apply_boundary(st, f, b), f)
    f = np.where(ncm[None] == ci, coll(f), f)
    for i, b in enumerate(post, start=ci + 1):
        r = apply_boundary(st, f, b)        # may touch f in place first, like the reference
        f = np.where(ncm[None] == i, r, f)
    return f


STRATEGIES = {"NO_STREAMING": (False, False), "PRE_STREAMING": (True, False),
              "POST_STREAMING": (False, True), "DOUBLE_STREAMING": (True, True)}


def step(st, f, collision, pre=(), post=(), ncm=None, nsm=None, strategy="POST_STREAMING"):
    """One time step for the given StreamingStrategy (lettuce/_simulation.py:149-166)."""
    pre_s, post_s = STRATEGIES[strategy]
    if pre_s:
        f = stream(st, f, nsm)
    f = collide(st, f, collision, pre, post, ncm)
    if post_s:
        f = stream(st, f, nsm)
    return f


def step_parallel(st, f, collision, strategy="POST_STREAMING", pool=None, chunks=8):
    """`step` for flows WITHOUT boundaries on several host cores: the node-local collide phase is split
    into x-chunks and the per-population rolls are spread over the thread pool `pool` (NumPy releases the
    GIL inside its loops).  Used by bench.py's CPU baseline so that the port uses all host threads, as the
    reference's torch path does through OpenMP."""
    if pool is None:
        return step(st, f, collision, strategy=strategy)
    pre_s, post_s = STRATEGIES[strategy]
    d = st["d"]
    args = {k: v for k, v in collision.items() if k != "kind"}
    coll = COLLISIONS[collision["kind"]]

    def stream_all(g):
        out = np.empty_like(g)
        out[0] = g[0]

        def one(i):
            out[i] = np.roll(g[i], shift=tuple(int(c) for c in st["e"][i]), axis=tuple(range(d)))
        list(pool.map(one, range(1, st["q"])))
        return out

    def collide_all(g):
        out = np.empty_like(g)
        edges = np.linspace(0, g.shape[1], chunks + 1).astype(int)

        def one(k):
            a, b = edges[k], edges[k + 1]
            if b > a:
                out[:, a:b] = coll(st, g[:, a:b], **args)
        list(pool.map(one, range(chunks)))
        return out

    if pre_s:
        f = stream_all(f)
    f = collide_all(f)
    if post_s:
        f = stream_all(f)
    return f


def run(st, f, nsteps, collision, pre=(), post=(), strategy="POST_STREAMING"):
    """`Simulation.__call__` without reporters (lettuce/_simulation.py:311-323)."""
    res = f.shape[1:]
    ncm, nsm = build_masks(st, res, list(pre), list(post))
    for _ in range(nsteps):
        f = step(st, f, collision, pre, post, ncm, nsm, strategy)
    return f


# --------------------------------------------------------------------------
# link-wise bounce-back boundaries applied AFTER streaming
# (examples/advanced_projects/efficient_bounce_back_obstacle; `ebb/` below is that directory)
# --------------------------------------------------------------------------
def solid_fluid_links(st, mask, periodicity=None, other_solid=None):
    """All (q_in, solid node, fluid node) triples of a solid `mask`: q_in points from the fluid node into the solid
    node.  Restates the neighbour search of ebb/boundary/fullway_bounce_back_boundary.py:50-123 and
    halfway_bounce_back_boundary.py:65-152 INCLUDING its index quirk: a neighbour index of -1 on a non-periodic
    axis silently wraps (NumPy negative indexing) while an index of n raises IndexError and is skipped.
    `other_solid` marks nodes that belong to other solid boundaries (no link towards them).  Order: solid nodes in
    C order, q ascending (the reference's loop order)."""
    mask = np.asarray(mask, dtype=bool)
    res, d = mask.shape, st["d"]
    periodicity = tuple(periodicity) if periodicity is not None else (False,) * d
    blocked = mask if other_solid is None else (mask | np.asarray(other_solid, dtype=bool))
    solid = np.argwhere(mask)                                    # [n, d], C order
    q_in, s_nodes, f_nodes = [], [], []
    for i in range(st["q"]):
        nb = solid + st["e"][i][None, :]
        ok = np.ones(len(solid), dtype=bool)
        for a in range(d):
            if periodicity[a]:
                nb[:, a] %= res[a]
            else:
                ok &= nb[:, a] < res[a]                          # IndexError in the reference: skipped
                nb[:, a] = np.where(nb[:, a] < 0, nb[:, a] + res[a], nb[:, a])   # negative index wraps
        nbc = np.where(ok[:, None], nb, 0)
        ok &= ~blocked[tuple(nbc.T)]
        q_in.append(np.full(int(ok.sum()), st["opposite"][i]))
        s_nodes.append(solid[ok])
        f_nodes.append(nb[ok])
    q_in, s_nodes, f_nodes = np.concatenate(q_in), np.concatenate(s_nodes), np.concatenate(f_nodes)
    # reference order: outer loop over solid nodes (C order), inner loop over stencil direction i
    flat = np.ravel_multi_index(tuple(s_nodes.T), res)
    order = np.lexsort((st["opposite"][q_in], flat))            # i = opposite[q_in]
    return q_in[order], s_nodes[order], f_nodes[order]


def cylinder_wall_distance(st, q_in, f_nodes, x_center, y_center, radius):
    """Distance d in (0, 1] (in units of the link length) from the fluid node to a circular cylinder's surface along
    link q_in (ebb/flow/obstacle_cylinder.py:322-365: the p-q formula; the first root that is <= 1 wins)."""
    c = st["e"][q_in][:, :2].astype(float)
    px, py = f_nodes[:, 0].astype(float), f_nodes[:, 1].astype(float)
    cc = c[:, 0] ** 2 + c[:, 1] ** 2
    h1 = (px * c[:, 0] + py * c[:, 1] - c[:, 0] * x_center - c[:, 1] * y_center) / cc
    h2 = (px * px + py * py + x_center ** 2 + y_center ** 2 - 2 * px * x_center - 2 * py * y_center - radius ** 2) / cc
    root = np.sqrt(h1 * h1 - h2)
    d1, d2 = -h1 + root, -h1 - root
    return np.where(d1 <= 1, d1, d2)


def fullway_links(st, mask, periodicity=None, global_solid_mask=None):
    """dict(kind='fullway'): links stored on the SOLID node (fullway_bounce_back_boundary.py:20-123)."""
    mask = np.asarray(mask, dtype=bool)
    other = None if global_solid_mask is None else np.where(~mask, np.asarray(global_solid_mask, dtype=bool), False)
    q_in, s_nodes, _ = solid_fluid_links(st, mask, periodicity, other)
    return dict(kind="fullway", mask=mask, q=q_in, nodes=s_nodes)


def halfway_links(st, mask, periodicity=None, global_solid_mask=None):
    """dict(kind='halfway'): links stored on the FLUID node, legacy neighbour search
    (halfway_bounce_back_boundary.py:60-152)."""
    mask = np.asarray(mask, dtype=bool)
    other = None if global_solid_mask is None else np.asarray(global_solid_mask, dtype=bool)
    q_in, _, f_nodes = solid_fluid_links(st, mask, periodicity, other)
    return dict(kind="halfway", mask=mask, q=q_in, nodes=f_nodes)


def cylinder_links(st, mask, x_center, y_center, radius, kind="interpolated"):
    """Link lists of ObstacleCylinder.make_ibb_index_lists (obstacle_cylinder.py:283-466): every axis is treated
    as periodic in the neighbour search, d <= 0.5 links first ('lt'), then d > 0.5 ('gt'), each in loop order.
    kind='halfway' uses the same links without the distances (halfway_bounce_back_boundary.py:48-56)."""
    mask = np.asarray(mask, dtype=bool)
    q_in, _, f_nodes = solid_fluid_links(st, mask, (True,) * st["d"])
    dist = cylinder_wall_distance(st, q_in, f_nodes, x_center, y_center, radius)
    lt = dist <= 0.5
    order = np.concatenate([np.flatnonzero(lt), np.flatnonzero(~lt)])
    return dict(kind=kind, mask=mask, q=q_in[order], nodes=f_nodes[order], d=dist[order])


def ebb_masks(st, res, pre, post, post_streaming):
    """Masks of EbbSimulation (ebb/simulation/ebb_simulation.py:22-57): the post-streaming boundaries continue
    the label numbering; fullway has no no-streaming mask, halfway / interpolated freeze every slot of their solid
    nodes."""
    ncm, nsm = build_masks(st, res, list(pre), list(post))
    if not post_streaming:
        return ncm, nsm
    ci = len(pre)
    if ncm is None:
        ncm = np.full(res, ci, dtype=np.uint8)
        nsm = np.full((st["q"], *res), ci, dtype=np.uint8)
    for i, b in enumerate(post_streaming, start=ci + 1 + len(post)):
        ncm[b["mask"]] = i
        if b["kind"] != "fullway":
            nsm |= b["mask"].astype(np.uint8)[None]
    return ncm, nsm


def apply_links(st, f, fc, b):
    """One post-streaming boundary on the streamed populations `f` (in place); `fc` are the populations between
    collision and streaming.  Returns the momentum-exchange force [d] in lattice units
    (fullway_bounce_back_boundary.py:132-170, halfway_bounce_back_boundary.py:167-215,
    linear_interpolated_bounce_back_boundary.py:57-143)."""
    q, nodes = b["q"], tuple(b["nodes"].T)
    opp = st["opposite"][q]
    e = st["e"][q].astype(f.dtype)
    if b["kind"] == "fullway":
        incoming = f[(q,) + nodes]
        force = 2 * (incoming[:, None] * e).sum(axis=0)
        f[(opp,) + nodes] = incoming
        return force
    fcq = fc[(q,) + nodes]
    if b["kind"] == "halfway":
        force = 2 * (fcq[:, None] * e).sum(axis=0)
        f[(opp,) + nodes] = fcq
        return force
    d = b["d"].astype(f.dtype)
    lt = b["d"] <= 0.5
    bounced_lt = 2 * d * fcq + (1 - 2 * d) * f[(q,) + nodes]
    with np.errstate(divide="ignore", invalid="ignore"):
        bounced_gt = (1 / (2 * d)) * fcq + (1 - 1 / (2 * d)) * fc[(opp,) + nodes]
    f[(opp[lt],) + tuple(c[lt] for c in nodes)] = bounced_lt[lt]          # d <= 0.5 first, then d > 0.5
    f[(opp[~lt],) + tuple(c[~lt] for c in nodes)] = bounced_gt[~lt]
    bounced = f[(opp,) + nodes]
    return ((fcq + bounced)[:, None] * e).sum(axis=0)


def ebb_step(st, f, collision, pre=(), post=(), post_streaming=(), ncm=None, nsm=None):
    """One EbbSimulation step: collide, remember the collided populations, stream, post-streaming boundaries in
    order (ebb_simulation.py:71-104).  Returns (f, [force of each post-streaming boundary])."""
    fc = collide(st, f, collision, pre, post, ncm)
    f = stream(st, fc, nsm)
    forces = [apply_links(st, f, fc, b) for b in post_streaming]
    return f, forces


def ebb_run(st, f, nsteps, collision, pre=(), post=(), post_streaming=()):
    ncm, nsm = ebb_masks(st, f.shape[1:], pre, post, post_streaming)
    forces = []
    for _ in range(nsteps):
        f, forces = ebb_step(st, f, collision, pre, post, post_streaming, ncm, nsm)
    return f, forces


# --------------------------------------------------------------------------
# reporter reductions (lettuce/ext/_reporter/observable_reporter.py:27-68,140-158)
# --------------------------------------------------------------------------
_W6 = (-1 / 60, 3 / 20, -3 / 4, 3 / 4, -3 / 20, 1 / 60)
_S6 = (3, 2, 1, -1, -2, -3)


def gradient6(a, dx=1.0):
    """6th-order periodic central difference along every axis
    (lettuce/util/utility.py:37-99, order=6); returns [ndim, *a.shape]."""
    out = []
    for ax in range(a.ndim):
        g = sum(wk * np.roll(a, sk, axis=ax) for wk, sk in zip(_W6, _S6))
        out.append(g * (1.0 / dx))
    return np.stack(out)


def incompressible_kinetic_energy(st, f, units):
    """observable_reporter.py:34-42 with flow.incompressible_energy (_flow.py:200-204)."""
    uu = u(st, f)
    e_lu = (0.5 * (uu * uu).sum(axis=0)).sum()
    return units.incompressible_energy_to_pu(e_lu) * units.length_to_pu(1.0) ** st["d"]


def maximum_velocity(st, f, units):
    """observable_reporter.py:27-31."""
    up = units.velocity_to_pu(u(st, f))
    return np.sqrt((up * up).sum(axis=0)).max()


def enstrophy(st, f, units):
    """observable_reporter.py:45-68."""
    up = units.velocity_to_pu(u(st, f))
    dx = units.length_to_pu(1.0)
    g = [gradient6(up[a], dx) for a in range(st["d"])]
    vort = ((g[0][1] - g[1][0]) ** 2).sum()
    if st["d"] == 3:
        vort += ((g[2][1] - g[1][2]) ** 2 + (g[0][2] - g[2][0]) ** 2).sum()
    return vort * dx ** st["d"]


def energy_spectrum(st, f, units):
    """observable_reporter.py:71-137: shell sums of 0.5 |fft(u_pu)/norm|^2 over k-0.5 < |k| <= k+0.5 for
    k = 0 .. int(max |k|) - 1; `norm` uses the FIRST resolution entry (:86-89)."""
    res = f.shape[1:]
    d = st["d"]
    dx = units.length_to_pu(1.0)
    freqs = [np.fft.fftfreq(n, d=1.0 / n) for n in res]
    knorm = np.sqrt(sum(k * k for k in np.meshgrid(*freqs, indexing="ij")))
    norm = res[0] * np.sqrt(2 * np.pi) / dx ** 2 if d == 3 else res[0] / dx
    up = units.velocity_to_pu(u(st, f))
    uh = np.stack([np.fft.fftn(up[a]) for a in range(d)]) / norm
    ekin = (0.5 * (uh.imag ** 2 + uh.real ** 2)).sum(axis=0)
    shells = np.arange(int(knorm.max()))
    return np.array([ekin[(knorm > k - 0.5) & (knorm <= k + 0.5)].sum() for k in shells])


def mass(f, no_mass_mask=None):
    """observable_reporter.py:140-158 (drops the border of the last two axes)."""
    m = f[..., 1:-1, 1:-1].sum()
    if no_mass_mask is not None:
        m -= (f * no_mass_mask.astype(np.float32)).sum()
    return m


# --------------------------------------------------------------------------
# flows used by the configs (initial conditions, host side)
# --------------------------------------------------------------------------
def tgv_units(res, reynolds_number, mach_number):
    """TaylorGreenVortex.make_units (lettuce/ext/_flows/taylorgreen.py:43-50)."""
    return Units(reynolds_number, mach_number, characteristic_length_lu=res[0] / (2 * np.pi))


def tgv_initial(st, res, reynolds_number, mach_number, dtype=np.float64, fneq=True):
    """TaylorGreenVortex initial populations (taylorgreen.py:52-94, _flow.py:127-143,341-367)."""
    d = st["d"]
    units = tgv_units(res, reynolds_number, mach_number)
    axes = [np.linspace(0, 2 * np.pi * (1 - 1 / n), n, dtype=dtype) for n in res]
    g = np.meshgrid(*axes, indexing="ij")
    if d == 2:
        u_pu = np.stack([np.cos(g[0]) * np.sin(g[1]), -np.sin(g[0]) * np.cos(g[1])])
        p_pu = -0.25 * (np.cos(2 * g[0]) + np.cos(2 * g[1]))
    else:
        u_pu = np.stack([np.sin(g[0]) * np.cos(g[1]) * np.cos(g[2]),
                         -np.cos(g[0]) * np.sin(g[1]) * np.cos(g[2]),
                         np.zeros_like(g[0])])
        p_pu = 1 / 16.0 * (np.cos(2 * g[0]) + np.cos(2 * g[1])) * (np.cos(2 * g[2]) + 2)
    rho0 = units.pressure_pu_to_density_lu(p_pu).astype(dtype)
    u0 = units.velocity_to_lu(u_pu).astype(dtype)
    f = equilibrium(st, rho0, u0).astype(dtype)
    if fneq:
        f = initialize_f_neq(st, f, units.tau).astype(dtype)
    return f, units


def initialize_f_neq(st, f, tau):
    """First-order non-equilibrium initialisation (lettuce/_flow.py:341-367)."""
    d = st["d"]
    r = rho(f)
    uu = u(st, f)
    S = np.stack([gradient6(uu[a], 1.0) for a in range(d)])        # S[a, b] = d_b u_a
    Pi1 = 1.0 * tau * r * S / CS2
    e = st["e"].astype(f.dtype)
    # quirk restated faithfully: the reference builds the identity with torch.eye's
    # default dtype (float32) before scaling by cs^2 (_flow.py:358-360), so the
    # trace part of Q carries an fp32-rounded cs^2 even in an fp64 run.
    Q = np.einsum("ia,ib->iab", e, e) - (np.eye(d, dtype=np.float32) * np.float32(CS2)).astype(f.dtype)
    PiQ = np.einsum("ab...,iab->i...", Pi1, Q)
    fneq = st["w"].astype(f.dtype).reshape((-1,) + (1,) * d) * PiQ
    return equilibrium(st, r, uu) - fneq


def obstacle_setup(st, res, reynolds_number=100, mach_number=0.05, dtype=np.float64):
    """The C4/C5 obstacle flow of BASELINE.md section 5 (`make_obstacle`):
    lt.Obstacle (lettuce/ext/_flows/obstacle.py:54-105) with post_boundaries
    [EquilibriumBoundaryPU(x==0, u=U e_x), EquilibriumOutletP(+x, rho=1), BounceBack(mask)].

    Returns (f0, units, post_boundaries, solid_mask)."""
    d = st["d"]
    D = res[1] / 8
    domain_length_x = res[0] / D
    l_lu = res[0] / domain_length_x * 1
    units = Units(reynolds_number, mach_number, characteristic_length_lu=l_lu)
    # quirk restated: Obstacle.grid divides an int64 arange by a Python float, which torch
    # evaluates in its default dtype float32 (obstacle.py:101-105), so the solid mask is
    # decided in fp32 regardless of the context dtype.
    grid = np.meshgrid(*[np.arange(n).astype(np.float32) * np.float32(1.0) / np.float32(l_lu) for n in res],
                       indexing="ij")
    c = [np.float32(0.25) * grid[0].max()] + [np.float32(0.5) * gi.max() for gi in grid[1:]]
    solid = sum((gi - ci) ** 2 for gi, ci in zip(grid, c)) < 0.5 ** 2
    ex = np.zeros(d); ex[0] = 1.0
    # quirk restated: initial_pu builds u from torch.eye (float32) and converts it to lattice
    # units before the context dtype is applied (obstacle.py:94-99, _flow.py:131-134), so the
    # initial velocity is the fp32 rounding of Ma*cs.
    u0_lu = ((~solid)[None] * ex.reshape((d,) + (1,) * d)).astype(np.float32) / np.float32(1.0) * np.float32(units.u_lu)
    rho0 = units.pressure_pu_to_density_lu(np.zeros((1, *res)))
    f0 = equilibrium(st, rho0[0].astype(dtype), u0_lu.astype(dtype)).astype(dtype)
    inlet = np.abs(grid[0]) < 1e-6
    post = [equilibrium_pu(inlet, units.pressure_pu_to_density_lu(np.zeros((1,) * (d + 1))),
                           units.velocity_to_lu(ex).reshape((d,) + (1,) * d)),
            outlet_p([1] + [0] * (d - 1), 1.0),
            bounce_back(solid)]
    return f0, units, post, solid
