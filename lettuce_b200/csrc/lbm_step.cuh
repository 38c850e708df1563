// lbm_step.cuh -- fused stream+collide kernels.
//
// One launch = one lettuce time step (lettuce/_simulation.py:149-166, 241-305): every population is
// read once and written once.  PULL gathers f_q(x - e_q) (stream-before-collide), PUSH scatters the
// post-collision value to x + e_q (stream-after-collide); the four StreamingStrategy values are the four
// (PULL, PUSH) combinations.
//
//   step_kernel          bulk kernel over EVERY node (`node_update`), one node per thread or two neighbours per
//                        thread as a float2 on the packed fp32 pipe.  Run-time options: the in-kernel lock step of
//                        multi-GPU slabs (SlabSync) and the reporters' moment reductions (sum 0.5|u|^2, max |u|^2)
//                        of the state the step writes or reads.
//   general_nodes_kernel masked runs: sparse kernel over the precomputed list of general nodes (boundaries, frozen
//                        slots): `general_node` restates Appendix A.2 of SURVEY.md node by node.  The bulk kernel
//                        treats every node as plain fluid; this kernel, launched BEHIND it with programmatic
//                        dependent launch, gathers and collides next to it and overwrites the general nodes' slots
//                        once the bulk grid has completed (griddepcontrol.wait).  Every slot the bulk kernel writes
//                        wrongly (from or into a general node) has a general node as its rightful writer: nodes
//                        that stream into a frozen slot carry the general bit, frozen slots are rewritten by their
//                        owner.  (Measured against two variants with a label test in the bulk kernel -- label first,
//                        and label loaded together with the populations -- this one was fastest on every masked
//                        configuration: profiles/r2_masked_modes.md.)
//   link_gather_kernel / link_scatter_kernel
//                        link-wise bounce-back boundaries applied AFTER streaming (fullway / halfway / linearly
//                        interpolated, momentum-exchange force), sparse over the list of boundary links
//
// Consecutive steps are chained with programmatic dependent launch (lbm_step_n): every step kernel first waits
// for the grid in front of it to COMPLETE (griddepcontrol.wait, a no-op after an ordinary launch), then releases
// the kernel behind it (griddepcontrol.launch_dependents), so the next kernel's CTAs are resident and past their
// index arithmetic when the data dependency resolves; nothing reads populations before the previous step is done.
//
// Planes x = -1 and x = n0 resolve to the wrapped plane of the same buffer (single GPU) or to a
// peer-mapped plane of the neighbour rank's buffer (in_plane / out_plane).
#pragma once
#include "lbm_core.cuh"

namespace lbm {

constexpr uint8_t kLabelGeneral = 0x80;  // label bit 7: node needs the general path

template <class R>
struct OpDev {
    int kind, axis, side, _pad;  // axis is an INTERNAL axis (0..2)
    R a, b;                      // collision scalars / rho_outlet in a
    const R *rho;
    const R *u;
    int64_t rho_stride[3];       // internal axis order
    int64_t u_stride[4];         // component, internal axes 0..2
};

// In-kernel lock step of neighbouring slabs (multi-GPU, unmasked runs): boundary-plane CTAs wait for the
// neighbour's progress counter before they touch peer memory and publish this rank's counter when the
// last of them has finished -- no separate synchronisation kernel, interior CTAs never wait.
struct SlabSync {
    unsigned long long *sig_lo, *sig_hi;   // peer-mapped: neighbour's wait slot for this rank
    const unsigned long long *wait;        // local: [0] written by lo neighbour, [1] by hi neighbour
    unsigned long long *done;              // local: [0],[1] finished boundary CTAs per side, [2] finished CTAs of the
                                           // sparse kernel (zeroed per call)
    unsigned long long wait_value, signal_value;
    unsigned long long timeout_cycles;     // a spin on a peer counter gives up (trap) after this many clocks
    unsigned int ctas_per_side;
    int on;
    int publish;                           // 0: the sparse kernel behind this grid publishes the counters
};

// Per-population base pointers of the bulk kernel, built on the host (fill_params).  Entry [k][q] already
// contains q * stride and the x-shift of population q, so that the address of a load / store is
//     table[k][q] + (x * plane + in-plane index)
// with k = 0 for interior planes, 1 for the first plane (its lower neighbour plane is the wrapped or
// peer-mapped `lo` plane), 2 for the last plane (`hi`), 3 when the slab is one plane thick.
constexpr int kQMax = 27;
template <class R>
struct AddrTables {
    const R *ld[4][kQMax];
    R *st[4][kQMax];
};

template <class R>
struct StepParams {
    const R *in;
    R *out;
    // planes x = -1 and x = n0 (peer-mapped in multi-GPU runs, wrapped otherwise)
    const R *in_lo, *in_hi;
    R *out_lo, *out_hi;
    int64_t in_lo_qs, in_hi_qs, out_lo_qs, out_hi_qs;
    int n0, n1, n2;
    int64_t N;  // n0*n1*n2
    const uint8_t *labels, *labels_lo, *labels_hi;
    const uint32_t *frozen, *frozen_lo, *frozen_hi;
    const int32_t *general_nodes;  // flat indices of the nodes with label bit 7 set
    int n_general;
    int n_ops, collision_index;
    int nested_outlets;  // two or more active outlets: an outlet's neighbour may lie on another outlet's plane
    R ca, cb;  // scalars of the collision entry
    ForceArgs<R> force;  // LBM_OP_BGK_FORCED only
    SlabSync sync;
    // fused moment reductions (kReduceNone / kReduceOutput / kReduceInput): one (sum 0.5|u|^2, max |u|^2) pair per
    // CTA, sums in [0, reduce_slots), maxima in [reduce_slots, 2 reduce_slots); the sparse kernel's CTAs use the
    // slots from reduce_sparse_offset on
    double *energy_partials;
    int reduce_mode, reduce_slots, reduce_sparse_offset;
    // Sweep direction of the bulk kernel over the x-planes.  Consecutive steps alternate it (lbm_api.cu), so that the
    // planes the previous step wrote LAST are the ones this step reads FIRST: they are still in the 126 MB L2.
    // Worth up to L2 size / population buffer size of the read traffic (4096x1024 D2Q9 fp32: 151 MB per buffer).
    int reverse_sweep;
    AddrTables<R> tbl;
    OpDev<R> ops[LBM_MAX_OPS];
};

// pointer to population 0 of plane `x` (x in [-1, n0]) plus its q stride
template <class R, class P>
struct Plane {
    P *p;
    int64_t qs;
};

// (branch-free selects: x is block-uniform, so these stay on the uniform datapath)
template <class R>
LBM_D Plane<R, const R> in_plane(const StepParams<R> &p, int x) {
    const bool lo = x < 0, hi = x >= p.n0;
    const R *inner = p.in + (int64_t)(lo || hi ? 0 : x) * p.n1 * p.n2;
    return {lo ? p.in_lo : (hi ? p.in_hi : inner), lo ? p.in_lo_qs : (hi ? p.in_hi_qs : p.N)};
}
template <class R>
LBM_D Plane<R, R> out_plane(const StepParams<R> &p, int x) {
    const bool lo = x < 0, hi = x >= p.n0;
    R *inner = p.out + (int64_t)(lo || hi ? 0 : x) * p.n1 * p.n2;
    return {lo ? p.out_lo : (hi ? p.out_hi : inner), lo ? p.out_lo_qs : (hi ? p.out_hi_qs : p.N)};
}
template <class R>
LBM_D const uint32_t *frozen_plane(const StepParams<R> &p, int x) {
    if (x < 0) return p.frozen_lo;
    if (x >= p.n0) return p.frozen_hi;
    return p.frozen + (int64_t)x * p.n1 * p.n2;
}
template <class R>
LBM_D const uint8_t *label_plane(const StepParams<R> &p, int x) {
    if (x < 0) return p.labels_lo;
    if (x >= p.n0) return p.labels_hi;
    return p.labels + (int64_t)x * p.n1 * p.n2;
}

LBM_D int wrap(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

// the collision entry of the transformer list applied to the node (V = R) or the two nodes (V = float2) held in f
template <class S, class V, int COLL, class P>
LBM_D void collide_lanes(const P &p, V (&f)[S::Q]) {
    if constexpr (COLL == LBM_OP_BGK_FORCED) collide_bgk_forced<S, V>(f, p.ca, p.force);
    else Collide<S, V, COLL>::apply(f, p.ca, p.cb);
}
template <class S, class R, int COLL>
LBM_D void collide_node(const StepParams<R> &p, R (&f)[S::Q]) { collide_lanes<S, R, COLL>(p, f); }

// ---------------------------------------------------------------------------
// general path pieces
// ---------------------------------------------------------------------------

// populations of node (x,y,z) as the collide phase sees them: after the optional
// pre-streaming gather with the destination-side frozen-slot rule
// (lettuce/_simulation.py:245-256).  x may be -1 or n0 (halo plane) ONLY when
// PULL is false.
template <class S, class R, bool PULL>
__device__ void gather_node(const StepParams<R> &p, int x, int y, int z, R (&f)[S::Q]) {
    uint32_t fr = 0;
    if (PULL && p.frozen) fr = frozen_plane(p, x)[(int64_t)y * p.n2 + z];
    ForQ<S::Q>::run([&]<int q>() {
        int xs = x, ys = y, zs = z;
        if (PULL && !((fr >> q) & 1u)) {
            xs = x - S::e(q, 0);
            ys = wrap(y - S::e(q, 1), p.n1);
            zs = wrap(z - S::e(q, 2), p.n2);
        }
        const auto pl = in_plane(p, xs);
        f[q] = pl.p[q * pl.qs + (int64_t)ys * p.n2 + zs];
    });
}

template <class S, class R>
LBM_D void bounce_back(R (&f)[S::Q]) {
    // lettuce/ext/_boundary/bounce_back_boundary.py:17-18 : f <- f[opposite]
    ForQ<S::Q>::run([&]<int q>() {
        constexpr int o = S::opp(q);
        if constexpr (q < o) {
            const R t = f[q];
            f[q] = f[o];
            f[o] = t;
        }
    });
}

template <class S, class R>
LBM_D void equilibrium_boundary(const OpDev<R> &op, int x, int y, int z, R (&f)[S::Q]) {
    // lettuce/ext/_boundary/equilibrium_boundary_pu.py:79-84 (values pre-converted to lattice units)
    const R rho = op.rho[x * op.rho_stride[0] + y * op.rho_stride[1] + z * op.rho_stride[2]];
    R u[3] = {R(0), R(0), R(0)};
#pragma unroll
    for (int c = 0; c < S::D; ++c)
        u[S::axis_of(c)] = op.u[c * op.u_stride[0] + x * op.u_stride[1] + y * op.u_stride[2] + z * op.u_stride[3]];
    equilibrium_all<S, R>(rho, u, f);
}

// applies transformer entry `i` to node-local populations if the node carries label i.
// Outlet kinds are handled by the caller (they need a neighbour).
// (one out-of-line copy of the collision per translation unit for the nested pipelines of multi-outlet lattices)
template <class S, class R, int COLL>
__device__ __noinline__ void collide_node_out_of_line(const StepParams<R> &p, R (&f)[S::Q]) {
    collide_node<S, R, COLL>(p, f);
}

template <class S, class R, int COLL, bool OUT_OF_LINE = false>
LBM_D void apply_local_op(const StepParams<R> &p, int i, int label, int x, int y, int z, R (&f)[S::Q]) {
    if (label != i) return;
    const OpDev<R> &op = p.ops[i];
    if (i == p.collision_index) {
        if constexpr (OUT_OF_LINE) collide_node_out_of_line<S, R, COLL>(p, f);
        else collide_node<S, R, COLL>(p, f);
    } else if (op.kind == LBM_OP_BOUNCE_BACK) {
        bounce_back<S, R>(f);
    } else if (op.kind == LBM_OP_EQUILIBRIUM) {
        equilibrium_boundary<S, R>(op, x, y, z, f);
    }
}

LBM_D bool in_plane_of(int axis, int side, int x, int y, int z, int n0, int n1, int n2) {
    const int c = axis == 0 ? x : (axis == 1 ? y : z);
    const int n = axis == 0 ? n0 : (axis == 1 ? n1 : n2);
    if (side == 0) return false;  // plane owned by another slab
    return c == (side > 0 ? n - 1 : 0);
}

// Populations of node (x,y,z) as the reference's collide phase sees them when transformer entry `n_end` is about to
// run: gathered (pre-streaming), then entries 0 .. n_end-1 applied in order.  An outlet entry rewrites its plane from
// the velocity of the neighbour node one step inside the domain, taken from the populations after the entries BEFORE
// that outlet -- itself such a prefix, one level down.  An outlet's neighbour lies on the plane of an EARLIER outlet
// only where outlet planes of different axes meet (or on a three-plane axis with outlets at both ends), so the
// nesting never exceeds the number of outlets minus one (kMaxOutletDepth, checked on the host).  Lattices with at most
// one active outlet take the inline two-level form (NESTED = false, DEPTH = 1).
constexpr int kMaxOutletDepth = 2;

template <class R>
struct NodeState {
    R rho, u[3];
};
// (out of line: one copy per translation unit instead of one per nesting level and kernel -- inlined, the four levels
// multiplied the build time by seven; only nodes on outlet planes ever get here)
template <class S, class R, int COLL, bool PULL, int DEPTH>
__device__ __noinline__ NodeState<R> neighbour_prefix(const StepParams<R> &p, int n_end, int x, int y, int z);

template <class S, class R, int COLL, bool PULL, int DEPTH, bool NESTED>
__device__ void pipeline_prefix(const StepParams<R> &p, int n_end, int x, int y, int z, int label, R (&f)[S::Q]) {
    constexpr int Q = S::Q;
    gather_node<S, R, PULL>(p, x, y, z, f);

    for (int i = 0; i < n_end; ++i) {
        const OpDev<R> &op = p.ops[i];
        if (op.kind == LBM_OP_OUTLET_P || op.kind == LBM_OP_ANTI_BOUNCE_BACK) {
            if (!in_plane_of(op.axis, op.side, x, y, z, p.n0, p.n1, p.n2)) continue;
            if constexpr (DEPTH == 0) continue;      // (unreachable: the host picks NESTED whenever planes can meet
                                                     // and bounds the number of outlets)
            R rho_n = R(1), u_n[3] = {R(0), R(0), R(0)};
            if constexpr (DEPTH > 0) {
                // velocity of the neighbour as the reference sees it when boundary i runs
                const int xn = x - (op.axis == 0 ? op.side : 0);
                const int yn = y - (op.axis == 1 ? op.side : 0);
                const int zn = z - (op.axis == 2 ? op.side : 0);
                if constexpr (NESTED) {
                    const NodeState<R> nb = neighbour_prefix<S, R, COLL, PULL, DEPTH - 1>(p, i, xn, yn, zn);
                    rho_n = nb.rho;
                    u_n[0] = nb.u[0]; u_n[1] = nb.u[1]; u_n[2] = nb.u[2];
                } else {
                    R g[Q], j[3];
                    const int label_n = label_plane(p, xn)[(int64_t)yn * p.n2 + zn] & 0x7f;
                    pipeline_prefix<S, R, COLL, PULL, DEPTH - 1, false>(p, i, xn, yn, zn, label_n, g);
                    moments<S, R>(g, rho_n, j);
                    const R inv = R(1) / rho_n;
                    u_n[0] = j[0] * inv; u_n[1] = j[1] * inv; u_n[2] = j[2] * inv;
                }
            }
            if (op.kind == LBM_OP_OUTLET_P) {
                // equilibrium_outlet_p.py:63-73: whole plane <- feq(rho_outlet, u[neighbour]),
                // irrespective of the label
                equilibrium_all<S, R>(op.a, u_n, f);
            } else {
                // anti_bounce_back_outlet.py:71-91, in place on the whole plane
                R rho_h, j_h[3];
                moments<S, R>(f, rho_h, j_h);
                const R inv = R(1) / rho_h;
                R uw[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const R uh = j_h[a] * inv;
                    uw[a] = uh + R(0.5) * (uh - u_n[a]);
                }
                const R uw2 = uw[0] * uw[0] + uw[1] * uw[1] + uw[2] * uw[2];
                R fnew[Q];
                ForQ<Q>::run([&]<int q>() {
                    fnew[q] = f[q];
                });
                ForQ<Q>::run([&]<int q>() {
                    constexpr int o = S::opp(q);
                    // q leaves through the plane when e_q . direction == 1
                    const int en = (op.axis == 0 ? S::e(q, 0) : (op.axis == 1 ? S::e(q, 1) : S::e(q, 2))) * op.side;
                    if (en == 1) {
                        R eu = R(0);
#pragma unroll
                        for (int a = 0; a < 3; ++a) eu += R(S::e(q, a)) * uw[a];
                        fnew[o] = -f[q] + R(S::w(q)) * rho_h *
                                              (R(2) + eu * eu * R(1.0 / (kCs2 * kCs2)) - uw2 * R(1.0 / kCs2));
                    }
                });
                ForQ<Q>::run([&]<int q>() { f[q] = fnew[q]; });
            }
        } else {
            apply_local_op<S, R, COLL, NESTED>(p, i, label, x, y, z, f);
        }
    }
}

// (one instantiation per nesting level: the stack each needs is known at compile time)
template <class S, class R, int COLL, bool PULL, int DEPTH>
__device__ __noinline__ NodeState<R> neighbour_prefix(const StepParams<R> &p, int n_end, int x, int y, int z) {
    R g[S::Q];
    const int label = label_plane(p, x)[(int64_t)y * p.n2 + z] & 0x7f;
    pipeline_prefix<S, R, COLL, PULL, DEPTH, true>(p, n_end, x, y, z, label, g);
    NodeState<R> out;
    R j[3];
    moments<S, R>(g, out.rho, j);
    const R inv = R(1) / out.rho;
    out.u[0] = j[0] * inv; out.u[1] = j[1] * inv; out.u[2] = j[2] * inv;
    return out;
}

template <class S, class R, int COLL, bool PULL>
__device__ __noinline__ void nested_node_pipeline(const StepParams<R> &p, int x, int y, int z, int label,
                                                  R (&f)[S::Q]) {
    pipeline_prefix<S, R, COLL, PULL, kMaxOutletDepth + 1, true>(p, p.n_ops, x, y, z, label, f);
}

// populations of node (x,y,z) after the collide phase (all transformer entries), before any post-streaming.
// NESTED = false (at most one active outlet: no outlet's neighbour lies on an outlet plane): node and neighbour
// inline, as fast as it gets; NESTED = true: every level out of line.
template <class S, class R, int COLL, bool PULL, bool NESTED>
__device__ void node_pipeline(const StepParams<R> &p, int x, int y, int z, int label, R (&f)[S::Q]) {
    if constexpr (NESTED) nested_node_pipeline<S, R, COLL, PULL>(p, x, y, z, label, f);
    else pipeline_prefix<S, R, COLL, PULL, 1, false>(p, p.n_ops, x, y, z, label, f);
}

// scatter of a general node's populations with the destination-side frozen-slot rule (_simulation.py:252-255):
// slot (q, dst) takes the streamed value unless it is frozen, in which case the node's own value stays.
template <class S, class R, bool PUSH>
__device__ void scatter_general(const StepParams<R> &p, int x, int y, int z, const R (&f)[S::Q]) {
    constexpr int Q = S::Q;
    const int64_t row = (int64_t)y * p.n2 + z;
    if (!PUSH) {
        const auto pl = out_plane(p, x);
        ForQ<Q>::run([&]<int q>() { pl.p[q * pl.qs + row] = f[q]; });
        return;
    }
    const uint32_t own = p.frozen ? frozen_plane(p, x)[row] : 0u;
    ForQ<Q>::run([&]<int q>() {
        if constexpr (q == 0) {
            const auto pl = out_plane(p, x);
            pl.p[row] = f[0];
        } else {
            if ((own >> q) & 1u) {
                const auto pl = out_plane(p, x);
                pl.p[q * pl.qs + row] = f[q];
            }
            const int xd = x + S::e(q, 0);
            const int yd = wrap(y + S::e(q, 1), p.n1);
            const int zd = wrap(z + S::e(q, 2), p.n2);
            const int64_t drow = (int64_t)yd * p.n2 + zd;
            uint32_t dfr = 0;
            if (p.frozen) {
                const uint32_t *fp = frozen_plane(p, xd);
                dfr = fp ? fp[drow] : 0u;
            }
            if (!((dfr >> q) & 1u)) {
                const auto pl = out_plane(p, xd);
                pl.p[q * pl.qs + drow] = f[q];
            }
        }
    });
}

// ---------------------------------------------------------------------------
// bulk kernel: LANES nodes per thread (1, or 2 neighbours along the contiguous axis as one float2 on the packed
// fp32 pipe), threadIdx.x along the contiguous axis.
// ---------------------------------------------------------------------------

// moment reductions fused into the step (StepParams::reduce_mode): of the state the step WRITES (steps that do not
// push: the node's output is still in registers, and collisions conserve rho and j) or of the state it READS (steps
// that only push: the input node is the previous step's output node)
enum : int { kReduceNone = 0, kReduceOutput = 1, kReduceInput = 2 };

template <class R, int LANES>
struct LaneVec {
    using type = R;
};
template <>
struct LaneVec<float, 2> {
    using type = float2;
};

template <class S, class R, int COLL, int LANES>
constexpr int bulk_threads() {
    return LANES == 2 ? 128 : 256;
}
template <class S, class R, int COLL, int LANES>
constexpr int min_blocks_per_sm() {
    // fp32 KBC on D3Q27, one node per thread: cap at 85 registers (3 CTAs of 256 threads); the push variant
    // otherwise hoists 27 store addresses into 108 registers.  Two nodes per thread: 128 registers, 4 CTAs of 128.
    if (sizeof(R) == 4 && COLL == LBM_OP_KBC && S::Q == 27) return LANES == 2 ? 4 : 3;
    return 0;  // no constraint
}

// 0.5 |u|^2 of the node(s) in f, per lane, accumulated into (sum, max of |u|^2)
template <class S, class V>
LBM_D void accumulate_kinetic(const V (&f)[S::Q], const bool (&take)[2], double &sum, double &mx) {
    V rho, u[3];
    density_velocity<S, V>(f, rho, u);
    const V uu = vfma(u[2], u[2], vfma(u[1], u[1], vmul(u[0], u[0])));
#pragma unroll
    for (int l = 0; l < VecTraits<V>::lanes; ++l) {
        if (take[l]) {
            const double v = (double)vlane(uu, l);
            sum += 0.5 * v;
            mx = fmax(mx, v);
        }
    }
}

// LANES nodes (x, y, z .. z+LANES-1): gather, collide, scatter.  An address is a block-uniform table entry
// (AddrTables, read from the constant bank) plus one 32-bit node index per thread, chosen from nine precomputed
// (row, column) combinations: one IMAD.WIDE per access.  With two lanes, populations that do not move along the
// contiguous axis are loaded / stored as one aligned 64-bit access; the others as two 32-bit ones.
//
// Masked runs update EVERY node here and general_nodes_kernel overwrites its nodes afterwards.  `single_writer`
// (block-uniform; the cut planes of a multi-GPU slab) makes the general nodes skip their stores instead: a wrong store
// into the NEIGHBOUR's memory could land after the neighbour's sparse kernel has written the right value there
// (nothing orders the two ranks' kernels within a step), so across a cut every slot keeps exactly one writer.
template <class S, class R, int COLL, bool PULL, bool PUSH, int LANES, bool EXTRAS>   // EXTRAS: labels / reductions
LBM_D void node_update(const StepParams<R> &p, int x, int y, int z, bool single_writer, double &e_sum, double &e_max) {
    constexpr int Q = S::Q;
    using V = typename LaneVec<R, LANES>::type;
    // neighbour rows / columns with periodic wrap (torch.roll, _simulation.py:241-243)
    const int ym = (y == 0 ? p.n1 : y) - 1, yp = (y + 1 == p.n1) ? 0 : y + 1;
    const int zm = (z == 0 ? p.n2 : z) - 1, zp = (z + LANES == p.n2) ? 0 : z + LANES;   // left of lane 0, right of the last
    const int xoff = x * (p.n1 * p.n2);
    const int rowm = xoff + ym * p.n2, row0 = xoff + y * p.n2, rowp = xoff + yp * p.n2;
    const int k = (x == 0 ? 1 : 0) | (x == p.n0 - 1 ? 2 : 0);

    // the labels are needed to keep the general nodes out of a fused reduction (the sparse kernel contributes
    // those) and for single_writer
    bool mine[2] = {true, true};
    if constexpr (EXTRAS) {
        if ((p.reduce_mode != kReduceNone || single_writer) && p.labels != nullptr) {
#pragma unroll
            for (int l = 0; l < LANES; ++l) mine[l] = !(__ldg(p.labels + (row0 + z + l)) & kLabelGeneral);
        }
    }

    V f[Q];
    ForQ<Q>::run([&]<int q>() {
        constexpr int e1 = S::e(q, 1), e2 = S::e(q, 2);
        const int rs = (!PULL || e1 == 0) ? row0 : (e1 == 1 ? rowm : rowp);
        const R *src = p.tbl.ld[k][q] + rs;
        if constexpr (LANES == 1) {
            const int zs = (!PULL || e2 == 0) ? z : (e2 == 1 ? zm : zp);
            f[q] = __ldg(src + zs);
        } else {
            if constexpr (!PULL || e2 == 0) f[q] = __ldg(reinterpret_cast<const float2 *>(src + z));
            else if constexpr (e2 == 1) f[q] = make_float2(__ldg(src + zm), __ldg(src + z));          // from z-1, z
            else f[q] = make_float2(__ldg(src + (z + 1)), __ldg(src + zp));                          // from z+1, z+2
        }
    });
    if constexpr (EXTRAS) {
        if (p.reduce_mode == kReduceInput) accumulate_kinetic<S, V>(f, mine, e_sum, e_max);
    }

    collide_lanes<S, V, COLL>(p, f);

    if (EXTRAS && single_writer && !(mine[0] && mine[LANES - 1])) {
        // a general node on a cut plane: only the other lane's node (if it is plain fluid) stores, lane by lane
#pragma unroll
        for (int l = 0; l < LANES; ++l) {
            if (!mine[l]) continue;
            const int zl = z + l;
            const int zlp = zl + 1 == p.n2 ? 0 : zl + 1, zlm = (zl == 0 ? p.n2 : zl) - 1;
            ForQ<Q>::run([&]<int q>() {
                constexpr int e1 = S::e(q, 1), e2 = S::e(q, 2);
                const int rd = (!PUSH || e1 == 0) ? row0 : (e1 == 1 ? rowp : rowm);
                const int zd = (!PUSH || e2 == 0) ? zl : (e2 == 1 ? zlp : zlm);
                p.tbl.st[k][q][rd + zd] = vlane(f[q], l);
            });
        }
    } else {
        ForQ<Q>::run([&]<int q>() {
            constexpr int e1 = S::e(q, 1), e2 = S::e(q, 2);
            const int rd = (!PUSH || e1 == 0) ? row0 : (e1 == 1 ? rowp : rowm);
            R *dst = p.tbl.st[k][q] + rd;
            if constexpr (LANES == 1) {
                const int zd = (!PUSH || e2 == 0) ? z : (e2 == 1 ? zp : zm);
                dst[zd] = f[q];
            } else {
                if constexpr (!PUSH || e2 == 0) {
                    *reinterpret_cast<float2 *>(dst + z) = f[q];
                } else if constexpr (e2 == 1) {        // to z+1, z+2
                    dst[z + 1] = f[q].x;
                    dst[zp] = f[q].y;
                } else {                               // to z-1, z
                    dst[zm] = f[q].x;
                    dst[z] = f[q].y;
                }
            }
        });
    }
    if constexpr (EXTRAS) {
        if (p.reduce_mode == kReduceOutput) accumulate_kinetic<S, V>(f, mine, e_sum, e_max);
    }
}

// Slab lock step (multi-GPU): W boundary planes per side take part (2 when the step both pulls and pushes, because
// plane 1 then reads slots the neighbour pushed into plane 0 and pushes into plane 0 itself); their CTAs are
// scheduled FIRST so that the progress counters go out early and the neighbour's next step never stalls.
template <bool PULL, bool PUSH>
LBM_D int sync_plane(int zb, int n0, bool reverse) {
    constexpr int W = (PULL && PUSH) ? 2 : 1;
    if (zb < W) return zb;                                  // host guarantees n0 >= 2 W
    if (zb < 2 * W) return n0 - 2 * W + zb;
    return reverse ? n0 - 1 - (zb - W) : zb - W;            // interior planes W .. n0-W-1, in either direction
}

LBM_D void spin_until(const volatile unsigned long long *w, unsigned long long value, unsigned long long limit) {
    const long long t0 = clock64();
    while (*w < value) {
        // a neighbour that never arrives must not hang the GPU for ever: after `limit` cycles the kernel gives up
        // with a trap (sticky error on this rank, reported by the next CUDA call); LBM_B200_PEER_TIMEOUT_S
        if ((unsigned long long)(clock64() - t0) > limit) __trap();
        __nanosleep(100);
    }
}

// one partial (sum, max) per CTA: warp shuffles, one shared-memory exchange; deterministic (no atomics)
LBM_D void store_cta_partials(double e_sum, double e_max, double *partials, int slot, int n_slots) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        e_sum += __shfl_down_sync(0xffffffffu, e_sum, o);
        e_max = fmax(e_max, __shfl_down_sync(0xffffffffu, e_max, o));
    }
    __shared__ double warp_part[2][8];
    const int t = threadIdx.y * blockDim.x + threadIdx.x, nwarps = (blockDim.x * blockDim.y + 31) >> 5;
    if ((t & 31) == 0) {
        warp_part[0][t >> 5] = e_sum;
        warp_part[1][t >> 5] = e_max;
    }
    __syncthreads();
    if (t == 0) {
        double s = 0.0, m = 0.0;
        for (int w = 0; w < nwarps; ++w) {
            s += warp_part[0][w];
            m = fmax(m, warp_part[1][w]);
        }
        partials[slot] = s;
        partials[n_slots + slot] = m;
    }
}

// Three instantiations, so that the run-time options cost the plain kernel neither instructions nor registers:
//   kStepPlain   nothing but gather, collide, scatter
//   kStepSync    + the slab lock step (multi-GPU, unmasked lattices, no reductions): what a slab step normally is
//   kStepFull    + fused reductions and, on masked slabs, single-writer cut planes
enum : int { kStepPlain = 0, kStepSync = 1, kStepFull = 2 };

template <class S, class R, int COLL, bool PULL, bool PUSH, int LANES, int MODE>
__global__ void __launch_bounds__((bulk_threads<S, R, COLL, LANES>()), (min_blocks_per_sm<S, R, COLL, LANES>()))
    step_kernel(const __grid_constant__ StepParams<R> p) {
    // (no-op unless THIS grid was launched programmatically behind the previous step: then the grid in front has to
    // be complete, and its writes visible, before anything below reads populations)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // The kernel behind this one on the stream may have been launched with programmatic stream serialization
    // (masked runs: general_nodes_kernel; chained steps of lbm_step_n: the next step): it may become resident once
    // every CTA of this grid has got here, i.e. once the previous step is known to be complete.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const bool reverse = p.reverse_sweep != 0;
    const int z = (blockIdx.x * blockDim.x + threadIdx.x) * LANES;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    double e_sum = 0.0, e_max = 0.0;
    if constexpr (MODE == kStepPlain) {
        const int x = reverse ? p.n0 - 1 - (int)blockIdx.z : (int)blockIdx.z;
        if (z < p.n2 && y < p.n1) node_update<S, R, COLL, PULL, PUSH, LANES, false>(p, x, y, z, false, e_sum, e_max);
    } else {
        const bool sync = p.sync.on != 0;
        const int x = sync ? sync_plane<PULL, PUSH>(blockIdx.z, p.n0, reverse)
                           : (reverse ? p.n0 - 1 - (int)blockIdx.z : (int)blockIdx.z);
        constexpr int W = (PULL && PUSH) ? 2 : 1;
        const bool lo = sync && x < W, hi = sync && x >= p.n0 - W;
        const bool leader = threadIdx.x == 0 && threadIdx.y == 0;
        if (lo || hi) {
            // boundary-plane CTAs wait for the neighbour's progress counter before they touch peer memory
            if (leader) {
                spin_until(p.sync.wait + (lo ? 0 : 1), p.sync.wait_value, p.sync.timeout_cycles);
                __threadfence_system();
            }
            __syncthreads();
        }
        if (z < p.n2 && y < p.n1) {
            if constexpr (MODE == kStepFull)
                node_update<S, R, COLL, PULL, PUSH, LANES, true>(p, x, y, z, (lo || hi) && p.labels != nullptr, e_sum,
                                                                 e_max);
            else
                node_update<S, R, COLL, PULL, PUSH, LANES, false>(p, x, y, z, false, e_sum, e_max);
        }
        if (lo || hi) {
            // ... and the last of them to finish publishes this rank's counter (unless the sparse kernel of a masked
            // step still has to touch the boundary planes: then IT publishes, general_nodes_kernel)
            __threadfence_system();   // this thread's loads from / stores to the neighbour are performed
            __syncthreads();
            if (leader && p.sync.publish) {
                const unsigned long long old = atomicAdd(p.sync.done + (lo ? 0 : 1), 1ULL);
                if ((old + 1) % p.sync.ctas_per_side == 0) {
                    __threadfence_system();
                    *(volatile unsigned long long *)(lo ? p.sync.sig_lo : p.sync.sig_hi) = p.sync.signal_value;
                }
            }
        }
        if constexpr (MODE == kStepFull) {
            if (p.reduce_mode != kReduceNone)
                store_cta_partials(e_sum, e_max, p.energy_partials,
                                   (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x, p.reduce_slots);
        }
    }
}

// ---------------------------------------------------------------------------
// sparse kernel: one thread per node of the precomputed list of general nodes
// (boundaries, frozen slots), launched behind step_kernel on the same stream.
// ---------------------------------------------------------------------------
template <class S, class R, int COLL, bool PULL, bool PUSH, bool NESTED>
__global__ void __launch_bounds__(128) general_nodes_kernel(const __grid_constant__ StepParams<R> p) {
    // Launched programmatically behind the bulk kernel, which released this grid only after the previous step
    // had completed: the input populations are final.  The next step's bulk kernel may become resident right away
    // (it waits for THIS grid's completion before it reads anything).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr int Q = S::Q;
    if (p.sync.on) {
        // slabs: the gather below may read the neighbours' planes before the bulk kernel's boundary CTAs have seen
        // the neighbours' progress counters
        if (threadIdx.x == 0) {
            spin_until(p.sync.wait + 0, p.sync.wait_value, p.sync.timeout_cycles);
            spin_until(p.sync.wait + 1, p.sync.wait_value, p.sync.timeout_cycles);
            __threadfence_system();
        }
        __syncthreads();
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < p.n_general;
    int x = 0, y = 0, z = 0;
    R f[Q];
    double e_sum = 0.0, e_max = 0.0;
    const bool take[2] = {true, false};
    if (active) {
        const int n = p.general_nodes[i];
        z = n % p.n2;
        y = (n / p.n2) % p.n1;
        x = n / (p.n1 * p.n2);
        if (p.reduce_mode == kReduceInput) {        // the node as the previous step left it (steps that only push)
            gather_node<S, R, false>(p, x, y, z, f);
            accumulate_kinetic<S, R>(f, take, e_sum, e_max);
        }
        node_pipeline<S, R, COLL, PULL, NESTED>(p, x, y, z, p.labels[n] & 0x7f, f);
    }
    // the bulk kernel updated every node as if it were plain fluid: wait for it to finish before overwriting the
    // slots that belong to the general nodes (every thread passes the wait, so that completion of this grid
    // implies completion of the bulk grid)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (active) {
        scatter_general<S, R, PUSH>(p, x, y, z, f);
        if (p.reduce_mode == kReduceOutput) accumulate_kinetic<S, R>(f, take, e_sum, e_max);
    }
    if (p.sync.on) {
        // the last CTA to finish publishes this rank's progress to both neighbours (the bulk kernel did not)
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long old = atomicAdd(p.sync.done + 2, 1ULL);
            if ((old + 1) % gridDim.x == 0) {
                __threadfence_system();
                *(volatile unsigned long long *)p.sync.sig_lo = p.sync.signal_value;
                *(volatile unsigned long long *)p.sync.sig_hi = p.sync.signal_value;
            }
        }
    }
    if (p.reduce_mode != kReduceNone)
        store_cta_partials(e_sum, e_max, p.energy_partials, p.reduce_sparse_offset + blockIdx.x, p.reduce_slots);
}

// ---------------------------------------------------------------------------
// Link-wise bounce-back boundaries applied after streaming: the "efficient bounce-back" boundaries of the
// reference's example project examples/advanced_projects/efficient_bounce_back_obstacle (ebb/ below).  A link is
// (node, q): q points from a fluid node into a solid node.  FULLWAY links live on the solid node and reverse the
// population that streamed in (ebb/boundary/fullway_bounce_back_boundary.py:132-155); HALFWAY and INTERPOLATED
// links live on the fluid node and need the populations between collision and streaming, fc
// (halfway_bounce_back_boundary.py:167-182, linear_interpolated_bounce_back_boundary.py:57-100).  The step kernels
// do not keep fc (it streams away, or is dropped at the frozen solid node), so the gather kernel re-evaluates the
// node's collide phase from the step's INPUT buffer, which the two-buffer scheme leaves intact.  Two kernels because
// the reference evaluates every right-hand side before it assigns (a node with links q and opposite(q) swaps).
// ---------------------------------------------------------------------------
template <class R>
struct LinkArgs {
    int kind;                 // lbm_link_kind
    int n;                    // number of links
    const int32_t *node;      // flat node index
    const uint8_t *q;         // population index pointing into the solid
    const R *d;               // INTERPOLATED: wall distance in link lengths, (0, 1]
    R *bounced;               // scratch [n]: value of slot (opposite(q), node) after the boundary
    double *partials;         // [3][gridDim.x] momentum-exchange partial sums (internal axis order) or nullptr
};

constexpr int kLinkThreads = 128;

template <class S, class R, int COLL, bool NESTED>
__global__ void __launch_bounds__(kLinkThreads) link_gather_kernel(const __grid_constant__ StepParams<R> p,
                                                                     const LinkArgs<R> a) {
    constexpr int Q = S::Q;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double force[3] = {0.0, 0.0, 0.0};
    if (i < a.n) {
        const int n = a.node[i];
        const int q = a.q[i];
        const R streamed = p.out[(int64_t)q * p.N + n];            // slot (q, node) after streaming
        R val, mom;                                                 // new slot value; coefficient of e_q in the force
        if (a.kind == LBM_LINK_FULLWAY) {
            val = streamed;
            mom = R(2) * streamed;                                  // fullway_bounce_back_boundary.py:157-170
        } else {
            const int z = n % p.n2;
            const int y = (n / p.n2) % p.n1;
            const int x = n / (p.n1 * p.n2);
            R f[Q];
            const int label = p.labels ? (p.labels[n] & 0x7f) : p.collision_index;
            node_pipeline<S, R, COLL, false, NESTED>(p, x, y, z, label, f);
            R fcq = R(0), fco = R(0);                               // fc[q], fc[opposite(q)] of this node
            ForQ<Q>::run([&]<int k>() {
                if (k == q) {
                    fcq = f[k];
                    fco = f[S::opp(k)];
                }
            });
            if (a.kind == LBM_LINK_HALFWAY) {
                val = fcq;
                mom = R(2) * fcq;                                   // halfway_bounce_back_boundary.py:206-215
            } else {
                const R d = a.d[i];
                if (d <= R(0.5)) val = R(2) * d * fcq + (R(1) - R(2) * d) * streamed;
                else val = (R(1) / (R(2) * d)) * fcq + (R(1) - R(1) / (R(2) * d)) * fco;
                mom = fcq + val;                                    // linear_interpolated_...py:116-143
            }
        }
        a.bounced[i] = val;
        ForQ<Q>::run([&]<int k>() {
            if (k == q) {
                force[0] = (double)(R(S::e(k, 0)) * mom);
                force[1] = (double)(R(S::e(k, 1)) * mom);
                force[2] = (double)(R(S::e(k, 2)) * mom);
            }
        });
    }
    if (a.partials == nullptr) return;
    __shared__ double warp_sum[3][kLinkThreads / 32];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double v = force[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) warp_sum[c][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int w = 0; w < kLinkThreads / 32; ++w) s += warp_sum[threadIdx.x][w];
        a.partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
}

template <class S, class R>
__global__ void __launch_bounds__(kLinkThreads) link_scatter_kernel(R *out, int64_t N, const LinkArgs<R> a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int q = a.q[i];
    int o = 0;
    ForQ<S::Q>::run([&]<int k>() {
        if (k == q) o = S::opp(k);
    });
    out[(int64_t)o * N + a.node[i]] = a.bounced[i];
}

}  // namespace lbm
