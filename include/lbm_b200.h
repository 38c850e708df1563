/* lbm_b200.h -- C ABI of the B200-native lattice-Boltzmann stream+collide engine.
 *
 * This is the drop-in boundary for lettuce's hot path.  It replaces the run-time
 * generated torch extension of the reference:
 *
 *   reference interface                                           replaced by
 *   -----------------------------------------------------------   ------------------------
 *   cuda_native.lettuce(f, [ncm,] tau_inv..., [nsm,] f_next)      lbm_step()
 *     (lettuce/cuda_native/_template.py:42-53,72-81; generated
 *      kernel body lettuce/cuda_native/_default_code_gen.py:285-316)
 *   invoke(simulation) python shim (_template.py:35-39)           lettuce_b200.native.invoke()
 *   Simulation.__init__ mask build (_simulation.py:100-146)       lbm_pack_masks()
 *   Flow.rho / Flow.j / Flow.u (lettuce/_flow.py:157-193)         lbm_moments()
 *   reporter reductions (ext/_reporter/observable_reporter.py     lbm_reduce(), lbm_step_moments()
 *     :27-68,140-158; util/utility.py:37-99 order=6)
 *   Simulation.__call__ with host-resident populations            lbm_run_host()
 *     (_simulation.py:311-323)
 *
 * Conventions: plain pointers and sizes, no torch types.  All pointers named
 * `d_*` or documented as "device" are CUDA device pointers of the current
 * device; `stream` is a cudaStream_t passed as void* (NULL = legacy default
 * stream).  Functions never synchronise the device unless stated, never throw,
 * and return LBM_OK (0) or a negative lbm_status.  Buffers are borrowed for
 * the duration of the call only (ownership stays with the caller, as in the
 * reference where Python owns every tensor, _default_code_gen.py:117-145).
 *
 * Memory layout: populations are `real f[q][nx][ny][nz]`, q slowest, z fastest,
 * exactly lettuce's `flow.f` (lettuce/_flow.py:92); 2-D lattices use nz = 1 and
 * `f[q][nx][ny]`.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_ABI_VERSION 5
#define LBM_MAX_OPS 8 /* transformer list length: pre_boundaries + collision + post_boundaries */

typedef enum lbm_status {
    LBM_OK = 0,
    LBM_ERR_BAD_ARGUMENT = -1, /* NULL pointer, non-positive extent, bad enum value */
    LBM_ERR_UNSUPPORTED = -2,  /* valid request the engine has no kernel for (e.g. KBC on D3Q19) */
    LBM_ERR_CUDA = -3,         /* a CUDA runtime call failed; see lbm_last_cuda_error() */
    LBM_ERR_ALIASING = -4,     /* f_in and f_out overlap */
    LBM_ERR_TOO_LARGE = -5     /* nx*ny*nz does not fit 31 bits */
} lbm_status;

/* lettuce/ext/_stencil/{d2q9,d3q19,d3q27}.py; velocity order, weights and opposite
 * tables are lettuce's. */
typedef enum lbm_stencil { LBM_D2Q9 = 0, LBM_D3Q19 = 1, LBM_D3Q27 = 2 } lbm_stencil;

typedef enum lbm_dtype { LBM_F32 = 0, LBM_F64 = 1 } lbm_dtype;

/* Same bit values as lettuce's StreamingStrategy (cuda_native/_default_code_gen.py:13-25):
 * bit 1 = stream before the collide phase, bit 0 = stream after it. */
typedef enum lbm_streaming {
    LBM_NO_STREAMING = 0,
    LBM_POST_STREAMING = 1,
    LBM_PRE_STREAMING = 2,
    LBM_DOUBLE_STREAMING = 3
} lbm_streaming;

typedef enum lbm_op_kind {
    LBM_OP_NO_COLLISION = 0, /* ext/_collision/no_collision.py:9-11 */
    LBM_OP_BGK = 1,          /* ext/_collision/bgk_collision.py:17-22 (force = None); p0 = tau */
    LBM_OP_TRT = 2,          /* ext/_collision/trt_collision.py:16-27; p0 = tau_plus, p1 = tau_minus */
    LBM_OP_KBC = 3,          /* ext/_collision/kbc_collision.py:96-160; p0 = tau (units.relaxation_parameter_lu) */
    LBM_OP_REGULARIZED = 4,  /* ext/_collision/regularized_collision.py:17-43; p0 = tau (units.relaxation_parameter_lu) */
    LBM_OP_SMAGORINSKY = 5,  /* ext/_collision/smagorinsky_collision.py:22-40 (force = None); p0 = tau, p1 = constant */
    LBM_OP_BGK_FORCED = 6,   /* BGK with a body force (bgk_collision.py:17-22 with ext/_force/guo.py or shan_chen.py):
                                p0 = tau, force[] = acceleration in lattice units, ueq_scale = 0.5 (Guo) or tau
                                (ShanChen), source_scale = 1 - 1/(2 tau_force) (Guo) or 0 (ShanChen) */
    LBM_OP_BOUNCE_BACK = 16, /* ext/_boundary/bounce_back_boundary.py:10-32 */
    LBM_OP_EQUILIBRIUM = 17, /* ext/_boundary/equilibrium_boundary_pu.py:79-84, values already in lattice units */
    LBM_OP_OUTLET_P = 18,    /* ext/_boundary/equilibrium_outlet_p.py:63-73; p0 = rho_outlet */
    LBM_OP_ANTI_BOUNCE_BACK = 19, /* ext/_boundary/anti_bounce_back_outlet.py:71-91 */
    LBM_OP_IDENTITY = 20     /* boundary entry that leaves its nodes untouched; used for stream-only passes (the
                                frozen-slot rule still applies), see Engine.step in lettuce_b200/native.py */
} lbm_op_kind;

/* One entry of the transformer list (lettuce/_simulation.py:70).  Entry i acts on
 * the nodes whose label equals i; outlet kinds additionally act on their whole
 * boundary plane irrespective of the label. */
typedef struct lbm_op {
    int32_t kind;  /* lbm_op_kind */
    int32_t axis;  /* outlets: 0 = x, 1 = y, 2 = z */
    int32_t side;  /* outlets: +1 = plane n-1 (neighbour n-2), -1 = plane 0 (neighbour 1) */
    int32_t _pad;
    double p0, p1; /* see lbm_op_kind */
    /* LBM_OP_EQUILIBRIUM: density and velocity in lattice units, dtype of the lattice.
     * Element strides per spatial axis x,y,z (0 broadcasts that axis); u has an
     * additional component stride.  This mirrors checked_tensor's broadcast rules
     * (equilibrium_boundary_pu.py:23-69). */
    const void *rho;
    const void *u;
    int64_t rho_stride[3];
    int64_t u_stride[4]; /* component, x, y, z */
    /* LBM_OP_BGK_FORCED: u_eq = u + ueq_scale * force / rho enters the equilibrium, and the source term
     * source_scale * w_q [ (e_q - u_eq)/cs^2 + (e_q.u_eq) e_q/cs^4 ] . force is added after relaxation */
    double force[3];
    double ueq_scale, source_scale;
} lbm_op;

/* Geometry of the slab of lattice this call works on. */
typedef struct lbm_lattice {
    int32_t stencil; /* lbm_stencil */
    int32_t dtype;   /* lbm_dtype */
    int32_t nx, ny, nz; /* nz = 1 for D2Q9 */
    int32_t _pad;
} lbm_lattice;

/* Planes adjacent to the slab in x.  Single GPU / periodic: leave all NULL and the
 * engine wraps around inside the buffer.  Multi-GPU x-slabs: `lo` is the plane at
 * local x = -1 (last plane of the left neighbour), `hi` the plane at x = nx (first
 * plane of the right neighbour); they may be peer-mapped device pointers, the
 * kernel then loads / stores them directly over NVLink.  Each pointer addresses
 * population 0 of that plane; population q is `q_stride` elements further. */
typedef struct lbm_halo {
    const void *in_lo;
    const void *in_hi;
    void *out_lo;
    void *out_hi;
    int64_t in_lo_qstride, in_hi_qstride, out_lo_qstride, out_hi_qstride;
    /* labels / frozen-slot words of the neighbour planes (masked runs only) */
    const uint8_t *label_lo;
    const uint8_t *label_hi;
    const uint32_t *frozen_lo;
    const uint32_t *frozen_hi;
} lbm_halo;

typedef struct lbm_step_desc {
    lbm_lattice lat;
    int32_t streaming; /* lbm_streaming */
    int32_t n_ops;     /* 1 <= n_ops <= LBM_MAX_OPS */
    int32_t collision_index; /* index of the collision entry in ops (= number of pre-boundaries) */
    int32_t variant;   /* bulk kernel: 0 = library default; 1 or 2 = LDG/STG kernel with that many nodes per thread (two
                          neighbouring nodes as one float2 on the packed fp32 pipe; used where that kernel exists:
                          fp32, even contiguous extent, PRE / POST streaming); 3 = TMA-staged kernel (csrc/lbm_tma.cuh:
                          fp32, every streaming mode but DOUBLE, contiguous extent a multiple of 64;
                          LBM_ERR_UNSUPPORTED otherwise).  All
                          give bit-identical results (csrc/lbm_vec.cuh). */
    lbm_op ops[LBM_MAX_OPS];
    /* Masked runs (any boundary present): per-node label byte produced by
     * lbm_pack_masks() and one frozen-slot word per node (bit q set = slot (q,node) is
     * not overwritten by streaming).  Both NULL for unmasked runs. */
    const uint8_t *labels;
    const uint32_t *frozen;
    /* flat node indices (x*ny*nz + y*nz + z) of all nodes whose label has bit 7 set, from
     * lbm_list_general_nodes(); they are stepped by a separate sparse kernel. */
    const int32_t *general_nodes;
    int64_t n_general;
    lbm_halo halo;
} lbm_step_desc;

/* One time step: f_out = Step(f_in) for desc->streaming, reading f_in once and writing
 * f_out once.  f_in and f_out must not overlap.  Launches on `stream`, does not sync. */
int lbm_step(const lbm_step_desc *desc, const void *d_f_in, void *d_f_out, void *stream);

/* One time step with the reporters' moment reductions fused into the step kernels (a reporter with interval 1
 * then costs no second pass over the populations): d_result[0] = sum over nodes of 0.5|u|^2 and d_result[1] = max
 * over nodes of |u|^2, both in lattice units -- IncompressibleKineticEnergy and MaximumVelocity
 * (ext/_reporter/observable_reporter.py:27-42 with lettuce/_flow.py:200-204).  Which state they describe depends
 * on the streaming strategy (lbm_step_moments_state):
 *   LBM_MOMENTS_OF_OUTPUT  NO_STREAMING, PRE_STREAMING: the state this step WRITES (the node's output is still in
 *                          registers; collisions conserve rho and j)
 *   LBM_MOMENTS_OF_INPUT   POST_STREAMING: the state this step READS, i.e. what the previous step left (a pushing
 *                          step assembles its output node from several source nodes, but its input node is whole):
 *                          the reduction for the report after step k rides on step k+1
 *   LBM_MOMENTS_UNAVAILABLE DOUBLE_STREAMING (use lbm_step + lbm_reduce)
 * Boundaries are allowed (the sparse general-nodes kernel contributes its nodes).  `d_scratch` needs
 * lbm_step_moments_scratch_bytes(desc) bytes (0 = not available for desc).  Deterministic (no atomics). */
typedef enum lbm_moments_state {
    LBM_MOMENTS_UNAVAILABLE = 0, LBM_MOMENTS_OF_OUTPUT = 1, LBM_MOMENTS_OF_INPUT = 2
} lbm_moments_state;
int lbm_step_moments_state(const lbm_step_desc *desc);
size_t lbm_step_moments_scratch_bytes(const lbm_step_desc *desc);
int lbm_step_moments(const lbm_step_desc *desc, const void *d_f_in, void *d_f_out, void *d_scratch,
                     size_t scratch_bytes, double *d_result, void *stream);
/* n consecutive steps (populations alternate between d_f_a and d_f_b like lbm_step_n), every one with the fused
 * reductions: d_results[2 k], d_results[2 k + 1] = (sum 0.5|u|^2, max |u|^2) of the state AFTER step k + 1, for
 * k < n when the steps describe the state they write (LBM_MOMENTS_OF_OUTPUT) and for k < n - 1 when they describe the
 * state they read (LBM_MOMENTS_OF_INPUT: the last state's moments would need a step n + 1; the caller reduces it with
 * lbm_reduce or lets the next batch deliver it).  A reporter with interval 1
 * (lettuce/ext/_reporter/observable_reporter.py:185-200) costs one library call per batch instead of one per step. */
int lbm_step_moments_n(const lbm_step_desc *desc, void *d_f_a, void *d_f_b, int64_t n, void *d_scratch,
                       size_t scratch_bytes, double *d_results, void *stream);

/* Link-wise bounce-back boundaries applied AFTER streaming -- the "efficient bounce-back" boundaries of the
 * reference's example project examples/advanced_projects/efficient_bounce_back_obstacle: its EbbSimulation runs
 * collide, stream, then every post_streaming_boundary (simulation/ebb_simulation.py:71-104).  A link is (node, q)
 * with population q pointing from a fluid node into a solid node:
 *   FULLWAY       link stored on the SOLID node: slot (opposite(q), node) <- slot (q, node) after streaming
 *                 (boundary/fullway_bounce_back_boundary.py:132-155); force 2 sum e_q f_q (:157-170)
 *   HALFWAY       link stored on the FLUID node: slot (opposite(q), node) <- fc_q(node), fc = populations between
 *                 collision and streaming (boundary/halfway_bounce_back_boundary.py:167-182); force 2 sum e_q fc_q
 *   INTERPOLATED  Bouzidi's linear interpolation with wall distance d in (0, 1]
 *                 (boundary/linear_interpolated_bounce_back_boundary.py:57-100); force sum e_q (fc_q + bounced)
 * All right-hand sides are evaluated before any slot is written, as in the reference. */
typedef enum lbm_link_kind { LBM_LINK_FULLWAY = 0, LBM_LINK_HALFWAY = 1, LBM_LINK_INTERPOLATED = 2 } lbm_link_kind;

typedef struct lbm_links {
    int32_t kind; /* lbm_link_kind */
    int32_t _pad;
    int64_t n;            /* number of links (0 is allowed: nothing happens, force = 0) */
    const int32_t *node;  /* flat node index (x slowest) */
    const uint8_t *q;     /* population index of the link */
    const void *d;        /* INTERPOLATED: n wall distances, dtype of the lattice; else NULL */
    void *bounced;        /* scratch: n reals */
    double *force_scratch; /* scratch: lbm_links_scratch_doubles(n) doubles, or NULL when no force is wanted */
    double *force;        /* device, 3 doubles: momentum-exchange force in lattice units (x, y, z), or NULL */
} lbm_links;

int64_t lbm_links_scratch_doubles(int64_t n);

/* Applies one post-streaming boundary to `d_f_post`, the output of lbm_step(desc, d_f_pre, d_f_post): `d_f_pre`
 * must still hold the step's input (HALFWAY / INTERPOLATED re-evaluate the collide phase of their fluid nodes from
 * it).  desc->streaming must be LBM_POST_STREAMING (collide, then stream, then this call). */
int lbm_apply_links(const lbm_step_desc *desc, const lbm_links *links, const void *d_f_pre, void *d_f_post,
                    void *stream);

/* `n` time steps of EbbSimulation.__call__ without returning to the caller (ebb_simulation.py:93-104 when no reporter
 * is due): each step is lbm_step(desc, a, b) followed by lbm_apply_links for the `n_boundaries` entries of `links` in
 * order, then the buffers swap roles.  The newest populations end up in d_f_b if n is odd, in d_f_a if n is even. */
int lbm_step_links_n(const lbm_step_desc *desc, const lbm_links *links, int32_t n_boundaries, void *d_f_a,
                     void *d_f_b, int64_t n, void *stream);

/* `n` consecutive steps ping-ponging between two buffers (a -> b -> a ...), without returning
 * to the caller in between: the loop `for _ in range(num_steps)` of Simulation.__call__
 * (lettuce/_simulation.py:317-318) when no reporter is due.  The newest populations end up in
 * d_f_b if n is odd and in d_f_a if n is even.  On lattices of up to LBM_B200_GRAPH_MAX_NODES nodes
 * (environment variable, default 2^20, 0 = off) without boundaries, batches of >= 32 steps are replayed from
 * a cached CUDA graph of 32 steps (launch-latency bound regime); results are identical to n lbm_step calls.
 * Consecutive steps are chained with programmatic dependent launch (LBM_B200_PDL=0 switches that off). */
int lbm_step_n(const lbm_step_desc *desc, void *d_f_a, void *d_f_b, int64_t n, void *stream);

/* Builds the per-node label byte and frozen-slot word from lettuce's masks
 * (`no_collision_mask` uint8 [nx,ny,nz] and `no_streaming_mask` uint8 [q,nx,ny,nz],
 * lettuce/_simulation.py:100-146).  label = ncm value, with bit 7 set on every node
 * that needs the general path: a label other than the collision's, a frozen slot of its own,
 * a neighbour slot it would stream into that is frozen, or membership in an outlet plane of `desc`.
 * Set-up call: it waits for `stream` before returning (its small staging buffer lives on the caller's stack). */
int lbm_pack_masks(const lbm_step_desc *desc, const uint8_t *d_ncm, const uint8_t *d_nsm,
                   uint8_t *d_labels, uint32_t *d_frozen, void *stream);

/* Compacts the indices of the nodes with label bit 7 set into d_list (capacity entries) and writes
 * their total number to *d_count (device int64; may exceed capacity, then only `capacity` entries
 * were stored -- call with capacity 0 first to size the list).  Order is unspecified. */
int lbm_list_general_nodes(const lbm_lattice *lat, const uint8_t *d_labels, int32_t *d_list, int64_t capacity,
                           int64_t *d_count, void *stream);

/* Density [nx,ny,nz] and velocity [d,nx,ny,nz] fields (either may be NULL). */
int lbm_moments(const lbm_lattice *lat, const void *d_f, void *d_rho, void *d_u, void *stream);

/* Equilibrium populations f_q = feq_q(rho, u) for every node (QuadraticEquilibrium.__call__,
 * lettuce/ext/_equilibrium/quadratic_equilibrium.py:11-24, as used by Flow.initialize,
 * lettuce/_flow.py:127-143) written straight into d_f_out without full-size temporaries.
 * rho / u are fields in lattice units with element strides per spatial axis x,y,z (0 broadcasts
 * the axis); u has a leading component stride. */
int lbm_equilibrium(const lbm_lattice *lat, const void *d_rho, const int64_t rho_stride[3], const void *d_u,
                    const int64_t u_stride[4], void *d_f_out, void *stream);

/* Initial populations with the first-order non-equilibrium part, f_q = feq_q(rho, u) - w_q Q_q : Pi1 with
 * Pi1 = tau rho grad(u) / cs^2 from 6th-order periodic differences (initialize_f_neq, lettuce/_flow.py:341-367),
 * written in ONE pass from the density field d_rho [nx,ny,nz] and the velocity field d_u [d,nx,ny,nz] (lattice
 * units, e.g. from lbm_moments) without the reference's full-size temporaries.  `eye_cs2` is what the reference
 * subtracts on the diagonal of Q: cs^2 rounded to float32 (torch.eye in torch's default dtype, _flow.py:358-360). */
int lbm_initialize_fneq(const lbm_lattice *lat, const void *d_rho, const void *d_u, double tau, double eye_cs2,
                        void *d_f_out, void *stream);

typedef enum lbm_reduction {
    LBM_SUM_HALF_U2 = 0, /* sum over nodes of 0.5*|u|^2 in lattice units (_flow.py:200-204) */
    LBM_MAX_U = 1,       /* max over nodes of |u| in lattice units (observable_reporter.py:27-31) */
    LBM_SUM_F = 2,       /* sum of all populations */
    LBM_SUM_F_INNER = 3, /* sum of f[..., 1:-1, 1:-1] (observable_reporter.py:155) */
    LBM_SUM_F_MASKED = 4,/* sum over q and nodes of f*mask, mask uint8 [nx,ny,nz] (observable_reporter.py:157) */
    LBM_ENSTROPHY = 5    /* sum of |curl u|^2 with 6th-order periodic differences, lattice units,
                            dx = 1 (observable_reporter.py:45-68, util/utility.py:56-58,88-98);
                            needs d_u from lbm_moments; an optional uint8 node mask restricts the sum
                            (used for slabs extended by the stencil radius) */
} lbm_reduction;

/* Bytes of scratch lbm_reduce needs for this lattice. */
size_t lbm_reduce_scratch_bytes(const lbm_lattice *lat);

/* Deterministic two-stage reduction.  `d_in` is f for the SUM_F*, SUM_HALF_U2 and MAX_U
 * kinds and the velocity field [d,nx,ny,nz] for LBM_ENSTROPHY.  The result is written as
 * ONE double to d_out (device). */
int lbm_reduce(const lbm_lattice *lat, int what, const void *d_in, const uint8_t *d_mask,
               void *d_scratch, double *d_out, void *stream);

/* End-to-end entry with HOST buffers: uploads h_f (q*nx*ny*nz reals) to the device,
 * runs `nsteps` time steps of `desc` (desc->labels/frozen and op field pointers are
 * device pointers as for lbm_step; halo must be all NULL), downloads the final
 * populations into h_f_out and synchronises.  If h_energy is non-NULL it must hold
 * nsteps doubles and receives sum 0.5|u|^2 after every step (a reporter with
 * interval 1; reduced inside the step kernels, lbm_step_moments).  h_f_out may equal h_f.  The device work space
 * (two population buffers + scratch) is kept for the next call with the same size on the same device;
 * lbm_run_host_release() frees it. */
int lbm_run_host(const lbm_step_desc *desc, const void *h_f, void *h_f_out, int64_t nsteps,
                 double *h_energy);
int lbm_run_host_release(void);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU x-slabs, one process per GPU (no counterpart in the reference, which is single-device:
 * lettuce/_context.py:60 stores one torch.device).  Each rank owns nx_local planes; the step kernel
 * of a rank loads (pull) or stores (push) the neighbour ranks' boundary planes DIRECTLY through
 * peer-mapped pointers over NVLink -- there are no ghost planes, no pack kernels and no separate
 * exchange step.  Buffers that peers must see are allocated with lbm_ipc_alloc and opened on the
 * neighbour with lbm_ipc_open (CUDA IPC).  Ranks stay in lock step through one 8-byte progress
 * counter per neighbour written over NVLink after every step.
 * ------------------------------------------------------------------------------------------- */
#define LBM_IPC_HANDLE_BYTES 64

/* cudaMalloc + zero fill + cudaIpcGetMemHandle.  `handle` receives LBM_IPC_HANDLE_BYTES bytes to ship
 * to the neighbour process. */
int lbm_ipc_alloc(size_t bytes, void **d_ptr, void *handle);
/* Map a neighbour's allocation into this process (cudaIpcOpenMemHandle, peer access enabled lazily). */
int lbm_ipc_open(const void *handle, void **d_ptr);
int lbm_ipc_close(void *d_ptr);
int lbm_ipc_free(void *d_ptr);

typedef struct lbm_slab {
    /* peer-mapped base addresses of the neighbours' two population buffers, in the same (a, b) order
     * as the arguments of lbm_slab_step_n.  lo = rank owning the planes below x = 0, hi = the rank
     * owning the planes from x = nx_local on (periodic ring). */
    void *lo_a, *lo_b, *hi_a, *hi_b;
    int32_t lo_nx, hi_nx;          /* slab thickness of the neighbours */
    /* peer-mapped addresses where this rank publishes the number of steps it has completed: slot 1 of
     * the lo neighbour's counter pair and slot 0 of the hi neighbour's */
    uint64_t *signal_lo, *signal_hi;
    /* this rank's own counter block (in lbm_ipc_alloc memory, at least 8 words): [0] written by lo,
     * [1] written by hi, [2]..[4] scratch of the in-kernel lock step */
    const uint64_t *wait_slots;
    uint64_t epoch;                /* steps completed before this call; identical on all ranks */
} lbm_slab;

/* n lock-stepped time steps on an x-slab.  Ranks publish the number of steps they have completed to both
 * neighbours and never run a boundary plane of step k+1 before both neighbours have completed step k,
 * which orders the peer reads/writes of consecutive steps.  This happens INSIDE the step kernels
 * (boundary-plane CTAs are scheduled first, wait for the neighbour's counter, and the last of them -- or, with
 * boundaries, the last CTA of the sparse general-nodes kernel -- publishes this rank's counter; interior CTAs never
 * wait); slabs thinner than two boundary layers use a 1-thread kernel per step.  A spin on a peer counter gives up
 * with a trap after LBM_B200_PEER_TIMEOUT_S seconds (default 600).  desc->halo's population pointers
 * are ignored (they are derived from `slab`); its label/frozen pointers are used as given. */
int lbm_slab_step_n(const lbm_step_desc *desc, const lbm_slab *slab, void *d_f_a, void *d_f_b, int64_t n,
                    void *stream);

/* ONE lock-stepped slab step (d_f_a -> d_f_b) with the fused reductions of lbm_step_moments over THIS rank's nodes
 * (d_result[0] = sum 0.5|u|^2, d_result[1] = max |u|^2; the caller combines the ranks' values). */
int lbm_slab_step_moments(const lbm_step_desc *desc, const lbm_slab *slab, void *d_f_a, void *d_f_b,
                          void *d_scratch, size_t scratch_bytes, double *d_result, void *stream);

/* Introspection. */
int lbm_abi_version(void);
const char *lbm_status_string(int status);
const char *lbm_last_cuda_error(void);
/* Number of kernel launches this library has issued since load (all streams). */
int64_t lbm_launch_count(void);
/* Name of the kernel variant lbm_step would launch for desc (static string). */
const char *lbm_step_variant_name(const lbm_step_desc *desc);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
