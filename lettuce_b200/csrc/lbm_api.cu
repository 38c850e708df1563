// lbm_api.cu -- the C ABI declared in include/lbm_b200.h: argument validation,
// descriptor -> kernel-parameter translation, dispatch, mask packing, host-buffer
// end-to-end entry.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include <cudaTypedefs.h>

#include "lbm_launch.cuh"

namespace lbm {

std::atomic<int64_t> g_launch_count{0};
static thread_local char g_cuda_error[256] = "";

template <class S, class R>
int launch_moments(const R *f, R *rho, R *u, int64_t N, cudaStream_t st);
template <class S, class R>
int launch_reduce(int what, const R *in, const uint8_t *mask, int n0, int n1, int n2, double *partials, double *out,
                  cudaStream_t st);
template <class S, class R>
int launch_init_fneq(const R *rho, const R *u, double tau_over_cs2, double eye_cs2, int n0, int n1, int n2, R *f,
                     cudaStream_t stream);
size_t reduce_scratch_bytes();
int launch_fold_sum(const double *partials, int n, double *stage, double *out, cudaStream_t st);
size_t fold_pair_stage_bytes();
int launch_fold_pair(const double *partials, int n, double *stage, double *out, cudaStream_t st);
template <class S, class R>
int launch_equilibrium(const R *rho, const int64_t *rs, const R *u, const int64_t *us, int n0, int n1, int n2, R *f,
                       cudaStream_t stream);

// collision dispatch over the separately compiled (stencil, dtype, collision) units
template <class S, class R>
static int launch_step(const StepParams<R> &p, int coll, int streaming, const LaunchOptions &opt, cudaStream_t stream) {
    switch (coll) {
        case LBM_OP_NO_COLLISION: return launch_step_coll<S, R, LBM_OP_NO_COLLISION>(p, streaming, opt, stream);
        case LBM_OP_BGK: return launch_step_coll<S, R, LBM_OP_BGK>(p, streaming, opt, stream);
        case LBM_OP_TRT: return launch_step_coll<S, R, LBM_OP_TRT>(p, streaming, opt, stream);
        case LBM_OP_KBC:
            // KBC exists for D2Q9 and D3Q27 only (kbc_collision.py:101,116)
            if constexpr (S::ID == LBM_D3Q19) return LBM_ERR_UNSUPPORTED;
            else return launch_step_coll<S, R, LBM_OP_KBC>(p, streaming, opt, stream);
        case LBM_OP_REGULARIZED: return launch_step_coll<S, R, LBM_OP_REGULARIZED>(p, streaming, opt, stream);
        case LBM_OP_SMAGORINSKY: return launch_step_coll<S, R, LBM_OP_SMAGORINSKY>(p, streaming, opt, stream);
        case LBM_OP_BGK_FORCED: return launch_step_coll<S, R, LBM_OP_BGK_FORCED>(p, streaming, opt, stream);
    }
    return LBM_ERR_BAD_ARGUMENT;
}

template <class S, class R>
static int launch_links(const StepParams<R> &p, const LinkArgs<R> &a, int coll, cudaStream_t stream) {
    switch (coll) {
        case LBM_OP_NO_COLLISION: return launch_links_coll<S, R, LBM_OP_NO_COLLISION>(p, a, stream);
        case LBM_OP_BGK: return launch_links_coll<S, R, LBM_OP_BGK>(p, a, stream);
        case LBM_OP_TRT: return launch_links_coll<S, R, LBM_OP_TRT>(p, a, stream);
        case LBM_OP_KBC:
            if constexpr (S::ID == LBM_D3Q19) return LBM_ERR_UNSUPPORTED;
            else return launch_links_coll<S, R, LBM_OP_KBC>(p, a, stream);
        case LBM_OP_REGULARIZED: return launch_links_coll<S, R, LBM_OP_REGULARIZED>(p, a, stream);
        case LBM_OP_SMAGORINSKY: return launch_links_coll<S, R, LBM_OP_SMAGORINSKY>(p, a, stream);
        case LBM_OP_BGK_FORCED: return launch_links_coll<S, R, LBM_OP_BGK_FORCED>(p, a, stream);
    }
    return LBM_ERR_BAD_ARGUMENT;
}

// Nodes per thread of the bulk kernel: lbm_step_desc::variant (1 or 2), else the environment variable
// LBM_B200_LANES (A/B measurements), else the default: two nodes per thread (packed fp32 arithmetic, lbm_vec.cuh)
// wherever that kernel exists -- fp32, even contiguous extent, PRE / POST streaming.  Both give the same bits.
static int chosen_lanes(const lbm_step_desc *d) {
    const int n2 = d->lat.stencil == LBM_D2Q9 ? d->lat.ny : d->lat.nz;
    if (!lanes2_available(d->lat.dtype, d->streaming, n2)) return 1;
    int want = d->variant;
    if (want != 1 && want != 2) {
        const char *e = getenv("LBM_B200_LANES");
        want = e ? atoi(e) : 0;
    }
    if (want == 1 || want == 2) return want;
    return 2;
}

// ---------------------------------------------------------------------------------------------------------
// TMA-staged kernel (lbm_tma.cuh): which steps take it, and the tensor maps of their buffers.
// lbm_step_desc::variant 3 asks for it, 1 / 2 for the LDG kernel; otherwise LBM_B200_TMA=0|1 decides, else the
// default at the end of tma_wanted.  Steps that carry the slab lock step or fused reductions always run the LDG kernel.
// ---------------------------------------------------------------------------------------------------------
static bool tma_wanted(const lbm_step_desc *d, bool slab = false) {
    const bool two_d = d->lat.stencil == LBM_D2Q9;
    const int n2 = two_d ? d->lat.ny : d->lat.nz;
    const int64_t nodes = (int64_t)d->lat.nx * d->lat.ny * d->lat.nz;
    if (!tma_available(d->lat.dtype, nodes, n2) || d->streaming == LBM_DOUBLE_STREAMING) return false;
    const lbm_halo &h = d->halo;
    // slabs: the neighbours' planes are not in the tensor -- the staged kernel takes the interior planes only, of
    // unmasked pulling steps (lbm_step_inst.cu); otherwise the peer planes rule it out
    if (slab) {
        if (d->labels || d->streaming != LBM_PRE_STREAMING || d->lat.nx < 4) return false;
    } else if (h.in_lo || h.in_hi || h.out_lo || h.out_hi) {
        return false;
    }
    if (d->variant == 3) return true;
    if (d->variant == 1 || d->variant == 2) return false;
    if (const char *e = getenv("LBM_B200_TMA")) return e[0] != '0';
    // Default: where it measured faster than the LDG kernel under sustained load (profiles/r2_tma_sweep.md) -- the
    // entropic operator on D3Q27, whose LDG kernel is bound by its issue rate and by the power cap, not by HBM
    // (512^3: 0.95 of the copy bandwidth instead of 0.86).  The bandwidth-bound operators are equal on D3Q27 and a
    // few percent slower on the smaller velocity sets (shorter tiles, same per-tile overheads).
    return d->lat.stencil == LBM_D3Q27 && d->ops[d->collision_index].kind == LBM_OP_KBC;
}

namespace {
struct TmaMapSlot {
    const void *ptr = nullptr;
    int n0 = 0, n1 = 0, n2 = 0, q = 0, b0 = 0, b1 = 0, b2 = 0, device = -1;
    CUtensorMap map;
    uint64_t used = 0;
};
std::mutex g_tma_mutex;
TmaMapSlot g_tma_slots[32];
uint64_t g_tma_clock = 0;

PFN_cuTensorMapEncodeTiled tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (PFN_cuTensorMapEncodeTiled)p;
    }();
    return fn;
}

// tensor map of one population buffer: fp32 [q][n0][n1][n2], box = (b0, b1, b2, 1) along (z, y, x, q); small LRU
// cache (a simulation alternates between two buffers)
int tensor_map_of(const void *ptr, int n0, int n1, int n2, int q, int b0, int b1, int b2, CUtensorMap *out) {
    int device = -1;
    if (cudaGetDevice(&device)) return LBM_ERR_CUDA;
    std::lock_guard<std::mutex> lock(g_tma_mutex);
    TmaMapSlot *victim = &g_tma_slots[0];
    for (TmaMapSlot &s : g_tma_slots) {
        if (s.ptr == ptr && s.n0 == n0 && s.n1 == n1 && s.n2 == n2 && s.q == q && s.b0 == b0 && s.b1 == b1 &&
            s.b2 == b2 && s.device == device) {
            s.used = ++g_tma_clock;
            *out = s.map;
            return LBM_OK;
        }
        if (s.used < victim->used) victim = &s;
    }
    PFN_cuTensorMapEncodeTiled encode = tensor_map_encoder();
    if (!encode) return LBM_ERR_UNSUPPORTED;
    const cuuint64_t dims[4] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)n0, (cuuint64_t)q};
    const cuuint64_t strides[3] = {(cuuint64_t)n2 * 4, (cuuint64_t)n1 * n2 * 4, (cuuint64_t)n0 * n1 * n2 * 4};
    const cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap m;
    const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return LBM_ERR_UNSUPPORTED;
    victim->ptr = ptr; victim->n0 = n0; victim->n1 = n1; victim->n2 = n2; victim->q = q;
    victim->b0 = b0; victim->b1 = b1; victim->b2 = b2;
    victim->device = device; victim->map = m; victim->used = ++g_tma_clock;
    *out = m;
    return LBM_OK;
}

// tile counters of the TMA-staged kernel: two words per (device, stream), zero between launches (the kernel rearms
// them itself), never freed
unsigned *tma_counters_of(cudaStream_t stream) {
    static std::mutex mutex;
    static std::map<std::pair<int, cudaStream_t>, unsigned *> slots;
    int device = -1;
    if (cudaGetDevice(&device)) return nullptr;
    std::lock_guard<std::mutex> lock(mutex);
    auto it = slots.find({device, stream});
    if (it != slots.end()) return it->second;
    unsigned *p = nullptr;
    if (cudaMalloc(&p, 2 * sizeof(unsigned)) || cudaMemset(p, 0, 2 * sizeof(unsigned))) return nullptr;
    slots[{device, stream}] = p;
    return p;
}

int device_sm_count() {
    static int cached[64] = {};
    int device = 0;
    if (cudaGetDevice(&device) || device < 0 || device >= 64) return 148;
    if (!cached[device]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) || n <= 0) n = 148;
        cached[device] = n;
    }
    return cached[device];
}
}  // namespace

static const char *step_variant_name(const lbm_step_desc *d, bool masked) {
    if (tma_wanted(d)) return masked ? "step_tma_kernel<TMA-staged, 2 nodes/thread, packed fp32> + general_nodes" :
                                       "step_tma_kernel<TMA-staged, 2 nodes/thread, packed fp32>";
    if (chosen_lanes(d) == 2) return masked ? "step_kernel<2 nodes/thread, packed fp32> + general_nodes" :
                                              "step_kernel<2 nodes/thread, packed fp32>";
    return masked ? "step_kernel<1 node/thread> + general_nodes" : "step_kernel<1 node/thread>";
}

// spins on peer progress counters give up after this long (LBM_B200_PEER_TIMEOUT_S, default 600 s), in SM clocks
static unsigned long long peer_timeout_cycles() {
    const char *e = getenv("LBM_B200_PEER_TIMEOUT_S");
    double s = e ? atof(e) : 600.0;
    if (!(s > 0)) s = 600.0;
    return (unsigned long long)(s * 2.0e9);
}

static int cuda_fail(int e) {
    if (e <= 0) return e;  // already an lbm_status
    snprintf(g_cuda_error, sizeof g_cuda_error, "%s: %s", cudaGetErrorName((cudaError_t)e),
             cudaGetErrorString((cudaError_t)e));
    return LBM_ERR_CUDA;
}

int cuda_fail_public(int e) { return cuda_fail(e); }
unsigned long long peer_timeout_cycles_public() { return peer_timeout_cycles(); }

struct Dims {
    int n0, n1, n2, d, q;
};

static int lattice_dims(const lbm_lattice *lat, Dims &dm) {
    if (!lat) return LBM_ERR_BAD_ARGUMENT;
    if (lat->dtype != LBM_F32 && lat->dtype != LBM_F64) return LBM_ERR_BAD_ARGUMENT;
    if (lat->nx < 1 || lat->ny < 1 || lat->nz < 1) return LBM_ERR_BAD_ARGUMENT;
    switch (lat->stencil) {
        case LBM_D2Q9:
            if (lat->nz != 1) return LBM_ERR_BAD_ARGUMENT;
            dm = {lat->nx, 1, lat->ny, 2, 9};  // (x, -, y): contiguous axis last
            break;
        case LBM_D3Q19: dm = {lat->nx, lat->ny, lat->nz, 3, 19}; break;
        case LBM_D3Q27: dm = {lat->nx, lat->ny, lat->nz, 3, 27}; break;
        default: return LBM_ERR_BAD_ARGUMENT;
    }
    if ((int64_t)dm.n0 * dm.n1 * dm.n2 >= (int64_t)1 << 31) return LBM_ERR_TOO_LARGE;
    return LBM_OK;
}

static bool is_collision(int kind) { return kind >= LBM_OP_NO_COLLISION && kind <= LBM_OP_BGK_FORCED; }
static bool is_outlet(int kind) { return kind == LBM_OP_OUTLET_P || kind == LBM_OP_ANTI_BOUNCE_BACK; }

static int validate_desc(const lbm_step_desc *d, Dims &dm) {
    if (!d) return LBM_ERR_BAD_ARGUMENT;
    int rc = lattice_dims(&d->lat, dm);
    if (rc) return rc;
    if (d->streaming < 0 || d->streaming > 3) return LBM_ERR_BAD_ARGUMENT;
    if (d->n_ops < 1 || d->n_ops > LBM_MAX_OPS) return LBM_ERR_BAD_ARGUMENT;
    if (d->collision_index < 0 || d->collision_index >= d->n_ops) return LBM_ERR_BAD_ARGUMENT;
    if (!is_collision(d->ops[d->collision_index].kind)) return LBM_ERR_BAD_ARGUMENT;
    for (int i = 0; i < d->n_ops; ++i) {
        const lbm_op &op = d->ops[i];
        if (i != d->collision_index && is_collision(op.kind)) return LBM_ERR_BAD_ARGUMENT;
        if (i != d->collision_index && op.kind != LBM_OP_BOUNCE_BACK && op.kind != LBM_OP_EQUILIBRIUM &&
            op.kind != LBM_OP_IDENTITY && !is_outlet(op.kind))
            return LBM_ERR_BAD_ARGUMENT;
        if (op.kind == LBM_OP_EQUILIBRIUM && (!op.rho || !op.u)) return LBM_ERR_BAD_ARGUMENT;
        if (is_outlet(op.kind)) {
            if (op.axis < 0 || op.axis >= dm.d || (op.side != 1 && op.side != -1 && op.side != 0))
                return LBM_ERR_BAD_ARGUMENT;
            const int n = op.axis == 0 ? d->lat.nx : (op.axis == 1 ? d->lat.ny : d->lat.nz);
            if (n < 2) return LBM_ERR_BAD_ARGUMENT;
            // (an outlet's neighbour may lie on the plane of an earlier outlet -- planes of different axes meet, or
            // both ends of a three-plane axis: the sparse kernel evaluates that neighbour's own outlet first,
            // pipeline_prefix; the nesting is bounded by the number of active outlets)
        }
    }
    int active_outlets = 0;
    for (int i = 0; i < d->n_ops; ++i) active_outlets += is_outlet(d->ops[i].kind) && d->ops[i].side != 0;
    if (active_outlets > kMaxOutletDepth + 1) return LBM_ERR_UNSUPPORTED;
    if (d->n_ops > 1 && (!d->labels || !d->frozen)) return LBM_ERR_BAD_ARGUMENT;
    if ((d->labels == nullptr) != (d->frozen == nullptr)) return LBM_ERR_BAD_ARGUMENT;
    if (d->n_general < 0 || (d->n_general > 0 && !d->general_nodes)) return LBM_ERR_BAD_ARGUMENT;
    if (d->ops[d->collision_index].kind == LBM_OP_KBC && d->lat.stencil == LBM_D3Q19) return LBM_ERR_UNSUPPORTED;
    return LBM_OK;
}

template <class S, class R>
static void fill_params(const lbm_step_desc *d, const Dims &dm, const void *f_in, void *f_out, StepParams<R> &p) {
    memset(&p, 0, sizeof p);
    p.in = (const R *)f_in;
    p.out = (R *)f_out;
    p.n0 = dm.n0; p.n1 = dm.n1; p.n2 = dm.n2;
    p.N = (int64_t)dm.n0 * dm.n1 * dm.n2;
    const int64_t plane = (int64_t)dm.n1 * dm.n2;
    const lbm_halo &h = d->halo;
    // periodic wrap inside the buffer unless the caller supplied neighbour planes
    p.in_lo = h.in_lo ? (const R *)h.in_lo : p.in + (dm.n0 - 1) * plane;
    p.in_lo_qs = h.in_lo ? h.in_lo_qstride : p.N;
    p.in_hi = h.in_hi ? (const R *)h.in_hi : p.in;
    p.in_hi_qs = h.in_hi ? h.in_hi_qstride : p.N;
    p.out_lo = h.out_lo ? (R *)h.out_lo : p.out + (dm.n0 - 1) * plane;
    p.out_lo_qs = h.out_lo ? h.out_lo_qstride : p.N;
    p.out_hi = h.out_hi ? (R *)h.out_hi : p.out;
    p.out_hi_qs = h.out_hi ? h.out_hi_qstride : p.N;
    // address tables of the bulk kernel (AddrTables): entry [k][q] + (x * plane + row + column)
    {
        const bool pull = d->streaming & LBM_PRE_STREAMING, push = d->streaming & LBM_POST_STREAMING;
        const int64_t last = (int64_t)(dm.n0 - 1) * plane;
        for (int k = 0; k < 4; ++k) {
            for (int q = 0; q < S::Q; ++q) {
                const int e0 = S::e(q, 0);
                const R *ld = p.in + q * p.N - (pull ? e0 * plane : 0);
                if (pull && (k & 1) && e0 == 1) ld = p.in_lo + q * p.in_lo_qs;              // x = 0 reads plane -1
                if (pull && (k & 2) && e0 == -1) ld = p.in_hi + q * p.in_hi_qs - last;      // x = n0-1 reads plane n0
                p.tbl.ld[k][q] = ld;
                R *st = p.out + q * p.N + (push ? e0 * plane : 0);
                if (push && (k & 1) && e0 == -1) st = p.out_lo + q * p.out_lo_qs;           // x = 0 writes plane -1
                if (push && (k & 2) && e0 == 1) st = p.out_hi + q * p.out_hi_qs - last;     // x = n0-1 writes plane n0
                p.tbl.st[k][q] = st;
            }
        }
    }
    p.labels = d->labels;
    p.frozen = d->frozen;
    p.labels_lo = h.label_lo ? h.label_lo : (d->labels ? d->labels + (dm.n0 - 1) * plane : nullptr);
    p.labels_hi = h.label_hi ? h.label_hi : d->labels;
    p.frozen_lo = h.frozen_lo ? h.frozen_lo : (d->frozen ? d->frozen + (dm.n0 - 1) * plane : nullptr);
    p.frozen_hi = h.frozen_hi ? h.frozen_hi : d->frozen;
    p.general_nodes = d->general_nodes;
    p.n_general = (int)d->n_general;
    p.n_ops = d->n_ops;
    p.collision_index = d->collision_index;
    int active_outlets = 0;
    for (int i = 0; i < d->n_ops; ++i) active_outlets += is_outlet(d->ops[i].kind) && d->ops[i].side != 0;
    p.nested_outlets = active_outlets >= 2 ? 1 : 0;
    const lbm_op &c = d->ops[d->collision_index];
    collision_scalars<R>(c.kind, c.p0, c.p1, p.ca, p.cb);
    for (int a = 0; a < 3; ++a) p.force.a[a] = R(0);
    for (int c2 = 0; c2 < S::D; ++c2) p.force.a[S::axis_of(c2)] = (R)c.force[c2];
    p.force.ueq_scale = (R)c.ueq_scale;
    p.force.src_scale = (R)c.source_scale;
    for (int i = 0; i < d->n_ops; ++i) {
        const lbm_op &o = d->ops[i];
        OpDev<R> &t = p.ops[i];
        t.kind = o.kind;
        t.axis = (o.axis >= 0 && o.axis < S::D) ? S::axis_of(o.axis) : 0;
        t.side = o.side;
        collision_scalars<R>(o.kind, o.p0, o.p1, t.a, t.b);
        if (o.kind == LBM_OP_OUTLET_P) t.a = (R)o.p0;
        t.rho = (const R *)o.rho;
        t.u = (const R *)o.u;
        for (int a = 0; a < 3; ++a) t.rho_stride[a] = 0;
        for (int a = 0; a < 4; ++a) t.u_stride[a] = 0;
        t.u_stride[0] = o.u_stride[0];
        for (int c2 = 0; c2 < S::D; ++c2) {
            t.rho_stride[S::axis_of(c2)] = o.rho_stride[c2];
            t.u_stride[1 + S::axis_of(c2)] = o.u_stride[1 + c2];
        }
    }
}

template <class S, class R>
static int step_typed(const lbm_step_desc *d, const Dims &dm, const void *f_in, void *f_out, const StepExtras &x,
                      cudaStream_t st) {
    StepParams<R> p;
    fill_params<S, R>(d, dm, f_in, f_out, p);
    if (x.sync) p.sync = *x.sync;
    p.energy_partials = x.partials;
    p.reduce_mode = x.partials ? x.reduce_mode : kReduceNone;
    // consecutive steps sweep the x-planes in opposite directions (L2 reuse, StepParams::reverse_sweep); the result
    // does not depend on the order.  LBM_B200_ALTERNATE_SWEEP=0 switches it off (A/B).
    static std::atomic<unsigned> sweep{0};
    static const bool alternate = [] {
        const char *e = getenv("LBM_B200_ALTERNATE_SWEEP");
        return !(e && e[0] == '0');
    }();
    p.reverse_sweep = alternate ? (int)(sweep.fetch_add(1) & 1u) : 0;
    LaunchOptions opt;
    opt.lanes = chosen_lanes(d);
    opt.chained = x.chained;
    TmaMaps maps;
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &capturing);      // (graph replays may run concurrently: they keep the LDG kernel)
    const bool slab_step = x.sync != nullptr && x.sync->on;
    // (fused reductions: the staged kernel has them for unmasked pulling steps on one GPU)
    const bool reduce_ok = !x.partials || (!slab_step && x.reduce_mode == kReduceOutput && !d->labels &&
                                           d->streaming == LBM_PRE_STREAMING);      // (the pushing kernel has none)
    opt.slots_used = x.slots_used;
    if (reduce_ok && tma_wanted(d, slab_step) && capturing == cudaStreamCaptureStatusNone) {
        const int tz = tma_row_extent(dm.n2), rows = tma_tile_rows(dm.n2);
        const bool boxable = tma_rows_boxable(dm.n0, dm.n1, dm.n2);
        const int by = dm.n1 > 1 ? rows : 1, bx = dm.n1 > 1 ? 1 : rows;      // rows run along y (3-D) or x (2-D)
        int rc = tensor_map_of(f_in, dm.n0, dm.n1, dm.n2, dm.q, tz, 1, 1, &maps.in_row);
        if (!rc) rc = tensor_map_of(f_out, dm.n0, dm.n1, dm.n2, dm.q, tz, 1, 1, &maps.out_row);
        if (boxable) {
            if (!rc) rc = tensor_map_of(f_in, dm.n0, dm.n1, dm.n2, dm.q, tz, by, bx, &maps.in_box);
            if (!rc) rc = tensor_map_of(f_in, dm.n0, dm.n1, dm.n2, dm.q, 4, by, bx, &maps.in_halo);
            if (!rc) rc = tensor_map_of(f_out, dm.n0, dm.n1, dm.n2, dm.q, tz, by, bx, &maps.out_box);
        } else {
            maps.in_box = maps.in_halo = maps.in_row;                         // never used
            maps.out_box = maps.out_row;
        }
        if (!rc && !(opt.tma_counters = tma_counters_of(st))) rc = LBM_ERR_CUDA;
        if (!rc) {
            opt.tma = &maps;
            opt.tma_interior = slab_step;
            opt.tma_boxable = boxable ? 1 : 0;
            opt.sm_count = device_sm_count();
        } else if (d->variant == 3) {
            return rc;                      // asked for explicitly: no silent change of kernel
        }
    } else if (d->variant == 3 && !x.sync && !x.partials) {
        return LBM_ERR_UNSUPPORTED;
    }
    return cuda_fail(launch_step<S, R>(p, d->ops[d->collision_index].kind, d->streaming, opt, st));
}

#define LBM_DISPATCH(stencil, dtype, ...)                                   \
    do {                                                                      \
        if ((dtype) == LBM_F32) {                                             \
            using R = float;                                                  \
            if ((stencil) == LBM_D2Q9) { using S = D2Q9; __VA_ARGS__; }              \
            else if ((stencil) == LBM_D3Q19) { using S = D3Q19; __VA_ARGS__; }       \
            else { using S = D3Q27; __VA_ARGS__; }                                   \
        } else {                                                              \
            using R = double;                                                 \
            if ((stencil) == LBM_D2Q9) { using S = D2Q9; __VA_ARGS__; }              \
            else if ((stencil) == LBM_D3Q19) { using S = D3Q19; __VA_ARGS__; }       \
            else { using S = D3Q27; __VA_ARGS__; }                                   \
        }                                                                     \
    } while (0)

// ---------------------------------------------------------------------------
// mask packing (lettuce/_simulation.py:100-146 -> label byte + frozen-slot word)
// ---------------------------------------------------------------------------
template <class S>
__global__ void pack_masks_kernel(const uint8_t *__restrict__ ncm, const uint8_t *__restrict__ nsm, int n0, int n1,
                                  int n2, int n_ops, int collision_index, const int *__restrict__ op_kind,
                                  const int *__restrict__ op_axis, const int *__restrict__ op_side,
                                  uint8_t *labels, uint32_t *frozen) {
    const int64_t N = (int64_t)n0 * n1 * n2;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(n % n2);
        const int y = (int)((n / n2) % n1);
        const int x = (int)(n / ((int64_t)n1 * n2));
        uint32_t own = 0;
        bool general = false;
        ForQ<S::Q>::run([&]<int q>() {
            if (nsm[q * N + n] == 1) own |= 1u << q;  // torch.eq(no_streaming_mask[i], 1), _simulation.py:254
            const int xd = wrap(x + S::e(q, 0), n0), yd = wrap(y + S::e(q, 1), n1), zd = wrap(z + S::e(q, 2), n2);
            if (nsm[q * N + ((int64_t)xd * n1 + yd) * n2 + zd] == 1) general = true;
        });
        if (own) general = true;
        if (ncm[n] != collision_index) general = true;
        for (int i = 0; i < n_ops; ++i) {
            if (op_kind[i] != LBM_OP_OUTLET_P && op_kind[i] != LBM_OP_ANTI_BOUNCE_BACK) continue;
            if (op_side[i] != 0 && in_plane_of(op_axis[i], op_side[i], x, y, z, n0, n1, n2)) general = true;
        }
        frozen[n] = own;
        labels[n] = (uint8_t)((ncm[n] & 0x7f) | (general ? kLabelGeneral : 0));
    }
}

template <class S>
static int pack_typed(const lbm_step_desc *d, const Dims &dm, const uint8_t *ncm, const uint8_t *nsm, uint8_t *labels,
                      uint32_t *frozen, cudaStream_t st) {
    int host[3][LBM_MAX_OPS];
    for (int i = 0; i < LBM_MAX_OPS; ++i) {
        host[0][i] = i < d->n_ops ? d->ops[i].kind : 0;
        const int ax = i < d->n_ops ? d->ops[i].axis : 0;
        host[1][i] = (ax >= 0 && ax < S::D) ? S::axis_of(ax) : 0;
        host[2][i] = i < d->n_ops ? d->ops[i].side : 0;
    }
    int *dev = nullptr;
    int e = (int)cudaMallocAsync(&dev, sizeof host, st);
    if (e) return e;
    e = (int)cudaMemcpyAsync(dev, host, sizeof host, cudaMemcpyHostToDevice, st);
    if (e) {
        cudaFreeAsync(dev, st);
        return e;
    }
    const int64_t N = (int64_t)dm.n0 * dm.n1 * dm.n2;
    int64_t b = (N + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    pack_masks_kernel<S><<<(int)b, 256, 0, st>>>(ncm, nsm, dm.n0, dm.n1, dm.n2, d->n_ops, d->collision_index, dev,
                                                  dev + LBM_MAX_OPS,
                                                  dev + 2 * LBM_MAX_OPS, labels, frozen);
    ++g_launch_count;
    e = (int)cudaGetLastError();
    cudaFreeAsync(dev, st);
    // the staging copy reads `host` from this stack frame
    cudaStreamSynchronize(st);
    return e;
}

__global__ void list_general_nodes_kernel(const uint8_t *__restrict__ labels, int64_t N, int32_t *list,
                                          long long capacity, unsigned long long *count) {
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        if (labels[n] & kLabelGeneral) {
            const unsigned long long slot = atomicAdd(count, 1ULL);
            if ((long long)slot < capacity) list[slot] = (int32_t)n;
        }
    }
}

}  // namespace lbm

using namespace lbm;

extern "C" {

int lbm_abi_version(void) { return LBM_ABI_VERSION; }

const char *lbm_status_string(int s) {
    switch (s) {
        case LBM_OK: return "ok";
        case LBM_ERR_BAD_ARGUMENT: return "bad argument (null pointer, bad extent or enum)";
        case LBM_ERR_UNSUPPORTED: return "unsupported combination (no kernel for this request)";
        case LBM_ERR_CUDA: return "CUDA runtime error (see lbm_last_cuda_error)";
        case LBM_ERR_ALIASING: return "f_in and f_out overlap";
        case LBM_ERR_TOO_LARGE: return "lattice has 2^31 nodes or more";
    }
    return "unknown status";
}

const char *lbm_last_cuda_error(void) { return g_cuda_error; }

int64_t lbm_launch_count(void) { return g_launch_count.load(); }

int lbm_step(const lbm_step_desc *desc, const void *d_f_in, void *d_f_out, void *stream) {
    return step_general(desc, d_f_in, d_f_out, StepExtras(), stream);
}

}  // extern "C"

namespace lbm {
// lbm_step plus the optional in-kernel slab lock step (lbm_slab_step_n), fused reductions (lbm_step_moments) and
// programmatic chaining behind the previous step (lbm_step_n)
int step_general(const lbm_step_desc *desc, const void *d_f_in, void *d_f_out, const StepExtras &x, void *stream) {
    Dims dm;
    int rc = validate_desc(desc, dm);
    if (rc) return rc;
    if (!d_f_in || !d_f_out) return LBM_ERR_BAD_ARGUMENT;
    const size_t bytes = (size_t)dm.q * dm.n0 * dm.n1 * dm.n2 * (desc->lat.dtype == LBM_F32 ? 4 : 8);
    const char *a = (const char *)d_f_in, *b = (const char *)d_f_out;
    if (a < b + bytes && b < a + bytes) return LBM_ERR_ALIASING;
    LBM_DISPATCH(desc->lat.stencil, desc->lat.dtype,
                 return (step_typed<S, R>(desc, dm, d_f_in, d_f_out, x, (cudaStream_t)stream)));
    return LBM_ERR_BAD_ARGUMENT;
}

// partial pairs a step with fused reductions may write (either bulk geometry)
static int max_reduce_slots(const lbm_step_desc *desc, const Dims &dm) {
    const int a = reduce_slots_for(dm.n0, dm.n1, dm.n2, 1, desc->n_general);
    const int b = reduce_slots_for(dm.n0, dm.n1, dm.n2, 2, desc->n_general);
    const int c = 1024;                       // the staged kernel: one pair per persistent CTA
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

// one step with fused reductions into (partials, stage) -> d_result[2]; shared by lbm_step_moments and the slab entry
int step_moments_general(const lbm_step_desc *desc, const void *d_f_in, void *d_f_out, const SlabSync *sync,
                         void *d_scratch, size_t scratch_bytes, double *d_result, bool chained, void *stream) {
    if (!d_scratch || !d_result) return LBM_ERR_BAD_ARGUMENT;
    const int state = lbm_step_moments_state(desc);
    if (state == LBM_MOMENTS_UNAVAILABLE) return LBM_ERR_UNSUPPORTED;
    const size_t need = lbm_step_moments_scratch_bytes(desc);
    if (need == 0 || scratch_bytes < need) return LBM_ERR_BAD_ARGUMENT;
    Dims dm;
    int rc = validate_desc(desc, dm);
    if (rc) return rc;
    StepExtras x;
    x.sync = sync;
    x.partials = (double *)d_scratch;
    x.reduce_mode = state == LBM_MOMENTS_OF_OUTPUT ? kReduceOutput : kReduceInput;
    x.chained = chained;
    int used = 0;
    x.slots_used = &used;
    rc = step_general(desc, d_f_in, d_f_out, x, stream);
    if (rc) return rc;
    const int slots = used > 0 ? used : reduce_slots_for(dm.n0, dm.n1, dm.n2, chosen_lanes(desc), desc->n_general);
    return cuda_fail(launch_fold_pair((const double *)d_scratch, slots, (double *)d_scratch + 2 * (size_t)max_reduce_slots(desc, dm),
                                      d_result, (cudaStream_t)stream));
}
}  // namespace lbm

extern "C" {

int lbm_step_moments_state(const lbm_step_desc *desc) {
    if (!desc) return LBM_MOMENTS_UNAVAILABLE;
    switch (desc->streaming) {
        case LBM_NO_STREAMING:
        case LBM_PRE_STREAMING: return LBM_MOMENTS_OF_OUTPUT;   // the node's output is still in registers
        case LBM_POST_STREAMING: return LBM_MOMENTS_OF_INPUT;   // the node's input is the previous step's output
    }
    return LBM_MOMENTS_UNAVAILABLE;                             // DOUBLE_STREAMING: neither
}

size_t lbm_step_moments_scratch_bytes(const lbm_step_desc *desc) {
    Dims dm;
    if (!desc || lattice_dims(&desc->lat, dm) || lbm_step_moments_state(desc) == LBM_MOMENTS_UNAVAILABLE) return 0;
    // one (sum, max) pair per CTA of the step's kernels + the staging area of the two-stage fold
    return 2 * sizeof(double) * (size_t)max_reduce_slots(desc, dm) + fold_pair_stage_bytes();
}

int lbm_step_moments(const lbm_step_desc *desc, const void *d_f_in, void *d_f_out, void *d_scratch,
                     size_t scratch_bytes, double *d_result, void *stream) {
    return step_moments_general(desc, d_f_in, d_f_out, nullptr, d_scratch, scratch_bytes, d_result, false, stream);
}

int lbm_step_moments_n(const lbm_step_desc *desc, void *d_f_a, void *d_f_b, int64_t n, void *d_scratch,
                       size_t scratch_bytes, double *d_results, void *stream) {
    if (n < 0 || !desc || !d_results) return LBM_ERR_BAD_ARGUMENT;
    const int state = lbm_step_moments_state(desc);
    if (state == LBM_MOMENTS_UNAVAILABLE) return LBM_ERR_UNSUPPORTED;
    void *a = d_f_a, *b = d_f_b;
    for (int64_t k = 0; k < n; ++k) {
        int rc;
        if (state == LBM_MOMENTS_OF_OUTPUT) {
            rc = step_moments_general(desc, a, b, nullptr, d_scratch, scratch_bytes, d_results + 2 * k, k > 0, stream);
        } else if (k == 0) {
            rc = lbm_step(desc, a, b, stream);       // describes the caller's state: nothing to report
        } else {
            rc = step_moments_general(desc, a, b, nullptr, d_scratch, scratch_bytes, d_results + 2 * (k - 1), true,
                                      stream);
        }
        if (rc) return rc;
        void *t = a; a = b; b = t;
    }
    return LBM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// CUDA-graph replay of step batches on small lattices.  Below a few hundred thousand nodes one step takes a few
// microseconds and the loop is bound by launch latency; kGraphSteps consecutive steps (a -> b -> a ...) are
// captured once per (descriptor, buffer pair, device) into an executable graph and replayed (measured on B200:
// D2Q9 fp64 256^2 4.3 -> 2.4 us/step, D3Q19 fp64 64^3 11.5 -> 9.4 us/step; bit-identical populations).
// LBM_B200_GRAPH_MAX_NODES=<largest lattice, in nodes, that takes this path> (default 2^20; 0 = never).
// ---------------------------------------------------------------------------------------------------------
namespace {

constexpr int kGraphSteps = 32;   // even: a replay leaves the newest populations where they were before it
constexpr int kGraphSlots = 4;

struct GraphSlot {
    lbm_step_desc desc;
    void *a = nullptr, *b = nullptr;
    int device = -1;
    cudaGraphExec_t exec = nullptr;
    uint64_t used = 0;
};
std::mutex g_graph_mutex;
GraphSlot g_graph_slots[kGraphSlots];
uint64_t g_graph_clock = 0;

bool pdl_enabled() {                   // LBM_B200_PDL=0 switches the programmatic chaining of steps off (A/B)
    const char *e = getenv("LBM_B200_PDL");
    return !(e && e[0] == '0');
}

int64_t graph_max_nodes() {           // read per call: cheap, and a host program may switch it at run time
    const char *e = getenv("LBM_B200_GRAPH_MAX_NODES");
    return e ? (int64_t)atoll(e) : (int64_t)1 << 20;
}

// captures kGraphSteps steps on a private stream (the caller's may be the legacy default stream, which cannot
// be captured) and instantiates them; returns 0 or a cudaError / lbm_status
int capture_steps(const lbm_step_desc *desc, void *a, void *b, bool chained, cudaGraphExec_t *exec) {
    cudaStream_t cs = nullptr;
    int e = (int)cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    if (e) return e;
    const int64_t launches_before = lbm::g_launch_count.load();
    e = (int)cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed);
    int rc = 0;
    cudaGraph_t graph = nullptr;
    if (!e) {
        void *x = a, *y = b;
        for (int k = 0; k < kGraphSteps && !rc; ++k) {
            lbm::StepExtras extras;
            extras.chained = chained && k > 0;      // programmatic edges between consecutive steps
            rc = lbm::step_general(desc, x, y, extras, cs);
            void *t = x; x = y; y = t;
        }
        e = (int)cudaStreamEndCapture(cs, &graph);
    }
    lbm::g_launch_count.store(launches_before);      // captured, not launched
    if (!e && !rc) e = (int)cudaGraphInstantiate(exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    cudaStreamDestroy(cs);
    return rc ? rc : e;
}

// replays as many whole graphs as fit into n; returns the number of steps done (0 = path not taken) or < 0
int64_t graph_steps(const lbm_step_desc *desc, void *a, void *b, int64_t n, cudaStream_t stream, int *status) {
    *status = LBM_OK;
    lbm::Dims dm;
    if (n < kGraphSteps || desc->n_ops != 1 || desc->labels || graph_max_nodes() <= 0 || lbm::lattice_dims(&desc->lat, dm)) return 0;
    if ((int64_t)dm.n0 * dm.n1 * dm.n2 > graph_max_nodes()) return 0;
    const lbm_halo &h = desc->halo;
    if (h.in_lo || h.in_hi || h.out_lo || h.out_hi) return 0;
    int device = -1;
    if (cudaGetDevice(&device)) return 0;
    std::lock_guard<std::mutex> lock(g_graph_mutex);
    GraphSlot *slot = nullptr, *victim = &g_graph_slots[0];
    for (GraphSlot &s : g_graph_slots) {
        if (s.exec && s.a == a && s.b == b && s.device == device && !memcmp(&s.desc, desc, sizeof *desc)) slot = &s;
        if (s.used < victim->used) victim = &s;
    }
    if (!slot) {
        cudaGraphExec_t exec = nullptr;
        int e = capture_steps(desc, a, b, pdl_enabled(), &exec);
        if (e && pdl_enabled()) {
            cudaGetLastError();                      // programmatic edges refused by this driver: plain capture
            e = capture_steps(desc, a, b, false, &exec);
        }
        if (e) {
            *status = lbm::cuda_fail_public(e);
            return -1;
        }
        if (victim->exec) cudaGraphExecDestroy(victim->exec);
        victim->desc = *desc;
        victim->a = a;
        victim->b = b;
        victim->device = device;
        victim->exec = exec;
        slot = victim;
    }
    slot->used = ++g_graph_clock;
    int64_t done = 0;
    while (n - done >= kGraphSteps) {
        const int e = (int)cudaGraphLaunch(slot->exec, stream);
        if (e) {
            *status = lbm::cuda_fail_public(e);
            return -1;
        }
        lbm::g_launch_count += kGraphSteps;
        done += kGraphSteps;
    }
    return done;
}

}  // namespace

// three rows of per-CTA partials + the staging row of the two-stage fold
int64_t lbm_links_scratch_doubles(int64_t n) {
    return n > 0 ? 3 * (int64_t)link_blocks(n) + (int64_t)(reduce_scratch_bytes() / sizeof(double)) : 0;
}

}  // extern "C"

template <class S, class R>
static int links_typed(const lbm_step_desc *desc, const Dims &dm, const lbm_links *l, const void *f_pre, void *f_post,
                       cudaStream_t st) {
    StepParams<R> p;
    fill_params<S, R>(desc, dm, f_pre, f_post, p);
    LinkArgs<R> a;
    a.kind = l->kind;
    a.n = (int)l->n;
    a.node = l->node;
    a.q = l->q;
    a.d = (const R *)l->d;
    a.bounced = (R *)l->bounced;
    a.partials = l->force ? l->force_scratch : nullptr;
    int e = launch_links<S, R>(p, a, desc->ops[desc->collision_index].kind, st);
    if (e || !l->force) return cuda_fail(e);
    // fold the per-CTA partials in a fixed order; internal axis order -> (x, y, z) of the caller
    const int blocks = link_blocks(l->n);
    for (int c = 0; c < 3; ++c) {
        if (c < S::D && l->n > 0) {
            e = launch_fold_sum(l->force_scratch + (size_t)S::axis_of(c) * blocks, blocks,
                                l->force_scratch + (size_t)3 * blocks, l->force + c, st);
        } else {
            e = (int)cudaMemsetAsync(l->force + c, 0, sizeof(double), st);
        }
        if (e) return cuda_fail(e);
    }
    return LBM_OK;
}

extern "C" {

int lbm_apply_links(const lbm_step_desc *desc, const lbm_links *links, const void *d_f_pre, void *d_f_post,
                    void *stream) {
    Dims dm;
    int rc = validate_desc(desc, dm);
    if (rc) return rc;
    if (!links || !d_f_pre || !d_f_post || links->n < 0) return LBM_ERR_BAD_ARGUMENT;
    if (links->kind < LBM_LINK_FULLWAY || links->kind > LBM_LINK_INTERPOLATED) return LBM_ERR_BAD_ARGUMENT;
    if (desc->streaming != LBM_POST_STREAMING) return LBM_ERR_UNSUPPORTED;
    if (links->n > 0 && (!links->node || !links->q || !links->bounced)) return LBM_ERR_BAD_ARGUMENT;
    if (links->n > 0 && links->kind == LBM_LINK_INTERPOLATED && !links->d) return LBM_ERR_BAD_ARGUMENT;
    if (links->force && links->n > 0 && !links->force_scratch) return LBM_ERR_BAD_ARGUMENT;
    if (links->n > (int64_t)1 << 30) return LBM_ERR_TOO_LARGE;
    const lbm_halo &h = desc->halo;
    if (h.in_lo || h.in_hi || h.out_lo || h.out_hi) return LBM_ERR_UNSUPPORTED;      // single-GPU entry point
    LBM_DISPATCH(desc->lat.stencil, desc->lat.dtype,
                 return (links_typed<S, R>(desc, dm, links, d_f_pre, d_f_post, (cudaStream_t)stream)));
    return LBM_ERR_BAD_ARGUMENT;
}

int lbm_step_links_n(const lbm_step_desc *desc, const lbm_links *links, int32_t n_boundaries, void *d_f_a,
                     void *d_f_b, int64_t n, void *stream) {
    if (n < 0 || n_boundaries < 0 || (n_boundaries > 0 && !links)) return LBM_ERR_BAD_ARGUMENT;
    void *a = d_f_a, *b = d_f_b;
    for (int64_t k = 0; k < n; ++k) {
        int rc = lbm_step(desc, a, b, stream);
        for (int32_t i = 0; i < n_boundaries && !rc; ++i) rc = lbm_apply_links(desc, &links[i], a, b, stream);
        if (rc) return rc;
        void *t = a; a = b; b = t;
    }
    return LBM_OK;
}

int lbm_step_n(const lbm_step_desc *desc, void *d_f_a, void *d_f_b, int64_t n, void *stream) {
    if (n < 0 || !desc) return LBM_ERR_BAD_ARGUMENT;
    void *a = d_f_a, *b = d_f_b;
    if (n >= kGraphSteps && graph_max_nodes() > 0) {
        lbm::Dims dm;
        int rc = lbm::validate_desc(desc, dm);
        if (rc) return rc;
        const int64_t done = graph_steps(desc, a, b, n, (cudaStream_t)stream, &rc);
        if (done < 0) return rc;
        n -= done;                                   // kGraphSteps is even: a still holds the newest populations
    }
    // consecutive steps are chained with programmatic dependent launch: step k+1's CTAs become resident while step
    // k drains and wait (griddepcontrol.wait) for its completion before they read anything
    const bool chain = pdl_enabled();
    for (int64_t k = 0; k < n; ++k) {
        lbm::StepExtras extras;
        extras.chained = chain && k > 0;
        const int rc = lbm::step_general(desc, a, b, extras, stream);
        if (rc) return rc;
        void *t = a; a = b; b = t;
    }
    return LBM_OK;
}

const char *lbm_step_variant_name(const lbm_step_desc *desc) {
    Dims dm;
    if (validate_desc(desc, dm)) return "invalid";
    return step_variant_name(desc, desc->n_ops > 1 || desc->labels != nullptr);
}

int lbm_pack_masks(const lbm_step_desc *desc, const uint8_t *d_ncm, const uint8_t *d_nsm, uint8_t *d_labels,
                   uint32_t *d_frozen, void *stream) {
    if (!desc || !d_ncm || !d_nsm || !d_labels || !d_frozen) return LBM_ERR_BAD_ARGUMENT;
    Dims dm;
    int rc = lattice_dims(&desc->lat, dm);
    if (rc) return rc;
    if (desc->n_ops < 1 || desc->n_ops > LBM_MAX_OPS) return LBM_ERR_BAD_ARGUMENT;
    switch (desc->lat.stencil) {
        case LBM_D2Q9: return cuda_fail(pack_typed<D2Q9>(desc, dm, d_ncm, d_nsm, d_labels, d_frozen, (cudaStream_t)stream));
        case LBM_D3Q19: return cuda_fail(pack_typed<D3Q19>(desc, dm, d_ncm, d_nsm, d_labels, d_frozen, (cudaStream_t)stream));
        default: return cuda_fail(pack_typed<D3Q27>(desc, dm, d_ncm, d_nsm, d_labels, d_frozen, (cudaStream_t)stream));
    }
}

int lbm_list_general_nodes(const lbm_lattice *lat, const uint8_t *d_labels, int32_t *d_list, int64_t capacity,
                           int64_t *d_count, void *stream) {
    Dims dm;
    int rc = lattice_dims(lat, dm);
    if (rc) return rc;
    if (!d_labels || !d_count || capacity < 0 || (capacity > 0 && !d_list)) return LBM_ERR_BAD_ARGUMENT;
    const int64_t N = (int64_t)dm.n0 * dm.n1 * dm.n2;
    int e = (int)cudaMemsetAsync(d_count, 0, sizeof(int64_t), (cudaStream_t)stream);
    if (e) return cuda_fail(e);
    int64_t b = (N + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    list_general_nodes_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(d_labels, N, d_list, (long long)capacity,
                                                                       (unsigned long long *)d_count);
    ++g_launch_count;
    return cuda_fail((int)cudaGetLastError());
}

int lbm_moments(const lbm_lattice *lat, const void *d_f, void *d_rho, void *d_u, void *stream) {
    Dims dm;
    int rc = lattice_dims(lat, dm);
    if (rc) return rc;
    if (!d_f || (!d_rho && !d_u)) return LBM_ERR_BAD_ARGUMENT;
    const int64_t N = (int64_t)dm.n0 * dm.n1 * dm.n2;
    LBM_DISPATCH(lat->stencil, lat->dtype,
                 return cuda_fail((launch_moments<S, R>((const R *)d_f, (R *)d_rho, (R *)d_u, N, (cudaStream_t)stream))));
    return LBM_ERR_BAD_ARGUMENT;
}

int lbm_equilibrium(const lbm_lattice *lat, const void *d_rho, const int64_t rho_stride[3], const void *d_u,
                    const int64_t u_stride[4], void *d_f_out, void *stream) {
    Dims dm;
    int rc = lattice_dims(lat, dm);
    if (rc) return rc;
    if (!d_rho || !d_u || !d_f_out || !rho_stride || !u_stride) return LBM_ERR_BAD_ARGUMENT;
    LBM_DISPATCH(lat->stencil, lat->dtype,
                 return cuda_fail((launch_equilibrium<S, R>((const R *)d_rho, rho_stride, (const R *)d_u, u_stride, dm.n0,
                                                            dm.n1, dm.n2, (R *)d_f_out, (cudaStream_t)stream))));
    return LBM_ERR_BAD_ARGUMENT;
}

int lbm_initialize_fneq(const lbm_lattice *lat, const void *d_rho, const void *d_u, double tau, double eye_cs2,
                        void *d_f_out, void *stream) {
    Dims dm;
    int rc = lattice_dims(lat, dm);
    if (rc) return rc;
    if (!d_rho || !d_u || !d_f_out) return LBM_ERR_BAD_ARGUMENT;
    LBM_DISPATCH(lat->stencil, lat->dtype,
                 // the reference divides by the stencil's cs^2 in the context dtype (torch_stencil.cs ** 2)
                 return cuda_fail((launch_init_fneq<S, R>((const R *)d_rho, (const R *)d_u,
                                                          (double)((R)tau / ((R)kCs * (R)kCs)), eye_cs2, dm.n0, dm.n1,
                                                          dm.n2, (R *)d_f_out, (cudaStream_t)stream))));
    return LBM_ERR_BAD_ARGUMENT;
}

size_t lbm_reduce_scratch_bytes(const lbm_lattice *) { return reduce_scratch_bytes(); }

int lbm_reduce(const lbm_lattice *lat, int what, const void *d_in, const uint8_t *d_mask, void *d_scratch,
               double *d_out, void *stream) {
    Dims dm;
    int rc = lattice_dims(lat, dm);
    if (rc) return rc;
    if (!d_in || !d_scratch || !d_out) return LBM_ERR_BAD_ARGUMENT;
    LBM_DISPATCH(lat->stencil, lat->dtype,
                 return cuda_fail((launch_reduce<S, R>(what, (const R *)d_in, d_mask, dm.n0, dm.n1, dm.n2,
                                                       (double *)d_scratch, d_out, (cudaStream_t)stream))));
    return LBM_ERR_BAD_ARGUMENT;
}

// device work space of lbm_run_host, kept between calls (allocating and freeing 2 x the lattice costs more than
// the steps of a short run)
namespace {
struct HostRunWorkspace {
    int device = -1;
    size_t bytes = 0, scratch_bytes = 0;
    void *a = nullptr, *b = nullptr, *scratch = nullptr;
    double *d_e = nullptr;
    cudaStream_t stream = nullptr;
    void release() {
        if (device >= 0) {
            int cur = -1;
            cudaGetDevice(&cur);
            cudaSetDevice(device);
            if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
            cudaFree(a); cudaFree(b); cudaFree(scratch); cudaFree(d_e);
            if (cur >= 0) cudaSetDevice(cur);
        }
        *this = HostRunWorkspace();
    }
};
HostRunWorkspace g_host_ws;
std::mutex g_host_ws_mutex;
}  // namespace

int lbm_run_host_release(void) {
    std::lock_guard<std::mutex> lock(g_host_ws_mutex);
    g_host_ws.release();
    return LBM_OK;
}

int lbm_run_host(const lbm_step_desc *desc, const void *h_f, void *h_f_out, int64_t nsteps, double *h_energy) {
    Dims dm;
    int rc = validate_desc(desc, dm);
    if (rc) return rc;
    if (!h_f || !h_f_out || nsteps < 0) return LBM_ERR_BAD_ARGUMENT;
    const lbm_halo &h = desc->halo;
    if (h.in_lo || h.in_hi || h.out_lo || h.out_hi) return LBM_ERR_BAD_ARGUMENT;
    const size_t bytes = (size_t)dm.q * dm.n0 * dm.n1 * dm.n2 * (desc->lat.dtype == LBM_F32 ? 4 : 8);
    const size_t fused_bytes = lbm_step_moments_scratch_bytes(desc);
    const size_t scratch_bytes = fused_bytes > reduce_scratch_bytes() ? fused_bytes : reduce_scratch_bytes();
    int device = -1;
    int e = (int)cudaGetDevice(&device);
    if (e) return cuda_fail(e);
    std::lock_guard<std::mutex> lock(g_host_ws_mutex);
    HostRunWorkspace &ws = g_host_ws;
    if (ws.device != device || ws.bytes != bytes || ws.scratch_bytes < scratch_bytes) {
        ws.release();
        ws.device = device;
        if ((e = (int)cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking)) ||
            (e = (int)cudaMalloc(&ws.a, bytes)) || (e = (int)cudaMalloc(&ws.b, bytes)) ||
            (e = (int)cudaMalloc(&ws.scratch, scratch_bytes)) || (e = (int)cudaMalloc(&ws.d_e, 2 * sizeof(double)))) {
            ws.release();
            return cuda_fail(e);
        }
        ws.bytes = bytes;
        ws.scratch_bytes = scratch_bytes;
    }
    cudaStream_t st = ws.stream;
    void *a = ws.a, *b = ws.b;
    auto fail = [&](int code) {
        cudaStreamSynchronize(st);
        return code;
    };
    if ((e = (int)cudaMemcpyAsync(a, h_f, bytes, cudaMemcpyHostToDevice, st))) return fail(cuda_fail(e));
    // Every step's kinetic energy is reduced inside the step kernels (lbm_step_moments).  Steps that describe the
    // state they WRITE deliver the value of step k with step k; steps that describe the state they READ
    // (POST_STREAMING) deliver it with step k + 1, and the last one comes from a stand-alone reduction.
    const int state = h_energy ? lbm_step_moments_state(desc) : LBM_MOMENTS_UNAVAILABLE;
    for (int64_t k = 0; k < nsteps; ++k) {
        const bool fused = state == LBM_MOMENTS_OF_OUTPUT || (state == LBM_MOMENTS_OF_INPUT && k > 0);
        rc = fused ? step_moments_general(desc, a, b, nullptr, ws.scratch, ws.scratch_bytes, ws.d_e, k > 0, st)
                   : lbm_step(desc, a, b, st);
        if (rc) return fail(rc);
        void *t = a; a = b; b = t;
        if (!h_energy) continue;
        int64_t slot = k;                       // which step's energy ws.d_e holds now
        if (state == LBM_MOMENTS_OF_INPUT) {
            if (k == 0) continue;
            slot = k - 1;
        } else if (state == LBM_MOMENTS_UNAVAILABLE) {
            rc = lbm_reduce(&desc->lat, LBM_SUM_HALF_U2, a, nullptr, ws.scratch, ws.d_e, st);
            if (rc) return fail(rc);
        }
        // result of that step read back to the host (reporter with interval 1)
        if ((e = (int)cudaMemcpyAsync(h_energy + slot, ws.d_e, sizeof(double), cudaMemcpyDeviceToHost, st)))
            return fail(cuda_fail(e));
    }
    if (h_energy && state == LBM_MOMENTS_OF_INPUT && nsteps > 0) {
        rc = lbm_reduce(&desc->lat, LBM_SUM_HALF_U2, a, nullptr, ws.scratch, ws.d_e, st);
        if (rc) return fail(rc);
        if ((e = (int)cudaMemcpyAsync(h_energy + nsteps - 1, ws.d_e, sizeof(double), cudaMemcpyDeviceToHost, st)))
            return fail(cuda_fail(e));
    }
    if ((e = (int)cudaMemcpyAsync(h_f_out, a, bytes, cudaMemcpyDeviceToHost, st))) return fail(cuda_fail(e));
    e = (int)cudaStreamSynchronize(st);
    return e ? cuda_fail(e) : LBM_OK;
}

}  // extern "C"
