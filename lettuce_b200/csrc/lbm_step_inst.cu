// lbm_step_inst.cu -- instantiates the step kernels for ONE (stencil, dtype, collision) triple.
// Compile with -DLBM_INST_STENCIL=D3Q19 -DLBM_INST_REAL=float -DLBM_INST_COLL=1 (see build.py).
#include <cstdlib>

#include "lbm_launch.cuh"

#ifndef LBM_INST_STENCIL
#error "define LBM_INST_STENCIL (D2Q9 | D3Q19 | D3Q27)"
#endif
#ifndef LBM_INST_REAL
#error "define LBM_INST_REAL (float | double)"
#endif
#ifndef LBM_INST_COLL
#error "define LBM_INST_COLL (lbm_op_kind collision value 0..6)"
#endif

namespace lbm {

namespace {

// launches `kernel` behind the previous launch on `stream` with programmatic stream serialization: it may start
// as soon as the previous grid has executed griddepcontrol.launch_dependents (or exited) in every CTA
template <class R>
int launch_dependent(void (*kernel)(const StepParams<R>), dim3 grid, dim3 block, const StepParams<R> &p,
                     cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    const int e = (int)cudaLaunchKernelEx(&cfg, kernel, p);
    ++g_launch_count;
    return e;
}

// One time step: the bulk kernel over every node and, in masked runs, the sparse kernel behind it (released by the
// bulk kernel once the previous step is known to be complete; it gathers and collides next to the bulk kernel and
// stores after the bulk grid has completed).  One node per thread measured fastest for D2Q9 in round 1 when the
// two nodes were computed with scalar arithmetic (4096x1024 fp32: 65.8 GLUPS vs 64.6); LANES = 2 here means the
// packed float2 arithmetic of lbm_vec.cuh, which halves the fp32 issue slots.
template <class S, class R, int COLL, bool PULL, bool PUSH, int LANES>
int launch_lanes(const StepParams<R> &p_in, const LaunchOptions &opt, cudaStream_t stream, bool cut_planes_only = false) {
    StepParams<R> p = p_in;
    dim3 grid, block;
    bulk_geometry(p.n0, p.n1, p.n2, LANES, bulk_threads<S, R, COLL, LANES>(), grid, block);
    // (slab lock step: the first 2 W grid layers are the cut planes, sync_plane)
    if (cut_planes_only) grid.z = (PULL && PUSH) ? 4 : 2;
    const bool sparse = p.labels != nullptr && p.n_general > 0;
    const int bulk_ctas = (int)(grid.x * grid.y * grid.z);
    if (p.reduce_mode != kReduceNone) {
        if (PULL && PUSH) return LBM_ERR_UNSUPPORTED;      // the output node is assembled from several source nodes
        p.reduce_sparse_offset = bulk_ctas;
        p.reduce_slots = bulk_ctas + (sparse ? sparse_blocks(p.n_general) : 0);
    }
    if (p.sync.on) {
        p.sync.ctas_per_side = grid.x * grid.y * ((PULL && PUSH) ? 2 : 1);
        p.sync.publish = sparse ? 0 : 1;
    }
    // the plain kernel unless the step carries the slab lock step (kStepSync) or fused reductions / a masked slab's
    // single-writer cut planes (kStepFull)
    void (*bulk)(const StepParams<R>) =
        (p.reduce_mode != kReduceNone || (p.sync.on && p.labels != nullptr))
            ? step_kernel<S, R, COLL, PULL, PUSH, LANES, kStepFull>
            : (p.sync.on ? step_kernel<S, R, COLL, PULL, PUSH, LANES, kStepSync>
                         : step_kernel<S, R, COLL, PULL, PUSH, LANES, kStepPlain>);
    int e;
    if (opt.chained) {
        e = launch_dependent<R>(bulk, grid, block, p, stream);
    } else {
        bulk<<<grid, block, 0, stream>>>(p);
        ++g_launch_count;
        e = (int)cudaGetLastError();
    }
    if (e || !sparse) return e;
    return launch_dependent<R>(p.nested_outlets ? general_nodes_kernel<S, R, COLL, PULL, PUSH, true>
                                                : general_nodes_kernel<S, R, COLL, PULL, PUSH, false>,
                               dim3(sparse_blocks(p.n_general)), dim3(kSparseThreads), p, stream);
}

// the TMA-staged persistent kernel (lbm_tma.cuh): one CTA per SM, tiles of kTmaTileNodes nodes dealt round-robin
template <class S, int COLL, bool PULL, bool REDUCE = false>
int launch_tma(const StepParams<float> &p, const LaunchOptions &opt, cudaStream_t stream) {
    auto kernel = step_tma_kernel<S, COLL, PULL, REDUCE>;
    // CTAs per SM and stages per CTA: as many stages as fit next to each other in the SM's 227 KB (1 KB per CTA is
    // reserved by the system); LBM_B200_TMA_CTAS / LBM_B200_TMA_STAGES override (measurements)
    static const int env_ctas = [] { const char *e = getenv("LBM_B200_TMA_CTAS"); return e ? atoi(e) : 0; }();
    static const int env_stages = [] { const char *e = getenv("LBM_B200_TMA_STAGES"); return e ? atoi(e) : 0; }();
    // measured (profiles/r2_tma_sweep.md): as many stages as fit, up to five; D2Q9's short tiles want two CTAs per SM
    int ctas = S::Q == 9 ? 2 : 1;
    if (env_ctas >= 1 && env_ctas <= tma_ctas_per_sm<S>()) ctas = env_ctas;
    const size_t per_cta = (227 * 1024) / ctas - (ctas > 1 ? 1024 : 0) - 1024;    // static part: barriers, tile ring
    int max_stages = (int)(per_cta / (tma_stage_floats<S>() * sizeof(float)));
    if (max_stages > kTmaMaxStages) max_stages = kTmaMaxStages;
    int stages = max_stages < 5 ? max_stages : 5;
    if (env_stages >= 2 && env_stages <= max_stages) stages = env_stages;
    if (stages < 2) return LBM_ERR_UNSUPPORTED;
    const size_t smem = tma_smem_bytes<S>(stages);
    static size_t configured = 0;            // per instantiation: largest window asked for so far
    if (smem > configured) {
        const int e = (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return e;
        configured = smem;
    }
    TmaParams t = {};
    t.in = p.in; t.out = p.out; t.N = p.N;
    t.counters = opt.tma_counters;
    t.n0 = p.n0; t.n1 = p.n1; t.n2 = p.n2;
    t.tz = tma_row_extent(p.n2);
    t.tz_log2 = 0;
    while ((1 << t.tz_log2) < t.tz) ++t.tz_log2;
    t.rows = kTmaTileNodes / t.tz;
    t.zchunks = p.n2 / t.tz;
    // a slab's cut planes (x = 0 and x = n0 - 1) are left to the lock-step kernel
    t.row_begin = opt.tma_interior ? p.n1 : 0;
    t.n_rows = p.n0 * p.n1 - t.row_begin;
    t.n_tiles = ((t.n_rows - t.row_begin + t.rows - 1) / t.rows) * t.zchunks;
    t.skip_wait = opt.tma_interior ? 1 : 0;
    t.stages = stages;
    t.boxable = opt.tma_boxable;
    t.reverse = p.reverse_sweep;
    t.ca = p.ca; t.cb = p.cb; t.force = p.force;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(t.n_tiles < opt.sm_count * ctas ? t.n_tiles : opt.sm_count * ctas);
    if constexpr (REDUCE) {
        t.partials = p.energy_partials;
        t.reduce_slots = (int)cfg.gridDim.x;
        if (opt.slots_used) *opt.slots_used = t.reduce_slots;
    }
    cfg.blockDim = dim3(kTmaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = (opt.chained || opt.tma_interior) ? 1 : 0;
    const int e = (int)cudaLaunchKernelEx(&cfg, kernel, *opt.tma, t);
    ++g_launch_count;
    return e;
}

// the pushing step on the staged machinery (step_tma_push_kernel)
template <class S, int COLL>
int launch_tma_push(const StepParams<float> &p, const LaunchOptions &opt, cudaStream_t stream) {
    auto kernel = step_tma_push_kernel<S, COLL>;
    static const int env_ctas = [] { const char *e = getenv("LBM_B200_TMA_CTAS"); return e ? atoi(e) : 0; }();
    static const int env_stages = [] { const char *e = getenv("LBM_B200_TMA_STAGES"); return e ? atoi(e) : 0; }();
    int ctas = S::Q == 9 ? 2 : 1;
    if (env_ctas >= 1 && env_ctas <= tma_ctas_per_sm<S>()) ctas = env_ctas;
    const size_t per_cta = (227 * 1024) / ctas - (ctas > 1 ? 1024 : 0) - 1024;
    int max_stages = (int)(per_cta / (tma_push_stage_floats<S>() * sizeof(float)));
    if (max_stages > kTmaMaxStages) max_stages = kTmaMaxStages;
    int stages = max_stages < 5 ? max_stages : 5;
    if (env_stages >= 2 && env_stages <= max_stages) stages = env_stages;
    if (stages < 2) return LBM_ERR_UNSUPPORTED;
    const size_t smem = tma_push_smem_bytes<S>(stages);
    static size_t configured = 0;
    if (smem > configured) {
        const int e = (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return e;
        configured = smem;
    }
    TmaParams t = {};
    t.in = p.in; t.out = p.out; t.N = p.N;
    t.counters = opt.tma_counters;
    t.n0 = p.n0; t.n1 = p.n1; t.n2 = p.n2;
    t.tz = tma_row_extent(p.n2);
    t.tz_log2 = 0;
    while ((1 << t.tz_log2) < t.tz) ++t.tz_log2;
    t.rows = kTmaTileNodes / t.tz;
    t.zchunks = p.n2 / t.tz;
    t.row_begin = 0;
    t.n_rows = p.n0 * p.n1;
    t.n_tiles = ((t.n_rows + t.rows - 1) / t.rows) * t.zchunks;
    t.stages = stages;
    t.boxable = opt.tma_boxable;
    t.reverse = p.reverse_sweep;
    t.ca = p.ca; t.cb = p.cb; t.force = p.force;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(t.n_tiles < opt.sm_count * ctas ? t.n_tiles : opt.sm_count * ctas);
    cfg.blockDim = dim3(kTmaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = opt.chained ? 1 : 0;
    const int e = (int)cudaLaunchKernelEx(&cfg, kernel, *opt.tma, t);
    ++g_launch_count;
    return e;
}

template <class S, class R, int COLL, bool PULL, bool PUSH>
int by_lanes(const StepParams<R> &p, const LaunchOptions &opt, cudaStream_t stream) {
    if constexpr (sizeof(R) == 4 && !PULL && PUSH) {
        if (opt.tma && !p.sync.on && p.reduce_mode == kReduceNone) {
            int e = launch_tma_push<S, COLL>(p, opt, stream);
            if (e || !(p.labels != nullptr && p.n_general > 0)) return e;
            return launch_dependent<R>(p.nested_outlets ? general_nodes_kernel<S, R, COLL, PULL, PUSH, true>
                                                        : general_nodes_kernel<S, R, COLL, PULL, PUSH, false>,
                                       dim3(sparse_blocks(p.n_general)), dim3(kSparseThreads), p, stream);
        }
    }
    if constexpr (sizeof(R) == 4 && !PUSH) {
        if (opt.tma && opt.tma_interior && p.sync.on && p.reduce_mode == kReduceNone && p.labels == nullptr) {
            // Multi-GPU slab: the two cut planes by the lock-step LDG kernel (it waits for the neighbours' progress
            // counters, reads their planes over NVLink and publishes this rank's counter), the interior planes by the
            // staged kernel, launched programmatically behind it: it becomes resident as soon as every cut-plane CTA
            // has started (i.e. the previous step is complete) and runs next to them.
            LaunchOptions o = opt;
            o.chained = false;
            int e;
            if (o.lanes == 2 && p.n2 % 2 == 0) e = launch_lanes<S, R, COLL, PULL, PUSH, 2>(p, o, stream, true);
            else e = launch_lanes<S, R, COLL, PULL, PUSH, 1>(p, o, stream, true);
            if (e) return e;
            return launch_tma<S, COLL, PULL>(p, opt, stream);
        }
        if constexpr (PULL) {
            // fused reductions of the state a pulling step writes (unmasked lattices): the staged kernel's consumers
            if (opt.tma && !p.sync.on && p.reduce_mode == kReduceOutput && p.labels == nullptr)
                return launch_tma<S, COLL, PULL, true>(p, opt, stream);
        }
        if (opt.tma && !p.sync.on && p.reduce_mode == kReduceNone) {
            int e = launch_tma<S, COLL, PULL>(p, opt, stream);
            if (e || !(p.labels != nullptr && p.n_general > 0)) return e;
            // masked runs: the sparse kernel behind it, as after the LDG bulk kernel (launch_lanes)
            return launch_dependent<R>(p.nested_outlets ? general_nodes_kernel<S, R, COLL, PULL, PUSH, true>
                                                        : general_nodes_kernel<S, R, COLL, PULL, PUSH, false>,
                                       dim3(sparse_blocks(p.n_general)), dim3(kSparseThreads), p, stream);
        }
    }
    if constexpr (sizeof(R) == 4 && (PULL != PUSH)) {
        if (opt.lanes == 2 && p.n2 % 2 == 0) return launch_lanes<S, R, COLL, PULL, PUSH, 2>(p, opt, stream);
    }
    return launch_lanes<S, R, COLL, PULL, PUSH, 1>(p, opt, stream);
}

}  // namespace

template <class S, class R, int COLL>
int launch_step_coll(const StepParams<R> &p, int streaming, const LaunchOptions &opt, cudaStream_t stream) {
    switch (streaming) {
        case LBM_NO_STREAMING: return by_lanes<S, R, COLL, false, false>(p, opt, stream);
        case LBM_POST_STREAMING: return by_lanes<S, R, COLL, false, true>(p, opt, stream);
        case LBM_PRE_STREAMING: return by_lanes<S, R, COLL, true, false>(p, opt, stream);
        case LBM_DOUBLE_STREAMING: return by_lanes<S, R, COLL, true, true>(p, opt, stream);
    }
    return LBM_ERR_BAD_ARGUMENT;
}

template <class S, class R, int COLL>
int launch_links_coll(const StepParams<R> &p, const LinkArgs<R> &a, cudaStream_t stream) {
    if (a.n <= 0) return 0;
    const int blocks = link_blocks(a.n);
    if (p.nested_outlets) link_gather_kernel<S, R, COLL, true><<<blocks, kLinkThreads, 0, stream>>>(p, a);
    else link_gather_kernel<S, R, COLL, false><<<blocks, kLinkThreads, 0, stream>>>(p, a);
    ++g_launch_count;
    int e = (int)cudaGetLastError();
    if (e) return e;
    link_scatter_kernel<S, R><<<blocks, kLinkThreads, 0, stream>>>(p.out, p.N, a);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

template int launch_links_coll<LBM_INST_STENCIL, LBM_INST_REAL, LBM_INST_COLL>(const StepParams<LBM_INST_REAL> &,
                                                                               const LinkArgs<LBM_INST_REAL> &,
                                                                               cudaStream_t);

template int launch_step_coll<LBM_INST_STENCIL, LBM_INST_REAL, LBM_INST_COLL>(const StepParams<LBM_INST_REAL> &, int,
                                                                              const LaunchOptions &, cudaStream_t);

}  // namespace lbm
