// lbm_step_inst.cu -- instantiates the step kernels for ONE (stencil, dtype, collision) triple.
// Compile with -DLBM_INST_STENCIL=D3Q19 -DLBM_INST_REAL=float -DLBM_INST_COLL=1 (see build.py).
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

// One node per thread.  (A 2- and 4-nodes-per-thread variant for D2Q9 was measured on B200 -- C4, 4096x1024 fp32,
// PRE: 65.8 GLUPS with one node per thread, 64.6 with two, 55.5 with four -- and removed: occupancy pays, more
// loads in flight per thread do not.)
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

template <class S, class R, int COLL, bool PULL, bool PUSH, int MODE>
int launch_scalar(const StepParams<R> &p, cudaStream_t stream) {
    constexpr bool MASKED = MODE != kUnmasked;
    dim3 grid, block;
    bulk_geometry(p.n0, p.n1, p.n2, grid, block);
    void (*bulk)(const StepParams<R>) = step_scalar_kernel<S, R, COLL, PULL, PUSH, MODE>;
    if (p.energy_partials) {       // lbm_step_energy: step + kinetic energy of the output in one kernel
        if constexpr (!MASKED && !PUSH) {
            if (p.sync.on) return LBM_ERR_UNSUPPORTED;
            step_energy_kernel<S, R, COLL, PULL><<<grid, block, 0, stream>>>(p);
            ++g_launch_count;
            return (int)cudaGetLastError();
        } else {
            return LBM_ERR_UNSUPPORTED;
        }
    }
    if constexpr (!MASKED) {
        if (p.sync.on) {       // multi-GPU slab with in-kernel lock step (lbm_slab_step_n)
            StepParams<R> ps = p;
            ps.sync.ctas_per_side = grid.x * grid.y * ((PULL && PUSH) ? 2 : 1);
            step_sync_kernel<S, R, COLL, PULL, PUSH><<<grid, block, 0, stream>>>(ps);
            ++g_launch_count;
            return (int)cudaGetLastError();
        }
    }
    if (p.sync.on) return LBM_ERR_UNSUPPORTED;
    if (MASKED && p.n_general > 0) {
        const dim3 sgrid((p.n_general + 127) / 128), sblock(128);
        if constexpr (MODE == kMaskedOverwrite) {
            // bulk kernel over every node first; the sparse kernel is released by the bulk kernel's first
            // instruction, gathers and collides next to it and stores once the bulk grid has completed
            bulk<<<grid, block, 0, stream>>>(p);
            ++g_launch_count;
            const int e = (int)cudaGetLastError();
            if (e) return e;
            return launch_dependent<R>(general_nodes_kernel<S, R, COLL, PULL, PUSH, true>, sgrid, sblock, p, stream);
        } else {
            // The sparse kernel is a chain of dependent loads on a handful of CTAs (~10 us at 15 k nodes); it
            // and the bulk kernel write disjoint slots, so the bulk kernel is launched with programmatic
            // dependent launch right behind it and overlaps it completely (the sparse kernel releases its
            // dependents in its first instruction; the bulk kernel waits for it in its LAST instruction).
            general_nodes_kernel<S, R, COLL, PULL, PUSH, false><<<sgrid, sblock, 0, stream>>>(p);
            ++g_launch_count;
            const int e = (int)cudaGetLastError();
            if (e) return e;
            return launch_dependent<R>(bulk, grid, block, p, stream);
        }
    }
    bulk<<<grid, block, 0, stream>>>(p);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

template <class S, class R, int COLL, int MODE>
int by_streaming(const StepParams<R> &p, int streaming, cudaStream_t stream) {
    switch (streaming) {
        case LBM_NO_STREAMING: return launch_scalar<S, R, COLL, false, false, MODE>(p, stream);
        case LBM_POST_STREAMING: return launch_scalar<S, R, COLL, false, true, MODE>(p, stream);
        case LBM_PRE_STREAMING: return launch_scalar<S, R, COLL, true, false, MODE>(p, stream);
        case LBM_DOUBLE_STREAMING: return launch_scalar<S, R, COLL, true, true, MODE>(p, stream);
    }
    return LBM_ERR_BAD_ARGUMENT;
}

template <class S, class R, int COLL>
int by_mask(const StepParams<R> &p, int streaming, bool masked, int variant, cudaStream_t stream) {
    if (!masked) return by_streaming<S, R, COLL, kUnmasked>(p, streaming, stream);
    switch (variant) {
        case kMaskedLabelFirst: return by_streaming<S, R, COLL, kMaskedLabelFirst>(p, streaming, stream);
        case kMaskedOverwrite: return by_streaming<S, R, COLL, kMaskedOverwrite>(p, streaming, stream);
        default: return by_streaming<S, R, COLL, kMaskedSpeculative>(p, streaming, stream);
    }
}

}  // namespace

template <class S, class R, int COLL>
int launch_step_coll(const StepParams<R> &p, int streaming, bool masked, int variant, cudaStream_t stream) {
    return by_mask<S, R, COLL>(p, streaming, masked, variant, stream);
}

template <class S, class R, int COLL>
int launch_links_coll(const StepParams<R> &p, const LinkArgs<R> &a, cudaStream_t stream) {
    if (a.n <= 0) return 0;
    const int blocks = link_blocks(a.n);
    link_gather_kernel<S, R, COLL><<<blocks, kLinkThreads, 0, stream>>>(p, a);
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
                                                                              bool, int, cudaStream_t);

}  // namespace lbm
