// lbm_step_inst.cu -- instantiates the step kernels for ONE (stencil, dtype) pair.
// Compile with -DLBM_INST_STENCIL=D3Q19 -DLBM_INST_REAL=float (see build.py).
#include "lbm_launch.cuh"

#ifndef LBM_INST_STENCIL
#error "define LBM_INST_STENCIL (D2Q9 | D3Q19 | D3Q27)"
#endif
#ifndef LBM_INST_REAL
#error "define LBM_INST_REAL (float | double)"
#endif

namespace lbm {

namespace {

// nodes per thread: small stencils need more loads in flight per thread (measured on B200:
// D2Q9 fp32 at 4096x1024 runs at 0.72 of the HBM roofline with 1 node per thread)
template <class S, class R>
constexpr int nodes_per_thread() {
    return S::Q == 9 ? (sizeof(R) == 4 ? 4 : 2) : 1;
}

template <class S, class R, int COLL, bool PULL, bool PUSH, bool MASKED>
int launch_scalar(const StepParams<R> &p, cudaStream_t stream) {
    constexpr int NPT = nodes_per_thread<S, R>();
    // threadIdx.x runs along the contiguous axis; fill the block up to 256 threads with rows.
    int tz = 32;
    while (tz * NPT < p.n2 && tz < 256) tz <<= 1;
    int ty = 256 / tz;
    while (ty > 1 && ty / 2 >= p.n1) ty >>= 1;
    dim3 block(tz, ty, 1);
    dim3 grid((p.n2 + tz * NPT - 1) / (tz * NPT), (p.n1 + ty - 1) / ty, p.n0);
    if constexpr (NPT == 1) step_scalar_kernel<S, R, COLL, PULL, PUSH, MASKED><<<grid, block, 0, stream>>>(p);
    else step_multi_kernel<S, R, COLL, PULL, PUSH, MASKED, NPT><<<grid, block, 0, stream>>>(p);
    ++g_launch_count;
    if (MASKED && p.n_general > 0) {
        general_nodes_kernel<S, R, COLL, PULL, PUSH><<<(p.n_general + 127) / 128, 128, 0, stream>>>(p);
        ++g_launch_count;
    }
    return (int)cudaGetLastError();
}

template <class S, class R, int COLL, bool MASKED>
int by_streaming(const StepParams<R> &p, int streaming, cudaStream_t stream) {
    switch (streaming) {
        case LBM_NO_STREAMING: return launch_scalar<S, R, COLL, false, false, MASKED>(p, stream);
        case LBM_POST_STREAMING: return launch_scalar<S, R, COLL, false, true, MASKED>(p, stream);
        case LBM_PRE_STREAMING: return launch_scalar<S, R, COLL, true, false, MASKED>(p, stream);
        case LBM_DOUBLE_STREAMING: return launch_scalar<S, R, COLL, true, true, MASKED>(p, stream);
    }
    return LBM_ERR_BAD_ARGUMENT;
}

template <class S, class R, int COLL>
int by_mask(const StepParams<R> &p, int streaming, bool masked, cudaStream_t stream) {
    return masked ? by_streaming<S, R, COLL, true>(p, streaming, stream)
                  : by_streaming<S, R, COLL, false>(p, streaming, stream);
}

}  // namespace

template <class S, class R>
int launch_step(const StepParams<R> &p, int coll, int streaming, bool masked, int variant, cudaStream_t stream) {
    (void)variant;
    switch (coll) {
        case LBM_OP_NO_COLLISION: return by_mask<S, R, LBM_OP_NO_COLLISION>(p, streaming, masked, stream);
        case LBM_OP_BGK: return by_mask<S, R, LBM_OP_BGK>(p, streaming, masked, stream);
        case LBM_OP_TRT: return by_mask<S, R, LBM_OP_TRT>(p, streaming, masked, stream);
        case LBM_OP_KBC:
            // KBC exists for D2Q9 and D3Q27 only (kbc_collision.py:101,116)
            if constexpr (S::ID == LBM_D3Q19) return LBM_ERR_UNSUPPORTED;
            else return by_mask<S, R, LBM_OP_KBC>(p, streaming, masked, stream);
    }
    return LBM_ERR_BAD_ARGUMENT;
}

template <class S, class R>
const char *step_variant_name(const StepParams<R> &, int, int, bool masked, int) {
    return masked ? "scalar_masked+general_nodes" : "scalar";
}

template int launch_step<LBM_INST_STENCIL, LBM_INST_REAL>(const StepParams<LBM_INST_REAL> &, int, int, bool, int,
                                                          cudaStream_t);
template const char *step_variant_name<LBM_INST_STENCIL, LBM_INST_REAL>(const StepParams<LBM_INST_REAL> &, int, int,
                                                                        bool, int);

}  // namespace lbm
