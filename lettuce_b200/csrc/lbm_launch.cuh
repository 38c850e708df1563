// lbm_launch.cuh -- declaration of the per-(stencil, dtype, collision) step launchers.  Each
// triple is compiled in its own translation unit (lbm_step_inst.cu with -DLBM_INST_STENCIL /
// -DLBM_INST_REAL / -DLBM_INST_COLL) so the 40 units build in parallel.
#pragma once
#include <atomic>

#include "lbm_step.cuh"
#include "lbm_tma.cuh"

namespace lbm {

extern std::atomic<int64_t> g_launch_count;  // kernels launched by this library (lbm_launch_count)

constexpr int kSparseThreads = 128;          // general_nodes_kernel

// launch geometry of the bulk kernel: threadIdx.x along the contiguous axis (`lanes` nodes per thread), blocks of up
// to `threads` threads filled with rows, one grid layer per x-plane
inline void bulk_geometry(int n0, int n1, int n2, int lanes, int threads, dim3 &grid, dim3 &block) {
    const int zthreads = (n2 + lanes - 1) / lanes;
    int tz = 32;
    while (tz < zthreads && tz < threads) tz <<= 1;
    int ty = threads / tz;
    while (ty > 1 && ty / 2 >= n1) ty >>= 1;
    block = dim3(tz, ty, 1);
    grid = dim3((zthreads + tz - 1) / tz, (n1 + ty - 1) / ty, n0);
}

// threads per CTA of the bulk kernel (bulk_threads<> in lbm_step.cuh: 256 for one node per thread, 128 for two)
inline int bulk_threads_for(int lanes) { return lanes == 2 ? 128 : 256; }

inline int sparse_blocks(int64_t n_general) { return (int)((n_general + kSparseThreads - 1) / kSparseThreads); }

// number of (sum, max) partial pairs a step with fused reductions writes: one per bulk CTA + one per sparse CTA
inline int reduce_slots_for(int n0, int n1, int n2, int lanes, int64_t n_general) {
    dim3 grid, block;
    bulk_geometry(n0, n1, n2, lanes, bulk_threads_for(lanes), grid, block);
    return (int)(grid.x * grid.y * grid.z) + sparse_blocks(n_general);
}

struct LaunchOptions {
    const TmaMaps *tma = nullptr;   // tensor maps of (f_in, f_out): run the TMA-staged kernel (lbm_tma.cuh) when the
                                    // step carries neither the slab lock step nor fused reductions
    int tma_boxable = 0;            // the box / halo maps are valid (tma_rows_boxable)
    bool tma_interior = false;      // multi-GPU slab: cut planes by the LDG lock-step kernel, interior planes staged
    unsigned *tma_counters = nullptr;   // two zeroed words in device memory: the kernel's tile counter (lbm_tma.cuh)
    int *slots_used = nullptr;          // out: partial pairs a step with fused reductions wrote, when that is not
                                        // reduce_slots_for (the staged kernel: one pair per persistent CTA)
    int sm_count = 148;
    int lanes;        // 1, or 2 (fp32, even n2, PRE / POST streaming): nodes per thread of the bulk kernel
    bool chained;     // the previous launch on the stream is a step kernel of this library: launch behind it with
                      // programmatic stream serialization (it released its dependents, see step_kernel)
};

// true when the TMA-staged kernel can run this lattice: fp32, contiguous extent a multiple of 64 (the tile rows are
// power-of-two boxes that divide it; 16-byte global strides), and enough nodes to fill the persistent grid
inline bool tma_available(int dtype, int64_t nodes, int n2) {
    return dtype == LBM_F32 && n2 % 64 == 0 && nodes >= (int64_t)kTmaTileNodes * 148;
}
inline int tma_row_extent(int n2) {                 // largest power of two <= 256 that divides n2
    int tz = 64;
    while (tz < 256 && n2 % (2 * tz) == 0) tz *= 2;
    return tz;
}

inline int tma_tile_rows(int n2) { return kTmaTileNodes / tma_row_extent(n2); }
// the rows of a full tile are consecutive along ONE axis (y in 3-D, x in 2-D) and never straddle it
inline bool tma_rows_boxable(int n0, int n1, int n2) {
    const int rows = tma_tile_rows(n2);
    return n1 > 1 ? n1 % rows == 0 : n0 % rows == 0;
}

// true when the two-nodes-per-thread kernel exists for this request (instantiated for float, PRE and POST streaming)
inline bool lanes2_available(int dtype, int streaming, int n2) {
    return dtype == LBM_F32 && (streaming == LBM_PRE_STREAMING || streaming == LBM_POST_STREAMING) && n2 % 2 == 0;
}

// everything a step launch can carry besides the descriptor (lbm_api.cu: step_general)
struct StepExtras {
    const SlabSync *sync = nullptr;     // in-kernel slab lock step
    double *partials = nullptr;         // fused reductions: 2 * reduce_slots doubles ...
    int reduce_mode = kReduceNone;      // ... of the written (kReduceOutput) or the read (kReduceInput) state
    bool chained = false;               // launch behind the previous step kernel with programmatic serialization
    int *slots_used = nullptr;          // out, see LaunchOptions
};
// lbm_step plus the optional in-kernel slab lock step (lbm_slab_step_n), fused reductions (lbm_step_moments) and
// programmatic chaining behind the previous step (lbm_step_n); returns an lbm_status
int step_general(const lbm_step_desc *desc, const void *d_f_in, void *d_f_out, const StepExtras &x, void *stream);
// one step with fused reductions: partials + fold -> d_result[2] (lbm_step_moments, lbm_slab_step_moments)
int step_moments_general(const lbm_step_desc *desc, const void *d_f_in, void *d_f_out, const SlabSync *sync,
                         void *d_scratch, size_t scratch_bytes, double *d_result, bool chained, void *stream);
int cuda_fail_public(int e);                       // cudaError -> lbm_status, message kept for lbm_last_cuda_error
unsigned long long peer_timeout_cycles_public();   // LBM_B200_PEER_TIMEOUT_S in SM clocks

// returns cudaError_t as int (0 = success) or a negative lbm_status; defined in lbm_step_inst.cu
template <class S, class R, int COLL>
int launch_step_coll(const StepParams<R> &p, int streaming, const LaunchOptions &opt, cudaStream_t stream);

// link-wise post-streaming boundary (lbm_apply_links): gather kernel, then scatter kernel
inline int link_blocks(int64_t n) { return (int)((n + kLinkThreads - 1) / kLinkThreads); }
template <class S, class R, int COLL>
int launch_links_coll(const StepParams<R> &p, const LinkArgs<R> &a, cudaStream_t stream);

}  // namespace lbm
