// lbm_launch.cuh -- declaration of the per-(stencil, dtype, collision) step launchers.  Each
// triple is compiled in its own translation unit (lbm_step_inst.cu with -DLBM_INST_STENCIL /
// -DLBM_INST_REAL / -DLBM_INST_COLL) so the 40 units build in parallel.
#pragma once
#include <atomic>

#include "lbm_step.cuh"

namespace lbm {

extern std::atomic<int64_t> g_launch_count;  // kernels launched by this library (lbm_launch_count)

// launch geometry of the bulk kernels: threadIdx.x along the contiguous axis, blocks of up to 256 threads
// filled with rows, one grid layer per x-plane
inline void bulk_geometry(int n0, int n1, int n2, dim3 &grid, dim3 &block) {
    int tz = 32;
    while (tz < n2 && tz < 256) tz <<= 1;
    int ty = 256 / tz;
    while (ty > 1 && ty / 2 >= n1) ty >>= 1;
    block = dim3(tz, ty, 1);
    grid = dim3((n2 + tz - 1) / tz, (n1 + ty - 1) / ty, n0);
}

// returns cudaError_t as int (0 = success) or a negative lbm_status; defined in lbm_step_inst.cu
template <class S, class R, int COLL>
int launch_step_coll(const StepParams<R> &p, int streaming, bool masked, int variant, cudaStream_t stream);

// link-wise post-streaming boundary (lbm_apply_links): gather kernel, then scatter kernel
inline int link_blocks(int64_t n) { return (int)((n + kLinkThreads - 1) / kLinkThreads); }
template <class S, class R, int COLL>
int launch_links_coll(const StepParams<R> &p, const LinkArgs<R> &a, cudaStream_t stream);

}  // namespace lbm
