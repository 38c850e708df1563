// lbm_launch.cuh -- declaration of the per-(stencil, dtype) step launchers.  Each
// (stencil, dtype) pair is compiled in its own translation unit
// (lbm_step_inst.cu with -DLBM_INST_STENCIL / -DLBM_INST_REAL) so the 6 units
// build in parallel.
#pragma once
#include "lbm_step.cuh"

namespace lbm {

extern int64_t g_launch_count;

// returns cudaError_t as int (0 = success) or LBM_ERR_UNSUPPORTED
template <class S, class R>
int launch_step(const StepParams<R> &p, int coll, int streaming, bool masked, int variant, cudaStream_t stream);

template <class S, class R>
const char *step_variant_name(const StepParams<R> &p, int coll, int streaming, bool masked, int variant);

}  // namespace lbm
