// lbm_launch.cuh -- declaration of the per-(stencil, dtype, collision) step launchers.  Each
// triple is compiled in its own translation unit (lbm_step_inst.cu with -DLBM_INST_STENCIL /
// -DLBM_INST_REAL / -DLBM_INST_COLL) so the 40 units build in parallel.
#pragma once
#include "lbm_step.cuh"

namespace lbm {

extern int64_t g_launch_count;

// returns cudaError_t as int (0 = success) or a negative lbm_status; defined in lbm_step_inst.cu
template <class S, class R, int COLL>
int launch_step_coll(const StepParams<R> &p, int streaming, bool masked, int variant, cudaStream_t stream);

}  // namespace lbm
