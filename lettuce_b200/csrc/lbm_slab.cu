// lbm_slab.cu -- multi-GPU x-slab plumbing of the C ABI: CUDA-IPC allocation, peer progress
// counters and the lock-stepped slab step loop.  The halo "exchange" itself is not here: it is the
// step kernel reading/writing the neighbour's boundary planes through the peer-mapped pointers that
// lbm_slab_step_n puts into the descriptor (csrc/lbm_step.cuh, in_plane / out_plane).
#include <cstring>

#include "lbm_launch.cuh"

namespace lbm {

// One thread: publish `epoch` to both neighbours, then wait until both neighbours have published it.
// Everything this rank's step kernel wrote (also into peer memory) is complete when this kernel starts
// (stream order); the system-scope fences order the counter against those writes for the peers.
__global__ void peer_signal_wait_kernel(unsigned long long *sig_lo, unsigned long long *sig_hi,
                                        const unsigned long long *wait_slots, unsigned long long epoch,
                                        unsigned long long timeout_cycles) {
    if (threadIdx.x != 0) return;
    __threadfence_system();
    *(volatile unsigned long long *)sig_lo = epoch;
    *(volatile unsigned long long *)sig_hi = epoch;
    __threadfence_system();
    // a neighbour that never arrives (crashed rank) must not hang the GPU for ever (LBM_B200_PEER_TIMEOUT_S)
    spin_until(wait_slots + 0, epoch, timeout_cycles);
    spin_until(wait_slots + 1, epoch, timeout_cycles);
    __threadfence_system();
}

// Wait only: used once at the end of a batch of in-kernel-synchronised steps, so that when the batch has
// completed on this rank's stream the neighbours have also finished the last step's boundary planes --
// their pushes into this rank's planes have landed and their pulls from them are done -- before a
// reporter reads, or the caller overwrites, the populations.
__global__ void peer_wait_kernel(const unsigned long long *wait_slots, unsigned long long epoch,
                                 unsigned long long timeout_cycles) {
    if (threadIdx.x != 0) return;
    spin_until(wait_slots + 0, epoch, timeout_cycles);
    spin_until(wait_slots + 1, epoch, timeout_cycles);
    __threadfence_system();
}

}  // namespace lbm

using namespace lbm;

extern "C" {

int lbm_ipc_alloc(size_t bytes, void **d_ptr, void *handle) {
    if (!d_ptr || !handle || bytes == 0) return LBM_ERR_BAD_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == LBM_IPC_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    int e = (int)cudaMalloc(&p, bytes);
    if (e) return cuda_fail_public(e);
    if ((e = (int)cudaMemset(p, 0, bytes))) { cudaFree(p); return cuda_fail_public(e); }
    cudaIpcMemHandle_t h;
    if ((e = (int)cudaIpcGetMemHandle(&h, p))) { cudaFree(p); return cuda_fail_public(e); }
    memcpy(handle, &h, sizeof h);
    *d_ptr = p;
    return LBM_OK;
}

int lbm_ipc_open(const void *handle, void **d_ptr) {
    if (!handle || !d_ptr) return LBM_ERR_BAD_ARGUMENT;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    const int e = (int)cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e) return cuda_fail_public(e);
    *d_ptr = p;
    return LBM_OK;
}

int lbm_ipc_close(void *d_ptr) {
    if (!d_ptr) return LBM_ERR_BAD_ARGUMENT;
    return cuda_fail_public((int)cudaIpcCloseMemHandle(d_ptr));
}

int lbm_ipc_free(void *d_ptr) {
    if (!d_ptr) return LBM_ERR_BAD_ARGUMENT;
    return cuda_fail_public((int)cudaFree(d_ptr));
}

// n lock-stepped steps; when d_result is given, the LAST one carries the fused reductions
static int slab_steps(const lbm_step_desc *desc, const lbm_slab *slab, void *d_f_a, void *d_f_b, int64_t n,
                      void *d_scratch, size_t scratch_bytes, double *d_result, void *stream) {
    if (!desc || !slab || !d_f_a || !d_f_b || n < 0) return LBM_ERR_BAD_ARGUMENT;
    if (!slab->lo_a || !slab->lo_b || !slab->hi_a || !slab->hi_b || !slab->signal_lo || !slab->signal_hi ||
        !slab->wait_slots || slab->lo_nx < 1 || slab->hi_nx < 1)
        return LBM_ERR_BAD_ARGUMENT;
    const size_t es = desc->lat.dtype == LBM_F32 ? 4 : 8;
    const int64_t plane = (int64_t)desc->lat.ny * desc->lat.nz;
    lbm_step_desc d = *desc;
    void *a = d_f_a, *b = d_f_b;
    char *lo_in = (char *)slab->lo_a, *lo_out = (char *)slab->lo_b;
    char *hi_in = (char *)slab->hi_a, *hi_out = (char *)slab->hi_b;
    unsigned long long epoch = slab->epoch;
    const unsigned long long timeout = peer_timeout_cycles_public();
    // Slabs thick enough for separate lo / hi boundary layers synchronise inside the step kernels; thinner ones use
    // the one-thread signal/wait kernel after each step.
    const int w = (desc->streaming == LBM_DOUBLE_STREAMING) ? 2 : 1;
    const bool fused = desc->lat.nx >= 2 * w;
    unsigned long long *counters = (unsigned long long *)slab->wait_slots;   // [0],[1] peers; [2]..[4] scratch
    if (fused) {
        const int e = (int)cudaMemsetAsync(counters + 2, 0, 3 * sizeof(unsigned long long), (cudaStream_t)stream);
        if (e) return cuda_fail_public(e);
    }
    for (int64_t k = 0; k < n; ++k) {
        // x = -1 is the LAST plane of the lo neighbour, x = nx is the FIRST plane of the hi neighbour
        d.halo.in_lo = lo_in + (size_t)(slab->lo_nx - 1) * plane * es;
        d.halo.out_lo = lo_out + (size_t)(slab->lo_nx - 1) * plane * es;
        d.halo.in_lo_qstride = d.halo.out_lo_qstride = (int64_t)slab->lo_nx * plane;
        d.halo.in_hi = hi_in;
        d.halo.out_hi = hi_out;
        d.halo.in_hi_qstride = d.halo.out_hi_qstride = (int64_t)slab->hi_nx * plane;
        const bool reduce = d_result != nullptr && k == n - 1;
        SlabSync sync = {};
        if (fused) {
            sync.sig_lo = (unsigned long long *)slab->signal_lo;
            sync.sig_hi = (unsigned long long *)slab->signal_hi;
            sync.wait = counters;
            sync.done = counters + 2;
            sync.wait_value = epoch;          // neighbours have completed the previous step
            sync.signal_value = epoch + 1;
            sync.timeout_cycles = timeout;
            sync.ctas_per_side = 0;           // filled in by the launcher, like publish
            sync.on = 1;
        }
        int rc;
        if (reduce) {
            rc = step_moments_general(&d, a, b, fused ? &sync : nullptr, d_scratch, scratch_bytes, d_result, false, stream);
        } else {
            StepExtras x;
            x.sync = fused ? &sync : nullptr;
            rc = step_general(&d, a, b, x, stream);
        }
        if (rc) return rc;
        ++epoch;
        if (!fused) {
            peer_signal_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long *)slab->signal_lo,
                                                                        (unsigned long long *)slab->signal_hi,
                                                                        (const unsigned long long *)slab->wait_slots,
                                                                        epoch, timeout);
            ++g_launch_count;
            const int e = (int)cudaGetLastError();
            if (e) return cuda_fail_public(e);
        }
        void *t = a; a = b; b = t;
        char *c = lo_in; lo_in = lo_out; lo_out = c;
        c = hi_in; hi_in = hi_out; hi_out = c;
    }
    if (fused && n > 0) {
        peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long *)slab->wait_slots, epoch,
                                                             timeout);
        ++g_launch_count;
        const int e = (int)cudaGetLastError();
        if (e) return cuda_fail_public(e);
    }
    return LBM_OK;
}

int lbm_slab_step_n(const lbm_step_desc *desc, const lbm_slab *slab, void *d_f_a, void *d_f_b, int64_t n,
                    void *stream) {
    return slab_steps(desc, slab, d_f_a, d_f_b, n, nullptr, 0, nullptr, stream);
}

int lbm_slab_step_moments(const lbm_step_desc *desc, const lbm_slab *slab, void *d_f_a, void *d_f_b,
                          void *d_scratch, size_t scratch_bytes, double *d_result, void *stream) {
    if (!d_scratch || !d_result) return LBM_ERR_BAD_ARGUMENT;
    return slab_steps(desc, slab, d_f_a, d_f_b, 1, d_scratch, scratch_bytes, d_result, stream);
}

}  // extern "C"
