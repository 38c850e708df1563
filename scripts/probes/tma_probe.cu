// Stand-alone probe of the TMA path used by csrc/lbm_tma.cuh: one CTA loads a box of a 4-D fp32 tensor into shared
// memory with cp.async.bulk.tensor, copies it back with a bulk tensor store.  Variants by argv[1]:
//   0  descriptor as __grid_constant__ kernel parameter, box 32      1  same, box 256
//   2  descriptor in global memory, box 32                           3  shifted coordinate (-1) with box 32
//   6 / 7 / 8 / 9  coordinate +1 / -4 / n2-4 / +3        10 / 11  shared-memory box 16 / 64 bytes off a 128-byte boundary
//   4  1-D bulk copy of 16 bytes                                      5  lane-parallel issue (27 lanes), box 32
// nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/probes/tma_probe.cu -o variants/tma_probe   (variants/ is git-ignored and travels with gpurun)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                      const CUtensorMap *gmaps, const float *in, float *out, int mode, int box, int c0) {
    extern __shared__ __align__(1024) float buf[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = smem_addr(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(mode == 5 ? 27 : 1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        printf("smem base %u (mod 128 = %u)\n", smem_addr(buf), smem_addr(buf) & 127u);
    }
    __syncthreads();
    const int soff = mode == 10 ? 4 : (mode == 11 ? 16 : 0);      // shared-memory offset in floats: 16 B, 64 B
    const CUtensorMap *im = mode == 2 ? gmaps : &in_map, *om = mode == 2 ? gmaps + 1 : &out_map;
    const bool issuer = mode == 5 ? threadIdx.x < 27 : threadIdx.x == 0;
    if (issuer) {
        const int q = mode == 5 ? threadIdx.x : 0;
        const unsigned bytes = mode == 4 ? 16u : (unsigned)box * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        if (mode == 4) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                         ::"r"(smem_addr(buf)), "l"(in + 4), "r"(b) : "memory");
        } else {
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                ::"r"(smem_addr(buf + q * box + soff)), "l"(im), "r"(c0), "r"(1), "r"(2), "r"(q), "r"(b) : "memory");
        }
    }
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(b), "r"(0) : "memory");
    }
    if (mode == 4) {
        if (threadIdx.x < 4) out[threadIdx.x] = buf[threadIdx.x];
        return;
    }
    for (int i = threadIdx.x; i < box * (mode == 5 ? 27 : 1); i += blockDim.x) buf[i + soff] += 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (issuer) {
        const int q = mode == 5 ? threadIdx.x : 0;
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(om), "r"(smem_addr(buf + q * box + soff)), "r"(0), "r"(1), "r"(2), "r"(q) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int box = (mode == 1) ? 256 : 32;
    const int n2 = 512, n1 = 6, n0 = 5, Q = 27;
    const size_t N = (size_t)n0 * n1 * n2;
    std::vector<float> h(Q * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000003);
    float *in, *out;
    CK(cudaMalloc(&in, h.size() * 4));
    CK(cudaMalloc(&out, h.size() * 4));
    CK(cudaMemcpy(in, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(out, 0, h.size() * 4));
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    printf("entry point %p query %d\n", fp, (int)qr);
    auto encode = (PFN_cuTensorMapEncodeTiled)fp;
    CUtensorMap maps[2];
    for (int k = 0; k < 2; ++k) {
        const cuuint64_t dims[4] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)n0, (cuuint64_t)Q};
        const cuuint64_t strides[3] = {(cuuint64_t)n2 * 4, (cuuint64_t)n1 * n2 * 4, (cuuint64_t)N * 4};
        const cuuint32_t bx[4] = {(cuuint32_t)box, 1, 1, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&maps[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, k ? (void *)out : (void *)in, dims, strides, bx,
                            es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d -> %d\n", k, (int)r);
    }
    CUtensorMap *gmaps;
    CK(cudaMalloc(&gmaps, sizeof maps));
    CK(cudaMemcpy(gmaps, maps, sizeof maps, cudaMemcpyHostToDevice));
    const int c0 = mode == 3 ? -1 : (mode == 6 ? 1 : (mode == 7 ? -4 : (mode == 8 ? n2 - 4 : (mode == 9 ? 3 : 0))));
    const size_t smem = 27 * 256 * 4 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<<<1, 64, smem>>>(maps[0], maps[1], gmaps, in, out, mode, box, c0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: kernel -> %s\n", mode, cudaGetErrorString(e));
    if (e) return 2;
    std::vector<float> o(h.size());
    CK(cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost));
    // element (q, x=2, y=1, z): expect in[z + c0] + 1 (0 + 1 outside the tensor)
    int bad = 0;
    const int nq = mode == 5 ? 27 : 1;
    if (mode == 4) {
        for (int i = 0; i < 4; ++i) bad += o[i] != h[4 + i];
    } else {
        for (int q = 0; q < nq; ++q)
            for (int z = 0; z < box; ++z) {
                const size_t at = q * N + (size_t)(2 * n1 + 1) * n2;
                const float want = (z + c0 >= 0 && z + c0 < n2 ? h[at + z + c0] : 0.0f) + 1.0f;
                bad += o[at + z] != want;
            }
    }
    printf("mode %d: %d wrong values\n", mode, bad);
    return bad ? 3 : 0;
}
