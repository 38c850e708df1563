// collide_host.cu -- TEST INFRASTRUCTURE, not part of the product: runs the collision operators of
// lettuce_b200/csrc/lbm_core.cuh on the HOST (the operators are written once for host and device, lbm_vec.cuh), so
// that their arithmetic can be checked against the golden vectors and the oracle without a GPU -- including the
// two-nodes-per-value float2 instantiation the packed kernels use.
//
// Build (tests/test_collide_host.py does this): nvcc -std=c++20 -O2 --expt-relaxed-constexpr -gencode
// arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off -shared -o libcollide_host.so collide_host.cu
#include "../../lettuce_b200/csrc/lbm_core.cuh"

using namespace lbm;

namespace {

template <class S, class V, int COLL>
void collide_one(V (&f)[S::Q], double p0, double p1, const double *force, double ueq_scale, double src_scale) {
    using T = scalar_t<V>;
    T a, b;
    collision_scalars<T>(COLL, p0, p1, a, b);
    if constexpr (COLL == LBM_OP_BGK_FORCED) {
        ForceArgs<T> fa;
        for (int c = 0; c < 3; ++c) fa.a[c] = T(0);
        for (int c = 0; c < S::D; ++c) fa.a[S::axis_of(c)] = (T)force[c];
        fa.ueq_scale = (T)ueq_scale;
        fa.src_scale = (T)src_scale;
        collide_bgk_forced<S, V>(f, a, fa);
    } else {
        Collide<S, V, COLL>::apply(f, a, b);
    }
}

// f: [Q][n] (population-major, like flow.f flattened), updated in place
template <class S, class T, int COLL>
int run(T *f, long n, int packed, double p0, double p1, const double *force, double ueq, double src) {
    constexpr int Q = S::Q;
    if (!packed) {
        for (long i = 0; i < n; ++i) {
            T g[Q];
            for (int q = 0; q < Q; ++q) g[q] = f[q * n + i];
            collide_one<S, T, COLL>(g, p0, p1, force, ueq, src);
            for (int q = 0; q < Q; ++q) f[q * n + i] = g[q];
        }
        return 0;
    }
    if constexpr (sizeof(T) == 4) {
        if (n % 2) return -1;
        for (long i = 0; i < n; i += 2) {
            float2 g[Q];
            for (int q = 0; q < Q; ++q) g[q] = make_float2(f[q * n + i], f[q * n + i + 1]);
            collide_one<S, float2, COLL>(g, p0, p1, force, ueq, src);
            for (int q = 0; q < Q; ++q) {
                f[q * n + i] = g[q].x;
                f[q * n + i + 1] = g[q].y;
            }
        }
        return 0;
    }
    return -1;
}

template <class S, class T>
int by_collision(int coll, T *f, long n, int packed, double p0, double p1, const double *force, double ueq,
                 double src) {
    switch (coll) {
        case LBM_OP_NO_COLLISION: return run<S, T, LBM_OP_NO_COLLISION>(f, n, packed, p0, p1, force, ueq, src);
        case LBM_OP_BGK: return run<S, T, LBM_OP_BGK>(f, n, packed, p0, p1, force, ueq, src);
        case LBM_OP_TRT: return run<S, T, LBM_OP_TRT>(f, n, packed, p0, p1, force, ueq, src);
        case LBM_OP_KBC:
            if constexpr (S::ID == LBM_D3Q19) return -2;
            else return run<S, T, LBM_OP_KBC>(f, n, packed, p0, p1, force, ueq, src);
        case LBM_OP_REGULARIZED: return run<S, T, LBM_OP_REGULARIZED>(f, n, packed, p0, p1, force, ueq, src);
        case LBM_OP_SMAGORINSKY: return run<S, T, LBM_OP_SMAGORINSKY>(f, n, packed, p0, p1, force, ueq, src);
        case LBM_OP_BGK_FORCED: return run<S, T, LBM_OP_BGK_FORCED>(f, n, packed, p0, p1, force, ueq, src);
    }
    return -3;
}

template <class T>
int by_stencil(int stencil, int coll, T *f, long n, int packed, double p0, double p1, const double *force, double ueq,
               double src) {
    switch (stencil) {
        case LBM_D2Q9: return by_collision<D2Q9, T>(coll, f, n, packed, p0, p1, force, ueq, src);
        case LBM_D3Q19: return by_collision<D3Q19, T>(coll, f, n, packed, p0, p1, force, ueq, src);
        case LBM_D3Q27: return by_collision<D3Q27, T>(coll, f, n, packed, p0, p1, force, ueq, src);
    }
    return -3;
}

}  // namespace

extern "C" int collide_host(int stencil, int dtype, int coll, int packed, void *f, long n, double p0, double p1,
                            const double *force, double ueq_scale, double src_scale) {
    static const double zero[3] = {0, 0, 0};
    if (!force) force = zero;
    if (dtype == LBM_F32) return by_stencil<float>(stencil, coll, (float *)f, n, packed, p0, p1, force, ueq_scale, src_scale);
    return by_stencil<double>(stencil, coll, (double *)f, n, packed, p0, p1, force, ueq_scale, src_scale);
}
