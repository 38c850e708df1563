// lbm_vec.cuh -- the value types the node arithmetic is written for, and their explicit operations.
//
//   float, double   one lattice node per thread
//   float2          TWO nodes per thread (z and z + 1) on Blackwell's packed fp32 pipe: FADD2 / FMUL2 / FFMA2 work on
//                   a register pair in one issue slot.  sm_100 issues a scalar FFMA every second cycle per SM
//                   sub-partition; the packed forms are what reaches the fp32 peak, and the collision operators
//                   (several hundred fp32 operations per node) are bound by exactly that pipe.
//
// Every operation is spelled out (vadd / vsub / vmul / vfma ...) and maps to a round-to-nearest intrinsic, which the
// compiler neither contracts nor reassociates.  Lane k of a float2 evaluation therefore yields the same bits as the
// float evaluation of that node: the one-node and two-node kernels, the sparse general-nodes kernel and the
// link kernels are interchangeable bit for bit.  The same code compiles for the host (tests/csrc/collide_host.cu
// checks the operators against the golden vectors without a GPU; host results may differ from the device in the
// last bit where the device uses an approximate reciprocal).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace lbm {

#define LBM_HD __host__ __device__ __forceinline__
#define LBM_D __device__ __forceinline__

template <class V>
struct VecTraits;
template <>
struct VecTraits<float> {
    using scalar = float;
    static constexpr int lanes = 1;
};
template <>
struct VecTraits<double> {
    using scalar = double;
    static constexpr int lanes = 1;
};
template <>
struct VecTraits<float2> {
    using scalar = float;
    static constexpr int lanes = 2;
};
template <class V>
using scalar_t = typename VecTraits<V>::scalar;

// ---- broadcast -------------------------------------------------------------------------------------------
template <class V>
LBM_HD V vset(double c);
template <>
LBM_HD float vset<float>(double c) { return (float)c; }
template <>
LBM_HD double vset<double>(double c) { return c; }
template <>
LBM_HD float2 vset<float2>(double c) { return make_float2((float)c, (float)c); }

template <class V>
LBM_HD V vsplat(scalar_t<V> s);
template <>
LBM_HD float vsplat<float>(float s) { return s; }
template <>
LBM_HD double vsplat<double>(double s) { return s; }
template <>
LBM_HD float2 vsplat<float2>(float s) { return make_float2(s, s); }

// ---- lanes -----------------------------------------------------------------------------------------------
LBM_HD float vlane(float a, int) { return a; }
LBM_HD double vlane(double a, int) { return a; }
LBM_HD float vlane(float2 a, int k) { return k ? a.y : a.x; }

// ---- arithmetic: a + b, a - b, a * b, a * b + c, -a ----------------------------------------------------------
#ifdef __CUDA_ARCH__
LBM_HD float vadd(float a, float b) { return __fadd_rn(a, b); }
LBM_HD float vsub(float a, float b) { return __fsub_rn(a, b); }
LBM_HD float vmul(float a, float b) { return __fmul_rn(a, b); }
LBM_HD float vfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
LBM_HD double vadd(double a, double b) { return __dadd_rn(a, b); }
LBM_HD double vsub(double a, double b) { return __dsub_rn(a, b); }
LBM_HD double vmul(double a, double b) { return __dmul_rn(a, b); }
LBM_HD double vfma(double a, double b, double c) { return __fma_rn(a, b, c); }
LBM_HD float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }       // folds into the operand modifier
LBM_HD float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
LBM_HD float2 vsub(float2 a, float2 b) { return __fadd2_rn(a, vneg(b)); }
LBM_HD float2 vmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
LBM_HD float2 vfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
// host build: compile with -ffp-contract=off so that a * b + c below is NOT fused behind our back
LBM_HD float vadd(float a, float b) { return a + b; }
LBM_HD float vsub(float a, float b) { return a - b; }
LBM_HD float vmul(float a, float b) { return a * b; }
LBM_HD float vfma(float a, float b, float c) { return fmaf(a, b, c); }
LBM_HD double vadd(double a, double b) { return a + b; }
LBM_HD double vsub(double a, double b) { return a - b; }
LBM_HD double vmul(double a, double b) { return a * b; }
LBM_HD double vfma(double a, double b, double c) { return fma(a, b, c); }
LBM_HD float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }
LBM_HD float2 vadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
LBM_HD float2 vsub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
LBM_HD float2 vmul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
LBM_HD float2 vfma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
LBM_HD float vneg(float a) { return -a; }
LBM_HD double vneg(double a) { return -a; }
// c - a * b
template <class V>
LBM_HD V vfnma(V a, V b, V c) { return vfma(vneg(a), b, c); }

// ---- 1 / x, correctly rounded (the reference divides: u = j / rho, lettuce/_flow.py:178-193) ---------------------
LBM_HD float vrecip(float a) {
#ifdef __CUDA_ARCH__
    return __frcp_rn(a);
#else
    return 1.0f / a;
#endif
}
LBM_HD double vrecip(double a) {
#ifdef __CUDA_ARCH__
    return __drcp_rn(a);
#else
    return 1.0 / a;
#endif
}
LBM_HD float2 vrecip(float2 a) { return make_float2(vrecip(a.x), vrecip(a.y)); }

// a / b, correctly rounded
LBM_HD float vdiv(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
LBM_HD double vdiv(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
LBM_HD float2 vdiv(float2 a, float2 b) { return make_float2(vdiv(a.x, b.x), vdiv(a.y, b.y)); }

// ---- a / b where the quotient feeds a noise-limited ratio (KBC's entropic sums, see Collide<KBC>): fp32 uses the
// approximate reciprocal (one MUFU.RCP, ~1 ulp; the divisor feq is O(1e-3..1), far from the denormal range), fp64 a
// reciprocal seed refined by two Newton steps to full precision; both avoid the IEEE division's slow-path
// subroutine 27 times per node ------------------------------------------------------------------------------------
LBM_HD float vdiv_fast(float a, float b) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return __fmul_rn(a, r);
#else
    return a * (1.0f / b);
#endif
}
LBM_HD double vdiv_fast(double a, double b) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = __fma_rn(__fma_rn(-b, r, 1.0), r, r);
    r = __fma_rn(__fma_rn(-b, r, 1.0), r, r);
    return __dmul_rn(a, r);
#else
    return a * (1.0 / b);
#endif
}
LBM_HD float2 vdiv_fast(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(b.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(b.y));
    return __fmul2_rn(a, r);
#else
    return make_float2(a.x * (1.0f / b.x), a.y * (1.0f / b.y));
#endif
}

// ---- x >= threshold ? x : alternative, per lane; a NaN takes the alternative -----------------------------------------
LBM_HD float vkeep_ge(float x, float thr, float alt) { return x >= thr ? x : alt; }
LBM_HD double vkeep_ge(double x, double thr, double alt) { return x >= thr ? x : alt; }
LBM_HD float2 vkeep_ge(float2 x, float thr, float alt) {
    return make_float2(x.x >= thr ? x.x : alt, x.y >= thr ? x.y : alt);
}

}  // namespace lbm
