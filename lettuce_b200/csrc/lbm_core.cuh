// lbm_core.cuh -- velocity sets, equilibrium and collision operators as
// compile-time-unrolled device code.  Everything here works on a register array
// `R f[Q]` holding the populations of ONE lattice node.
//
// Semantics follow lettuce's torch path (the parity oracle), cited per function.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lbm_b200.h"

namespace lbm {

#define LBM_HD __host__ __device__ __forceinline__
#define LBM_D __device__ __forceinline__

// ---------------------------------------------------------------------------
// Velocity sets.  Internally every lattice is three-dimensional with extents
// (n0, n1, n2), n2 fastest.  A 2-D lattice [nx, ny] is stored as (nx, 1, ny) so
// that x is always axis 0 (the slab axis) and the contiguous axis is always
// axis 2; D2Q9's velocity components are mapped accordingly (e_y -> axis 2).
// Orders, weights and opposite tables: lettuce/ext/_stencil/d2q9.py:8-10,
// d3q19.py:8-13, d3q27.py:8-12.
// ---------------------------------------------------------------------------
struct D2Q9 {
    static constexpr int Q = 9, D = 2, ID = LBM_D2Q9;
    LBM_HD static constexpr int e(int q, int a) {  // a: internal axis 0..2
        constexpr int t[9][3] = {{0, 0, 0}, {1, 0, 0}, {0, 0, 1}, {-1, 0, 0}, {0, 0, -1},
                                 {1, 0, 1}, {-1, 0, 1}, {-1, 0, -1}, {1, 0, -1}};
        return t[q][a];
    }
    LBM_HD static constexpr int opp(int q) {
        constexpr int t[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
        return t[q];
    }
    LBM_HD static constexpr double w(int q) { return q == 0 ? 4.0 / 9.0 : (q < 5 ? 1.0 / 9.0 : 1.0 / 36.0); }
    // user-facing velocity component c (0..D-1) -> internal axis
    LBM_HD static constexpr int axis_of(int c) { return c == 0 ? 0 : 2; }
};

struct D3Q19 {
    static constexpr int Q = 19, D = 3, ID = LBM_D3Q19;
    LBM_HD static constexpr int e(int q, int a) {
        constexpr int t[19][3] = {{0, 0, 0},  {1, 0, 0},   {-1, 0, 0}, {0, 1, 0},  {0, -1, 0},
                                  {0, 0, 1},  {0, 0, -1},  {0, 1, 1},  {0, -1, -1}, {0, 1, -1},
                                  {0, -1, 1}, {1, 0, 1},   {-1, 0, -1}, {1, 0, -1}, {-1, 0, 1},
                                  {1, 1, 0},  {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}};
        return t[q][a];
    }
    LBM_HD static constexpr int opp(int q) { return q == 0 ? 0 : ((q & 1) ? q + 1 : q - 1); }
    LBM_HD static constexpr double w(int q) { return q == 0 ? 1.0 / 3.0 : (q < 7 ? 1.0 / 18.0 : 1.0 / 36.0); }
    LBM_HD static constexpr int axis_of(int c) { return c; }
};

struct D3Q27 {
    static constexpr int Q = 27, D = 3, ID = LBM_D3Q27;
    LBM_HD static constexpr int e(int q, int a) {
        constexpr int t[27][3] = {{0, 0, 0},   {1, 0, 0},   {-1, 0, 0},  {0, 1, 0},   {0, -1, 0},  {0, 0, 1},
                                  {0, 0, -1},  {0, 1, 1},   {0, -1, -1}, {0, 1, -1},  {0, -1, 1},  {1, 0, 1},
                                  {-1, 0, -1}, {1, 0, -1},  {-1, 0, 1},  {1, 1, 0},   {-1, -1, 0}, {1, -1, 0},
                                  {-1, 1, 0},  {1, 1, 1},   {-1, -1, -1}, {1, 1, -1}, {-1, -1, 1}, {1, -1, 1},
                                  {-1, 1, -1}, {1, -1, -1}, {-1, 1, 1}};
        return t[q][a];
    }
    LBM_HD static constexpr int opp(int q) { return q == 0 ? 0 : ((q & 1) ? q + 1 : q - 1); }
    LBM_HD static constexpr double w(int q) {
        return q == 0 ? 8.0 / 27.0 : (q < 7 ? 2.0 / 27.0 : (q < 19 ? 1.0 / 54.0 : 1.0 / 216.0));
    }
    LBM_HD static constexpr int axis_of(int c) { return c; }
};

// cs^2 exactly as the reference computes it: the square of the rounded 1/sqrt(3)
// (lettuce/_stencil.py:19, quadratic_equilibrium.py:19-22).
constexpr double kCs = 0.57735026918962584;  // 1/sqrt(3) rounded to double
constexpr double kCs2 = kCs * kCs;

// ---------------------------------------------------------------------------
// moments (lettuce/_flow.py:157-193)
// ---------------------------------------------------------------------------
template <class S, class R>
LBM_D void moments(const R (&f)[S::Q], R &rho, R (&j)[3]) {
    rho = R(0);
    j[0] = j[1] = j[2] = R(0);
#pragma unroll
    for (int q = 0; q < S::Q; ++q) {
        rho += f[q];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (S::e(q, a) == 1) j[a] += f[q];
            if (S::e(q, a) == -1) j[a] -= f[q];
        }
    }
}

// feq_q = w_q rho ((2 e.u - u.u)/(2 cs^2) + (e.u/cs^2)^2/2 + 1)
// (lettuce/ext/_equilibrium/quadratic_equilibrium.py:11-24)
template <class S, class R>
struct Equilibrium {
    R rho, u[3], base;  // base = 1 - u.u/(2 cs^2)
    LBM_D Equilibrium(R rho_, const R (&u_)[3]) : rho(rho_) {
        u[0] = u_[0]; u[1] = u_[1]; u[2] = u_[2];
        const R uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
        base = R(1) - uu * R(1.0 / (2.0 * kCs2));
    }
    template <int q>
    LBM_D R get() const {
        R eu = R(0);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (S::e(q, a) == 1) eu += u[a];
            if (S::e(q, a) == -1) eu -= u[a];
        }
        // 1 + eu/cs2 + eu^2/(2 cs2^2) - uu/(2 cs2)
        const R poly = base + eu * (R(1.0 / kCs2) + eu * R(0.5 / (kCs2 * kCs2)));
        return R(S::w(q)) * rho * poly;
    }
    // feq of q and of its opposite: they share the even part base + (e.u)^2/(2 cs^4)
    template <int q>
    LBM_D void pair(R &fq, R &fo) const {
        R eu = R(0);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (S::e(q, a) == 1) eu += u[a];
            if (S::e(q, a) == -1) eu -= u[a];
        }
        const R wr = R(S::w(q)) * rho;
        const R even = base + (eu * eu) * R(0.5 / (kCs2 * kCs2));
        const R odd = eu * R(1.0 / kCs2);
        fq = wr * (even + odd);
        fo = wr * (even - odd);
    }
#if defined(LBM_KBC_PACKED)
    // (feq_q, feq_opposite(q)) as one float2: wr * (even + odd * (1, -1))
    template <int q>
    LBM_D float2 pair2(const float2 pm) const {
        R eu = R(0);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (S::e(q, a) == 1) eu += u[a];
            if (S::e(q, a) == -1) eu -= u[a];
        }
        const float wr = float(R(S::w(q)) * rho);
        const float even = float(base + (eu * eu) * R(0.5 / (kCs2 * kCs2)));
        const float odd = float(eu * R(1.0 / kCs2));
        return __fmul2_rn(make_float2(wr, wr), __ffma2_rn(make_float2(odd, odd), pm, make_float2(even, even)));
    }
#endif
    // feq_q + feq_opposite(q)
    template <int q>
    LBM_D R pair_sum() const {
        R eu = R(0);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (S::e(q, a) == 1) eu += u[a];
            if (S::e(q, a) == -1) eu -= u[a];
        }
        return (R(2.0 * S::w(q)) * rho) * (base + (eu * eu) * R(0.5 / (kCs2 * kCs2)));
    }
};

// compile-time loop helper: body.template operator()<q>() for q in [0, Q)
template <int Q, int q = 0>
struct ForQ {
    template <class F>
    LBM_D static void run(F &&fn) {
        fn.template operator()<q>();
        ForQ<Q, q + 1>::run(fn);
    }
};
template <int Q>
struct ForQ<Q, Q> {
    template <class F>
    LBM_D static void run(F &&) {}
};

template <class S, class R>
LBM_D void equilibrium_all(R rho, const R (&u)[3], R (&feq)[S::Q]) {
    Equilibrium<S, R> eq(rho, u);
    ForQ<S::Q>::run([&]<int q>() { feq[q] = eq.template get<q>(); });
}

// ---------------------------------------------------------------------------
// collisions.  COLL is an lbm_op_kind collision value.
// ---------------------------------------------------------------------------
template <class S, class R, int COLL>
struct Collide;

template <class S, class R>
struct Collide<S, R, LBM_OP_NO_COLLISION> {
    LBM_D static void apply(R (&)[S::Q], R, R) {}
};

// f - (f - feq)/tau   (lettuce/ext/_collision/bgk_collision.py:17-22)
template <class S, class R>
struct Collide<S, R, LBM_OP_BGK> {
    LBM_D static void apply(R (&f)[S::Q], R inv_tau, R) {
        R rho, j[3];
        moments<S, R>(f, rho, j);
        const R inv_rho = R(1) / rho;
        const R u[3] = {j[0] * inv_rho, j[1] * inv_rho, j[2] * inv_rho};
        Equilibrium<S, R> eq(rho, u);
        ForQ<S::Q>::run([&]<int q>() { f[q] = f[q] - inv_tau * (f[q] - eq.template get<q>()); });
    }
};

// f - [ (f+ - feq+)/tau+ + (f- - feq-)/tau- ]   (lettuce/ext/_collision/trt_collision.py:16-27)
// a = 1/(2 tau+), b = 1/(2 tau-)
template <class S, class R>
struct Collide<S, R, LBM_OP_TRT> {
    LBM_D static void apply(R (&f)[S::Q], R a, R b) {
        R rho, j[3];
        moments<S, R>(f, rho, j);
        const R inv_rho = R(1) / rho;
        const R u[3] = {j[0] * inv_rho, j[1] * inv_rho, j[2] * inv_rho};
        Equilibrium<S, R> eq(rho, u);
        ForQ<S::Q>::run([&]<int q>() {
            constexpr int o = S::opp(q);
            if constexpr (q == 0) {
                const R fe = eq.template get<0>();
                f[0] = f[0] - ((f[0] + f[0]) - (fe + fe)) * a;
            } else if constexpr (q < o) {
                const R fq = f[q], fo = f[o];
                const R eq_q = eq.template get<q>(), eq_o = eq.template get<o>();
                const R even = ((fq + fo) - (eq_q + eq_o)) * a;
                const R odd = ((fq - fo) - (eq_q - eq_o)) * b;
                f[q] = fq - (even + odd);
                f[o] = fo - (even - odd);
            }
        });
    }
};

// dh / feq inside KBC's entropic sums.  gamma is the ratio of two sums that are dominated by
// rounding noise in smooth flow (see tests/test_gpu_parity.py), so fp32 uses the 2-ulp fast
// approximate reciprocal (one MUFU.RCP, ~1 ulp; feq is O(1e-3..1), far from the denormal range)
// instead of 27 IEEE divisions with their slow-path calls per node.
LBM_D float kbc_div(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return a * r;
}
// fp64: reciprocal seed (MUFU.RCP64H, ~20 bits) refined by two Newton steps to full precision; avoids the
// IEEE division's slow-path subroutine 27 times per node.
LBM_D double kbc_div(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    return a * r;
}

// Entropic KBC in the closed form of SURVEY.md Appendix A.3
// (lettuce/ext/_collision/kbc_collision.py:22-160).  beta = 1/(2 tau).
//
// Register plan: only f itself lives in registers.  Moments come from opposite-pair sums and
// differences, feq is re-evaluated per opposite pair where it is needed (the pair shares the even
// part; the second moments need only that part), and delta_s is one of ten scalars derived from
// the six second moments.
template <class S, class R>
struct Collide<S, R, LBM_OP_KBC> {
    // shear part of population q from the precomputed moment combinations (kbc_collision.py:44-94)
    template <int q>
    LBM_D static R ds_of(const R (&c)[7]) {
        if constexpr (q == 0) return c[0];
        if constexpr (S::D == 2) {
            // c[1] = (T+N)/4 (x axis pops 1,3), c[2] = (T-N)/4 (y axis pops 2,4), c[3] = Pxy/4
            if constexpr (q == 1 || q == 3) return c[1];
            else if constexpr (q == 2 || q == 4) return c[2];
            else if constexpr (q == 5 || q == 7) return c[3];
            else return -c[3];
        } else {
            if constexpr (q <= 2) return c[1];
            else if constexpr (q <= 4) return c[2];
            else if constexpr (q <= 6) return c[3];
            else if constexpr (q <= 8) return c[4];
            else if constexpr (q <= 10) return -c[4];
            else if constexpr (q <= 12) return c[5];
            else if constexpr (q <= 14) return -c[5];
            else if constexpr (q <= 16) return c[6];
            else if constexpr (q <= 18) return -c[6];
            else return R(0);
        }
    }

    LBM_D static void apply(R (&f)[S::Q], R beta, R) {
        constexpr int Q = S::Q;
        // Pass 1a: density and momentum from opposite-pair sums and differences (even moments only see
        // f_q + f_o, odd ones only f_q - f_o).
        R rho = f[0], j[3] = {R(0), R(0), R(0)};
        ForQ<Q>::run([&]<int q>() {
            constexpr int o = S::opp(q);
            if constexpr (q != 0 && q < o) {
                constexpr int e0 = S::e(q, 0), e1 = S::e(q, 1), e2 = S::e(q, 2);
                const R d = f[q] - f[o];
                rho += f[q] + f[o];
                if constexpr (e0 == 1) j[0] += d;
                if constexpr (e0 == -1) j[0] -= d;
                if constexpr (e1 == 1) j[1] += d;
                if constexpr (e1 == -1) j[1] -= d;
                if constexpr (e2 == 1) j[2] += d;
                if constexpr (e2 == -1) j[2] -= d;
            }
        });
        const R inv_rho = R(1) / rho;
        const R u[3] = {j[0] * inv_rho, j[1] * inv_rho, j[2] * inv_rho};
        Equilibrium<S, R> eq(rho, u);
        // Pass 1b: raw second moments of f - feq.  They are even in e, so per opposite pair only
        // (f_q + f_o) - (feq_q + feq_o) is needed, and feq_q + feq_o = 2 w rho (base + (e.u)^2/(2 cs^4)) costs
        // three operations.  (Subtracting the closed-form equilibrium stress from the moments of f instead
        // is cheaper still but loses a digit: the differences are taken between O(rho/3) numbers.)
        R P00 = 0, P11 = 0, P22 = 0, P01 = 0, P02 = 0, P12 = 0;
        ForQ<Q>::run([&]<int q>() {
            constexpr int o = S::opp(q);
            if constexpr (q != 0 && q < o) {
                constexpr int e0 = S::e(q, 0), e1 = S::e(q, 1), e2 = S::e(q, 2);
                const R s = (f[q] + f[o]) - eq.template pair_sum<q>();
                if constexpr (e0 != 0) P00 += s;
                if constexpr (e1 != 0) P11 += s;
                if constexpr (e2 != 0) P22 += s;
                if constexpr (e0 * e1 == 1) P01 += s;
                if constexpr (e0 * e1 == -1) P01 -= s;
                if constexpr (e0 * e2 == 1) P02 += s;
                if constexpr (e0 * e2 == -1) P02 -= s;
                if constexpr (e1 * e2 == 1) P12 += s;
                if constexpr (e1 * e2 == -1) P12 -= s;
            }
        });
        R c[7];
        if constexpr (S::D == 2) {
            const R T = P00 + P22, N = P00 - P22;        // internal axis 2 is y
            c[0] = -T;
            c[1] = R(0.25) * (T + N);
            c[2] = R(0.25) * (T - N);
            c[3] = R(0.25) * P02;
            c[4] = c[5] = c[6] = R(0);
        } else {
            const R T = P00 + P11 + P22, Nxz = P00 - P22, Nyz = P11 - P22;
            c[0] = -T;
            c[1] = (R(2) * Nxz - Nyz + T) * R(1.0 / 6.0);
            c[2] = (R(2) * Nyz - Nxz + T) * R(1.0 / 6.0);
            c[3] = (-Nxz - Nyz + T) * R(1.0 / 6.0);
            c[4] = R(0.25) * P12;
            c[5] = R(0.25) * P02;
            c[6] = R(0.25) * P01;
        }
#if defined(LBM_KBC_PACKED)
        // EXPERIMENT (not the default build): passes 2 and 3 with Blackwell's packed fp32 arithmetic
        // (FADD2 / FMUL2 / FFMA2 on sm_100): an opposite pair (q, o) is one float2.  Per lane the operations and
        // their rounding are those of the scalar code below, except that <dh|dh> is accumulated per lane.
        if constexpr (sizeof(R) == 4) {
            float sum_s = 0.f;
            float2 sum_h2 = make_float2(0.f, 0.f);
            const float2 pm = make_float2(1.f, -1.f);
            {
                const float fe = eq.template get<0>();
                const float ds = ds_of<0>(c), dh = (f[0] - fe) - ds;
                const float r = kbc_div(dh, fe);
                sum_s += ds * r;
                sum_h2.x += dh * r;
            }
            ForQ<Q>::run([&]<int q>() {
                constexpr int o = S::opp(q);
                if constexpr (q != 0 && q < o) {
                    const float2 EQ = eq.template pair2<q>(pm);
                    const float ds = ds_of<q>(c);
                    const float2 F = make_float2(f[q], f[o]);
                    const float2 DH = __fadd2_rn(__ffma2_rn(EQ, make_float2(-1.f, -1.f), F), make_float2(-ds, -ds));
                    float2 RC;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(RC.x) : "f"(EQ.x));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(RC.y) : "f"(EQ.y));
                    const float2 RR = __fmul2_rn(DH, RC);
                    sum_s += ds * (RR.x + RR.y);
                    sum_h2 = __ffma2_rn(DH, RR, sum_h2);
                }
            });
            const float sum_h = sum_h2.x + sum_h2.y;
            const float inv_beta = 1.f / beta;
            float gamma = inv_beta - (2.f - inv_beta) * (sum_s / sum_h);
            if (!(gamma >= 1e-15f)) gamma = 2.f;
            const float a = 1.f - beta * gamma, na = beta * gamma, b = beta * (gamma - 2.f);
            f[0] = a * f[0] + (na * eq.template get<0>() + b * ds_of<0>(c));
            ForQ<Q>::run([&]<int q>() {
                constexpr int o = S::opp(q);
                if constexpr (q != 0 && q < o) {
                    const float2 EQ = eq.template pair2<q>(pm);
                    const float bds = b * ds_of<q>(c);
                    const float2 F = make_float2(f[q], f[o]);
                    const float2 OUT = __ffma2_rn(make_float2(a, a), F,
                                                  __ffma2_rn(make_float2(na, na), EQ, make_float2(bds, bds)));
                    f[q] = OUT.x;
                    f[o] = OUT.y;
                }
            });
            return;
        }
#endif
        // Pass 2: entropic stabiliser gamma = 1/beta - (2 - 1/beta) <ds|dh>/<dh|dh>, weights 1/feq,
        // with dh = (f - feq) - ds.  f itself stays in the registers.
        R sum_s = 0, sum_h = 0;
        ForQ<Q>::run([&]<int q>() {
            constexpr int o = S::opp(q);
            if constexpr (q == 0) {
                const R fe = eq.template get<0>();
                const R ds = ds_of<0>(c), dh = (f[0] - fe) - ds;
                const R r = kbc_div(dh, fe);
                sum_s += ds * r;
                sum_h += dh * r;
            } else if constexpr (q < o) {
                R eq_q, eq_o;
                eq.template pair<q>(eq_q, eq_o);
                const R ds = ds_of<q>(c);          // even in e: same for q and its opposite
                const R dh_q = (f[q] - eq_q) - ds, dh_o = (f[o] - eq_o) - ds;
                const R r_q = kbc_div(dh_q, eq_q), r_o = kbc_div(dh_o, eq_o);
                sum_s += ds * (r_q + r_o);
                sum_h += dh_q * r_q + dh_o * r_o;
            }
        });
        const R inv_beta = R(1) / beta;
        R gamma = inv_beta - (R(2) - inv_beta) * (sum_s / sum_h);
        // kbc_collision.py:154-157: gamma < 1e-15 -> 2 ; NaN -> 2
        if (!(gamma >= R(1e-15))) gamma = R(2);
        // Pass 3: f' = f - beta (2 ds + gamma dh) = a f + (1 - a) feq + b ds,  a = 1 - beta gamma, b = beta (gamma - 2)
        const R a = R(1) - beta * gamma, na = beta * gamma, b = beta * (gamma - R(2));
        ForQ<Q>::run([&]<int q>() {
            constexpr int o = S::opp(q);
            if constexpr (q == 0) {
                f[0] = a * f[0] + (na * eq.template get<0>() + b * ds_of<0>(c));
            } else if constexpr (q < o) {
                R eq_q, eq_o;
                eq.template pair<q>(eq_q, eq_o);
                const R bds = b * ds_of<q>(c);
                f[q] = a * f[q] + (na * eq_q + bds);
                f[o] = a * f[o] + (na * eq_o + bds);
            }
        });
    }
};

// raw second moments P_ab = sum_q g_q e_qa e_qb of (f - feq), with f <- f - feq done in place
template <class S, class R>
struct NonEqMoments {
    R P00 = 0, P11 = 0, P22 = 0, P01 = 0, P02 = 0, P12 = 0;
    LBM_D void accumulate(const R (&g)[S::Q]) {
        ForQ<S::Q>::run([&]<int q>() {
            constexpr int e0 = S::e(q, 0), e1 = S::e(q, 1), e2 = S::e(q, 2);
            if constexpr (e0 != 0) P00 += g[q];
            if constexpr (e1 != 0) P11 += g[q];
            if constexpr (e2 != 0) P22 += g[q];
            if constexpr (e0 * e1 == 1) P01 += g[q];
            if constexpr (e0 * e1 == -1) P01 -= g[q];
            if constexpr (e0 * e2 == 1) P02 += g[q];
            if constexpr (e0 * e2 == -1) P02 -= g[q];
            if constexpr (e1 * e2 == 1) P12 += g[q];
            if constexpr (e1 * e2 == -1) P12 -= g[q];
        });
    }
};

// Regularized LBM (Latt & Chopard 2006): f = feq + (1 - 1/tau) w_q (Q_q : Pi_neq) / (2 cs^4),
// Q_q = e_q e_q - cs^2 I  (lettuce/ext/_collision/regularized_collision.py:17-43).  a = 1 - 1/tau.
template <class S, class R>
struct Collide<S, R, LBM_OP_REGULARIZED> {
    LBM_D static void apply(R (&f)[S::Q], R a, R) {
        R rho, j[3];
        moments<S, R>(f, rho, j);
        const R inv_rho = R(1) / rho;
        const R u[3] = {j[0] * inv_rho, j[1] * inv_rho, j[2] * inv_rho};
        Equilibrium<S, R> eq(rho, u);
        R feq[S::Q];
        ForQ<S::Q>::run([&]<int q>() {
            feq[q] = eq.template get<q>();
            f[q] -= feq[q];
        });
        NonEqMoments<S, R> m;
        m.accumulate(f);
        const R trace = (m.P00 + m.P11 + m.P22) * R(kCs2);
        const R scale = a * R(1.0 / (2.0 * kCs2 * kCs2));
        ForQ<S::Q>::run([&]<int q>() {
            constexpr int e0 = S::e(q, 0), e1 = S::e(q, 1), e2 = S::e(q, 2);
            R qpi = -trace;
            if constexpr (e0 != 0) qpi += m.P00;
            if constexpr (e1 != 0) qpi += m.P11;
            if constexpr (e2 != 0) qpi += m.P22;
            if constexpr (e0 * e1 != 0) qpi += R(2 * e0 * e1) * m.P01;
            if constexpr (e0 * e2 != 0) qpi += R(2 * e0 * e2) * m.P02;
            if constexpr (e1 * e2 != 0) qpi += R(2 * e1 * e2) * m.P12;
            f[q] = feq[q] + scale * (R(S::w(q)) * qpi);
        });
    }
};

// Smagorinsky LES on BGK (lettuce/ext/_collision/smagorinsky_collision.py:22-40, force = None): strain from the
// non-equilibrium second moments, two fixed-point iterations for tau_eff, then BGK with tau_eff.
// a = tau, b = smagorinsky constant.
template <class S, class R>
struct Collide<S, R, LBM_OP_SMAGORINSKY> {
    LBM_D static void apply(R (&f)[S::Q], R tau, R constant) {
        R rho, j[3];
        moments<S, R>(f, rho, j);
        const R inv_rho = R(1) / rho;
        const R u[3] = {j[0] * inv_rho, j[1] * inv_rho, j[2] * inv_rho};
        Equilibrium<S, R> eq(rho, u);
        R fn[S::Q];
        ForQ<S::Q>::run([&]<int q>() { fn[q] = f[q] - eq.template get<q>(); });
        NonEqMoments<S, R> m;
        m.accumulate(fn);
        // S_shear = Pi_neq / (2 rho cs^2); sum over ALL a,b of S_ab^2 (off-diagonals count twice)
        const R k = R(1) / (R(2.0 * kCs2) * rho);
        const R ss0 = k * k * (m.P00 * m.P00 + m.P11 * m.P11 + m.P22 * m.P22 +
                               R(2) * (m.P01 * m.P01 + m.P02 * m.P02 + m.P12 * m.P12));
        const R nu = (tau - R(0.5)) / R(3);
        R tau_eff = tau;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const R ss = ss0 / (tau_eff * tau_eff);
            tau_eff = (nu + constant * constant * ss) * R(3) + R(0.5);
        }
        const R inv_tau = R(1) / tau_eff;
        ForQ<S::Q>::run([&]<int q>() { f[q] = f[q] - inv_tau * fn[q]; });
    }
};

// BGK with a body force (lettuce/ext/_collision/bgk_collision.py:17-22 with Guo, ext/_force/guo.py:16-38,
// or ShanChen, ext/_force/shan_chen.py:13-26): the equilibrium is evaluated at u + ueq_scale a / rho and the
// source term  src_scale w_q [ (e_q - u)/cs^2 + (e_q.u) e_q / cs^4 ] . a  is added after relaxation.
template <class R>
struct ForceArgs {
    R a[3];  // acceleration, internal axes
    R ueq_scale, src_scale;
};

template <class S, class R>
LBM_D void collide_bgk_forced(R (&f)[S::Q], R inv_tau, const ForceArgs<R> &fa) {
    R rho, j[3];
    moments<S, R>(f, rho, j);
    const R inv_rho = R(1) / rho;
    const R k = fa.ueq_scale * inv_rho;
    const R u[3] = {j[0] * inv_rho + k * fa.a[0], j[1] * inv_rho + k * fa.a[1], j[2] * inv_rho + k * fa.a[2]};
    Equilibrium<S, R> eq(rho, u);
    const R ua = u[0] * fa.a[0] + u[1] * fa.a[1] + u[2] * fa.a[2];
    ForQ<S::Q>::run([&]<int q>() {
        R eu = R(0), ea = R(0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (S::e(q, c) == 1) { eu += u[c]; ea += fa.a[c]; }
            if (S::e(q, c) == -1) { eu -= u[c]; ea -= fa.a[c]; }
        }
        // (e - u).a / cs^2 + (e.u)(e.a) / cs^4
        const R src = (ea - ua) * R(1.0 / kCs2) + eu * ea * R(1.0 / (kCs2 * kCs2));
        f[q] = f[q] - inv_tau * (f[q] - eq.template get<q>()) + fa.src_scale * (R(S::w(q)) * src);
    });
}

// parameters handed to Collide::apply for a given collision kind
template <class R>
LBM_HD void collision_scalars(int kind, double p0, double p1, R &a, R &b) {
    a = R(0); b = R(0);
    if (kind == LBM_OP_BGK || kind == LBM_OP_BGK_FORCED) a = R(1.0 / p0);
    if (kind == LBM_OP_TRT) { a = R(1.0 / (2.0 * p0)); b = R(1.0 / (2.0 * p1)); }
    if (kind == LBM_OP_KBC) a = R(1.0 / (2.0 * p0));
    if (kind == LBM_OP_REGULARIZED) a = R(1.0 - 1.0 / p0);
    if (kind == LBM_OP_SMAGORINSKY) { a = R(p0); b = R(p1); }
}

}  // namespace lbm
