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

struct CollisionParams {
    double p0, p1;
};

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

// Entropic KBC in the closed form of SURVEY.md Appendix A.3
// (lettuce/ext/_collision/kbc_collision.py:22-160).  beta = 1/(2 tau).
template <class S, class R>
struct Collide<S, R, LBM_OP_KBC> {
    LBM_D static void apply(R (&f)[S::Q], R beta, R) {
        constexpr int Q = S::Q;
        R rho, j[3];
        moments<S, R>(f, rho, j);
        const R inv_rho = R(1) / rho;
        const R u[3] = {j[0] * inv_rho, j[1] * inv_rho, j[2] * inv_rho};
        Equilibrium<S, R> eq(rho, u);
        R feq[Q], fn[Q];
        ForQ<Q>::run([&]<int q>() {
            feq[q] = eq.template get<q>();
            fn[q] = f[q] - feq[q];
        });
        // raw second moments of the non-equilibrium part; axes 0,1,2 are x,y,z in 3-D and
        // x,(unused),y in 2-D.
        R P00 = 0, P11 = 0, P22 = 0, P01 = 0, P02 = 0, P12 = 0;
        ForQ<Q>::run([&]<int q>() {
            constexpr int e0 = S::e(q, 0), e1 = S::e(q, 1), e2 = S::e(q, 2);
            if constexpr (e0 != 0) P00 += fn[q];
            if constexpr (e1 != 0) P11 += fn[q];
            if constexpr (e2 != 0) P22 += fn[q];
            if constexpr (e0 * e1 == 1) P01 += fn[q];
            if constexpr (e0 * e1 == -1) P01 -= fn[q];
            if constexpr (e0 * e2 == 1) P02 += fn[q];
            if constexpr (e0 * e2 == -1) P02 -= fn[q];
            if constexpr (e1 * e2 == 1) P12 += fn[q];
            if constexpr (e1 * e2 == -1) P12 -= fn[q];
        });
        R ds[Q];
        if constexpr (S::D == 2) {
            // kbc_collision.py:76-94 ; internal axis 2 is y
            const R T = P00 + P22, N = P00 - P22, Pxy = P02;
            ds[0] = -T;
            ds[1] = ds[3] = R(0.25) * (T + N);
            ds[2] = ds[4] = R(0.25) * (T - N);
            ds[5] = ds[7] = R(0.25) * Pxy;
            ds[6] = ds[8] = R(-0.25) * Pxy;
        } else {
            // kbc_collision.py:44-74 ; entries 19..26 stay zero
            const R T = P00 + P11 + P22, Nxz = P00 - P22, Nyz = P11 - P22;
            ds[0] = -T;
            ds[1] = ds[2] = (R(2) * Nxz - Nyz + T) * R(1.0 / 6.0);
            ds[3] = ds[4] = (R(2) * Nyz - Nxz + T) * R(1.0 / 6.0);
            ds[5] = ds[6] = (-Nxz - Nyz + T) * R(1.0 / 6.0);
            ds[7] = ds[8] = R(0.25) * P12;
            ds[9] = ds[10] = R(-0.25) * P12;
            ds[11] = ds[12] = R(0.25) * P02;
            ds[13] = ds[14] = R(-0.25) * P02;
            ds[15] = ds[16] = R(0.25) * P01;
            ds[17] = ds[18] = R(-0.25) * P01;
#pragma unroll
            for (int q = 19; q < Q; ++q) ds[q] = R(0);
        }
        R sum_s = 0, sum_h = 0;
        ForQ<Q>::run([&]<int q>() {
            const R dh = fn[q] - ds[q];
            const R r = dh / feq[q];
            sum_s += ds[q] * r;
            sum_h += dh * r;
        });
        const R inv_beta = R(1) / beta;
        R gamma = inv_beta - (R(2) - inv_beta) * (sum_s / sum_h);
        // kbc_collision.py:154-157: gamma < 1e-15 -> 2 ; NaN -> 2
        if (!(gamma >= R(1e-15))) gamma = R(2);
        ForQ<Q>::run([&]<int q>() {
            const R dh = fn[q] - ds[q];
            f[q] = f[q] - beta * (R(2) * ds[q] + gamma * dh);
        });
    }
};

// parameters handed to Collide::apply for a given collision kind
template <class R>
LBM_HD void collision_scalars(int kind, double p0, double p1, R &a, R &b) {
    a = R(0); b = R(0);
    if (kind == LBM_OP_BGK) a = R(1.0 / p0);
    if (kind == LBM_OP_TRT) { a = R(1.0 / (2.0 * p0)); b = R(1.0 / (2.0 * p1)); }
    if (kind == LBM_OP_KBC) a = R(1.0 / (2.0 * p0));
}

}  // namespace lbm
