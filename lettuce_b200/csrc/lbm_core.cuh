// lbm_core.cuh -- velocity sets, equilibrium and collision operators as compile-time-unrolled code on a register
// array `V f[Q]` holding the populations of ONE lattice node (V = float, double) or of TWO neighbouring nodes
// (V = float2, Blackwell's packed fp32 arithmetic; see lbm_vec.cuh).  Written once, with explicit round-to-nearest
// operations, so that every instantiation yields the same bits per node; compiles for the device and the host.
//
// Semantics follow lettuce's torch path (the parity oracle), cited per function.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lbm_b200.h"
#include "lbm_vec.cuh"

namespace lbm {

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

// compile-time loop helper: body.template operator()<q>() for q in [0, Q)
template <int Q, int q = 0>
struct ForQ {
    template <class F>
    LBM_HD static void run(F &&fn) {
        fn.template operator()<q>();
        ForQ<Q, q + 1>::run(fn);
    }
};
template <int Q>
struct ForQ<Q, Q> {
    template <class F>
    LBM_HD static void run(F &&) {}
};

// running sum whose first term is taken as is (no "0 + x"); every branch below is resolved at compile time
// once the loops over q are unrolled
template <class V>
struct Acc {
    V v;
    bool set = false;
    LBM_HD void add(V t) {
        v = set ? vadd(v, t) : t;
        set = true;
    }
    LBM_HD void sub(V t) {
        v = set ? vsub(v, t) : vneg(t);
        set = true;
    }
    LBM_HD void add_signed(int sign, V t) {
        if (sign > 0) add(t);
        if (sign < 0) sub(t);
    }
    LBM_HD V get() const { return set ? v : vset<V>(0.0); }
};

// e_q . u as a sum of +-u_a
template <class S, class V, int q>
LBM_HD V e_dot(const V (&u)[3]) {
    Acc<V> eu;
#pragma unroll
    for (int a = 0; a < 3; ++a) eu.add_signed(S::e(q, a), u[a]);
    return eu.get();
}

// ---------------------------------------------------------------------------
// moments (lettuce/_flow.py:157-193): rho = sum_q f_q, j = sum_q e_q f_q
// ---------------------------------------------------------------------------
template <class S, class V>
LBM_HD void moments(const V (&f)[S::Q], V &rho, V (&j)[3]) {
    Acc<V> r, m[3];
    ForQ<S::Q>::run([&]<int q>() {
        r.add(f[q]);
#pragma unroll
        for (int a = 0; a < 3; ++a) m[a].add_signed(S::e(q, a), f[q]);
    });
    rho = r.get();
    j[0] = m[0].get(); j[1] = m[1].get(); j[2] = m[2].get();
}

// rho and u = j / rho of one node (Flow.u without force correction)
template <class S, class V>
LBM_HD void density_velocity(const V (&f)[S::Q], V &rho, V (&u)[3]) {
    V j[3];
    moments<S, V>(f, rho, j);
    const V inv = vrecip(rho);
    u[0] = vmul(j[0], inv); u[1] = vmul(j[1], inv); u[2] = vmul(j[2], inv);
}

// feq_q = w_q rho ((2 e.u - u.u)/(2 cs^2) + (e.u/cs^2)^2/2 + 1)
//       = w_q rho (base + e.u (1/cs^2 + e.u / (2 cs^4))),  base = 1 - u.u/(2 cs^2)
// (lettuce/ext/_equilibrium/quadratic_equilibrium.py:11-24)
template <class S, class V>
struct Equilibrium {
    V rho, u[3], base;
    LBM_HD Equilibrium(V rho_, const V (&u_)[3]) : rho(rho_) {
        u[0] = u_[0]; u[1] = u_[1]; u[2] = u_[2];
        const V uu = vfma(u[2], u[2], vfma(u[1], u[1], vmul(u[0], u[0])));
        base = vfma(uu, vset<V>(-1.0 / (2.0 * kCs2)), vset<V>(1.0));
    }
    template <int q>
    LBM_HD V wrho() const { return vmul(vset<V>(S::w(q)), rho); }
    template <int q>
    LBM_HD V get() const {
        if constexpr (q == 0) {
            return vmul(wrho<0>(), base);
        } else {
            const V eu = e_dot<S, V, q>(u);
            const V poly = vfma(eu, vfma(eu, vset<V>(0.5 / (kCs2 * kCs2)), vset<V>(1.0 / kCs2)), base);
            return vmul(wrho<q>(), poly);
        }
    }
    // feq_q - g and g - feq_q with the last product fused into the subtraction.  (Spelled out because ptxas
    // contracts a packed mul.rn.f32x2 feeding an add.rn.f32x2 into FFMA2 although both carry a rounding modifier --
    // it never does that to the scalar forms -- so the one-node and two-node kernels give the same bits only if no
    // product is left to feed a sum: scripts/check_packed_contraction.py.)
    template <int q>
    LBM_HD V minus(V g) const {
        if constexpr (q == 0) {
            return vfma(wrho<0>(), base, vneg(g));
        } else {
            const V eu = e_dot<S, V, q>(u);
            const V poly = vfma(eu, vfma(eu, vset<V>(0.5 / (kCs2 * kCs2)), vset<V>(1.0 / kCs2)), base);
            return vfma(wrho<q>(), poly, vneg(g));
        }
    }
    template <int q>
    LBM_HD V subtracted_from(V g) const {
        if constexpr (q == 0) {
            return vfnma(wrho<0>(), base, g);
        } else {
            const V eu = e_dot<S, V, q>(u);
            const V poly = vfma(eu, vfma(eu, vset<V>(0.5 / (kCs2 * kCs2)), vset<V>(1.0 / kCs2)), base);
            return vfnma(wrho<q>(), poly, g);
        }
    }
    // even part of q and its opposite: base + (e.u)^2 / (2 cs^4); feq_q + feq_opposite = 2 w rho even
    template <int q>
    LBM_HD V even(V eu) const { return vfma(vmul(eu, eu), vset<V>(0.5 / (kCs2 * kCs2)), base); }
    // feq of q and of its opposite, sharing the even part: w rho (even +- e.u / cs^2)
    template <int q>
    LBM_HD void pair(V &fq, V &fo) const {
        const V eu = e_dot<S, V, q>(u);
        const V wr = wrho<q>();
        const V a = vmul(wr, even<q>(eu));
        const V b = vmul(wr, vset<V>(1.0 / kCs2));
        fq = vfma(b, eu, a);
        fo = vfnma(b, eu, a);
    }
};

template <class S, class V>
LBM_HD void equilibrium_all(V rho, const V (&u)[3], V (&feq)[S::Q]) {
    Equilibrium<S, V> eq(rho, u);
    ForQ<S::Q>::run([&]<int q>() { feq[q] = eq.template get<q>(); });
}

// ---------------------------------------------------------------------------
// collisions.  COLL is an lbm_op_kind collision value; a and b are the operator's scalars
// (collision_scalars below), the same for every lane.
// ---------------------------------------------------------------------------
template <class S, class V, int COLL>
struct Collide;

template <class S, class V>
struct Collide<S, V, LBM_OP_NO_COLLISION> {
    LBM_HD static void apply(V (&)[S::Q], scalar_t<V>, scalar_t<V>) {}
};

// f - (f - feq)/tau   (lettuce/ext/_collision/bgk_collision.py:17-22); a = 1/tau
template <class S, class V>
struct Collide<S, V, LBM_OP_BGK> {
    LBM_HD static void apply(V (&f)[S::Q], scalar_t<V> inv_tau, scalar_t<V>) {
        V rho, u[3];
        density_velocity<S, V>(f, rho, u);
        Equilibrium<S, V> eq(rho, u);
        const V omega = vsplat<V>(inv_tau);
        ForQ<S::Q>::run([&]<int q>() { f[q] = vfma(omega, eq.template minus<q>(f[q]), f[q]); });
    }
};

// f - [ (f+ - feq+)/tau+ + (f- - feq-)/tau- ]   (lettuce/ext/_collision/trt_collision.py:16-27)
// a = 1/(2 tau+), b = 1/(2 tau-)
template <class S, class V>
struct Collide<S, V, LBM_OP_TRT> {
    LBM_HD static void apply(V (&f)[S::Q], scalar_t<V> a_, scalar_t<V> b_) {
        V rho, u[3];
        density_velocity<S, V>(f, rho, u);
        Equilibrium<S, V> eq(rho, u);
        const V a = vsplat<V>(a_), b = vsplat<V>(b_);
        ForQ<S::Q>::run([&]<int q>() {
            constexpr int o = S::opp(q);
            if constexpr (q == 0) {
                const V d = eq.template subtracted_from<0>(f[0]);
                f[0] = vfnma(vadd(d, d), a, f[0]);
            } else if constexpr (q < o) {
                V eq_q, eq_o;
                eq.template pair<q>(eq_q, eq_o);
                const V fq = f[q], fo = f[o];
                const V even = vsub(vadd(fq, fo), vadd(eq_q, eq_o));
                const V odd = vsub(vsub(fq, fo), vsub(eq_q, eq_o));
                f[q] = vfnma(b, odd, vfnma(a, even, fq));
                f[o] = vfma(b, odd, vfnma(a, even, fo));
            }
        });
    }
};

// Entropic KBC in the closed form of SURVEY.md Appendix A.3
// (lettuce/ext/_collision/kbc_collision.py:22-160).  a = beta = 1/(2 tau).
//
// Register plan: only f itself lives in registers.  Moments come from opposite-pair sums and
// differences, feq is re-evaluated per opposite pair where it is needed (the pair shares the even
// part; the second moments need only that part), and delta_s is one of ten scalars derived from
// the six second moments.  In fp32 this is ~550 operations per node on the fp32 pipe: the one-node kernel is bound
// by that pipe (sm_100 issues a scalar FFMA every second cycle); the two-node float2 instantiation halves it.
template <class S, class V>
struct Collide<S, V, LBM_OP_KBC> {
    // shear part of population q from the precomputed moment combinations (kbc_collision.py:44-94):
    // index into c[] and sign; index < 0: no shear part
    LBM_HD static constexpr int ds_index(int q) {
        if (q == 0) return 0;
        if (S::D == 2) return q <= 4 ? (q & 1 ? 1 : 2) : 3;      // c[1] (x axis: 1,3), c[2] (y axis: 2,4), c[3] = Pxy/4
        return q <= 6 ? (q + 1) / 2 : (q <= 10 ? 4 : (q <= 14 ? 5 : (q <= 18 ? 6 : -1)));
    }
    LBM_HD static constexpr int ds_sign(int q) {
        if (S::D == 2) return (q == 6 || q == 8) ? -1 : 1;
        return (q == 9 || q == 10 || q == 13 || q == 14 || q == 17 || q == 18) ? -1 : 1;
    }
    template <int q>
    LBM_HD static V ds_of(const V (&c)[7]) {
        static_assert(ds_index(q) >= 0);
        return ds_sign(q) > 0 ? c[ds_index(q)] : vneg(c[ds_index(q)]);
    }

    LBM_HD static void apply(V (&f)[S::Q], scalar_t<V> beta_, scalar_t<V>) {
        constexpr int Q = S::Q;
        // Pass 1a: density and momentum from opposite-pair sums and differences (even moments only see
        // f_q + f_o, odd ones only f_q - f_o).
        Acc<V> r, m[3];
        r.add(f[0]);
        ForQ<Q>::run([&]<int q>() {
            constexpr int o = S::opp(q);
            if constexpr (q != 0 && q < o) {
                r.add(vadd(f[q], f[o]));
                const V d = vsub(f[q], f[o]);
#pragma unroll
                for (int a = 0; a < 3; ++a) m[a].add_signed(S::e(q, a), d);
            }
        });
        const V rho = r.get();
        const V inv_rho = vrecip(rho);
        const V u[3] = {vmul(m[0].get(), inv_rho), vmul(m[1].get(), inv_rho), vmul(m[2].get(), inv_rho)};
        Equilibrium<S, V> eq(rho, u);
        // Pass 1b: raw second moments of f - feq.  They are even in e, so per opposite pair only
        // (f_q + f_o) - (feq_q + feq_o) is needed, and feq_q + feq_o = 2 w rho (base + (e.u)^2/(2 cs^4)).
        // (Subtracting the closed-form equilibrium stress from the moments of f instead is cheaper still but
        // loses a digit: the differences are taken between O(rho/3) numbers.)
        Acc<V> P00, P11, P22, P01, P02, P12;
        ForQ<Q>::run([&]<int q>() {
            constexpr int o = S::opp(q);
            if constexpr (q != 0 && q < o) {
                constexpr int e0 = S::e(q, 0), e1 = S::e(q, 1), e2 = S::e(q, 2);
                const V eu = e_dot<S, V, q>(u);
                const V s = vfnma(vmul(vset<V>(2.0 * S::w(q)), rho), eq.template even<q>(eu), vadd(f[q], f[o]));
                if constexpr (e0 != 0) P00.add(s);
                if constexpr (e1 != 0) P11.add(s);
                if constexpr (e2 != 0) P22.add(s);
                P01.add_signed(e0 * e1, s);
                P02.add_signed(e0 * e2, s);
                P12.add_signed(e1 * e2, s);
            }
        });
        V c[7];
        if constexpr (S::D == 2) {
            const V T = vadd(P00.get(), P22.get()), N = vsub(P00.get(), P22.get());        // internal axis 2 is y
            c[0] = vneg(T);
            c[1] = vmul(vset<V>(0.25), vadd(T, N));
            c[2] = vmul(vset<V>(0.25), vsub(T, N));
            c[3] = vmul(vset<V>(0.25), P02.get());
            c[4] = c[5] = c[6] = vset<V>(0.0);
        } else {
            const V T = vadd(vadd(P00.get(), P11.get()), P22.get());
            const V Nxz = vsub(P00.get(), P22.get()), Nyz = vsub(P11.get(), P22.get());
            const V sixth = vset<V>(1.0 / 6.0);
            c[0] = vneg(T);
            c[1] = vmul(vadd(vsub(vadd(Nxz, Nxz), Nyz), T), sixth);      // (2 Nxz - Nyz + T) / 6
            c[2] = vmul(vadd(vsub(vadd(Nyz, Nyz), Nxz), T), sixth);      // (2 Nyz - Nxz + T) / 6
            c[3] = vmul(vsub(T, vadd(Nxz, Nyz)), sixth);                 // (-Nxz - Nyz + T) / 6
            c[4] = vmul(vset<V>(0.25), P12.get());
            c[5] = vmul(vset<V>(0.25), P02.get());
            c[6] = vmul(vset<V>(0.25), P01.get());
        }
        // Pass 2: entropic stabiliser gamma = 1/beta - (2 - 1/beta) <ds|dh>/<dh|dh>, weights 1/feq,
        // with dh = (f - feq) - ds.  f itself stays in the registers.  gamma is the ratio of two sums that are
        // dominated by rounding noise in smooth flow (tests/test_gpu_parity.py), hence vdiv_fast for the 27 weights.
        V sum_s, sum_h;
        {
            const V fe = eq.template get<0>();
            const V ds = ds_of<0>(c), dh = vsub(vsub(f[0], fe), ds);
            const V w = vdiv_fast(dh, fe);
            sum_s = vmul(ds, w);
            sum_h = vmul(dh, w);
        }
        ForQ<Q>::run([&]<int q>() {
            constexpr int o = S::opp(q);
            if constexpr (q != 0 && q < o) {
                V eq_q, eq_o;
                eq.template pair<q>(eq_q, eq_o);
                if constexpr (ds_index(q) >= 0) {
                    const V ds = ds_of<q>(c);          // even in e: same for q and its opposite
                    const V dh_q = vsub(vsub(f[q], eq_q), ds), dh_o = vsub(vsub(f[o], eq_o), ds);
                    const V w_q = vdiv_fast(dh_q, eq_q), w_o = vdiv_fast(dh_o, eq_o);
                    sum_s = vfma(ds, vadd(w_q, w_o), sum_s);
                    sum_h = vfma(dh_q, w_q, vfma(dh_o, w_o, sum_h));
                } else {                               // corner populations of D3Q27 carry no shear part
                    const V dh_q = vsub(f[q], eq_q), dh_o = vsub(f[o], eq_o);
                    sum_h = vfma(dh_q, vdiv_fast(dh_q, eq_q), vfma(dh_o, vdiv_fast(dh_o, eq_o), sum_h));
                }
            }
        });
        const V beta = vsplat<V>(beta_);
        const V inv_beta = vsplat<V>(scalar_t<V>(1) / beta_);
        V gamma = vfnma(vsub(vset<V>(2.0), inv_beta), vdiv(sum_s, sum_h), inv_beta);
        // kbc_collision.py:154-157: gamma < 1e-15 -> 2 ; NaN -> 2
        gamma = vkeep_ge(gamma, scalar_t<V>(1e-15), scalar_t<V>(2));
        // Pass 3: f' = f - beta (2 ds + gamma dh) = a f + na feq + b ds,  na = beta gamma, a = 1 - na,
        // b = beta (gamma - 2); per opposite pair na feq = G +- H with G = na w rho even, H = na w rho e.u / cs^2
        const V na = vmul(beta, gamma), a = vsub(vset<V>(1.0), na), b = vmul(beta, vsub(gamma, vset<V>(2.0)));
        const V nrho = vmul(na, rho);
        f[0] = vfma(a, f[0], vfma(vmul(vset<V>(S::w(0)), nrho), eq.base, vmul(b, ds_of<0>(c))));
        ForQ<Q>::run([&]<int q>() {
            constexpr int o = S::opp(q);
            if constexpr (q != 0 && q < o) {
                const V eu = e_dot<S, V, q>(u);
                const V nwr = vmul(vset<V>(S::w(q)), nrho);
                V g;
                if constexpr (ds_index(q) >= 0) g = vfma(nwr, eq.template even<q>(eu), vmul(b, ds_of<q>(c)));
                else g = vmul(nwr, eq.template even<q>(eu));
                const V h = vmul(nwr, vset<V>(1.0 / kCs2));
                f[q] = vfma(a, f[q], vfma(h, eu, g));
                f[o] = vfma(a, f[o], vfnma(h, eu, g));
            }
        });
    }
};

// raw second moments P_ab = sum_q g_q e_qa e_qb of a set of populations g
template <class S, class V>
struct SecondMoments {
    V P00, P11, P22, P01, P02, P12;
    LBM_HD explicit SecondMoments(const V (&g)[S::Q]) {
        Acc<V> a00, a11, a22, a01, a02, a12;
        ForQ<S::Q>::run([&]<int q>() {
            constexpr int e0 = S::e(q, 0), e1 = S::e(q, 1), e2 = S::e(q, 2);
            if constexpr (e0 != 0) a00.add(g[q]);
            if constexpr (e1 != 0) a11.add(g[q]);
            if constexpr (e2 != 0) a22.add(g[q]);
            a01.add_signed(e0 * e1, g[q]);
            a02.add_signed(e0 * e2, g[q]);
            a12.add_signed(e1 * e2, g[q]);
        });
        P00 = a00.get(); P11 = a11.get(); P22 = a22.get();
        P01 = a01.get(); P02 = a02.get(); P12 = a12.get();
    }
};

// Regularized LBM (Latt & Chopard 2006): f = feq + (1 - 1/tau) w_q (Q_q : Pi_neq) / (2 cs^4),
// Q_q = e_q e_q - cs^2 I  (lettuce/ext/_collision/regularized_collision.py:17-43).  a = 1 - 1/tau.
template <class S, class V>
struct Collide<S, V, LBM_OP_REGULARIZED> {
    LBM_HD static void apply(V (&f)[S::Q], scalar_t<V> a, scalar_t<V>) {
        V rho, u[3];
        density_velocity<S, V>(f, rho, u);
        Equilibrium<S, V> eq(rho, u);
        V feq[S::Q];
        ForQ<S::Q>::run([&]<int q>() {
            feq[q] = eq.template get<q>();
            f[q] = vsub(f[q], feq[q]);
        });
        const SecondMoments<S, V> m(f);
        const V trace = vmul(vadd(vadd(m.P00, m.P11), m.P22), vset<V>(kCs2));
        const V scale = vsplat<V>(a * scalar_t<V>(1.0 / (2.0 * kCs2 * kCs2)));
        ForQ<S::Q>::run([&]<int q>() {
            constexpr int e0 = S::e(q, 0), e1 = S::e(q, 1), e2 = S::e(q, 2);
            V qpi = vneg(trace);
            if constexpr (e0 != 0) qpi = vadd(qpi, m.P00);
            if constexpr (e1 != 0) qpi = vadd(qpi, m.P11);
            if constexpr (e2 != 0) qpi = vadd(qpi, m.P22);
            if constexpr (e0 * e1 != 0) qpi = vfma(vset<V>(2 * e0 * e1), m.P01, qpi);
            if constexpr (e0 * e2 != 0) qpi = vfma(vset<V>(2 * e0 * e2), m.P02, qpi);
            if constexpr (e1 * e2 != 0) qpi = vfma(vset<V>(2 * e1 * e2), m.P12, qpi);
            f[q] = vfma(scale, vmul(vset<V>(S::w(q)), qpi), feq[q]);
        });
    }
};

// Smagorinsky LES on BGK (lettuce/ext/_collision/smagorinsky_collision.py:22-40, force = None): strain from the
// non-equilibrium second moments, two fixed-point iterations for tau_eff, then BGK with tau_eff.
// a = tau, b = smagorinsky constant.
template <class S, class V>
struct Collide<S, V, LBM_OP_SMAGORINSKY> {
    LBM_HD static void apply(V (&f)[S::Q], scalar_t<V> tau_, scalar_t<V> constant_) {
        using T = scalar_t<V>;
        V rho, u[3];
        density_velocity<S, V>(f, rho, u);
        Equilibrium<S, V> eq(rho, u);
        V fn[S::Q];
        ForQ<S::Q>::run([&]<int q>() { fn[q] = vsub(f[q], eq.template get<q>()); });
        const SecondMoments<S, V> m(fn);
        // S_shear = Pi_neq / (2 rho cs^2); sum over ALL a,b of S_ab^2 (off-diagonals count twice)
        const V k = vrecip(vmul(vset<V>(2.0 * kCs2), rho));
        const V diag = vfma(m.P22, m.P22, vfma(m.P11, m.P11, vmul(m.P00, m.P00)));
        const V off = vfma(m.P12, m.P12, vfma(m.P02, m.P02, vmul(m.P01, m.P01)));
        const V ss0 = vmul(vmul(k, k), vfma(vset<V>(2.0), off, diag));
        const V nu = vsplat<V>((tau_ - T(0.5)) / T(3));
        const V c2 = vsplat<V>(constant_ * constant_);
        V tau_eff = vsplat<V>(tau_);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const V ss = vdiv(ss0, vmul(tau_eff, tau_eff));
            tau_eff = vfma(vfma(c2, ss, nu), vset<V>(3.0), vset<V>(0.5));
        }
        const V inv_tau = vrecip(tau_eff);
        ForQ<S::Q>::run([&]<int q>() { f[q] = vfnma(inv_tau, fn[q], f[q]); });
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

template <class S, class V>
LBM_HD void collide_bgk_forced(V (&f)[S::Q], scalar_t<V> inv_tau, const ForceArgs<scalar_t<V>> &fa) {
    V rho, j[3];
    moments<S, V>(f, rho, j);
    const V inv_rho = vrecip(rho);
    const V k = vmul(vsplat<V>(fa.ueq_scale), inv_rho);
    const V acc[3] = {vsplat<V>(fa.a[0]), vsplat<V>(fa.a[1]), vsplat<V>(fa.a[2])};
    const V u[3] = {vfma(k, acc[0], vmul(j[0], inv_rho)), vfma(k, acc[1], vmul(j[1], inv_rho)),
                    vfma(k, acc[2], vmul(j[2], inv_rho))};
    Equilibrium<S, V> eq(rho, u);
    const V ua = vfma(u[2], acc[2], vfma(u[1], acc[1], vmul(u[0], acc[0])));
    const V omega = vsplat<V>(inv_tau), src_scale = vsplat<V>(fa.src_scale);
    ForQ<S::Q>::run([&]<int q>() {
        const V eu = e_dot<S, V, q>(u), ea = e_dot<S, V, q>(acc);
        // (e - u).a / cs^2 + (e.u)(e.a) / cs^4
        const V src = vfma(vmul(eu, ea), vset<V>(1.0 / (kCs2 * kCs2)), vmul(vsub(ea, ua), vset<V>(1.0 / kCs2)));
        const V relaxed = vfma(omega, eq.template minus<q>(f[q]), f[q]);
        f[q] = vfma(src_scale, vmul(vset<V>(S::w(q)), src), relaxed);
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
