// lbm_moments.cu -- density / velocity fields and the reporters' reductions
// (lettuce/_flow.py:157-204, lettuce/ext/_reporter/observable_reporter.py:27-68,140-158,
// lettuce/util/utility.py:37-99 order=6).
//
// Reductions are deterministic: every block reduces a grid-stride slice with warp
// shuffles in double precision and writes one partial; a single-block second stage
// folds the partials in a fixed order.  No atomics.
#include "lbm_launch.cuh"

namespace lbm {

constexpr int kReduceBlocks = 148 * 8;  // 8 resident CTAs per SM on a 148-SM B200
constexpr int kReduceThreads = 256;

template <class S, class R>
__global__ void moments_kernel(const R *__restrict__ f, R *__restrict__ rho_out, R *__restrict__ u_out, int64_t N) {
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        R g[S::Q];
        ForQ<S::Q>::run([&]<int q>() { g[q] = __ldg(f + q * N + n); });
        R rho, j[3];
        moments<S, R>(g, rho, j);
        if (rho_out) rho_out[n] = rho;
        if (u_out) {
#pragma unroll
            for (int c = 0; c < S::D; ++c) u_out[c * N + n] = j[S::axis_of(c)] / rho;
        }
    }
}

struct FieldStrides {
    int64_t rho[3];  // internal axis order
    int64_t u[4];    // component, internal axes
};

template <class S, class R>
__global__ void equilibrium_kernel(const R *__restrict__ rho, const R *__restrict__ u, FieldStrides st, int n0, int n1,
                                   int n2, R *__restrict__ f) {
    const int64_t N = (int64_t)n0 * n1 * n2;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(n % n2);
        const int y = (int)((n / n2) % n1);
        const int x = (int)(n / ((int64_t)n1 * n2));
        const R r = rho[x * st.rho[0] + y * st.rho[1] + z * st.rho[2]];
        R v[3] = {R(0), R(0), R(0)};
#pragma unroll
        for (int c = 0; c < S::D; ++c) v[S::axis_of(c)] = u[c * st.u[0] + x * st.u[1] + y * st.u[2] + z * st.u[3]];
        Equilibrium<S, R> eq(r, v);
        ForQ<S::Q>::run([&]<int q>() { f[q * N + n] = eq.template get<q>(); });
    }
}

template <class S, class R>
int launch_equilibrium(const R *rho, const int64_t *rs, const R *u, const int64_t *us, int n0, int n1, int n2, R *f,
                       cudaStream_t stream) {
    FieldStrides st;
    for (int a = 0; a < 3; ++a) st.rho[a] = 0;
    for (int a = 0; a < 4; ++a) st.u[a] = 0;
    st.u[0] = us[0];
    for (int c = 0; c < S::D; ++c) {
        st.rho[S::axis_of(c)] = rs[c];
        st.u[1 + S::axis_of(c)] = us[1 + c];
    }
    const int64_t N = (int64_t)n0 * n1 * n2;
    int64_t b = (N + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
    equilibrium_kernel<S, R><<<(int)b, 256, 0, stream>>>(rho, u, st, n0, n1, n2, f);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

template <bool MAX>
LBM_D double warp_fold(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_down_sync(0xffffffffu, v, o);
        v = MAX ? fmax(v, w) : v + w;
    }
    return v;
}

template <bool MAX>
LBM_D void block_fold_store(double v, double *partials) {
    __shared__ double sm[kReduceThreads / 32];
    v = warp_fold<MAX>(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = threadIdx.x < kReduceThreads / 32 ? sm[threadIdx.x] : (MAX ? -1.0e300 : 0.0);
        w = warp_fold<MAX>(w);
        if (threadIdx.x == 0) partials[blockIdx.x] = w;
    }
}

template <bool MAX>
__global__ void fold_partials_kernel(const double *partials, int n, double *out) {
    double v = MAX ? -1.0e300 : 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v = MAX ? fmax(v, partials[i]) : v + partials[i];
    __shared__ double sm[kReduceThreads / 32];
    v = warp_fold<MAX>(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = threadIdx.x < kReduceThreads / 32 ? sm[threadIdx.x] : (MAX ? -1.0e300 : 0.0);
        w = warp_fold<MAX>(w);
        if (threadIdx.x == 0) *out = w;
    }
}

// sum 0.5|u|^2 or max |u| straight from the populations (one pass over f)
template <class S, class R, bool MAX>
__global__ void velocity_reduce_kernel(const R *__restrict__ f, int64_t N, double *partials) {
    double acc = MAX ? -1.0e300 : 0.0;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        R g[S::Q];
        ForQ<S::Q>::run([&]<int q>() { g[q] = __ldg(f + q * N + n); });
        R rho, j[3];
        moments<S, R>(g, rho, j);
        const R u0 = j[0] / rho, u1 = j[1] / rho, u2 = j[2] / rho;
        const R uu = u0 * u0 + u1 * u1 + u2 * u2;
        if (MAX) acc = fmax(acc, (double)sqrt(uu));
        else acc += (double)(R(0.5) * uu);
    }
    block_fold_store<MAX>(acc, partials);
}

// mode 0: all populations; 1: interior of the last two user axes; 2: weighted by a uint8 node mask
template <class R, int MODE>
__global__ void population_sum_kernel(const R *__restrict__ f, int q, int n0, int n1, int n2, int d,
                                      const uint8_t *__restrict__ mask, double *partials) {
    const int64_t N = (int64_t)n0 * n1 * n2;
    double acc = 0.0;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        bool take = true;
        double wgt = 1.0;
        if (MODE == 1) {
            const int z = (int)(n % n2);
            const int y = (int)((n / n2) % n1);
            const int x = (int)(n / ((int64_t)n1 * n2));
            // f[..., 1:-1, 1:-1]: last two axes are (y,z) in 3-D and (x,y)=(axis0,axis2) in 2-D
            if (d == 3) take = (y >= 1 && y < n1 - 1 && z >= 1 && z < n2 - 1);
            else take = (x >= 1 && x < n0 - 1 && z >= 1 && z < n2 - 1);
        }
        if (MODE == 2) {
            wgt = (double)mask[n];
            take = wgt != 0.0;
        }
        if (!take) continue;
        double s = 0.0;
        for (int k = 0; k < q; ++k) s += (double)__ldg(f + k * N + n);
        acc += s * wgt;
    }
    block_fold_store<false>(acc, partials);
}

// 6th-order periodic first derivative (util/utility.py:56-58, 88-98): weights
// -1/60, 3/20, -3/4, 3/4, -3/20, 1/60 on x-3 ... x+3.
template <class R>
LBM_D R d6(const R *__restrict__ a, int64_t base, int i, int n, int64_t stride) {
    auto at = [&](int k) {
        int m = i + k;
        m = m < 0 ? m + n : (m >= n ? m - n : m);
        return __ldg(a + base + (int64_t)(m - i) * stride);
    };
    return R(-1.0 / 60.0) * at(-3) + R(3.0 / 20.0) * at(-2) + R(-3.0 / 4.0) * at(-1) + R(3.0 / 4.0) * at(1) +
           R(-3.0 / 20.0) * at(2) + R(1.0 / 60.0) * at(3);
}

// sum |curl u|^2, u given as [d][n0*n1*n2] in USER component order
template <class R, int D>
__global__ void enstrophy_kernel(const R *__restrict__ u, int n0, int n1, int n2, const uint8_t *__restrict__ mask,
                                 double *partials) {
    const int64_t N = (int64_t)n0 * n1 * n2;
    const int64_t s0 = (int64_t)n1 * n2, s1 = n2, s2 = 1;
    double acc = 0.0;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[n]) continue;  // e.g. the halo planes of a slab extended for the stencil radius
        const int z = (int)(n % n2);
        const int y = (int)((n / n2) % n1);
        const int x = (int)(n / s0);
        if (D == 2) {
            // user axes (x, y) = internal (0, 2)
            const R du0_dy = d6(u, n, z, n2, s2);
            const R du1_dx = d6(u + N, n, x, n0, s0);
            const R w = du0_dy - du1_dx;
            acc += (double)(w * w);
        } else {
            const R du0_dy = d6(u, n, y, n1, s1), du0_dz = d6(u, n, z, n2, s2);
            const R du1_dx = d6(u + N, n, x, n0, s0), du1_dz = d6(u + N, n, z, n2, s2);
            const R du2_dx = d6(u + 2 * N, n, x, n0, s0), du2_dy = d6(u + 2 * N, n, y, n1, s1);
            const R a = du0_dy - du1_dx, b = du2_dy - du1_dz, c = du0_dz - du2_dx;
            acc += (double)(a * a + b * b + c * c);
        }
    }
    block_fold_store<false>(acc, partials);
}

// Initial populations with the first-order non-equilibrium part (lettuce/_flow.py:341-367, Krueger et al. 2017):
//   f_q = feq_q(rho, u) - w_q Q_q : Pi1,   Pi1_ab = tau rho d_b u_a / cs^2,   Q_q,ab = e_qa e_qb - delta_ab eye_cs2
// with 6th-order periodic differences (torch_gradient(order=6)).  rho [N] and u [D][N] (user component order) are
// fields; f is written in one pass, no full-size temporaries.  `eye_cs2` is the value the reference subtracts on
// the diagonal: cs^2 rounded to float32 (torch.eye in torch's default dtype, _flow.py:358-360).
template <class S, class R>
__global__ void init_fneq_kernel(const R *__restrict__ rho, const R *__restrict__ u, R tau_over_cs2, R eye_cs2, int n0,
                                 int n1, int n2, R *__restrict__ f) {
    constexpr int D = S::D;
    const int64_t N = (int64_t)n0 * n1 * n2;
    const int64_t stride[3] = {(int64_t)n1 * n2, n2, 1};
    const int extent[3] = {n0, n1, n2};
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const int idx[3] = {(int)(n / stride[0]), (int)((n / n2) % n1), (int)(n % n2)};
        const R r = rho[n];
        R v[3] = {R(0), R(0), R(0)};      // internal axis order for the equilibrium
        R pi[D][D];                       // user component / axis order
#pragma unroll
        for (int a = 0; a < D; ++a) {
            v[S::axis_of(a)] = u[a * N + n];
#pragma unroll
            for (int b = 0; b < D; ++b) {
                const int ax = S::axis_of(b);
                pi[a][b] = tau_over_cs2 * r * d6(u + a * N, n, idx[ax], extent[ax], stride[ax]);
            }
        }
        Equilibrium<S, R> eq(r, v);
        ForQ<S::Q>::run([&]<int q>() {
            R piq = R(0);
#pragma unroll
            for (int a = 0; a < D; ++a) {
#pragma unroll
                for (int b = 0; b < D; ++b) {
                    const R qab = R(S::e(q, S::axis_of(a)) * S::e(q, S::axis_of(b))) - (a == b ? eye_cs2 : R(0));
                    piq += pi[a][b] * qab;
                }
            }
            f[q * N + n] = eq.template get<q>() - R(S::w(q)) * piq;
        });
    }
}

template <class S, class R>
int launch_init_fneq(const R *rho, const R *u, double tau_over_cs2, double eye_cs2, int n0, int n1, int n2, R *f,
                     cudaStream_t stream) {
    const int64_t N = (int64_t)n0 * n1 * n2;
    int64_t b = (N + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
    init_fneq_kernel<S, R><<<(int)b, 256, 0, stream>>>(rho, u, (R)tau_over_cs2, (R)eye_cs2, n0, n1, n2, f);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// host-side launch helpers used by lbm_api.cu
// ---------------------------------------------------------------------------
static int grid_for(int64_t n) {
    int64_t b = (n + kReduceThreads - 1) / kReduceThreads;
    if (b > kReduceBlocks) b = kReduceBlocks;
    if (b < 1) b = 1;
    return (int)b;
}

template <class S, class R>
int launch_moments(const R *f, R *rho, R *u, int64_t N, cudaStream_t st) {
    int64_t b = (N + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
    moments_kernel<S, R><<<(int)b, 256, 0, st>>>(f, rho, u, N);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

template <class S, class R>
int launch_reduce(int what, const R *in, const uint8_t *mask, int n0, int n1, int n2, double *partials, double *out,
                  cudaStream_t st) {
    const int64_t N = (int64_t)n0 * n1 * n2;
    const int g = grid_for(N);
    bool is_max = false;
    switch (what) {
        case LBM_SUM_HALF_U2: velocity_reduce_kernel<S, R, false><<<g, kReduceThreads, 0, st>>>(in, N, partials); break;
        case LBM_MAX_U:
            velocity_reduce_kernel<S, R, true><<<g, kReduceThreads, 0, st>>>(in, N, partials);
            is_max = true;
            break;
        case LBM_SUM_F:
            population_sum_kernel<R, 0><<<g, kReduceThreads, 0, st>>>(in, S::Q, n0, n1, n2, S::D, nullptr, partials);
            break;
        case LBM_SUM_F_INNER:
            population_sum_kernel<R, 1><<<g, kReduceThreads, 0, st>>>(in, S::Q, n0, n1, n2, S::D, nullptr, partials);
            break;
        case LBM_SUM_F_MASKED:
            if (!mask) return LBM_ERR_BAD_ARGUMENT;
            population_sum_kernel<R, 2><<<g, kReduceThreads, 0, st>>>(in, S::Q, n0, n1, n2, S::D, mask, partials);
            break;
        case LBM_ENSTROPHY:
            enstrophy_kernel<R, S::D><<<g, kReduceThreads, 0, st>>>(in, n0, n1, n2, mask, partials);
            break;
        default: return LBM_ERR_BAD_ARGUMENT;
    }
    ++g_launch_count;
    int e = (int)cudaGetLastError();
    if (e) return e;
    if (is_max) fold_partials_kernel<true><<<1, kReduceThreads, 0, st>>>(partials, g, out);
    else fold_partials_kernel<false><<<1, kReduceThreads, 0, st>>>(partials, g, out);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

#define LBM_INSTANTIATE(S, R)                                                        \
    template int launch_equilibrium<S, R>(const R *, const int64_t *, const R *, const int64_t *, int, int, int, R *, \
                                          cudaStream_t);                              \
    template int launch_moments<S, R>(const R *, R *, R *, int64_t, cudaStream_t);   \
    template int launch_init_fneq<S, R>(const R *, const R *, double, double, int, int, int, R *, cudaStream_t); \
    template int launch_reduce<S, R>(int, const R *, const uint8_t *, int, int, int, double *, double *, cudaStream_t);
LBM_INSTANTIATE(D2Q9, float)
LBM_INSTANTIATE(D2Q9, double)
LBM_INSTANTIATE(D3Q19, float)
LBM_INSTANTIATE(D3Q19, double)
LBM_INSTANTIATE(D3Q27, float)
LBM_INSTANTIATE(D3Q27, double)

size_t reduce_scratch_bytes() { return sizeof(double) * (size_t)kReduceBlocks; }

// sum of n partials in a fixed order -> *out.  More than 16 per thread are first folded by kReduceBlocks CTAs into
// `stage` (kReduceBlocks doubles), so that one CTA never walks hundreds of dependent iterations.
__global__ void fold_stage_kernel(const double *__restrict__ partials, int n, double *stage) {
    double v = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += partials[i];
    block_fold_store<false>(v, stage);
}

int launch_fold_sum(const double *partials, int n, double *stage, double *out, cudaStream_t st) {
    if (n > 16 * kReduceThreads) {
        const int g = grid_for(n / 4);
        fold_stage_kernel<<<g, kReduceThreads, 0, st>>>(partials, n, stage);
        ++g_launch_count;
        int e = (int)cudaGetLastError();
        if (e) return e;
        partials = stage;
        n = g;
    }
    fold_partials_kernel<false><<<1, kReduceThreads, 0, st>>>(partials, n, out);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

// (sum, max) partial pairs of a step with fused reductions (StepParams::energy_partials: sums in [0, n), maxima in
// [n, 2n)) -> out[0] = sum in a fixed order, out[1] = max.  Two stages above 16 partials per thread, as above.
__global__ void fold_pair_stage_kernel(const double *__restrict__ partials, int n, double *stage) {
    double s = 0.0, m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        s += partials[i];
        m = fmax(m, partials[n + i]);
    }
    block_fold_store<false>(s, stage);
    __syncthreads();
    block_fold_store<true>(m, stage + gridDim.x);
}

__global__ void fold_pair_final_kernel(const double *__restrict__ partials, int n, double *out) {
    double s = 0.0, m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s += partials[i];
        m = fmax(m, partials[n + i]);
    }
    __shared__ double sm[2][kReduceThreads / 32];
    s = warp_fold<false>(s);
    m = warp_fold<true>(m);
    if ((threadIdx.x & 31) == 0) {
        sm[0][threadIdx.x >> 5] = s;
        sm[1][threadIdx.x >> 5] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < kReduceThreads / 32; ++w) {
            a += sm[0][w];
            b = fmax(b, sm[1][w]);
        }
        out[0] = a;
        out[1] = b;
    }
}

size_t fold_pair_stage_bytes() { return 2 * sizeof(double) * (size_t)kReduceBlocks; }

int launch_fold_pair(const double *partials, int n, double *stage, double *out, cudaStream_t st) {
    if (n > 16 * kReduceThreads) {
        const int g = grid_for(n / 4);
        fold_pair_stage_kernel<<<g, kReduceThreads, 0, st>>>(partials, n, stage);
        ++g_launch_count;
        const int e = (int)cudaGetLastError();
        if (e) return e;
        partials = stage;
        n = g;
    }
    fold_pair_final_kernel<<<1, kReduceThreads, 0, st>>>(partials, n, out);
    ++g_launch_count;
    return (int)cudaGetLastError();
}

}  // namespace lbm
